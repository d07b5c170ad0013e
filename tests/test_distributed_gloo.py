"""world_size-2 gloo test (CPU) of the N>1 path: batch sharding + the dot-sum exchange.
Each rank reduces only its own shard (here with the oracle standing in for the device
kernel — this test checks the host-side partition / all-reduce logic, not the kernels)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import oracle
    from tensor_b200 import distributed as D
    r, _, w = D.init_process_group("gloo")
    first, count = D.strong_shard(batch, w, r)
    n = 4
    out = {}
    for name, dc, npd in (("i64", oracle.I64, np.int64), ("f64", oracle.F64, np.float64)):
        # each rank generates ONLY its slice of the global counter-based stream
        x = oracle.fill(dc, oracle.SEED_A, 4, count * n, first_elem=first * n)
        y = oracle.fill(dc, oracle.SEED_B, 4, count * n, first_elem=first * n)
        part = oracle.dot_sum(x, y, n) if count else npd(0)
        out[name] = D.all_reduce_scalar(part, npd)
    out["tmax"] = D.max_over_ranks(10.0 + rank)
    out["shard"] = (first, count)
    D.barrier()
    q.put((rank, out))


def test_two_rank_shard_and_dot_sum():
    import oracle
    batch, world = 100003, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 4
    xi = oracle.fill(oracle.I64, oracle.SEED_A, 4, batch * n)
    yi = oracle.fill(oracle.I64, oracle.SEED_B, 4, batch * n)
    x = oracle.fill(oracle.F64, oracle.SEED_A, 4, batch * n)
    y = oracle.fill(oracle.F64, oracle.SEED_B, 4, batch * n)
    for r in range(world):
        assert res[r]["i64"] == oracle.dot_sum(xi, yi, n)            # Int64: order-independent, bit-exact
        assert abs(res[r]["f64"] - oracle.dot_sum(x, y, n)) <= 1e-12 * np.sum(np.abs(x * y))
        assert res[r]["tmax"] == 11.0
    (f0, c0), (f1, c1) = res[0]["shard"], res[1]["shard"]
    assert f0 == 0 and f1 == c0 and c0 + c1 == batch and c0 % 256 == 0
