"""Multi-GPU tests of the single-process, multi-device context (the way one Haskell
process drives 8 GPUs): the batch axis is sharded, no collective on the product /
elimination paths, and the dot-sum partials meet in a 1-element NCCL all-reduce.
Skipped when fewer than 2 GPUs are visible (run with `gpurun --gpus 2`)."""
import numpy as np
import pytest

import oracle
import tensor_b200 as tb
from tensor_b200 import _lib

pytestmark = pytest.mark.gpu


def ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def mctx():
    n = ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    c = tb.Context(list(range(min(n, 8))))
    yield c
    c.close()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


@pytest.mark.parametrize("batch", [1, 255, 100003, 1 << 20])
def test_sharded_ops_match_oracle(mctx, batch):
    assert mctx.ndev >= 2
    a = oracle.fill(oracle.F64, oracle.SEED_A, 1, batch * 9)
    b = oracle.fill(oracle.F64, oracle.SEED_B, 1, batch * 9)
    A, B = mctx.from_numpy(a, (3, 3)), mctx.from_numpy(b, (3, 3))
    assert np.array_equal(bits(tb.matmul(A, B).to_numpy()), bits(oracle.matmul(a, b, 3, 3, 3, threads=8)))
    assert np.array_equal(bits(tb.transpose(A).to_numpy()), bits(oracle.transpose(a, [3, 3])))
    am = a.reshape(batch, 3, 3).copy()
    oracle.plant_specials(am, 3, 0, swap_every=16, sing_every=128)
    inv, st = tb.inverse(mctx.from_numpy(am, (3, 3)))
    oinv, ost = oracle.inverse(am, 3, threads=8)
    assert np.array_equal(st.to_numpy(), ost) and np.array_equal(bits(inv.to_numpy()), bits(oinv))


@pytest.mark.parametrize("batch", [3, 100003])
def test_sharded_wider_surface_matches_oracle(mctx, batch):
    """The ops either side of the hot path on a sharded batch: every device runs its own range, results meet in
    one download - rowEchelonForm, 8x8 inverse (streamed-result kernel), long dot, staged transpose, split, pure."""
    from tensor_b200 import structure as st
    a4 = oracle.fill(oracle.F64, oracle.SEED_A, 31, batch * 20).reshape(batch, 4, 5).copy()
    a4[::5, 2] = a4[::5, 0]                                             # exact rank defects
    e, r = tb.row_echelon_form(mctx.from_numpy(a4, (4, 5)), 4)
    oe, orank = oracle.row_echelon(a4, 4, 5, 4, threads=8)
    assert np.array_equal(r.to_numpy(), orank) and np.array_equal(bits(e.to_numpy()), bits(oe))
    a8 = oracle.fill(oracle.F64, oracle.SEED_A, 32, batch * 64).reshape(batch, 8, 8).copy()
    oracle.plant_specials(a8, 8, 0, swap_every=16, sing_every=128)
    A8 = mctx.from_numpy(a8, (8, 8))
    inv, stt = tb.inverse(A8)
    oinv, ost = oracle.inverse(a8, 8, threads=8)
    assert np.array_equal(stt.to_numpy(), ost) and np.array_equal(bits(inv.to_numpy()), bits(oinv))
    b8 = oracle.fill(oracle.F64, oracle.SEED_B, 32, batch * 64)
    assert np.array_equal(bits(tb.prod(A8, mctx.from_numpy(b8, (8, 8)), 2).to_numpy()),
                          bits(oracle.prod(a8.ravel(), [8, 8], b8, [8, 8], 2)))
    a16 = oracle.fill(oracle.F32, oracle.SEED_A, 33, batch * 16 * 20)
    assert np.array_equal(bits(tb.transpose(mctx.from_numpy(a16, (16, 20))).to_numpy()), bits(oracle.transpose(a16, [16, 20])))
    l, rgt = st.split(2, 3, mctx.from_numpy(a4, (4, 5)))
    assert np.array_equal(l.to_numpy(), a4[:, :, :3]) and np.array_equal(rgt.to_numpy(), a4[:, :, 3:])
    u = tb.unit(mctx, 3, np.int64, batch)
    assert np.array_equal(u.to_numpy(), np.broadcast_to(np.eye(3, dtype=np.int64), (batch, 3, 3)))


def test_device_synthetic_stream_is_global(mctx):
    batch = 300001
    A = mctx.synthetic((4, 4), np.float32, batch, oracle.SEED_A, 3)
    assert np.array_equal(bits(A.to_numpy().ravel()), bits(oracle.fill(oracle.F32, oracle.SEED_A, 3, batch * 16)))


def test_dot_sum_allreduce(mctx):
    batch, n = 1 << 20, 4
    xi = oracle.fill(oracle.I64, oracle.SEED_A, 4, batch * n)
    yi = oracle.fill(oracle.I64, oracle.SEED_B, 4, batch * n)
    got = tb.dot_sum(mctx.from_numpy(xi, (n,)), mctx.from_numpy(yi, (n,)))
    assert got == oracle.dot_sum(xi, yi, n)               # Int64: order-independent, bit-exact
    assert _lib.lib().t5_dot_sum_used_nccl(mctx._h) == 1, "multi-device dot_sum must go through NCCL"
    x = oracle.fill(oracle.F64, oracle.SEED_A, 4, batch * n)
    y = oracle.fill(oracle.F64, oracle.SEED_B, 4, batch * n)
    got = tb.dot_sum(mctx.from_numpy(x, (n,)), mctx.from_numpy(y, (n,)))
    assert abs(got - oracle.dot_sum(x, y, n)) <= 1e-12 * np.sum(np.abs(x * y))


def test_host_entry_multi_device(mctx):
    batch = 500001
    a = oracle.fill(oracle.F32, oracle.SEED_A, 14, batch * 16)
    b = oracle.fill(oracle.F32, oracle.SEED_B, 14, batch * 16)
    out = np.empty(batch * 16, dtype=np.float32)
    tb.matmul_host(mctx, a, b, 4, 4, 4, out)
    assert np.array_equal(bits(out), bits(oracle.matmul(a, b, 4, 4, 4, threads=8).ravel()))


def test_host_entries_one_thread_per_device(mctx):
    """Host-array entries on a multi-device context: one host thread per device inside the call, every device's
    slots drained before it returns; the rank-3 transpose also reaches the shared index-table cache from them."""
    from tensor_b200 import host
    batch = 300_001
    a3 = oracle.fill(oracle.F64, oracle.SEED_A, 51, batch * 24)
    assert np.array_equal(bits(host.transpose(mctx, a3, (2, 3, 4))), bits(oracle.transpose(a3, [2, 3, 4], threads=8)))
    n = 4
    a = oracle.fill(oracle.F64, oracle.SEED_A, 52, batch * n * n).reshape(batch, n, n).copy()
    oracle.plant_specials(a, n, 0, swap_every=16, sing_every=128)
    b = oracle.fill(oracle.F64, oracle.SEED_B, 52, batch * n)
    x, s = host.solve(mctx, a, b, n)
    ox, os_ = oracle.solve(a, b, n, 1, threads=8)
    assert np.array_equal(s, os_) and np.array_equal(bits(x), bits(ox))
    inv, s = host.inverse(mctx, a, n)
    oinv, os_ = oracle.inverse(a, n, threads=8)
    assert np.array_equal(s, os_) and np.array_equal(bits(inv), bits(oinv))


def test_per_device_timer_and_partition(mctx):
    batch = 1 << 20
    A = mctx.synthetic((4, 4), np.float32, batch, oracle.SEED_A, 1)
    B = mctx.synthetic((4, 4), np.float32, batch, oracle.SEED_B, 1)
    mctx.sync()
    mctx.timer_begin()
    C = tb.matmul(A, B)
    ms = mctx.timer_end()
    per = mctx.timer_per_device()
    assert len(per) == mctx.ndev and max(per) == pytest.approx(ms, rel=1e-6) and min(per) > 0
    shares = [tb.partition(batch, mctx.ndev, g)[1] for g in range(mctx.ndev)]
    assert sum(shares) == batch and max(shares) - min(shares) <= 256 * mctx.ndev
    a, b = A.to_numpy().ravel(), B.to_numpy().ravel()
    assert np.array_equal(bits(C.to_numpy()), bits(oracle.matmul(a, b, 4, 4, 4, threads=8)))


def test_soa_dot_sum_and_detour_multi_device(mctx):
    batch = 100_003                                                   # ragged: every device has row padding
    x = oracle.fill(oracle.I64, oracle.SEED_A, 53, batch * 4)
    y = oracle.fill(oracle.I64, oracle.SEED_B, 53, batch * 4)
    assert tb.dot_sum(mctx.from_numpy(x, (4,), tb.SOA), mctx.from_numpy(y, (4,), tb.SOA)) == oracle.dot_sum(x, y, 4)
    m = 32                                                            # Int64: the detour's kernel is the exact blocked product
    a = oracle.fill(oracle.I64, oracle.SEED_A, 54, 999 * m * m)
    b = oracle.fill(oracle.I64, oracle.SEED_B, 54, 999 * m * m)
    got = tb.matmul(mctx.from_numpy(a, (m, m), tb.SOA), mctx.from_numpy(b, (m, m), tb.SOA))
    assert np.array_equal(bits(got.to_numpy()), bits(oracle.matmul(a, b, m, m, m, threads=8)))
    ad = oracle.fill(oracle.F64, oracle.SEED_A, 54, 999 * m * m)      # Double 32 x 32: four items to a DMMA tile on every device
    bd = oracle.fill(oracle.F64, oracle.SEED_B, 54, 999 * m * m)
    got = tb.matmul(mctx.from_numpy(ad, (m, m), tb.SOA), mctx.from_numpy(bd, (m, m), tb.SOA)).to_numpy().reshape(999, m, m)
    want = oracle.matmul(ad, bd, m, m, m, threads=8).reshape(999, m, m)
    assert np.all(np.abs(got - want) <= 1e-12 * np.einsum("bij,bjk->bik", np.abs(ad.reshape(999, m, m)), np.abs(bd.reshape(999, m, m))))


def test_multi_device_buffers_may_outlive_shutdown():
    import ctypes
    n = ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.t5_init(None, 0, ctypes.byref(h)) == _lib.OK
    b = ctypes.c_void_p()
    assert L.t5_alloc(h, _lib.F32, 16, 1 << 20, _lib.AOS, ctypes.byref(b)) == _lib.OK
    assert L.t5_fill_synthetic(b, 1, 1) == _lib.OK
    assert L.t5_shutdown(h) == _lib.OK
    assert L.t5_fill_synthetic(b, 1, 1) == _lib.ERR_INVALID
    assert L.t5_free(b) == _lib.OK
