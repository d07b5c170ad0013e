"""The host-array (end-to-end) entries t5_*_host: HOST buffers in, HOST buffers out, the batch
streamed through the GPU in chunks on three slots.  Sizes are chosen to span several chunks
(so that slot reuse, ragged last chunks and result ordering are exercised) and every result
is compared with the CPU oracle bit for bit."""
import numpy as np
import pytest

import oracle
import tensor_b200 as tb
from tensor_b200 import host

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def same(got, want, what):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.array_equal(bits(got), bits(want)), what


@pytest.mark.parametrize("dc", [oracle.F64, oracle.F32, oracle.I64])
def test_matmul_dot_prod_host(ctx, dc):
    batch = 400_003                                   # 4x4 Float: ~2.3 chunks of 174 592 items
    a = oracle.fill(dc, oracle.SEED_A, 21, batch * 16)
    b = oracle.fill(dc, oracle.SEED_B, 21, batch * 16)
    same(host.matmul(ctx, a, b, 4, 4, 4), oracle.matmul(a, b, 4, 4, 4, threads=8), "matmul_host 4x4")
    same(host.matmul(ctx, a, b[: batch * 4], 4, 4, 1), oracle.matmul(a, b[: batch * 4], 4, 4, 1, threads=8), "matvec_host")
    x, y = a[: batch * 4], b[: batch * 4]
    same(host.dot(ctx, x, y, 4), oracle.dot(x, y, 4, threads=8), "dot_host")
    small = 5_003
    same(host.prod(ctx, a[: small * 6], (2, 3), b[: small * 4], (2, 2), 0),
         oracle.prod(a[: small * 6], [2, 3], b[: small * 4], [2, 2], 0), "tensor product host")
    same(host.prod(ctx, a[: small * 6], (3, 2), b[: small * 4], (2, 2), 1),
         oracle.prod(a[: small * 6], [3, 2], b[: small * 4], [2, 2], 1), "prod 1 host")


def test_transpose_host(ctx):
    batch = 300_001
    a = oracle.fill(oracle.F32, oracle.SEED_A, 22, batch * 16)
    same(host.transpose(ctx, a, (4, 4)), oracle.transpose(a, [4, 4]), "transpose 4x4 host")
    a3 = oracle.fill(oracle.F64, oracle.SEED_A, 22, 70_001 * 24)
    same(host.transpose(ctx, a3, (2, 3, 4)), oracle.transpose(a3, [2, 3, 4]), "transpose rank 3 host (gather table)")
    v = oracle.fill(oracle.I64, oracle.SEED_A, 22, 1000 * 5)
    same(host.transpose(ctx, v, (5,)), v.reshape(1000, 5), "rank 1 is the identity")


@pytest.mark.parametrize("dc", [oracle.F64, oracle.F32])
@pytest.mark.parametrize("n", [3, 4, 6])
def test_det_inverse_solve_host(ctx, dc, n):
    batch = 150_001 if n <= 4 else 20_001
    a = oracle.fill(dc, oracle.SEED_A, 23, batch * n * n).reshape(batch, n, n).copy()
    oracle.plant_specials(a, n, 0, swap_every=16, sing_every=128)
    b = oracle.fill(dc, oracle.SEED_B, 23, batch * n)
    same(host.det(ctx, a, n), oracle.det(a, n, threads=8), f"det_host {n}")
    inv, st = host.inverse(ctx, a, n)
    oinv, ost = oracle.inverse(a, n, threads=8)
    assert np.array_equal(st, ost)
    same(inv, oinv, f"inverse_host {n}")
    x, st = host.solve(ctx, a, b, n, 1)
    ox, ost = oracle.solve(a, b, n, 1, threads=8)
    assert np.array_equal(st, ost)
    same(x, ox, f"solve_host {n}")
    assert ost.sum() == (batch + 127) // 128


def test_host_entries_pinned_and_out(ctx):
    batch = 200_000
    pa, pb, pc = (tb.PinnedArray((batch * 9,), np.float64) for _ in range(3))
    pa.array[:] = oracle.fill(oracle.F64, oracle.SEED_A, 24, batch * 9)
    pb.array[:] = oracle.fill(oracle.F64, oracle.SEED_B, 24, batch * 9)
    r = host.matmul(ctx, pa.array, pb.array, 3, 3, 3, out=pc.array)
    assert r is pc.array
    same(pc.array.reshape(batch, 3, 3), oracle.matmul(pa.array, pb.array, 3, 3, 3, threads=8), "pinned in/out")
    for p in (pa, pb, pc):
        p.free()


def test_host_entries_edge_cases_and_errors(ctx):
    e = np.empty((0,), dtype=np.float64)
    assert host.matmul(ctx, e, e, 3, 3, 3).shape == (0, 3, 3)           # empty batch
    assert host.det(ctx, e, 3).shape == (0,)
    one = np.arange(9, dtype=np.float64)
    same(host.matmul(ctx, one, one, 3, 3, 3), oracle.matmul(one, one, 3, 3, 3), "batch of one")
    with pytest.raises(tb.WrongListLength):
        host.matmul(ctx, np.zeros(17), np.zeros(18), 3, 3, 3)
    with pytest.raises(tb.WrongListLength):
        host.prod(ctx, np.zeros(6), (2, 3), np.zeros(4), (2, 2), 1)       # contracted extents differ
    with pytest.raises(tb.Unsupported):
        host.det(ctx, np.zeros(9, dtype=np.int64), 3)                    # needs Fractional
    with pytest.raises(tb.Unsupported):
        host.inverse(ctx, np.zeros(129 * 129 * 2), 129)                  # n > 128: no kernel, no CPU fallback
    with pytest.raises(TypeError):
        host.matmul(ctx, np.zeros(9, dtype=np.int32), np.zeros(9, dtype=np.int32), 3, 3, 3)


def test_host_entry_large_items_dmma(ctx):
    # 128x128 Double: 384 KiB per item, chunks of 256 items -> the DMMA kernel inside the pipeline
    batch = 300
    a = oracle.fill(oracle.F64, oracle.SEED_A, 25, batch * 128 * 128)
    b = oracle.fill(oracle.F64, oracle.SEED_B, 25, batch * 128 * 128)
    got = host.matmul(ctx, a, b, 128, 128, 128)
    A, B = a.reshape(batch, 128, 128), b.reshape(batch, 128, 128)
    want = np.einsum("bij,bjk->bik", A, B)
    bound = np.einsum("bij,bjk->bik", np.abs(A), np.abs(B))
    assert np.all(np.abs(got - want) <= 1e-12 * bound)


def test_host_entry_large_items_tcgen05(ctx):
    # 128x128 Float through the host pipeline: the tcgen05 3xTF32 kernel per chunk, a fresh pair of TMA
    # tensor maps per launch; 700 items = several chunks over the three slots
    batch = 700
    a = oracle.fill(oracle.F32, oracle.SEED_A, 26, batch * 128 * 128)
    b = oracle.fill(oracle.F32, oracle.SEED_B, 26, batch * 128 * 128)
    got = host.matmul(ctx, a, b, 128, 128, 128).astype(np.float64)
    A, B = a.reshape(batch, 128, 128).astype(np.float64), b.reshape(batch, 128, 128).astype(np.float64)
    want = np.einsum("bij,bjk->bik", A, B)
    bound = np.einsum("bij,bjk->bik", np.abs(A), np.abs(B))
    assert np.all(np.abs(got.reshape(want.shape) - want) <= 1e-5 * bound)
