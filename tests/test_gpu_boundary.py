"""Boundary behaviour of the C ABI that the reference-side binding relies on (SURVEY.md §8b Ownership / Errors):
buffers that outlive t5_shutdown, the host-array pipeline on records too large for a 256-item chunk, the bounded
index-table cache, SoA operands of AoS-only kernels (the in-library relayout detour) and a dot_sum that does not
depend on SoA padding contents.  Every numeric result is compared with the CPU oracle."""
import ctypes

import numpy as np
import pytest

import oracle
import tensor_b200 as tb
from tensor_b200 import _lib, host
from tensor_b200 import structure as st

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def test_buffers_may_outlive_shutdown():
    """t5_shutdown with live buffers: they become orphans (operations refused, late t5_free legal)."""
    L = _lib.lib()
    h = ctypes.c_void_p()
    dev = (ctypes.c_int * 1)(0)
    assert L.t5_init(dev, 1, ctypes.byref(h)) == _lib.OK
    bufs = []
    for layout in (_lib.AOS, _lib.SOA):
        b = ctypes.c_void_p()
        assert L.t5_alloc(h, _lib.F64, 9, 100_000, layout, ctypes.byref(b)) == _lib.OK
        bufs.append(b)
    assert L.t5_live_buffers(h) == 2
    assert L.t5_fill_synthetic(bufs[0], 1, 1) == _lib.OK          # work in flight on the buffer when the context goes
    assert L.t5_shutdown(h) == _lib.OK
    out = np.empty(900_000)
    assert L.t5_download(bufs[0], out.ctypes.data_as(ctypes.c_void_p), out.nbytes) == _lib.ERR_INVALID
    assert L.t5_fill_synthetic(bufs[1], 1, 1) == _lib.ERR_INVALID
    assert L.t5_shutdown(h) == _lib.ERR_INVALID                    # the struct lives on for the orphans: detected
    assert L.t5_free(bufs[0]) == _lib.OK
    assert L.t5_free(bufs[1]) == _lib.OK                           # the last orphan releases the context struct
    # and the device is still healthy afterwards
    with tb.Context([0]) as c:
        a = oracle.fill(oracle.F64, oracle.SEED_A, 1, 9 * 1000)
        assert np.array_equal(tb.transpose(c.from_numpy(a, (3, 3))).to_numpy(), oracle.transpose(a, [3, 3]))


def test_host_pipeline_large_records(ctx):
    """Records of 1 MiB: a chunk holds fewer than 256 items and staging is sized by the call, not by the cap."""
    n, batch = 512, 5
    a = oracle.fill(oracle.F32, oracle.SEED_A, 41, batch * n * n)
    b = oracle.fill(oracle.F32, oracle.SEED_B, 41, batch * n * n)
    got = host.matmul(ctx, a, b, n, n, n)
    a64, b64 = a.reshape(batch, n, n).astype(np.float64), b.reshape(batch, n, n).astype(np.float64)
    bound = 1e-5 * np.einsum("bij,bjk->bik", np.abs(a64), np.abs(b64))
    assert np.all(np.abs(got.astype(np.float64) - np.einsum("bij,bjk->bik", a64, b64)) <= bound)
    t = host.transpose(ctx, a, (n, n))
    assert np.array_equal(bits(t), bits(oracle.transpose(a, [n, n])))


def test_host_pipeline_error_leaves_nothing_in_flight(ctx):
    """An unsupported request inside the pipeline returns its code after draining; the context keeps working."""
    a = oracle.fill(oracle.F64, oracle.SEED_A, 42, 1000 * 9).view(np.int64)
    with pytest.raises((tb.Unsupported, TypeError, ValueError)):
        host.det(ctx, a, 3)                                         # det has no Int64 kernels
    x = oracle.fill(oracle.F64, oracle.SEED_A, 42, 1000 * 9)
    assert np.array_equal(bits(host.det(ctx, x, 3)), bits(oracle.det(x.reshape(1000, 3, 3), 3)))


def test_index_table_cache_is_bounded(ctx):
    """`ta i` over every row of a tall tensor: one table per i; beyond 256 cached tables the oldest are released
    and results stay correct (also when an evicted table is asked for again)."""
    rows, batch = 300, 64
    a = oracle.fill(oracle.F64, oracle.SEED_A, 43, batch * rows * 4).reshape(batch, rows, 4)
    A = ctx.from_numpy(a, (rows, 4))
    for i in list(range(1, rows + 1)) + [1, 2, 3]:
        got = st.ta(i, A)
        assert np.array_equal(got.to_numpy(), a[:, i - 1, :]), i
        got.free()


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.int64])
def test_soa_large_items_take_the_aos_kernels(ctx, dt):
    """SoA operands of shapes whose kernels are AoS-only run through the in-library relayout detour and give the
    AoS result (bit-identical for the exact kernels; the tensor-core shapes are compared with their own AoS run)."""
    code = tb.device.DTYPE_OF[np.dtype(dt)]
    oc = {tb.F64: oracle.F64, tb.F32: oracle.F32, tb.I64: oracle.I64}[code]
    for (m, k, n, batch) in [(32, 32, 32, 301), (17, 33, 9, 257), (64, 64, 64, 67), (128, 128, 128, 9), (24, 24, 24, 100)]:
        a = oracle.fill(oc, oracle.SEED_A, 44, batch * m * k)
        b = oracle.fill(oc, oracle.SEED_B, 44, batch * k * n)
        aos = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
        soa = tb.matmul(ctx.from_numpy(a, (m, k), tb.SOA), ctx.from_numpy(b, (k, n), tb.SOA))
        assert soa.layout == tb.SOA
        assert np.array_equal(bits(soa.to_numpy()), bits(aos)), (m, k, n)
        if tb.matmul_is_exact(code, m, k, n):
            assert np.array_equal(bits(aos), bits(oracle.matmul(a, b, m, k, n, threads=8))), (m, k, n)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_soa_elimination_beyond_the_tile_kernels(ctx, dt):
    oc = oracle.F64 if dt == np.float64 else oracle.F32
    batch = 1003
    for n, nrhs in [(6, 5), (7, 3), (8, 7)]:                       # column counts only the AoS row-per-lane kernel has
        a = oracle.fill(oc, oracle.SEED_A, 45, batch * n * n).reshape(batch, n, n).copy()
        oracle.plant_specials(a, n, 0, swap_every=16, sing_every=128)
        b = oracle.fill(oc, oracle.SEED_B, 45, batch * n * nrhs)
        x, s = tb.solve(ctx.from_numpy(a, (n, n), tb.SOA), ctx.from_numpy(b, (n, nrhs), tb.SOA))
        ox, os_ = oracle.solve(a, b, n, nrhs, threads=8)
        assert x.layout == tb.SOA
        assert np.array_equal(s.to_numpy(), os_) and np.array_equal(bits(x.to_numpy()), bits(ox)), (n, nrhs)
    a = oracle.fill(oc, oracle.SEED_A, 46, batch * 7 * 9)
    e, r = tb.row_echelon_form(ctx.from_numpy(a, (7, 9), tb.SOA))    # any-shape echelon kernel: AoS only
    oe, orank = oracle.row_echelon(a, 7, 9, 9, threads=8)
    assert np.array_equal(r.to_numpy(), orank) and np.array_equal(bits(e.to_numpy()), bits(oe))


@pytest.mark.parametrize("batch", [100, 4097])
def test_dot_sum_ignores_soa_padding(ctx, batch):
    """An element-wise op leaves 0 * inf = NaN in the 32-item row padding of a SoA buffer; the sum over the batch
    must not see it."""
    ones = np.ones((batch, 4))
    X = ctx.from_numpy(ones, (4,), tb.SOA)
    W = tb.scale(np.float64(np.inf), X)                               # data +inf, padding 0 * inf = NaN
    got = tb.dot_sum(W, X)
    assert got == np.inf
    xi = oracle.fill(oracle.I64, oracle.SEED_A, 47, batch * 4)
    yi = oracle.fill(oracle.I64, oracle.SEED_B, 47, batch * 4)
    Xi, Yi = ctx.from_numpy(xi, (4,), tb.SOA), ctx.from_numpy(yi, (4,), tb.SOA)
    Zi = tb.zip_with("+", Xi, Yi)                                     # (pad stays 0 + 0 here; the masked kernel is exercised)
    assert tb.dot_sum(Zi, Yi) == oracle.dot_sum((xi + yi), yi, 4)
