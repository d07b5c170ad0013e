"""GPU parity tests: every call goes through the C ABI (ctypes) into the CUDA kernels and
is compared with the CPU oracle on the same seeded inputs.  Integer work and every
thread-per-item floating-point kernel are compared BIT-EXACTLY (the kernels keep the
oracle's operation order and never fuse multiply-add); the DMMA 128x128 product uses
the stated tolerance (1e-12 relative to sum |a||b|)."""
import json
import os

import numpy as np
import pytest

import oracle
import tensor_b200 as tb
from _parity import assert_matmul_matches, matmul_is_exact

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NPD = {oracle.F64: np.float64, oracle.F32: np.float32, oracle.I64: np.int64}
ALL = [oracle.F64, oracle.F32, oracle.I64]
FRAC = [oracle.F64, oracle.F32]
RAGGED = [0, 1, 31, 127, 128, 129, 255, 257, 1000, 4099]


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def assert_bits_equal(got, want, what=""):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if not np.array_equal(bits(got), bits(want)):
        bad = np.argwhere(bits(got).ravel() != bits(want).ravel()).ravel()
        raise AssertionError(f"{what}: {bad.size} mismatching elements, first at {bad[:5]}: "
                             f"got {got.ravel()[bad[:5]]} want {want.ravel()[bad[:5]]}")


def inputs(dc, op_id, batch, da, db=None):
    a = oracle.fill(dc, oracle.SEED_A, op_id, batch * da)
    b = oracle.fill(dc, oracle.SEED_B, op_id, batch * db) if db else None
    return a, b


# ---- matmul (.*.) ----------------------------------------------------------------
@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("mkn", [(3, 3, 3), (4, 4, 4), (2, 2, 2), (4, 4, 1), (3, 3, 1), (1, 4, 4), (2, 5, 3), (7, 2, 6), (1, 9, 1), (8, 8, 8),
                                 (5, 5, 5), (6, 6, 6), (7, 7, 7), (10, 10, 10), (12, 12, 12), (16, 16, 16),
                                 (5, 5, 1), (8, 8, 1), (1, 7, 7), (1, 8, 8), (6, 6, 1),
                                 (9, 9, 9), (11, 11, 11), (13, 7, 19), (14, 14, 14), (17, 17, 17), (18, 18, 18), (19, 19, 19),
                                 (20, 20, 20), (24, 24, 24), (5, 24, 9), (23, 5, 6), (32, 32, 32), (48, 48, 48),
                                 (9, 9, 1), (16, 16, 1), (24, 24, 1), (32, 32, 1)])
@pytest.mark.parametrize("layout", [tb.AOS, tb.SOA])
def test_matmul_bit_exact(ctx, dc, mkn, layout):
    m, k, n = mkn
    for batch in (1, 257, 4099):
        a, b = inputs(dc, 1, batch, m * k, k * n)
        A = ctx.from_numpy(a, (m, k), layout)
        B = ctx.from_numpy(b, (k, n), layout)
        C = tb.matmul(A, B)                  # SoA: the tile kernels where they exist, the thread-per-item SoA contraction otherwise
        assert C.layout == layout
        assert C.shape == (m, n)
        # bit-exact, except the Double shapes the FP64 tensor pipe takes (24^3, 32^3, 48^3 here): stated tolerance
        assert_matmul_matches(C.to_numpy(), a, b, m, k, n, dc, f"matmul {mkn} dt{dc} batch{batch}")


@pytest.mark.parametrize("batch", RAGGED)
def test_matmul_ragged_batches(ctx, batch):
    for dc, (m, k, n) in ((oracle.F64, (3, 3, 3)), (oracle.F32, (4, 4, 4)), (oracle.I64, (3, 3, 3))):
        a, b = inputs(dc, 2, batch, m * k, k * n)
        C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n)))
        assert_bits_equal(C.to_numpy(), oracle.matmul(a, b, m, k, n), f"ragged {batch}")


def test_matmul_signed_zero_and_specials(ctx):
    # 0 + (-0) = +0: the fold starts from +0 exactly like the oracle
    a = np.array([[-0.0, 0.0, 0.0, 0.0, np.inf, 1.0, np.nan, 0.0, 1e308]] * 3, dtype=np.float64)
    b = np.array([[1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 10.0]] * 3, dtype=np.float64)
    C = tb.matmul(ctx.from_numpy(a, (3, 3)), ctx.from_numpy(b, (3, 3))).to_numpy()
    assert_bits_equal(C, oracle.matmul(a, b, 3, 3, 3), "specials")


def test_int64_wraparound(ctx):
    a = np.array([2 ** 62, 2, 0, 1], dtype=np.int64)
    b = np.array([4, 0, 0, 2 ** 63 - 1], dtype=np.int64)
    C = tb.matmul(ctx.from_numpy(a, (2, 2)), ctx.from_numpy(b, (2, 2))).to_numpy()
    assert C.ravel().tolist() == [0, -2, 0, 2 ** 63 - 1]


# ---- dot / dot_sum ---------------------------------------------------------------
@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 16, 33, 36, 64, 100, 200, 260])    # n > 4: the staged long-dot kernels (chunk tails at 33, 36, 100, 200;
def test_dot_bit_exact(ctx, dc, n):                                          # whole 16-byte groups per vector take the pipelined one)
    for batch in (1, 1000, 4099) + ((300001,) if n in (16, 36, 64) else ()):  # 300001: several tiles per CTA, the prefetch runs across tiles
        x, y = inputs(dc, 3, batch, n, n)
        D = tb.dot(ctx.from_numpy(x, (n,)), ctx.from_numpy(y, (n,)))
        assert D.shape == ()
        assert_bits_equal(D.to_numpy(), oracle.dot(x, y, n), f"dot n={n}")


@pytest.mark.parametrize("layout", [tb.AOS, tb.SOA])
def test_dot_sum(ctx, layout):
    batch, n = 100003, 4
    xi, yi = inputs(oracle.I64, 4, batch, n, n)
    assert tb.dot_sum(ctx.from_numpy(xi, (n,), layout), ctx.from_numpy(yi, (n,), layout)) == oracle.dot_sum(xi, yi, n)
    x, y = inputs(oracle.F64, 4, batch, n, n)
    got = tb.dot_sum(ctx.from_numpy(x, (n,), layout), ctx.from_numpy(y, (n,), layout))
    # tolerance: 1e-12 relative to sum |x||y|
    assert abs(got - oracle.dot_sum(x, y, n)) <= 1e-12 * np.sum(np.abs(x) * np.abs(y))
    xf, yf = inputs(oracle.F32, 4, batch, n, n)
    gotf = tb.dot_sum(ctx.from_numpy(xf, (n,), layout), ctx.from_numpy(yf, (n,), layout))
    assert abs(float(gotf) - float(oracle.dot_sum(xf, yf, n))) <= 1e-5 * np.sum(np.abs(xf) * np.abs(yf))


# ---- prod n / tensor product -----------------------------------------------------
@pytest.mark.parametrize("dc", ALL)
def test_prod_and_tensor_product(ctx, dc):
    batch = 300
    a, b = inputs(dc, 5, batch, 64, 64)
    A, B = ctx.from_numpy(a, (8, 8)), ctx.from_numpy(b, (8, 8))
    T = tb.tensor_product(A, B)
    assert T.shape == (8, 8, 8, 8)
    assert_bits_equal(T.to_numpy(), oracle.prod(a, [8, 8], b, [8, 8], 0), "8x8 (x) 8x8")
    assert_bits_equal(tb.prod(A, B, 1).to_numpy(), oracle.prod(a, [8, 8], b, [8, 8], 1), "prod 1")
    assert_bits_equal(tb.prod(A, B, 2).to_numpy(), oracle.prod(a, [8, 8], b, [8, 8], 2), "prod 2")
    a3, b3 = inputs(dc, 6, batch, 2 * 3 * 4, 4 * 5)
    P = tb.prod(ctx.from_numpy(a3, (2, 3, 4)), ctx.from_numpy(b3, (4, 5)), 1)
    assert P.shape == (2, 3, 5)
    assert_bits_equal(P.to_numpy(), oracle.prod(a3, [2, 3, 4], b3, [4, 5], 1), "rank3.rank2")
    # small outer products: the coalesced-run kernel (odd and even extents, ragged batch)
    for (da, db) in (((3, 3), (3, 3)), ((4,), (4,)), ((2, 2), (5,)), ((7,), (2, 3))):
        for bt in (1, 4099):
            ao, bo = inputs(dc, 8, bt, int(np.prod(da)), int(np.prod(db)))
            To = tb.tensor_product(ctx.from_numpy(ao, da), ctx.from_numpy(bo, db))
            assert To.shape == tuple(da) + tuple(db)
            assert_bits_equal(To.to_numpy(), oracle.prod(ao, list(da), bo, list(db), 0), f"{da} (x) {db}")
    # odd Q exercises the scalar (non-vectorised) path
    a4, b4 = inputs(dc, 7, batch, 3, 5)
    assert_bits_equal(tb.tensor_product(ctx.from_numpy(a4, (3,)), ctx.from_numpy(b4, (5,))).to_numpy(),
                      oracle.prod(a4, [3], b4, [5], 0), "3 (x) 5")


@pytest.mark.parametrize("dc", ALL)
def test_prod_and_tensor_product_soa(ctx, dc):
    """The same contractions on SoA batches: the staged-B contraction, the vectorised outer product, the fallbacks."""
    for batch in (1, 333, 4099):
        a, b = inputs(dc, 5, batch, 64, 64)
        A, B = ctx.from_numpy(a, (8, 8), tb.SOA), ctx.from_numpy(b, (8, 8), tb.SOA)
        T = tb.tensor_product(A, B)
        assert T.layout == tb.SOA and T.shape == (8, 8, 8, 8)
        assert_bits_equal(T.to_numpy(), oracle.prod(a, [8, 8], b, [8, 8], 0), "SoA 8x8 (x) 8x8")
        assert_bits_equal(tb.prod(A, B, 1).to_numpy(), oracle.prod(a, [8, 8], b, [8, 8], 1), "SoA prod 1")
        assert_bits_equal(tb.prod(A, B, 2).to_numpy(), oracle.prod(a, [8, 8], b, [8, 8], 2), "SoA prod 2")
        a4, b4 = inputs(dc, 7, batch, 3, 5)
        assert_bits_equal(tb.tensor_product(ctx.from_numpy(a4, (3,), tb.SOA), ctx.from_numpy(b4, (5,), tb.SOA)).to_numpy(),
                          oracle.prod(a4, [3], b4, [5], 0), "SoA 3 (x) 5")
        a3, b3 = inputs(dc, 6, batch, 2 * 3 * 4, 4 * 5)
        P = tb.prod(ctx.from_numpy(a3, (2, 3, 4), tb.SOA), ctx.from_numpy(b3, (4, 5), tb.SOA), 1)
        assert_bits_equal(P.to_numpy(), oracle.prod(a3, [2, 3, 4], b3, [4, 5], 1), "SoA rank3.rank2")
        for rows in (2, 3, 5):              # row counts of A that end inside the prefetch ring of the column-split 8x8 kernel
            a5, b5 = inputs(dc, 8, batch, rows * 8, 64)
            P = tb.prod(ctx.from_numpy(a5, (rows, 8), tb.SOA), ctx.from_numpy(b5, (8, 8), tb.SOA), 1)
            assert_bits_equal(P.to_numpy(), oracle.prod(a5, [rows, 8], b5, [8, 8], 1), "SoA %dx8 . 8x8" % rows)


def test_prod_shape_mismatch_is_wrong_list_length(ctx):
    A = ctx.empty((3, 4), np.float64, 4)
    B = ctx.empty((5, 2), np.float64, 4)
    with pytest.raises(tb.WrongListLength):
        tb.prod(A, B, 1)
    with pytest.raises(tb.WrongListLength):
        tb.matmul(A, B)


# ---- transpose ---------------------------------------------------------------------
def test_transpose_golden(ctx):
    with open(os.path.join(GOLD, "transpose_known.json")) as f:
        cases = json.load(f)
    for c in cases:
        for dt in (np.int64, np.float64, np.float32):
            a = np.asarray(c["in"], dtype=dt)
            T = tb.transpose(ctx.from_numpy(a, c["dims"]))
            assert list(T.shape) == list(reversed(c["dims"]))
            assert T.to_numpy().ravel().tolist() == np.asarray(c["out"], dtype=dt).tolist(), c["source"]


@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("dims", [(4, 4), (3, 3), (2, 3), (4, 1), (5,), (2, 3, 4), (8, 8), (3, 1, 2, 2), (16, 16),
                                  (17, 33), (32, 32), (20, 40), (64, 16), (16, 1, 31), (15, 40), (128, 128), (100, 130),
                                  (48, 50), (66, 190), (64, 64), (130, 48)])      # (4-byte elements, even extents >= 48: 64 x 64 tiles, 8-byte accesses)
@pytest.mark.parametrize("layout", [tb.AOS, tb.SOA])
def test_transpose_bit_exact_and_involution(ctx, dc, dims, layout):
    d = int(np.prod(dims))
    for batch in ((1, 513, 4099) if d <= 1024 else (1, 37)):   # extents >= 16: the tiled matrix transpose (ragged tiles at 17, 33, 100, 130)
        a, _ = inputs(dc, 8, batch, d)
        A = ctx.from_numpy(a, dims, layout)
        T = tb.transpose(A)                  # SoA beyond the tile kernels: a copy of whole rows
        assert T.layout == layout
        assert_bits_equal(T.to_numpy(), oracle.transpose(a, dims), f"transpose {dims}")
        assert_bits_equal(tb.transpose(T).to_numpy().reshape(batch, *dims), a.reshape(batch, *dims), "involution")


def test_int64_product_transpose_identity(ctx):
    batch, n = 5000, 4
    a, b = inputs(oracle.I64, 9, batch, n * n, n * n)
    A, B = ctx.from_numpy(a, (n, n)), ctx.from_numpy(b, (n, n))
    lhs = tb.transpose(A @ B).to_numpy()
    rhs = (tb.transpose(B) @ tb.transpose(A)).to_numpy()
    assert np.array_equal(lhs, rhs)


# ---- det / inverse / solve ---------------------------------------------------------
def elim_inputs(dc, n, batch, nrhs=1):
    a = oracle.fill(dc, oracle.SEED_A, 10, batch * n * n).reshape(batch, n, n).copy()
    oracle.plant_specials(a, n, 0, swap_every=16, sing_every=128)
    b = oracle.fill(dc, oracle.SEED_B, 10, batch * n * nrhs)
    return a, b


@pytest.mark.parametrize("dc", FRAC)
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("layout", [tb.AOS, tb.SOA])
def test_det_inverse_solve_bit_exact(ctx, dc, n, layout):
    for batch in (1, 300, 4099):
        a, b = elim_inputs(dc, n, batch)
        A = ctx.from_numpy(a, (n, n), layout)
        assert_bits_equal(tb.det(A).to_numpy(), oracle.det(a, n), f"det{n}")
        inv, st = tb.inverse(A)
        oinv, ost = oracle.inverse(a, n)
        assert np.array_equal(st.to_numpy(), ost)
        assert_bits_equal(inv.to_numpy(), oinv, f"inverse{n}")
        X, st = tb.solve(A, ctx.from_numpy(b, (n,), layout))
        ox, ost = oracle.solve(a, b, n, 1)
        assert np.array_equal(st.to_numpy(), ost)
        assert_bits_equal(X.to_numpy(), ox.reshape(batch, n), f"solve{n}")
        if batch >= 300:
            assert ost.sum() == (batch + 127) // 128          # every 128th matrix is singular
            assert np.all(oracle.det(a, n)[::128] == 0)


@pytest.mark.parametrize("dc", FRAC)
@pytest.mark.parametrize("n", [5, 6, 7, 8])
@pytest.mark.parametrize("layout", [tb.AOS, tb.SOA])
def test_solve_matrix_rhs_bit_exact(ctx, dc, n, layout):
    """Matrix right-hand sides for 5 <= n <= 8: 2, 3, 4 and n columns run record-per-thread (LU + lazily read
    right-hand side, AoS and SoA), any other count on the row-per-lane kernel (AoS; SoA operands reach it through
    the library's relayout detour)."""
    for nrhs in (2, 3, 4, n, 5 if n != 5 else 6):
        for batch in (1, 1500):
            a, _ = elim_inputs(dc, n, batch)
            b = oracle.fill(dc, oracle.SEED_B, 20 + nrhs, batch * n * nrhs)
            A = ctx.from_numpy(a, (n, n), layout)
            B = ctx.from_numpy(b, (n, nrhs), layout)
            X, st = tb.solve(A, B)
            ox, ost = oracle.solve(a, b, n, nrhs)
            assert np.array_equal(st.to_numpy(), ost), (n, nrhs)
            assert_bits_equal(X.to_numpy(), ox.reshape(batch, n, nrhs), f"solve n={n} nrhs={nrhs}")
            if batch > 128:
                assert ost.sum() == (batch + 127) // 128


@pytest.mark.parametrize("dc", FRAC)
@pytest.mark.parametrize("n", [4, 6, 8])
def test_solve_matrix_rhs_equals_inverse(ctx, dc, n):
    batch = 777
    a, _ = elim_inputs(dc, n, batch)
    eye = np.broadcast_to(np.eye(n, dtype=a.dtype), (batch, n, n)).copy()
    A = ctx.from_numpy(a, (n, n))
    X, st = tb.solve(A, ctx.from_numpy(eye, (n, n)))
    inv, st2 = tb.inverse(A)
    assert_bits_equal(X.to_numpy(), inv.to_numpy(), "solve(A, I) == inverse(A)")
    assert np.array_equal(st.to_numpy(), st2.to_numpy())


@pytest.mark.parametrize("n", [5, 8])
def test_row_per_lane_elimination_pivot_positions(ctx, n):
    """5 <= n <= 8 (record-per-thread and row-per-lane kernels): a zero pivot in EVERY column position with the first non-zero
    entry at every possible row below it, several swaps in one matrix, and singular columns at each k."""
    rng = np.random.default_rng(n)
    mats = []
    for k in range(n):
        for r in range(k + 1, n):                 # column k is zero in rows k..r-1: pivot must come from row r
            u = np.triu(rng.standard_normal((n, n))) + np.eye(n)   # start from an upper-triangular matrix: elimination
            u[k:r, k] = 0.0                                        # leaves its diagonal alone, so the zeros stay zero
            u[r, k] = 1.5
            mats.append(u)
        s = np.triu(rng.standard_normal((n, n))) + np.eye(n)
        s[k:, k] = 0.0                                             # no pivot in column k at all: singular
        mats.append(s)
    p = np.eye(n)[::-1].copy()                                     # anti-diagonal: n // 2 swaps
    mats.append(p)
    a = np.stack(mats).astype(np.float64)
    batch = a.shape[0]
    b = rng.standard_normal((batch, n))
    A = ctx.from_numpy(a, (n, n))
    assert_bits_equal(tb.det(A).to_numpy(), oracle.det(a, n), "det")
    inv, st = tb.inverse(A)
    oinv, ost = oracle.inverse(a, n)
    assert ost.sum() == n and np.array_equal(st.to_numpy(), ost)
    assert_bits_equal(inv.to_numpy(), oinv, "inverse")
    X, st = tb.solve(A, ctx.from_numpy(b, (n,)))
    ox, ost = oracle.solve(a, b, n, 1)
    assert np.array_equal(st.to_numpy(), ost)
    assert_bits_equal(X.to_numpy(), ox.reshape(batch, n), "solve")
    # matrix right-hand side: the row-per-lane kernel (det / inverse / vector solve above run record-per-thread)
    bm = rng.standard_normal((batch, n, 3))
    X, st = tb.solve(A, ctx.from_numpy(bm, (n, 3)))
    ox, ost = oracle.solve(a, bm, n, 3)
    assert np.array_equal(st.to_numpy(), ost)
    assert_bits_equal(X.to_numpy(), ox.reshape(batch, n, 3), "solve, 3 right-hand sides")


def test_det_rejects_int64_and_nonsquare(ctx):
    with pytest.raises(tb.Unsupported):
        tb.det(ctx.empty((3, 3), np.int64, 4))
    with pytest.raises(tb.WrongListLength):
        tb.det(ctx.empty((3, 4), np.float64, 4))


def test_elimination_residuals(ctx):
    batch, n = 20000, 4
    a = oracle.fill(oracle.F64, oracle.SEED_A, 11, batch * n * n).reshape(batch, n, n)
    b = oracle.fill(oracle.F64, oracle.SEED_B, 11, batch * n).reshape(batch, n)
    A = ctx.from_numpy(a, (n, n))
    X, st = tb.solve(A, ctx.from_numpy(b, (n,)))
    x = X.to_numpy()
    assert st.to_numpy().sum() == 0
    r = np.abs(np.einsum("bij,bj->bi", a, x) - b).max(axis=1)
    scale = np.abs(a).sum(axis=(1, 2)) * np.abs(x).max(axis=1) + np.abs(b).max(axis=1)
    cond = np.linalg.cond(a)
    assert np.all(r <= 1e-12 * scale * np.maximum(cond, 1.0))


def _canon_nan(x):
    """NaN sign / payload is not part of the contract (x86 and the GPU generate different
    default NaNs); everything else is compared bit for bit."""
    x = np.array(x, copy=True)
    x[np.isnan(x)] = np.nan
    return x


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8])     # n >= 5: LU with kept multipliers, streamed results, one-stage pipeline
@pytest.mark.parametrize("layout", [tb.AOS, tb.SOA])
def test_elimination_extreme_operands_bit_exact(ctx, n, layout):
    """The Double elimination kernels divide through a shared reciprocal with the compiler's
    own fast-path test; operands outside the fast path (denormal, huge, zero, Inf, NaN
    quotients, numerators or pivots) must take the full division and still match the
    oracle's IEEE `/` exactly."""
    rng = np.random.default_rng(1234 + n)
    batch = 6000
    a = rng.uniform(-1, 1, size=(batch, n, n))
    b = rng.uniform(-1, 1, size=(batch, n))
    # per-matrix and per-entry power-of-two scalings over the whole exponent range
    e_item = rng.integers(-1060, 1020, size=(batch, 1, 1))
    a = np.ldexp(a, np.where(rng.random((batch, 1, 1)) < 0.6, e_item, 0))
    e_ent = rng.integers(-1074, 1023, size=(batch, n, n))
    a = np.where(rng.random((batch, n, n)) < 0.15, np.ldexp(a, e_ent), a)
    b = np.ldexp(b, np.where(rng.random((batch, 1)) < 0.5, rng.integers(-1070, 1020, size=(batch, 1)), 0))
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308,
                         1.7976931348623157e308, -1.7976931348623157e308])
    hit = rng.random((batch, n, n)) < 0.04
    a = np.where(hit, rng.choice(specials, size=(batch, n, n)), a)
    hitb = rng.random((batch, n)) < 0.04
    b = np.where(hitb, rng.choice(specials, size=(batch, n)), b)
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    with np.errstate(all="ignore"):
        A = ctx.from_numpy(a, (n, n), layout)
        assert_bits_equal(_canon_nan(tb.det(A).to_numpy()), _canon_nan(oracle.det(a, n)), f"det{n} extremes")
        inv, st = tb.inverse(A)
        oinv, ost = oracle.inverse(a, n)
        assert np.array_equal(st.to_numpy(), ost)
        assert_bits_equal(_canon_nan(inv.to_numpy()), _canon_nan(oinv), f"inverse{n} extremes")
        X, st = tb.solve(A, ctx.from_numpy(b, (n,), layout))
        ox, ost = oracle.solve(a, b, n, 1)
        assert np.array_equal(st.to_numpy(), ost)
        assert_bits_equal(_canon_nan(X.to_numpy()), _canon_nan(ox.reshape(batch, n)), f"solve{n} extremes")
        # matrix right-hand side (n >= 5: the lazily read, streamed kernel) and the echelon form of the same items
        if layout == tb.AOS or n == 2 or n >= 5:        # (SoA: two columns have a kernel for n = 2 and n >= 5)
            b2 = np.ascontiguousarray(np.stack([b, b[::-1]], axis=2))
            X2, st2 = tb.solve(A, ctx.from_numpy(b2, (n, 2), layout))
            ox2, ost2 = oracle.solve(a, b2, n, 2)
            assert np.array_equal(st2.to_numpy(), ost2)
            assert_bits_equal(_canon_nan(X2.to_numpy()), _canon_nan(ox2.reshape(batch, n, 2)), f"solve{n} x 2 extremes")
        E, rk = tb.row_echelon_form(A)
        oe, ork = oracle.row_echelon(a, n, n)
        assert np.array_equal(rk.to_numpy(), ork)
        assert_bits_equal(_canon_nan(E.to_numpy()), _canon_nan(oe), f"rowEchelonForm{n} extremes")


# ---- regression fixtures through the device ----------------------------------------
def test_la_regression_fixtures_on_device(ctx):
    with open(os.path.join(GOLD, "la_regression.json")) as f:
        cases = json.load(f)
    npd = {"f64": np.float64, "f32": np.float32, "i64": np.int64}
    for c in cases:
        dt = npd[c["dtype"]]
        ins = [np.asarray(x, dtype=dt) for x in c["in"]]
        p = c["params"]
        if c["op"] == "matmul":
            A = ctx.from_numpy(ins[0], (p["m"], p["k"]))
            B = ctx.from_numpy(ins[1], (p["k"], p["n"]))
            outs = [tb.matmul(A, B).to_numpy()]
        elif c["op"] == "dot":
            outs = [tb.dot(ctx.from_numpy(ins[0], (p["n"],)), ctx.from_numpy(ins[1], (p["n"],))).to_numpy()]
        elif c["op"] == "prod":
            outs = [tb.prod(ctx.from_numpy(ins[0], p["dims_a"]), ctx.from_numpy(ins[1], p["dims_b"]), p["n"]).to_numpy()]
        elif c["op"] == "det":
            outs = [tb.det(ctx.from_numpy(ins[0], (p["n"], p["n"]))).to_numpy()]
        elif c["op"] == "inverse":
            o, s = tb.inverse(ctx.from_numpy(ins[0], (p["n"], p["n"])))
            outs = [o.to_numpy(), s.to_numpy()]
        else:
            o, s = tb.solve(ctx.from_numpy(ins[0], (p["n"], p["n"])), ctx.from_numpy(ins[1], (p["n"],)))
            outs = [o.to_numpy(), s.to_numpy()]
        for got, want in zip(outs, c["out"]):
            want = np.asarray(want, dtype=got.dtype)
            assert np.array_equal(bits(got).ravel(), bits(want).ravel()), (c["op"], c["dtype"], p)


# ---- marshalling, layouts, element-wise, errors ------------------------------------
def test_from_list_length_check(ctx):
    with pytest.raises(tb.WrongListLength):
        ctx.from_list([1.0] * 17, (3, 3), np.float64, 2)
    b = ctx.from_list(list(range(18)), (3, 3), np.int64, 2)
    assert b.to_list() == list(range(18))
    with pytest.raises(tb.WrongListLength):
        b.upload(np.zeros(17, dtype=np.int64))


@pytest.mark.parametrize("dc", ALL)
def test_relayout_roundtrip(ctx, dc):
    # odd and even counts either side of the 128-item tile, d even / odd, d <= 64 and beyond (4-byte elements with even d move
    # as 8-byte pairs: two elements of an item on the AoS side, two items of a row on the SoA side)
    for batch, shape in ((1, (3, 3)), (1000, (4, 4)), (4099, (3,)), (77, (8, 8, 2)), (1, (4, 4)), (129, (4, 4)), (255, (2, 3)),
                         (4097, (8, 8)), (127, (2,)), (65, (10, 13)), (193, (33, 2)), (64, (32, 32))):
        a, _ = inputs(dc, 12, batch, int(np.prod(shape)))
        A = ctx.from_numpy(a, shape)
        S = A.relayout(tb.SOA)
        assert_bits_equal(S.to_numpy().ravel(), a, "soa download")
        assert_bits_equal(S.relayout(tb.AOS).to_numpy().ravel(), a, "aos roundtrip")
        S.free(); A.free()
    # the SoA image is the one the SoA kernels read: a product of relayouted operands = the AoS product
    for batch in (1, 129, 1023):
        a, b = inputs(dc, 13, batch, 16, 16)
        A, B = ctx.from_numpy(a, (4, 4)), ctx.from_numpy(b, (4, 4))
        want = tb.matmul(A, B).to_numpy()
        As, Bs = A.relayout(tb.SOA), B.relayout(tb.SOA)
        assert_bits_equal(tb.matmul(As, Bs).to_numpy().ravel(), want.ravel(), "product of relayouted operands")


@pytest.mark.parametrize("dc", ALL)
def test_elementwise(ctx, dc):
    # 3001 x 9, 1 x 3, 5 x 3, 700001 x 1: element counts that end inside a 16-byte vector and inside the
    # unrolled-by-two main loop of the vectorised kernels
    for batch, d in ((3001, 9), (1, 3), (5, 3), (700001, 1), (2, 2)):
        a, b = inputs(dc, 13, batch, d, d)
        shape = (3, 3) if d == 9 else ((d,) if d > 1 else ())
        A, B = ctx.from_numpy(a, shape), ctx.from_numpy(b, shape)
        with np.errstate(over="ignore"):
            assert_bits_equal((A + B).to_numpy().ravel(), a + b, "+")
            assert_bits_equal((A - B).to_numpy().ravel(), a - b, "-")
            assert_bits_equal(tb.zip_with("*", A, B).to_numpy().ravel(), a * b, "*")
            k = a[min(3, a.size - 1)]
            assert_bits_equal(tb.scale(k, A).to_numpy().ravel(), k * a, "scale")


@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("layout", [tb.AOS, tb.SOA])
def test_pure_replicate_unit(ctx, dc, layout):
    """pure (Vector.hs:185-188) / replicate / unit: every item equals the record, in either layout."""
    dt = NPD[dc]
    for batch, shape in ((1, (3, 3)), (4099, (2, 3, 4)), (70001, (5,)), (300, ()), (0, (4, 4))):
        rec, _ = inputs(dc, 15, 1, int(np.prod(shape)) if shape else 1)
        R = tb.replicate(ctx, rec, shape, batch, dt, layout)
        assert R.shape == tuple(shape) and R.batch == batch and R.layout == layout
        assert_bits_equal(R.to_numpy().reshape(batch, rec.size), np.broadcast_to(rec, (batch, rec.size)).copy(), "replicate")
    P = tb.pure(ctx, dt(7), (2, 2), dt, 1000, layout)
    assert np.array_equal(P.to_numpy(), np.full((1000, 2, 2), 7, dtype=dt))
    U = tb.unit(ctx, 4, dt, 513, layout)
    assert np.array_equal(U.to_numpy(), np.broadcast_to(np.eye(4, dtype=dt), (513, 4, 4)))
    a, _ = inputs(dc, 16, 513, 16)
    A = ctx.from_numpy(a, (4, 4), layout)
    with np.errstate(over="ignore"):
        assert_bits_equal((A @ U).to_numpy().ravel(), oracle.matmul(a, np.broadcast_to(np.eye(4, dtype=dt), (513, 4, 4)).copy().ravel(), 4, 4, 4).ravel(), "A .*. unit")
    with pytest.raises(tb.WrongListLength):
        tb.replicate(ctx, np.zeros(5, dtype=dt), (2, 3), 10, dt, layout)


def test_device_synthetic_matches_oracle(ctx):
    for dc in ALL:
        for layout in (tb.AOS, tb.SOA):
            batch = 4099
            A = ctx.synthetic((3, 3), NPD[dc], batch, oracle.SEED_A, 7, layout)
            assert_bits_equal(A.to_numpy().ravel(), oracle.fill(dc, oracle.SEED_A, 7, batch * 9), "fill")
    A = ctx.synthetic((4, 4), np.float64, 70000, oracle.SEED_A, 3).plant_specials()
    ref = oracle.fill(oracle.F64, oracle.SEED_A, 3, 70000 * 16).reshape(70000, 4, 4).copy()
    oracle.plant_specials(ref, 4)
    assert_bits_equal(A.to_numpy(), ref, "plant")


def test_host_entry_matches_device_entry(ctx):
    batch = 300001
    a, b = inputs(oracle.F32, 14, batch, 16, 16)
    out = np.empty(batch * 16, dtype=np.float32)
    tb.matmul_host(ctx, a, b, 4, 4, 4, out)
    assert_bits_equal(out.reshape(batch, 4, 4), oracle.matmul(a, b, 4, 4, 4, threads=8), "host pipeline")


def test_launch_counter_counts_kernels(ctx):
    A = ctx.synthetic((3, 3), np.float64, 1000, 1, 1)
    B = ctx.synthetic((3, 3), np.float64, 1000, 2, 1)
    n0 = ctx.launch_count()
    for _ in range(5):
        tb.matmul(A, B)
    assert ctx.launch_count() - n0 == 5


# ---- full BASELINE sizes: size-independent properties --------------------------------
def test_full_size_config1_properties(ctx):
    batch = 1 << 20
    A = ctx.synthetic((3, 3), np.int64, batch, oracle.SEED_A, 1)
    B = ctx.synthetic((3, 3), np.int64, batch, oracle.SEED_B, 1)
    lhs = tb.transpose(A @ B).to_numpy()
    rhs = (tb.transpose(B) @ tb.transpose(A)).to_numpy()
    assert np.array_equal(lhs, rhs)
    # spot parity on a slice of the full batch, including the ragged tail region
    a = oracle.fill(oracle.I64, oracle.SEED_A, 1, batch * 9).reshape(batch, 3, 3)
    b = oracle.fill(oracle.I64, oracle.SEED_B, 1, batch * 9).reshape(batch, 3, 3)
    sl = slice(batch - 5000, batch)
    assert np.array_equal(lhs[sl], oracle.transpose(oracle.matmul(a[sl], b[sl], 3, 3, 3), [3, 3]))
    Ad = ctx.synthetic((3, 3), np.float64, batch, oracle.SEED_A, 2)
    eye = ctx.from_numpy(np.broadcast_to(np.eye(3), (batch, 3, 3)).copy(), (3, 3))
    assert_bits_equal((Ad @ eye).to_numpy(), Ad.to_numpy(), "A.I == A")


def test_full_size_config3_properties(ctx):
    batch = 1 << 22
    A = ctx.synthetic((4, 4), np.float64, batch, oracle.SEED_A, 3).plant_specials()
    inv, st = tb.inverse(A)
    s = st.to_numpy()
    assert s.sum() == batch // 65536 and np.all(s[::65536] == 1)
    d = tb.det(A).to_numpy()
    assert np.all(d[::65536] == 0) and np.count_nonzero(d == 0) == batch // 65536
    # A . A^-1 ~ I on the non-singular items (residual scaled by the condition number)
    prod = (A @ inv).to_numpy()
    ok = s == 0
    res = np.abs(prod[ok] - np.eye(4)).max(axis=(1, 2))
    a = A.to_numpy()[ok][:200000]
    cond = np.linalg.cond(a)
    assert np.all(res[:200000] <= 1e-12 * 100 * cond)
    # parity with the oracle on a slice that contains swaps and a singular item
    sl = slice(65536 - 2048, 65536 + 2048)
    ao = A.to_numpy()[sl]
    oinv, ost = oracle.inverse(ao, 4)
    assert_bits_equal(inv.to_numpy()[sl], oinv, "inverse slice")
    assert np.array_equal(s[sl], ost)


def test_full_size_config2_properties(ctx):
    """16Mi 4x4 Float (BASELINE configs[1]): (A.B)^T == B^T.A^T holds BIT-EXACTLY for floats too (the same
    products are added in the same j order), A.I == A, transpose is an involution; spot parity on slices."""
    batch = 1 << 24
    A = ctx.synthetic((4, 4), np.float32, batch, oracle.SEED_A, 1)
    B = ctx.synthetic((4, 4), np.float32, batch, oracle.SEED_B, 1)
    C = A @ B
    lhs = tb.transpose(C).to_numpy()
    At, Bt = tb.transpose(A), tb.transpose(B)
    rhs = (Bt @ At).to_numpy()
    assert_bits_equal(lhs, rhs, "(AB)^T == B^T A^T")
    del rhs
    assert_bits_equal(tb.transpose(At).to_numpy(), A.to_numpy(), "transpose involution")
    a = oracle.fill(oracle.F32, oracle.SEED_A, 1, batch * 16).reshape(batch, 4, 4)
    b = oracle.fill(oracle.F32, oracle.SEED_B, 1, batch * 16).reshape(batch, 4, 4)
    for sl in (slice(0, 4096), slice(batch // 2 - 1000, batch // 2 + 1000), slice(batch - 4099, batch)):
        want = oracle.transpose(oracle.matmul(a[sl], b[sl], 4, 4, 4), [4, 4]).reshape(-1, 4, 4)
        assert_bits_equal(lhs[sl], want, f"slice {sl}")
    # mat.vec and dot on 4-vectors at full size: linearity in the vector is exact for a power-of-two factor
    x = ctx.synthetic((4,), np.float32, batch, oracle.SEED_B, 2)
    y = (A @ x).to_numpy()
    y2 = (A @ tb.scale(np.float32(4.0), x)).to_numpy()
    assert_bits_equal(y2, 4.0 * y, "A.(4x) == 4 A.x")
    d = tb.dot(x, x).to_numpy()
    assert np.all(d >= 0) and d.shape == (batch,)


def test_full_size_config4_properties(ctx):
    """1Mi 8x8 (x) 8x8 -> 8^4 Double (BASELINE configs[3]; 34 GB result): the squared norm of an outer
    product factorises, sum(C_b^2) == sum(a_b^2) * sum(b_b^2), reduced on the device; prefix parity."""
    batch = 1 << 20
    A = ctx.synthetic((8, 8), np.float64, batch, oracle.SEED_A, 4)
    B = ctx.synthetic((8, 8), np.float64, batch, oracle.SEED_B, 4)
    C = tb.tensor_product(A, B)
    assert C.shape == (8, 8, 8, 8)
    import ctypes
    out = np.zeros(1)                                   # sum(C*C) reduced on the device (C viewed as 4096-vectors)
    tb.device.check(tb._lib.lib().t5_dot_sum(ctx._h, C.dtype, 4096, C._h, C._h, out.ctypes.data_as(ctypes.c_void_p)), ctx)
    a = A.to_numpy().reshape(batch, 64)
    b = B.to_numpy().reshape(batch, 64)
    want = float(np.sum(np.sum(a * a, axis=1) * np.sum(b * b, axis=1)))
    assert abs(out[0] - want) <= 1e-9 * want, (out[0], want)
    # n = 1 contraction at full size equals the matrix product (prod 1 == .*.), bit-exact
    P1 = tb.prod(A, B, 1).to_numpy()
    M = (A @ B).to_numpy()
    assert_bits_equal(P1, M, "prod 1 == .*.")
    sl = slice(batch - 300, batch)
    assert_bits_equal(M[sl].reshape(-1), oracle.matmul(a[sl].ravel(), b[sl].ravel(), 8, 8, 8).ravel(), "8x8 slice")


def test_full_size_config5_properties(ctx):
    """64Ki 128x128 Double on the FP64 tensor cores (BASELINE configs[4]), checked entirely on the device:
    A.(2I) == 2A exactly (one non-zero term per sum), so sum((A.(2I) - 2A)^2) must be exactly 0."""
    batch, n = 1 << 16, 128
    A = ctx.synthetic((n, n), np.float64, batch, oracle.SEED_A, 5)
    two_eye = ctx.from_numpy(np.broadcast_to(2.0 * np.eye(n), (256, n, n)).copy(), (n, n))
    a_slab = np.ascontiguousarray(A.to_numpy()[: 256])
    C = tb.matmul(ctx.from_numpy(a_slab, (n, n)), two_eye).to_numpy()
    assert_bits_equal(C, 2.0 * a_slab, "A.(2I) == 2A (slab)")
    # full batch: C = A.B against the tolerance bound on a strided sample, and a launch over every tile
    B = ctx.synthetic((n, n), np.float64, batch, oracle.SEED_B, 5)
    Cfull = A @ B
    ctx.sync()
    c = np.ascontiguousarray(Cfull.to_numpy()[:: 4099])      # copies, so each 8.6 GB download is dropped at once
    a = np.ascontiguousarray(A.to_numpy()[:: 4099])
    b = np.ascontiguousarray(B.to_numpy()[:: 4099])
    want = np.einsum("bij,bjk->bik", a, b)
    bound = 1e-12 * np.einsum("bij,bjk->bik", np.abs(a), np.abs(b))
    assert np.all(np.abs(c - want) <= bound)
    assert np.all(np.isfinite(c))


def test_full_size_float128_tcgen05_properties(ctx):
    """64Ki 128x128 Float on the tcgen05 tensor cores at the size of BASELINE configs[4] (443 tiles per CTA, both
    TMEM accumulators and every ring phase many times over): a strided sample against float64."""
    batch, n = 1 << 16, 128
    A = ctx.synthetic((n, n), np.float32, batch, oracle.SEED_A, 6)
    B = ctx.synthetic((n, n), np.float32, batch, oracle.SEED_B, 6)
    C = A @ B
    ctx.sync()
    c = np.ascontiguousarray(C.to_numpy()[:: 1021]).astype(np.float64)
    a = np.ascontiguousarray(A.to_numpy()[:: 1021]).astype(np.float64)
    b = np.ascontiguousarray(B.to_numpy()[:: 1021]).astype(np.float64)
    want = np.einsum("bij,bjk->bik", a, b)
    bound = 1e-5 * np.einsum("bij,bjk->bik", np.abs(a), np.abs(b))
    assert np.all(np.abs(c - want) <= bound)
    # the very last items too (tail of the persistent loop)
    tail = slice(batch - 3, batch)
    ct = C.to_numpy()[tail].astype(np.float64)
    at, bt = A.to_numpy()[tail].astype(np.float64), B.to_numpy()[tail].astype(np.float64)
    assert np.all(np.abs(ct - np.einsum("bij,bjk->bik", at, bt)) <= 1e-5 * np.einsum("bij,bjk->bik", np.abs(at), np.abs(bt)))


# ---- FP64 tensor-core (DMMA) product: tolerance stated by north_star -------------------
@pytest.mark.parametrize("mkn", [(128, 128, 128), (128, 64, 128), (256, 32, 128), (128, 96, 256),
                                 (64, 64, 64), (64, 32, 192), (192, 96, 64), (128, 160, 64),
                                 # ragged against the 128 / 64 tiles and the 32-deep K chunks (round 2: the item is a tensor-map
                                 # dimension, so the missing rows / columns / k-slices of a box arrive as zeros)
                                 (96, 96, 96), (100, 64, 104), (48, 32, 56), (130, 40, 160), (65, 128, 64), (200, 104, 96),
                                 (127, 48, 224), (257, 32, 64),
                                 # items of at most 32 x 32 (four to a tile) and 48 x 48 (two to a tile), the 96 tile, K tails
                                 # of 8 / 16 / 24 (skipped k-steps), K shorter than one chunk
                                 (32, 32, 32), (24, 24, 24), (17, 8, 24), (25, 40, 32), (32, 16, 32), (31, 56, 24),
                                 (48, 48, 48), (40, 40, 40), (33, 24, 48), (72, 72, 72), (80, 88, 80), (160, 48, 160), (192, 72, 96),
                                 # k and n that are only EVEN (one TMA box per 64-byte group, partial groups zero-filled, odd k-step counts)
                                 (100, 100, 100), (50, 50, 50), (30, 10, 26), (18, 22, 24), (66, 34, 70), (129, 14, 130), (32, 12, 30)])
def test_dmma_matmul_within_tolerance(ctx, mkn):
    m, k, n = mkn
    assert not matmul_is_exact(oracle.F64, m, k, n)
    for batch in (1, 2, 3, 7, 150, 601):     # > 148 tiles: every CTA takes work, some take two; 1 / 2 / 3 / 7: part-filled item tiles
        a, b = inputs(oracle.F64, 15, batch, m * k, k * n)
        C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
        want = oracle.matmul(a, b, m, k, n, threads=8)
        a3, b3 = np.abs(a.reshape(batch, m, k)), np.abs(b.reshape(batch, k, n))
        bound = 1e-12 * np.einsum("bij,bjk->bik", a3, b3)      # |delta| <= 1e-12 * sum_j |a_ij||b_jk|
        assert np.all(np.abs(C - want) <= bound), f"max excess {np.max(np.abs(C - want) / bound)}"
        # and it is not accidentally the generic kernel: fused k4 steps differ in the last bits somewhere
        # (informational only; equality would also be fine)


def test_mid_size_matmul_generic_is_bit_exact(ctx):
    for dc in ALL:
        for (m, k, n) in ((64, 64, 64), (16, 16, 16), (100, 20, 36)):
            batch = 9
            a, b = inputs(dc, 16, batch, m * k, k * n)
            C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
            if not matmul_is_exact(dc, m, k, n):
                continue        # tensor-core shapes (tolerance tests)
            assert_bits_equal(C, oracle.matmul(a, b, m, k, n), f"generic {m}x{k}x{n}")


# ---- exact blocked product (t5_xgemm.cu): oracle order, so bit-exact for every dtype and shape ----
@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("mkn", [(128, 128, 128), (64, 64, 64), (65, 17, 63), (32, 8, 32), (33, 100, 20), (100, 20, 36),
                                 (16, 16, 17), (200, 9, 40), (24, 31, 130), (48, 48, 48), (1, 64, 300), (300, 64, 1)])
def test_exact_blocked_matmul_bit_exact(ctx, dc, mkn):
    m, k, n = mkn
    if not matmul_is_exact(dc, m, k, n):
        pytest.skip("this shape takes a tensor-core path for this element type (tolerance tests)")
    for batch in (1, 5, 37):
        a, b = inputs(dc, 21, batch, m * k, k * n)
        C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
        assert_bits_equal(C, oracle.matmul(a, b, m, k, n), f"blocked {m}x{k}x{n} batch {batch}")


# ---- Float on tcgen05 (3xTF32, TMEM accumulators): tolerance stated by north_star for Float ----
@pytest.mark.parametrize("mkn", [(128, 128, 128), (128, 32, 128), (256, 64, 128), (128, 96, 256), (256, 160, 256),
                                 (64, 64, 64), (64, 32, 192), (192, 96, 64), (128, 160, 64),
                                 # ragged against the 128 / 64 tiles and the 32-deep K chunks: the tensor maps carry the item as
                                 # a dimension, so the missing rows / columns / k-slices of a box arrive as zeros
                                 (96, 96, 96), (100, 64, 96), (48, 32, 64), (130, 36, 160), (65, 128, 64), (200, 100, 96),
                                 (127, 44, 224), (257, 32, 64),
                                 # items of at most 32 x 32, four to a 128 tile (diagonal accumulator blocks), short and ragged K
                                 (32, 32, 32), (17, 40, 32), (25, 44, 32), (32, 64, 32), (31, 20, 32),
                                 # column counts that are not multiples of 32 (one TMA box per 128-byte column block, columns past
                                 # n zero-filled, the last block stored in part): packed and unpacked geometries
                                 (28, 28, 28), (26, 32, 28), (32, 24, 28), (48, 48, 48), (40, 40, 40), (56, 64, 44), (36, 32, 36), (100, 32, 100),
                                 (130, 36, 68), (64, 32, 36)])
def test_tf32x3_matmul_within_tolerance(ctx, mkn):
    m, k, n = mkn
    assert not matmul_is_exact(oracle.F32, m, k, n)
    for batch in (1, 2, 3, 7, 150, 301, 1201):          # > 148 tiles: every CTA works, some take several tiles (both TMEM accumulators)
        a, b = inputs(oracle.F32, 22, batch, m * k, k * n)
        C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
        a3 = a.reshape(batch, m, k).astype(np.float64)
        b3 = b.reshape(batch, k, n).astype(np.float64)
        want = np.einsum("bij,bjk->bik", a3, b3)
        bound = 1e-5 * np.einsum("bij,bjk->bik", np.abs(a3), np.abs(b3))      # |delta| <= 1e-5 * sum_j |a_ij||b_jk|
        err = np.abs(C.astype(np.float64) - want)
        assert np.all(err <= bound), f"max excess {np.max(err / bound)} at {np.unravel_index(np.argmax(err / bound), err.shape)}"
        # and far better than a single TF32 pass would be (~1e-3): the split terms are really there
        assert np.max(err / bound) < 0.5


@pytest.mark.parametrize("mkn", [(96, 100, 96), (30, 20, 32)])
def test_tf32x3_ragged_specials_and_neighbours(ctx, mkn):
    """A ragged shape (96 x 100 . 100 x 96: one 128-tile, K tail of 4) and a packed one (30 x 20 . 20 x 32: four items to
    a tile, the last tile half empty): items with inf / NaN are recomputed exactly, and the zero-filled part of a box
    must really be zeros - an item of huge values next to an item of tiny ones would show up at once if a box ever read
    its neighbour's rows (or, packed, if an epilogue warp read an off-diagonal accumulator block)."""
    (m, k, n), batch = mkn, 6
    assert not matmul_is_exact(oracle.F32, m, k, n)
    a, b = inputs(oracle.F32, 25, batch, m * k, k * n)
    a = a.copy(); b = b.copy()
    a[1 * m * k: 2 * m * k] *= np.float32(1e30)                 # neighbours of very different magnitude
    a[2 * m * k: 3 * m * k] *= np.float32(1e-30)
    a[4 * m * k + 5] = np.inf
    b[5 * k * n + 77] = np.nan
    C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy().reshape(batch, m, n)
    want = oracle.matmul(a, b, m, k, n).reshape(batch, m, n)
    for item in (4, 5):
        nan = np.isnan(want[item])
        assert np.array_equal(np.isnan(C[item]), nan), item
        assert_bits_equal(np.where(nan, 0, C[item]), np.where(nan, 0, want[item]), f"special item {item}")
    a3 = a.reshape(batch, m, k).astype(np.float64)
    b3 = b.reshape(batch, k, n).astype(np.float64)
    for item in (0, 1, 2, 3):
        err = np.abs(C[item].astype(np.float64) - a3[item] @ b3[item])
        assert np.all(err <= 1e-5 * (np.abs(a3[item]) @ np.abs(b3[item]))), item


@pytest.mark.parametrize("n", [128, 64])
def test_tf32x3_identity_and_specials(ctx, n):
    batch = 5
    a, _ = inputs(oracle.F32, 23, batch, n * n)
    eye = np.broadcast_to(np.eye(n, dtype=np.float32), (batch, n, n)).copy()
    A = ctx.from_numpy(a, (n, n))
    C = (A @ ctx.from_numpy(eye, (n, n))).to_numpy().ravel()
    assert np.allclose(C, a, rtol=2e-6, atol=0)                 # hi + lo recombine the operand to ~2^-21
    # non-finite operands: the tiles that contain them are recomputed with the exact sequential fold, so they
    # match the oracle bit for bit (inf stays inf where the reference has inf, NaN where it has NaN)
    a2 = a.copy(); a2[7] = np.inf; a2[n * n + 3] = np.nan; a2[3 * n * n + 130] = -np.inf
    b2, _ = inputs(oracle.F32, 24, batch, n * n)
    b2 = b2.copy(); b2[2 * n * n + 77] = np.inf
    C2 = (ctx.from_numpy(a2, (n, n)) @ ctx.from_numpy(b2, (n, n))).to_numpy().reshape(batch, n, n)
    want = oracle.matmul(a2, b2, n, n, n).reshape(batch, n, n)
    for item in (0, 1, 2, 3):
        nan = np.isnan(want[item])
        assert np.array_equal(np.isnan(C2[item]), nan), item
        assert_bits_equal(np.where(nan, 0, C2[item]), np.where(nan, 0, want[item]), f"special item {item}")
    assert np.isinf(want[0]).any() and np.isinf(C2[0]).any()
    err = np.abs(C2[4].astype(np.float64) - want[4])                      # the untouched item stays on the tensor cores
    assert np.all(err <= 1e-5 * (np.abs(a2.reshape(batch, n, n)[4]).astype(np.float64) @ np.abs(b2.reshape(batch, n, n)[4]).astype(np.float64)))


def test_exact_blocked_matmul_specials_and_k_tail(ctx):
    """inf / nan / signed zeros in the operands and a K that is not a multiple of the chunk: the K
    tail must be a loop bound, not zero padding (0 * inf would poison sums the oracle never forms)."""
    m, k, n, batch = 40, 19, 24, 3           # (k = 19: no tensor-core kernel takes it)
    assert matmul_is_exact(oracle.F64, m, k, n) and matmul_is_exact(oracle.F32, m, k, n)
    rng = np.random.default_rng(5)
    a = rng.standard_normal(batch * m * k)
    b = rng.standard_normal(batch * k * n)
    a[5] = np.inf; a[m * k + 7] = -0.0; a[2 * m * k + 11] = np.nan
    b[3] = -np.inf; b[k * n + 9] = 0.0; b[2 * k * n + 100] = 1e308

    def same(got, want, what):      # bit-equal, except that a NaN only has to be a NaN (generated NaNs differ in sign/payload between x86 and the GPU)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), what
        assert_bits_equal(np.where(nan, 0, got), np.where(nan, 0, want), what)

    C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
    same(C.ravel(), oracle.matmul(a, b, m, k, n).ravel(), "specials f64")
    af, bf = a.astype(np.float32), b.astype(np.float32)
    Cf = tb.matmul(ctx.from_numpy(af, (m, k)), ctx.from_numpy(bf, (k, n))).to_numpy()
    same(Cf.ravel(), oracle.matmul(af, bf, m, k, n).ravel(), "specials f32")


def test_dmma_ragged_neighbours_and_specials(ctx):
    """A ragged Double shape (96 x 104 . 104 x 96: one 128-tile, K tail of 8): neighbouring items of very different
    magnitude (a box that read its neighbour's rows instead of zeros would show at once) and inf / NaN operands, which
    the FP64 tensor pipe propagates like the sequential fold does (same inf / NaN pattern as the oracle)."""
    m, k, n, batch = 96, 104, 96, 6
    a, b = inputs(oracle.F64, 26, batch, m * k, k * n)
    a = a.copy(); b = b.copy()
    a[1 * m * k: 2 * m * k] *= 1e200
    a[2 * m * k: 3 * m * k] *= 1e-200
    a[4 * m * k + 5] = np.inf
    b[5 * k * n + 77] = np.nan
    C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy().reshape(batch, m, n)
    want = oracle.matmul(a, b, m, k, n).reshape(batch, m, n)
    for item in (4, 5):
        assert np.array_equal(np.isnan(C[item]), np.isnan(want[item])), item
        assert np.array_equal(np.isinf(C[item]), np.isinf(want[item])), item
    a3, b3 = np.abs(a.reshape(batch, m, k)), np.abs(b.reshape(batch, k, n))
    for item in (0, 1, 2, 3):
        assert np.all(np.abs(C[item] - want[item]) <= 1e-12 * (a3[item] @ b3[item])), item


def test_dmma_identity_and_linearity(ctx):
    batch, n = 20, 128
    a, _ = inputs(oracle.F64, 17, batch, n * n)
    eye = np.broadcast_to(np.eye(n), (batch, n, n)).copy()
    A = ctx.from_numpy(a, (n, n))
    assert_bits_equal((A @ ctx.from_numpy(eye, (n, n))).to_numpy().ravel(), a, "A.I == A (exact: one non-zero term per sum)")
    two = ctx.from_numpy(2.0 * eye, (n, n))
    assert_bits_equal((A @ two).to_numpy().ravel(), 2.0 * a, "A.(2I) == 2A")


@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("n", [6, 8, 16])
def test_row_per_thread_contraction_bit_exact(ctx, dc, n):
    for batch in (1, 31, 33, 1000, 4099):
        a, b = inputs(dc, 18, batch, n * n, n * n)
        C = tb.matmul(ctx.from_numpy(a, (n, n)), ctx.from_numpy(b, (n, n))).to_numpy()
        assert_bits_equal(C, oracle.matmul(a, b, n, n, n), f"{n}x{n} batch {batch}")
