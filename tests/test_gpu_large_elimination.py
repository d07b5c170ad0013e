"""det / inverse / solve for 9 <= n <= 128 (t5_lu.cuh: one CTA per matrix, the matrix 2D-cyclic in registers).
The kernels keep the oracle's operation order element for element, so every result - the determinant, the status
bytes and every entry of the inverse / solution - is compared BIT FOR BIT, with planted zero pivots (row swaps,
including swaps between rows of the same owner thread) and exactly singular items in every batch."""
import numpy as np
import pytest

import oracle
import tensor_b200 as tb
from tensor_b200 import host

pytestmark = pytest.mark.gpu

FRAC = [oracle.F64, oracle.F32]


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def inputs(dc, n, batch, seed_op):
    a = oracle.fill(dc, oracle.SEED_A, seed_op, batch * n * n).reshape(batch, n, n).copy()
    for i in range(0, batch, 3):                       # leading zero: swap with the first non-zero row below
        a[i, 0, 0] = 0
    for i in range(1, batch, 5):                       # several leading zeros in column 0 and a zero met later on the diagonal
        a[i, 0:3, 0] = 0
        a[i, 2, :] = a[i, 1, :] * 0.5
        a[i, 2, n - 1] += 1.0
    for i in range(2, batch, 7):                       # swap partner 8 / 16 rows away: rows of the SAME owner thread
        a[i, 0, 0] = 0
        a[i, 1:min(n, 17) - 1, 0] = 0
    for i in range(4, batch, 11):                      # exactly singular: a repeated row
        a[i, n - 1, :] = a[i, n // 2, :]
    return a


@pytest.mark.parametrize("dc", FRAC)
@pytest.mark.parametrize("n", [9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 24, 25, 32, 33, 48, 64, 65, 100, 128])   # 9..14: the record-per-thread kernels of t5_small_ops_9.cu
def test_large_det_inverse_solve_bit_exact(ctx, dc, n):
    batch = 67 if n <= 32 else (35 if n <= 64 else 23)
    a = inputs(dc, n, batch, 60 + n)
    A = ctx.from_numpy(a, (n, n))
    od = oracle.det(a, n, threads=8)
    assert np.array_equal(bits(tb.det(A).to_numpy()), bits(od)), f"det {n}"
    inv, st = tb.inverse(A)
    oinv, ost = oracle.inverse(a, n, threads=8)
    assert ost.sum() >= batch // 11 and ost.sum() < batch
    assert np.array_equal(st.to_numpy(), ost), f"inverse status {n}"
    assert np.array_equal(bits(inv.to_numpy()), bits(oinv)), f"inverse {n}"
    assert np.all(od[ost == 1] == 0)
    for nrhs in (1, 3, n // 2 + 1):
        b = oracle.fill(dc, oracle.SEED_B, 60 + nrhs, batch * n * nrhs)
        shape = (n,) if nrhs == 1 else (n, nrhs)
        X, st = tb.solve(A, ctx.from_numpy(b, shape))
        ox, ost2 = oracle.solve(a, b, n, nrhs, threads=8)
        assert np.array_equal(st.to_numpy(), ost2)
        assert np.array_equal(bits(X.to_numpy().reshape(batch, n, nrhs)), bits(ox)), f"solve {n} x {nrhs}"


@pytest.mark.parametrize("n,batch", [(9, 70001), (10, 40001), (12, 9001), (24, 7919), (40, 2999)])
def test_large_elimination_batches_beyond_one_wave(ctx, n, batch):
    """More items than the grid holds at once (888 CTAs x 8 groups for the 4 x 4-thread geometry): every CTA loops, the
    last pass is part empty - groups that share a warp with a group past the end of the batch recompute the last item
    and must store nothing - and zero-pivot / singular items sit next to ordinary ones in the same warp."""
    a = inputs(oracle.F64, n, batch, 80 + n)
    A = ctx.from_numpy(a, (n, n))
    assert np.array_equal(bits(tb.det(A).to_numpy()), bits(oracle.det(a, n, threads=8))), f"det {n}"
    inv, st = tb.inverse(A)
    oinv, ost = oracle.inverse(a, n, threads=8)
    assert np.array_equal(st.to_numpy(), ost)
    assert np.array_equal(bits(inv.to_numpy()), bits(oinv)), f"inverse {n}"
    b = oracle.fill(oracle.F64, oracle.SEED_B, 80 + n, batch * n)
    X, st = tb.solve(A, ctx.from_numpy(b, (n,)))
    ox, ost2 = oracle.solve(a, b, n, 1, threads=8)
    assert np.array_equal(st.to_numpy(), ost2)
    assert np.array_equal(bits(X.to_numpy().reshape(batch, n, 1)), bits(ox)), f"solve {n}"


@pytest.mark.parametrize("n", [16, 64])
def test_large_inverse_is_an_inverse(ctx, n):
    """Independent of the oracle: A . A^-1 is the identity to rounding for well-conditioned items (diagonally dominant)."""
    batch = 200
    a = oracle.fill(oracle.F64, oracle.SEED_A, 70, batch * n * n).reshape(batch, n, n).copy()
    a += np.eye(n) * n
    A = ctx.from_numpy(a, (n, n))
    inv, st = tb.inverse(A)
    assert st.to_numpy().sum() == 0
    P = tb.matmul(A, inv).to_numpy()
    assert np.abs(P - np.eye(n)).max() < 1e-12
    d = tb.det(A).to_numpy()
    assert np.allclose(d, np.linalg.det(a), rtol=1e-10)


def test_large_solve_wide_rhs_and_host_entry(ctx):
    """More right-hand-side columns than one pass holds (n = 9: 16 per pass), and the host-array entry on n > 8."""
    n, nrhs, batch = 9, 40, 301
    a = inputs(oracle.F64, n, batch, 71)
    b = oracle.fill(oracle.F64, oracle.SEED_B, 71, batch * n * nrhs)
    X, st = tb.solve(ctx.from_numpy(a, (n, n)), ctx.from_numpy(b, (n, nrhs)))
    ox, ost = oracle.solve(a, b, n, nrhs, threads=8)
    assert np.array_equal(st.to_numpy(), ost) and np.array_equal(bits(X.to_numpy()), bits(ox))
    inv, st = host.inverse(ctx, a, n)
    oinv, ost = oracle.inverse(a, n, threads=8)
    assert np.array_equal(st, ost) and np.array_equal(bits(inv), bits(oinv))
    n = 128
    a = inputs(oracle.F64, n, 9, 72)
    assert np.array_equal(bits(host.det(ctx, a, n)), bits(oracle.det(a, n)))


def test_small_n_with_many_right_hand_sides(ctx):
    """n <= 8 with a column count that has no record-per-thread / row-per-lane kernel lands on the CTA-per-matrix kernel."""
    for n, nrhs in [(4, 20), (3, 9), (8, 13), (1, 5), (2, 130)]:
        batch = 203
        a = inputs(oracle.F64, n, batch, 74) if n >= 3 else oracle.fill(oracle.F64, oracle.SEED_A, 74, batch * n * n).reshape(batch, n, n)
        b = oracle.fill(oracle.F64, oracle.SEED_B, 74, batch * n * nrhs)
        X, st = tb.solve(ctx.from_numpy(a, (n, n)), ctx.from_numpy(b, (n, nrhs)))
        ox, ost = oracle.solve(a, b, n, nrhs, threads=8)
        assert np.array_equal(st.to_numpy(), ost) and np.array_equal(bits(X.to_numpy()), bits(ox)), (n, nrhs)


def test_large_elimination_soa_and_limits(ctx):
    n, batch = 12, 100
    a = inputs(oracle.F32, n, batch, 73)
    inv, st = tb.inverse(ctx.from_numpy(a, (n, n), tb.SOA))          # SoA: through the relayout detour
    oinv, ost = oracle.inverse(a, n)
    assert inv.layout == tb.SOA and np.array_equal(st.to_numpy(), ost) and np.array_equal(bits(inv.to_numpy()), bits(oinv))
    with pytest.raises(tb.Unsupported):
        tb.det(ctx.empty((129, 129), np.float64, 2))                  # documented limit: n <= 128
    with pytest.raises(tb.Unsupported):
        tb.det(ctx.empty((12, 12), np.int64, 2))                      # needs Fractional
