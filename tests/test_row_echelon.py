"""rowEchelonForm / triangularSolve (SURVEY.md §8f-4 "rowEchelonForm-style outputs").  The class methods left
the reference with Data.Tensor.LinearAlgebra (CHANGELOG.md:24-29), so parity is unpinned: the CPU tests pin
the oracle's specification against hand-derived known answers (exact arithmetic), numpy and algebraic
properties; the GPU tests hold the CUDA kernel to the oracle bit for bit, through the C ABI."""
import json
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "row_echelon_regression.json")
NPD = {"f64": np.float64, "f32": np.float32}
CODE = {"f64": oracle.F64, "f32": oracle.F32}
SHAPES = [(1, 1), (2, 2), (3, 3), (4, 4), (3, 4), (4, 5), (4, 8), (5, 3), (6, 7), (8, 8), (8, 16), (2, 16), (8, 1), (7, 7), (8, 9), (6, 5), (7, 9)]


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def load():
    with open(GOLD) as f:
        return json.load(f)


def mats(dname, m, n, batch, op_id=11):
    """Seeded matrices; every 7th gets an exact rank defect (zero column / repeated row / doubled row / zero corner)."""
    a = oracle.fill(CODE[dname], oracle.SEED_A, op_id, batch * m * n).reshape(batch, m, n).copy()
    rng = np.random.default_rng(m * 100 + n)
    for i in range(0, batch, 7):
        kind = (i // 7) % 5
        if kind == 0:
            a[i, :, rng.integers(0, n)] = 0
        elif kind == 1 and m > 1:
            a[i, m - 1] = a[i, 0]
        elif kind == 2 and m > 1:
            a[i, 1] = 2 * a[i, 0]
        elif kind == 3:
            a[i, 0, 0] = 0
        else:
            a[i] = 0
    return a


# ---- CPU: the specification --------------------------------------------------------------
def test_known_answers_exact():
    n_known = 0
    for c in load():
        if not c["source"].startswith("hand-derived"):
            continue
        m, n = c["rows"], c["cols"]
        for dt in (np.float64, np.float32):
            a = np.asarray(c["in"], dtype=dt).reshape(1, m, n)
            e, r = oracle.row_echelon(a, m, n, c["pivot_cols"])
            assert e.ravel().tolist() == c["out"], c
            assert r.tolist() == c["rank"], c
        n_known += 1
    assert n_known >= 10


def test_regression_fixture_cpu():
    for c in load():
        dt = NPD[c["dtype"]]
        m, n = c["rows"], c["cols"]
        a = np.asarray(c["in"], dtype=dt).reshape(-1, m, n)
        e, r = oracle.row_echelon(a, m, n, c["pivot_cols"])
        assert np.array_equal(bits(e).ravel(), bits(np.asarray(c["out"], dtype=dt)).ravel()), (c["dtype"], m, n)
        assert r.tolist() == c["rank"]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 8])
def test_nonsingular_square_reduces_to_identity_and_solves(n):
    batch = 40
    a = oracle.fill(oracle.F64, oracle.SEED_A, 12, batch * n * n).reshape(batch, n, n)
    e, r = oracle.row_echelon(a, n, n)
    assert np.all(r == n)
    assert np.array_equal(e, np.broadcast_to(np.eye(n), e.shape))              # pivots are set to 1, cleared entries to 0
    b = oracle.fill(oracle.F64, oracle.SEED_B, 12, batch * n * 2).reshape(batch, n, 2)
    aug = np.concatenate([a, b], axis=2)
    e, r = oracle.row_echelon(aug, n, n + 2, n)                                # triangularSolve: [I | A^-1 b]
    assert np.all(r == n)
    want = np.linalg.solve(a, b)
    assert np.allclose(e[:, :, n:], want, rtol=1e-8, atol=1e-8)
    inv, _ = oracle.row_echelon(np.concatenate([a, np.broadcast_to(np.eye(n), a.shape)], axis=2), n, 2 * n, n)
    assert np.allclose(inv[:, :, n:], np.linalg.inv(a), rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("m,n", SHAPES)
def test_structure_rank_and_idempotence(m, n):
    a = mats("f64", m, n, 70)
    e, r = oracle.row_echelon(a, m, n)
    assert np.all(r <= min(m, n))
    for i in range(a.shape[0]):
        k = int(r[i])
        assert k == np.linalg.matrix_rank(a[i]), i                               # the planted defects are exact
        assert np.all(e[i, k:] == 0)                                             # zero rows at the bottom
        lead = [int(np.flatnonzero(e[i, row])[0]) for row in range(k)]
        assert lead == sorted(set(lead))                                         # pivot columns strictly increase
        for row, c in enumerate(lead):
            col = e[i, :, c]
            assert col[row] == 1 and np.count_nonzero(col) == 1                  # unit pivot columns
    e2, r2 = oracle.row_echelon(e, m, n)
    assert np.array_equal(e2, e) and np.array_equal(r2, r)                       # idempotent (values; a zero may change sign)


def test_pivot_cols_limits_the_search():
    a = np.array([[[1.0, 2, 5], [2, 4, 7]]])
    e, r = oracle.row_echelon(a, 2, 3, 2)
    assert r[0] == 1 and e[0].tolist() == [[1, 2, 5], [0, 0, -3]]
    e, r = oracle.row_echelon(a, 2, 3, 0)
    assert r[0] == 0 and np.array_equal(e, a)


# ---- GPU: the kernel against the oracle, bit for bit ---------------------------------------
def _tb():
    import tensor_b200 as tb
    return tb


@pytest.mark.gpu
@pytest.mark.parametrize("dname", ["f64", "f32"])
@pytest.mark.parametrize("m,n", SHAPES)
def test_row_echelon_bit_exact(ctx, dname, m, n):
    tb = _tb()
    for batch in (1, 300, 4099):
        a = mats(dname, m, n, batch)
        for pc in sorted({n, min(m, n), n // 2}):
            got, rank = tb.row_echelon_form(ctx.from_numpy(a, (m, n)), pc)
            want, wrank = oracle.row_echelon(a, m, n, pc)
            assert np.array_equal(rank.to_numpy(), wrank), (m, n, pc, batch)
            assert np.array_equal(bits(got.to_numpy()), bits(want)), (m, n, pc, batch)
            if pc == n and batch >= 300 and min(m, n) > 1:
                assert len(set(wrank.tolist())) > 1                              # rank-deficient items are present


@pytest.mark.gpu
@pytest.mark.parametrize("dname", ["f64", "f32"])
@pytest.mark.parametrize("m,n", [(2, 2), (3, 3), (4, 4), (2, 3), (3, 4), (4, 5), (2, 4), (3, 6), (4, 8),
                                 (5, 5), (6, 6), (7, 7), (8, 8), (5, 6), (6, 7), (7, 8), (8, 9)])
def test_row_echelon_register_shapes_soa_and_all_deficient(ctx, dname, m, n):
    """The shapes with a register-resident kernel: SoA layout, and batches in which EVERY item is rank-deficient
    (the run-time rank then selects the off-diagonal step instances for whole warps)."""
    tb = _tb()
    a = mats(dname, m, n, 1500, op_id=17)
    got, rank = tb.row_echelon_form(ctx.from_numpy(a, (m, n), tb.SOA))
    want, wrank = oracle.row_echelon(a, m, n)
    assert got.layout == tb.SOA
    assert np.array_equal(rank.to_numpy(), wrank)
    assert np.array_equal(bits(got.to_numpy()), bits(want))
    rng = np.random.default_rng(n)
    for variant in range(4):
        d = a.copy()
        if variant == 0:
            d[:, :, 0] = 0                                   # no pivot in the first column: the rank lags from the start
        elif variant == 1:
            d[:, :, min(1, n - 1)] = 0
        elif variant == 2:
            d[:, m - 1] = d[:, 0]                            # repeated row: an exact zero row at the end
        else:
            d[np.arange(1500), rng.integers(0, m, 1500)] = 0  # a zero row somewhere
        got, rank = tb.row_echelon_form(ctx.from_numpy(d, (m, n)))
        want, wrank = oracle.row_echelon(d, m, n)
        assert np.array_equal(rank.to_numpy(), wrank), variant
        assert np.array_equal(bits(got.to_numpy()), bits(want)), variant


@pytest.mark.gpu
def test_row_echelon_special_values(ctx):
    """Signed zeros, infinities, NaN and denormals go through the same operations as on the CPU."""
    tb = _tb()
    rng = np.random.default_rng(5)
    a = rng.standard_normal((400, 4, 5))
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -2.2e-308, 1e308])
    idx = rng.integers(0, a.size, size=600)
    a.ravel()[idx] = specials[rng.integers(0, len(specials), size=600)]
    got, rank = tb.row_echelon_form(ctx.from_numpy(a, (4, 5)))
    want, wrank = oracle.row_echelon(a, 4, 5)
    assert np.array_equal(rank.to_numpy(), wrank)
    g, w = got.to_numpy(), want
    assert np.array_equal(np.isnan(g), np.isnan(w))
    assert np.array_equal(bits(np.where(np.isnan(g), 0.0, g)), bits(np.where(np.isnan(w), 0.0, w)))
    # the any-shape kernel (two lanes per item, columns split by parity): odd and even column counts
    for m, n in ((6, 5), (5, 8), (7, 9)):
        a = rng.standard_normal((777, m, n))
        idx = rng.integers(0, a.size, size=1500)
        a.ravel()[idx] = specials[rng.integers(0, len(specials), size=1500)]
        got, rank = tb.row_echelon_form(ctx.from_numpy(a, (m, n)))
        want, wrank = oracle.row_echelon(a, m, n)
        assert np.array_equal(rank.to_numpy(), wrank), (m, n)
        g, w = got.to_numpy(), want
        assert np.array_equal(np.isnan(g), np.isnan(w)), (m, n)
        assert np.array_equal(bits(np.where(np.isnan(g), 0.0, g)), bits(np.where(np.isnan(w), 0.0, w))), (m, n)


@pytest.mark.gpu
@pytest.mark.parametrize("dname", ["f64", "f32"])
def test_triangular_solve(ctx, dname):
    tb = _tb()
    batch, n = 1000, 4
    a = mats(dname, n, n, batch, op_id=13)
    for bshape in ((n, 3), (n,)):
        b = oracle.fill(CODE[dname], oracle.SEED_B, 13, batch * int(np.prod(bshape))).reshape(batch, *bshape)
        ra, rb, rank = tb.triangular_solve(ctx.from_numpy(a, (n, n)), ctx.from_numpy(b, bshape))
        aug = np.concatenate([a, b.reshape(batch, n, -1)], axis=2)
        want, wrank = oracle.row_echelon(aug, n, aug.shape[2], n)
        assert ra.shape == (n, n) and rb.shape == tuple(bshape)
        assert np.array_equal(rank.to_numpy(), wrank)
        assert np.array_equal(bits(ra.to_numpy()), bits(np.ascontiguousarray(want[:, :, :n])))
        assert np.array_equal(bits(rb.to_numpy()), bits(np.ascontiguousarray(want[:, :, n:]).reshape(batch, *bshape)))


@pytest.mark.gpu
def test_row_echelon_fixture_on_device(ctx):
    tb = _tb()
    for c in load():
        dt = NPD[c["dtype"]]
        m, n = c["rows"], c["cols"]
        a = np.asarray(c["in"], dtype=dt).reshape(-1, m, n)
        got, rank = tb.row_echelon_form(ctx.from_numpy(a, (m, n)), c["pivot_cols"])
        assert np.array_equal(bits(got.to_numpy()).ravel(), bits(np.asarray(c["out"], dtype=dt)).ravel()), (c["dtype"], m, n)
        assert rank.to_numpy().tolist() == c["rank"]


@pytest.mark.gpu
def test_row_echelon_errors(ctx):
    tb = _tb()
    with pytest.raises(tb.Unsupported):
        tb.row_echelon_form(ctx.empty((3, 3), np.int64, 4))                      # needs Fractional
    with pytest.raises(tb.Unsupported):
        tb.row_echelon_form(ctx.empty((9, 4), np.float64, 4))                    # rows <= 8
    with pytest.raises(tb.Unsupported):
        tb.row_echelon_form(ctx.empty((4, 17), np.float64, 4))                   # cols <= 16
    a = oracle.fill(oracle.F64, oracle.SEED_A, 77, 4 * 15)
    got, rank = tb.row_echelon_form(ctx.from_numpy(a, (5, 3), tb.SOA))          # SoA beyond the register-resident shapes:
    want, wrank = oracle.row_echelon(a, 5, 3, 3)                                 # served through the AoS kernel (relayout detour)
    assert got.layout == tb.SOA and np.array_equal(bits(got.to_numpy()), bits(want)) and np.array_equal(rank.to_numpy(), wrank)
    with pytest.raises(tb.WrongListLength):
        tb.row_echelon_form(ctx.empty((3, 3), np.float64, 4), 4)                 # more pivot columns than columns
    with pytest.raises(tb.WrongListLength):
        tb.row_echelon_form(ctx.empty((3,), np.float64, 4))                      # not a matrix
    with pytest.raises(tb.WrongListLength):
        tb.triangular_solve(ctx.empty((3, 3), np.float64, 4), ctx.empty((4,), np.float64, 4))
    e, r = tb.row_echelon_form(ctx.empty((3, 3), np.float64, 0))                 # empty batch
    assert e.to_numpy().shape == (0, 3, 3) and r.to_numpy().shape == (0,)
