"""CPU tests of the oracle itself: reference known answers, a second derivation from the
GADT backend, numpy second opinion, metamorphic properties."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import gadt_model as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NPD = {"f64": np.float64, "f32": np.float32, "i64": np.int64}


# ---- reference known answers -------------------------------------------------------
def test_multiindex_known_answers():
    # tests/MultiIndex.hs:12-21 (toList of the GADT index) restated with the flat index maps
    # of Vector.hs:150-168: 1-based indices, row-major.
    assert oracle.unlinearize([], 0) == []                      # toList Nil == []
    assert oracle.unlinearize([1], 0) == [1]                    # OneCons Nil == [1]
    assert oracle.unlinearize([2], 1) == [2]                    # HeadSucc (OneCons Nil) == [2]
    assert oracle.unlinearize([1, 1], 0) == [1, 1]
    assert oracle.unlinearize([1, 2], 1) == [1, 2]
    assert oracle.unlinearize([2, 2], 3) == [2, 2]
    assert oracle.linearize([2, 2], [2, 2]) == 3


def test_linearize_roundtrip():
    ds = [2, 3, 4]
    for i in range(24):
        assert oracle.linearize(ds, oracle.unlinearize(ds, i)) == i
    assert oracle.unlinearize(ds, 23) == [2, 3, 4]
    assert oracle.unlinearize(ds, 0) == [1, 1, 1]


def test_transpose_golden():
    with open(os.path.join(GOLD, "transpose_known.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 15
    for c in cases:
        out = oracle.transpose(np.asarray(c["in"], dtype=np.int64), c["dims"])
        assert out.ravel().tolist() == c["out"], c["source"]
        assert list(out.shape[1:]) == list(reversed(c["dims"]))


@pytest.mark.parametrize("dims", [[], [3], [2, 3], [3, 2, 2], [2, 1, 3, 2], [4, 4]])
def test_transpose_matches_gadt_model(dims):
    d = int(np.prod(dims)) if dims else 1
    flat = list(range(100, 100 + d))
    t = G.transpose(G.from_list(dims, flat))
    assert G.shape(t) == list(reversed(dims))
    assert G.to_list(t) == oracle.transpose(np.asarray(flat, dtype=np.int64), dims).ravel().tolist()


def test_gadt_rev_unrev_properties():
    # the reference's own properties (tests/Tensor.hs:44-79) on the GADT restatement
    x = G.from_list([2, 3], [G.from_list([4], list(range(i * 4, i * 4 + 4))) for i in range(6)])
    assert G.rev(x) == G.unT0(G.concat(G.concat(x)))
    assert G.unrev(x) == G.tmap(G.unT0, G.unconcat(x))
    y = G.T0(G.from_list([2, 3], list(range(6))))
    assert G.rev(y) == G.unT0(y)
    assert G.unrev(y) == G.tmap(G.unT0, G.unconcat(G.unconcat(y)))


def test_transpose_involution_and_dtypes():
    rng = np.random.default_rng(1)
    for dt in (np.float64, np.float32, np.int64):
        a = rng.integers(-99, 99, size=(7, 3, 4, 2)).astype(dt)
        t = oracle.transpose(a, [3, 4, 2])
        assert np.array_equal(t, a.transpose(0, 3, 2, 1))
        assert np.array_equal(oracle.transpose(t, [2, 4, 3]), a)


# ---- regression fixtures -----------------------------------------------------------
def test_la_regression_fixtures():
    with open(os.path.join(GOLD, "la_regression.json")) as f:
        cases = json.load(f)
    assert len(cases) == 45
    for c in cases:
        dt = NPD[c["dtype"]]
        ins = [np.asarray(x, dtype=dt) for x in c["in"]]
        p = c["params"]
        if c["op"] == "matmul":
            outs = [oracle.matmul(ins[0], ins[1], p["m"], p["k"], p["n"])]
        elif c["op"] == "dot":
            outs = [oracle.dot(ins[0], ins[1], p["n"])]
        elif c["op"] == "prod":
            outs = [oracle.prod(ins[0], p["dims_a"], ins[1], p["dims_b"], p["n"])]
        elif c["op"] == "det":
            outs = [oracle.det(ins[0], p["n"])]
        elif c["op"] == "inverse":
            outs = list(oracle.inverse(ins[0], p["n"]))
        else:
            outs = list(oracle.solve(ins[0], ins[1], p["n"], p["nrhs"]))
        for got, want in zip(outs, c["out"]):
            want = np.asarray(want, dtype=got.dtype)
            assert np.array_equal(got.ravel(), want, equal_nan=True), (c["op"], c["dtype"], p)


# ---- numpy second opinion ----------------------------------------------------------
@pytest.mark.parametrize("dc,tol", [(oracle.F64, 1e-12), (oracle.F32, 1e-5)])
def test_against_numpy(dc, tol):
    B = 64
    for n in (2, 3, 4, 6):
        a = oracle.fill(dc, oracle.SEED_A, 7, B * n * n).reshape(B, n, n)
        b = oracle.fill(dc, oracle.SEED_B, 7, B * n * n).reshape(B, n, n)
        ref = np.einsum("bij,bjk->bik", a.astype(np.float64), b.astype(np.float64))
        scale = np.einsum("bij,bjk->bik", np.abs(a).astype(np.float64), np.abs(b).astype(np.float64))
        assert np.all(np.abs(oracle.matmul(a, b, n, n, n) - ref) <= tol * scale)
        # elimination without partial pivoting: compare through residuals
        a64 = a.astype(np.float64)
        d = oracle.det(a, n).astype(np.float64)
        dref = np.linalg.det(a64)
        cond = np.linalg.cond(a64)
        assert np.all(np.abs(d - dref) <= tol * 50 * cond * np.maximum(1.0, np.abs(dref)))
        inv, st = oracle.inverse(a, n)
        assert st.sum() == 0
        res = np.einsum("bij,bjk->bik", a64, inv.astype(np.float64)) - np.eye(n)
        assert np.all(np.abs(res).max(axis=(1, 2)) <= tol * 200 * cond)
        rhs = oracle.fill(dc, oracle.SEED_B, 8, B * n).reshape(B, n)
        x, st = oracle.solve(a, rhs, n, 1)
        r = np.einsum("bij,bj->bi", a64, x.astype(np.float64).reshape(B, n)) - rhs
        assert np.all(np.abs(r).max(axis=1) <= tol * 200 * cond)


# ---- metamorphic properties ----------------------------------------------------------
def test_int64_product_transpose_identity_exact():
    B, n = 33, 4
    a = oracle.fill(oracle.I64, oracle.SEED_A, 9, B * n * n)
    b = oracle.fill(oracle.I64, oracle.SEED_B, 9, B * n * n)
    ab_t = oracle.transpose(oracle.matmul(a, b, n, n, n), [n, n])
    bt_at = oracle.matmul(oracle.transpose(b, [n, n]), oracle.transpose(a, [n, n]), n, n, n)
    assert np.array_equal(ab_t, bt_at)


def test_int64_wraparound_known_answer():
    a = np.array([2 ** 62, 2, 0, 1], dtype=np.int64)
    b = np.array([4, 0, 0, 2 ** 63 - 1], dtype=np.int64)
    c = oracle.matmul(a, b, 2, 2, 2).ravel()
    # [[2^62*4 + 0, 0 + 2*(2^63-1)], [0, 2^63-1]] mod 2^64 as signed
    assert c.tolist() == [0, -2, 0, 2 ** 63 - 1]


@pytest.mark.parametrize("dc", [oracle.F64, oracle.F32, oracle.I64])
def test_identity_is_exact_unit(dc):
    B, n = 17, 3
    a = oracle.fill(dc, oracle.SEED_A, 10, B * n * n).reshape(B, n, n)
    eye = np.broadcast_to(np.eye(n, dtype=a.dtype), (B, n, n)).copy()
    assert np.array_equal(oracle.matmul(a, eye, n, n, n), a)
    assert np.array_equal(oracle.matmul(eye, a, n, n, n), a)


def test_prod_special_cases():
    B = 9
    a = oracle.fill(oracle.F64, oracle.SEED_A, 11, B * 12)
    b = oracle.fill(oracle.F64, oracle.SEED_B, 11, B * 20)
    # prod 1 on matrices == .*.
    assert np.array_equal(oracle.prod(a, [3, 4], b, [4, 5], 1), oracle.matmul(a, b, 3, 4, 5))
    # prod 0 == outer product, one multiply per output
    x = a.reshape(B, 12)[:, :4].copy()
    y = b.reshape(B, 20)[:, :4].copy()
    assert np.array_equal(oracle.prod(x, [4], y, [4], 0), x[:, :, None] * y[:, None, :])
    # full contraction of two vectors == dot
    assert np.array_equal(oracle.prod(x, [4], y, [4], 1).ravel(), oracle.dot(x, y, 4))
    # prod 2 of matrices == sum over both indices, row-major order
    c = oracle.prod(a, [3, 4], a, [3, 4], 2).ravel()
    assert np.allclose(c, (a.reshape(B, 12) ** 2).sum(axis=1), rtol=1e-14)


def test_pivot_rule_and_singularity():
    # A[0,0] = 0 forces the swap with the FIRST non-zero row below
    a = np.array([[0.0, 2.0, 1.0], [0.0, 3.0, 5.0], [4.0, 1.0, 1.0]])
    d = oracle.det(a, 3)[0]
    assert d == np.float64(-(4.0 * 3.0 * (1.0 - (2.0 / 3.0) * 5.0))) or abs(d - np.linalg.det(a)) < 1e-12
    assert abs(d - np.linalg.det(a)) < 1e-12
    s = np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0], [0.5, 0.25, 4.0]])
    assert oracle.det(s, 3)[0] == 0.0 and not np.signbit(oracle.det(s, 3)[0])
    inv, st = oracle.inverse(s, 3)
    assert st[0] == 1 and np.all(inv == 0)
    x, st = oracle.solve(s, np.ones(3), 3, 1)
    assert st[0] == 1 and np.all(x == 0)
    assert oracle.det(np.eye(4), 4)[0] == 1.0


@pytest.mark.parametrize("dc", [oracle.F64, oracle.F32])
def test_inverse_is_solve_with_identity(dc):
    B, n = 40, 4
    a = oracle.fill(dc, oracle.SEED_A, 12, B * n * n).reshape(B, n, n).copy()
    oracle.plant_specials(a, n, 0, swap_every=3, sing_every=8)
    eye = np.broadcast_to(np.eye(n, dtype=a.dtype), (B, n, n)).copy()
    inv, st = oracle.inverse(a, n)
    x, st2 = oracle.solve(a, eye, n, n)
    assert np.array_equal(inv, x) and np.array_equal(st, st2)
    assert st.sum() == 5 and st[0] == 1 and st[8] == 1


def test_dot_sum():
    B = 1000
    for dc in (oracle.F64, oracle.F32):
        x = oracle.fill(dc, oracle.SEED_A, 13, B * 4)
        y = oracle.fill(dc, oracle.SEED_B, 13, B * 4)
        ref = float(np.sum(x.astype(np.float64) * y.astype(np.float64)))
        assert abs(float(oracle.dot_sum(x, y, 4)) - ref) <= (1e-12 if dc == oracle.F64 else 1e-4) * B
    xi = oracle.fill(oracle.I64, oracle.SEED_A, 13, B * 4)
    yi = oracle.fill(oracle.I64, oracle.SEED_B, 13, B * 4)
    with np.errstate(over="ignore"):
        ref = np.sum(xi.view(np.uint64) * yi.view(np.uint64), dtype=np.uint64).view(np.int64)
    assert oracle.dot_sum(xi, yi, 4) == ref


def test_threads_agree():
    B = 5000
    a = oracle.fill(oracle.F64, oracle.SEED_A, 14, B * 9)
    b = oracle.fill(oracle.F64, oracle.SEED_B, 14, B * 9)
    assert np.array_equal(oracle.matmul(a, b, 3, 3, 3, threads=1), oracle.matmul(a, b, 3, 3, 3, threads=4))


def test_synthetic_generators_agree():
    import tensor_b200.synthetic as syn
    for dt, dc in ((np.float64, oracle.F64), (np.float32, oracle.F32), (np.int64, oracle.I64)):
        a = syn.fill(dt, syn.SEED_A, 5, 4099, first_elem=123457)
        b = oracle.fill(dc, oracle.SEED_A, 5, 4099, first_elem=123457)
        assert np.array_equal(a, b)
        if dt != np.int64:
            assert a.min() >= -1.0 and a.max() < 1.0
