"""tr / polyEval / charPoly / minPoly — the rest of the old SquareMatrix class
(CHANGELOG.md:28-29).  The reference checkout holds no code for them, so
parity is unpinned: the CPU tests pin the oracle's specification against numpy, against
hand-checkable matrices and against tests/golden/sqmat_regression.json; the GPU tests then
hold the CUDA kernels to the oracle bit for bit, through the C ABI."""
import json
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NPD = {"f64": np.float64, "f32": np.float32, "i64": np.int64}
CODE = {"f64": oracle.F64, "f32": oracle.F32, "i64": oracle.I64}


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def rand_mats(dc, n, batch, op_id=7):
    return oracle.fill(dc, oracle.SEED_A, op_id, batch * n * n).reshape(batch, n, n)


# ---- CPU: the oracle against numpy and known answers ------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8])
def test_trace_and_char_poly_against_numpy(n):
    a = rand_mats(oracle.F64, n, 50)
    assert np.allclose(oracle.trace(a, n), np.trace(a, axis1=1, axis2=2), rtol=0, atol=1e-14 * n)
    cp = oracle.char_poly(a, n)
    assert np.all(cp[:, n] == 1.0)
    for i in range(a.shape[0]):
        want = np.poly(a[i])[::-1]                      # numpy: high order first, via eigenvalues
        assert np.allclose(cp[i], want, rtol=1e-9, atol=1e-9)
    # Cayley-Hamilton through polyEval: p_A(A) = 0, and c0 = (-1)^n det A
    for i in range(a.shape[0]):
        z = oracle.poly_eval(a[i:i + 1], cp[i], n)[0]
        assert np.abs(z).max() <= 1e-10
    assert np.allclose(cp[:, 0], (-1) ** n * oracle.det(a, n), rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("dname", ["f64", "f32", "i64"])
def test_poly_eval_against_numpy(dname):
    n, batch = 3, 20
    a = rand_mats(CODE[dname], n, batch)
    if dname == "i64":
        a = a % 1000                       # keep numpy's int64 reference free of overflow
        coeffs = [3, -7, 11, 5]
    else:
        coeffs = [0.5, -1.25, 2.0, 0.75]
    got = oracle.poly_eval(a, coeffs, n)
    eye = np.eye(n, dtype=a.dtype)
    for i in range(batch):
        A = a[i]
        want = coeffs[0] * eye + coeffs[1] * A + coeffs[2] * (A @ A) + coeffs[3] * (A @ A @ A)
        if dname == "i64":
            assert np.array_equal(got[i], want)
        else:
            assert np.allclose(got[i], want, rtol=1e-5 if dname == "f32" else 1e-12, atol=1e-5 if dname == "f32" else 1e-12)
    assert np.array_equal(oracle.poly_eval(a, [], n), np.zeros_like(a))           # empty polynomial
    c0 = oracle.poly_eval(a, coeffs[:1], n)                                       # constant: c0 I
    assert np.array_equal(c0, np.broadcast_to(coeffs[0] * eye, a.shape).astype(a.dtype))
    assert np.array_equal(bits(oracle.poly_eval(a, [0, 1], n)), bits(oracle.matmul(np.broadcast_to(eye, a.shape).copy(), a, n, n, n)))


def test_int64_trace_and_poly_eval_wrap():
    a = np.array([[[2 ** 62, 1], [3, 2 ** 62]]], dtype=np.int64)
    assert oracle.trace(a, 2)[0] == np.int64(-2 ** 63)                            # 2^62 + 2^62 wraps
    got = oracle.poly_eval(a, [1, 0, 1], 2)[0]                                    # I + A^2 mod 2^64
    A = [[2 ** 62, 1], [3, 2 ** 62]]
    sq = [[sum(A[i][k] * A[k][j] for k in range(2)) for j in range(2)] for i in range(2)]
    want = [[(sq[i][j] + (1 if i == j else 0) + 2 ** 63) % 2 ** 64 - 2 ** 63 for j in range(2)] for i in range(2)]
    assert got.tolist() == want


def test_min_poly_known_answers():
    cases = [  # (matrix, degree, coefficients low order first)
        (np.eye(3) * 2, 1, [-2, 1]),                                        # x - 2
        (np.diag([2.0, 2, 3]), 2, [6, -5, 1]),                              # (x-2)(x-3)
        ([[0, 1, 0], [0, 0, 1], [0, 0, 0]], 3, [0, 0, 0, 1]),               # nilpotent Jordan block: x^3
        ([[0, 1, 0], [0, 0, 0], [0, 0, 0]], 2, [0, 0, 1]),                  # x^2
        (np.zeros((3, 3)), 1, [0, 1]),                                      # x
        ([[1, 0, 0], [0, 0, 0], [0, 0, 0]], 2, [0, -1, 1]),                 # projection: x^2 - x
        ([[4, 1, 0], [0, 4, 0], [0, 0, 4]], 2, [16, -8, 1]),                # (x-4)^2
    ]
    a = np.array([np.asarray(m, dtype=np.float64) for m, _, _ in cases])
    mp, deg = oracle.min_poly(a, 3)
    for i, (_, d, c) in enumerate(cases):
        assert deg[i] == d
        assert mp[i].tolist() == [float(x) for x in c] + [0.0] * (3 - d)
    # a generic matrix: degree n, coefficients = charPoly (bit for bit)
    g = rand_mats(oracle.F64, 4, 16)
    mp, deg = oracle.min_poly(g, 4)
    assert np.all(deg == 4)
    assert np.array_equal(bits(mp), bits(oracle.char_poly(g, 4)))
    # the minimal polynomial annihilates the matrix (exactly: small integers)
    mp3, deg3 = oracle.min_poly(a, 3)
    for i in range(a.shape[0]):
        z = oracle.poly_eval(a[i:i + 1], mp3[i][: deg3[i] + 1], 3)[0]
        assert np.array_equal(z, np.zeros((3, 3)))


def load_fixture():
    with open(os.path.join(GOLD, "sqmat_regression.json")) as f:
        return json.load(f)


def run_oracle_case(c):
    dt = NPD[c["dtype"]]
    n = c["params"]["n"]
    a = np.asarray(c["in"][0], dtype=dt).reshape(-1, n, n)
    if c["op"] == "trace":
        return [oracle.trace(a, n)]
    if c["op"] == "poly_eval":
        return [oracle.poly_eval(a, c["params"]["coeffs"], n)]
    if c["op"] == "char_poly":
        return [oracle.char_poly(a, n)]
    return list(oracle.min_poly(a, n))


def test_sqmat_regression_fixtures():
    cases = load_fixture()
    assert len(cases) == 38
    for c in cases:
        for got, want in zip(run_oracle_case(c), c["out"]):
            want = np.asarray(want, dtype=got.dtype)
            assert np.array_equal(bits(got).ravel(), bits(want).ravel()), (c["op"], c["dtype"], c["params"])


def test_threads_agree():
    a = rand_mats(oracle.F64, 4, 5000)
    assert np.array_equal(bits(oracle.char_poly(a, 4)), bits(oracle.char_poly(a, 4, threads=4)))
    assert np.array_equal(bits(oracle.poly_eval(a, [1, 2, 3], 4)), bits(oracle.poly_eval(a, [1, 2, 3], 4, threads=4)))


# ---- GPU: the CUDA kernels against the oracle, bit for bit ------------------------------------
def _tb():
    import tensor_b200 as tb
    return tb


@pytest.mark.gpu
@pytest.mark.parametrize("dname", ["f64", "f32", "i64"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_trace_and_poly_eval_bit_exact(ctx, dname, n, layout):
    tb = _tb()
    lay = tb.SOA if layout == "soa" else tb.AOS
    coeffs = [3, -7, 11, 5, -2] if dname == "i64" else [0.5, -1.25, 2.0, 0.75, -0.5]
    for batch in (1, 300, 4099):
        a = rand_mats(CODE[dname], n, batch)
        A = ctx.from_numpy(a, (n, n), lay)
        # (SoA beyond the tile-pipeline sizes 2..4 runs the AoS kernel through the library's relayout detour)
        assert np.array_equal(bits(tb.tr(A).to_numpy()), bits(oracle.trace(a, n))), f"tr n={n}"
        for nc in (0, 1, 2, 5):
            got = tb.poly_eval(A, coeffs[:nc]).to_numpy()
            assert np.array_equal(bits(got), bits(oracle.poly_eval(a, coeffs[:nc], n))), f"polyEval n={n} ncoef={nc}"
    if lay == tb.AOS:
        long_c = (np.arange(1, 24) % 5 - 2).astype(NPD[dname]) / (1 if dname == "i64" else 4)   # degree 22: device-side coefficients
        a = rand_mats(CODE[dname], n, 257)
        got = tb.poly_eval(ctx.from_numpy(a, (n, n)), long_c).to_numpy()
        assert np.array_equal(bits(got), bits(oracle.poly_eval(a, long_c, n)))


@pytest.mark.gpu
@pytest.mark.parametrize("dname", ["f64", "f32"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_char_poly_bit_exact(ctx, dname, n, layout):
    tb = _tb()
    lay = tb.SOA if layout == "soa" else tb.AOS
    for batch in (1, 300, 4099):
        a = rand_mats(CODE[dname], n, batch)
        A = ctx.from_numpy(a, (n, n), lay)
        got = tb.char_poly(A)
        assert got.shape == (n + 1,)
        assert np.array_equal(bits(got.to_numpy()), bits(oracle.char_poly(a, n))), f"charPoly n={n}"


@pytest.mark.gpu
@pytest.mark.parametrize("dname", ["f64", "f32"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 6])
def test_min_poly_bit_exact(ctx, dname, n):
    tb = _tb()
    dt = NPD[dname]
    rng = np.random.default_rng(n)
    a = rand_mats(CODE[dname], n, 600).copy()
    # every third matrix gets exact small-integer structure (diagonal with repeats, nilpotent, projection)
    for i in range(0, 600, 3):
        kind = (i // 3) % 4
        if kind == 0:
            a[i] = np.diag(rng.integers(1, 3, size=n)).astype(dt)
        elif kind == 1:
            a[i] = np.eye(n, k=1, dtype=dt) * (rng.integers(0, 2, size=(n, 1)))
        elif kind == 2:
            a[i] = np.diag((rng.random(n) < 0.5).astype(dt))
        else:
            a[i] = np.eye(n, dtype=dt) * rng.integers(-3, 4)
    got, deg = tb.min_poly(ctx.from_numpy(a, (n, n)))
    want, wdeg = oracle.min_poly(a, n)
    assert np.array_equal(deg.to_numpy(), wdeg)
    assert len(set(wdeg.tolist())) > 1 or n == 1
    assert np.array_equal(bits(got.to_numpy()), bits(want))


@pytest.mark.gpu
def test_sqmat_extras_shape_and_type_errors(ctx):
    tb = _tb()
    with pytest.raises(tb.WrongListLength):
        tb.tr(ctx.empty((3, 4), np.float64, 4))
    with pytest.raises(tb.Unsupported):
        tb.char_poly(ctx.empty((3, 3), np.int64, 4))          # needs Fractional
    with pytest.raises(tb.Unsupported):
        tb.min_poly(ctx.empty((7, 7), np.float64, 4))         # n <= 6
    with pytest.raises(tb.Unsupported):
        tb.poly_eval(ctx.empty((9, 9), np.float64, 4), [1.0, 2.0])


@pytest.mark.gpu
def test_sqmat_regression_fixtures_on_device(ctx):
    tb = _tb()
    for c in load_fixture():
        dt = NPD[c["dtype"]]
        n = c["params"]["n"]
        a = np.asarray(c["in"][0], dtype=dt).reshape(-1, n, n)
        A = ctx.from_numpy(a, (n, n))
        if c["op"] == "trace":
            outs = [tb.tr(A).to_numpy()]
        elif c["op"] == "poly_eval":
            outs = [tb.poly_eval(A, c["params"]["coeffs"]).to_numpy()]
        elif c["op"] == "char_poly":
            outs = [tb.char_poly(A).to_numpy()]
        else:
            o, d = tb.min_poly(A)
            outs = [o.to_numpy(), d.to_numpy()]
        for got, want in zip(outs, c["out"]):
            want = np.asarray(want, dtype=got.dtype)
            assert np.array_equal(bits(got).ravel(), bits(want).ravel()), (c["op"], c["dtype"], c["params"])
