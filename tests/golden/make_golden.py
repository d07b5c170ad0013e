"""Generates tests/golden/*.json.  Run from the repo root:  python tests/golden/make_golden.py

Two kinds of fixture:
  * transpose_known.json — pinned WITHOUT the oracle: (a) the known answers
    hand-derived in SURVEY.md §3a from src/Data/Tensor/Vector.hs:295-307, and
    (b) transposes produced by oracle/gadt_model.py, the restatement of the GADT
    backend the reference's own tests run on (tests/Tensor.hs:62-79).
  * la_regression.json — outputs of the C++ oracle on seeded inputs for the
    operations the reference checkout does not contain (parity unpinned,
    SURVEY.md §0 F1): a regression net so the specification cannot drift silently.
The Haskell reference cannot be executed here (no ghc), so no fixture is
"reference-executed".
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import gadt_model as G  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    known = [
        {"source": "SURVEY §3a (hand-derived from Vector.hs:295-307)", "dims": [2, 3],
         "in": [1, 2, 3, 4, 5, 6], "out": [1, 4, 2, 5, 3, 6]},
        {"source": "SURVEY §3a (hand-derived from Vector.hs:295-307)", "dims": [2, 3, 4],
         "in": list(range(24)),
         "out": [0, 12, 4, 16, 8, 20, 1, 13, 5, 17, 9, 21, 2, 14, 6, 18, 10, 22, 3, 15, 7, 19, 11, 23]},
        {"source": "SURVEY §3a: rank 1 is the identity", "dims": [5], "in": [9, 8, 7, 6, 5], "out": [9, 8, 7, 6, 5]},
    ]
    rng = np.random.default_rng(20261019)
    for dims in ([], [1], [4], [2, 2], [3, 3], [4, 4], [2, 3], [3, 1, 2], [2, 3, 4], [2, 2, 2, 2], [3, 2, 1, 4], [8, 8]):
        d = int(np.prod(dims)) if dims else 1
        flat = [int(x) for x in rng.integers(-1000, 1000, size=d)]
        t = G.transpose(G.from_list(dims, flat))
        known.append({"source": "oracle/gadt_model.py (Tensor.hs:229-255 unRev/unConcat; Indexable.hs:101-102)",
                      "dims": dims, "in": flat, "out": G.to_list(t), "out_dims": G.shape(t)})
    with open(os.path.join(HERE, "transpose_known.json"), "w") as f:
        json.dump(known, f, indent=1)

    reg = []
    batch = 5

    def rec(op, dtype, params, ins, outs):
        reg.append({"op": op, "dtype": dtype, "params": params,
                    "in": [np.asarray(x).ravel().tolist() for x in ins],
                    "out": [np.asarray(x).ravel().tolist() for x in outs]})

    for dname, dc in (("f64", oracle.F64), ("f32", oracle.F32), ("i64", oracle.I64)):
        for (m, k, n) in ((3, 3, 3), (4, 4, 4), (4, 4, 1), (2, 5, 3)):
            a = oracle.fill(dc, oracle.SEED_A, 1, batch * m * k)
            b = oracle.fill(dc, oracle.SEED_B, 1, batch * k * n)
            rec("matmul", dname, {"m": m, "k": k, "n": n}, [a, b], [oracle.matmul(a, b, m, k, n)])
        x = oracle.fill(dc, oracle.SEED_A, 2, batch * 4)
        y = oracle.fill(dc, oracle.SEED_B, 2, batch * 4)
        rec("dot", dname, {"n": 4}, [x, y], [oracle.dot(x, y, 4)])
        a = oracle.fill(dc, oracle.SEED_A, 4, batch * 4)
        b = oracle.fill(dc, oracle.SEED_B, 4, batch * 6)
        rec("prod", dname, {"dims_a": [2, 2], "dims_b": [2, 3], "n": 0}, [a, b], [oracle.prod(a, [2, 2], b, [2, 3], 0)])
        rec("prod", dname, {"dims_a": [2, 2], "dims_b": [2, 3], "n": 1}, [a, b], [oracle.prod(a, [2, 2], b, [2, 3], 1)])
    for dname, dc in (("f64", oracle.F64), ("f32", oracle.F32)):
        for n in (2, 3, 4, 5):
            a = oracle.fill(dc, oracle.SEED_A, 3, batch * n * n).reshape(batch, n, n).copy()
            oracle.plant_specials(a, n, 0, swap_every=2, sing_every=4)
            b = oracle.fill(dc, oracle.SEED_B, 3, batch * n)
            rec("det", dname, {"n": n}, [a], [oracle.det(a, n)])
            inv, st = oracle.inverse(a, n)
            rec("inverse", dname, {"n": n}, [a], [inv, st])
            x, st = oracle.solve(a, b, n, 1)
            rec("solve", dname, {"n": n, "nrhs": 1}, [a, b], [x, st])
    with open(os.path.join(HERE, "la_regression.json"), "w") as f:
        json.dump(reg, f)
    print("wrote", len(known), "transpose cases and", len(reg), "regression cases")

    # the rest of SquareMatrix (tr / polyEval / charPoly / minPoly): oracle-generated
    # regression net + hand-checkable structured matrices whose minimal polynomial is known exactly
    reg = []
    structured = {
        2: [np.eye(2) * 3, [[0, 1], [0, 0]], [[2, 0], [0, 5]], np.zeros((2, 2)), [[1, 1], [0, 1]]],
        3: [np.eye(3) * 2, np.diag([2.0, 2, 3]), [[0, 1, 0], [0, 0, 1], [0, 0, 0]], [[0, 1, 0], [0, 0, 0], [0, 0, 0]],
            np.zeros((3, 3)), [[1, 0, 0], [0, 0, 0], [0, 0, 0]], [[4, 1, 0], [0, 4, 0], [0, 0, 4]]],
        4: [np.eye(4), np.diag([1.0, 1, 2, 2]), np.diag([1.0, 2, 3, 4]), np.eye(4)[::-1].copy(), np.zeros((4, 4)),
            [[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0, 0, 0, 0]]],
    }
    for dname, dc in (("f64", oracle.F64), ("f32", oracle.F32)):
        npd = oracle.NP_DTYPE[dc]
        for n in (2, 3, 4, 5):
            rnd = oracle.fill(dc, oracle.SEED_A, 7, batch * n * n).reshape(batch, n, n)
            mats = [rnd[i] for i in range(batch)] + [np.asarray(m, dtype=npd) for m in structured.get(n, [])]
            a = np.ascontiguousarray(np.stack(mats), dtype=npd)
            coeffs = [0.5, -1.25, 2.0, 0.75, -0.5][: n + 1]
            rec("trace", dname, {"n": n}, [a], [oracle.trace(a, n)])
            rec("poly_eval", dname, {"n": n, "coeffs": coeffs}, [a], [oracle.poly_eval(a, coeffs, n)])
            rec("char_poly", dname, {"n": n}, [a], [oracle.char_poly(a, n)])
            mp, deg = oracle.min_poly(a, n)
            rec("min_poly", dname, {"n": n}, [a], [mp, deg])
    for n in (2, 3, 4):
        a = oracle.fill(oracle.I64, oracle.SEED_A, 7, batch * n * n).reshape(batch, n, n)
        rec("trace", "i64", {"n": n}, [a], [oracle.trace(a, n)])
        rec("poly_eval", "i64", {"n": n, "coeffs": [3, -7, 11, 2 ** 40]}, [a], [oracle.poly_eval(a, [3, -7, 11, 2 ** 40], n)])
    with open(os.path.join(HERE, "sqmat_regression.json"), "w") as f:
        json.dump(reg, f)
    print("wrote", len(reg), "SquareMatrix-extras regression cases")


if __name__ == "__main__":
    main()
