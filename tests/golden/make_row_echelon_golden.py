"""Generates tests/golden/row_echelon_regression.json (run from the repo root:
python tests/golden/make_row_echelon_golden.py).  rowEchelonForm is absent from the reference
checkout (CHANGELOG.md:24-29): the fixture holds (a) hand-checkable known answers whose
arithmetic is exact in binary floating point, written out below WITHOUT the oracle, and
(b) oracle outputs on seeded and structured inputs as a regression net (parity unpinned)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (rows, cols, pivot_cols, matrix, reduced row echelon form, rank): every quotient and product is exact
KNOWN = [
    (3, 3, 3, [2, 4, 6, 1, 4, 5, 3, 6, 9], [1, 0, 1, 0, 1, 1, 0, 0, 0], 2),                    # third row = 1.5 x first
    (2, 2, 2, [0, 1, 1, 0], [1, 0, 0, 1], 2),                                   # row switch
    (2, 4, 4, [0, 0, 2, 4, 0, 0, 1, 2], [0, 0, 1, 2, 0, 0, 0, 0], 1),           # pivot in the third column
    (3, 2, 2, [2, 4, 1, 2, 3, 7], [1, 0, 0, 1, 0, 0], 2),                       # tall: pivot rows move up
    (2, 3, 2, [2, 0, 6, 0, 4, 2], [1, 0, 3, 0, 1, 0.5], 2),                     # triangularSolve: [A | b], pivots in A only
    (2, 3, 2, [1, 2, 5, 2, 4, 7], [1, 2, 5, 0, 0, -3], 1),                      # singular A: b is carried, never a pivot
    (2, 3, 3, [1, 2, 5, 2, 4, 7], [1, 2, 0, 0, 0, 1], 2),                       # same matrix, every column eligible
    (3, 3, 3, [0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0], 0),
    (1, 1, 1, [4], [1], 1),
    (2, 2, 0, [3, 1, 2, 5], [3, 1, 2, 5], 0),                                   # no eligible column: unchanged
]


def structured(rng, npd, m, n, count):
    """Seeded matrices with exact rank defects: zero columns, repeated / doubled rows, zero leading entries."""
    out = []
    for i in range(count):
        a = rng.integers(-4, 5, size=(m, n)).astype(npd)
        kind = i % 5
        if kind == 0 and n > 1:
            a[:, rng.integers(0, n)] = 0
        elif kind == 1 and m > 1:
            a[m - 1] = a[0]
        elif kind == 2 and m > 1:
            a[1] = 2 * a[0]
        elif kind == 3:
            a[0, 0] = 0
        out.append(a)
    return np.stack(out)


def main():
    reg = []
    for (m, n, pc, a, e, r) in KNOWN:
        reg.append({"source": "hand-derived, exact arithmetic", "dtype": "f64", "rows": m, "cols": n, "pivot_cols": pc,
                    "in": [float(x) for x in a], "out": [float(x) for x in e], "rank": [r]})
    rng = np.random.default_rng(20261020)
    for dname, dc in (("f64", oracle.F64), ("f32", oracle.F32)):
        npd = oracle.NP_DTYPE[dc]
        for (m, n, pc) in ((2, 2, 2), (3, 3, 3), (4, 4, 4), (3, 4, 3), (4, 5, 4), (4, 8, 4), (5, 3, 3), (8, 8, 8), (6, 7, 5)):
            rnd = oracle.fill(dc, oracle.SEED_A, 9, 4 * m * n).reshape(4, m, n)
            a = np.ascontiguousarray(np.concatenate([rnd, structured(rng, npd, m, n, 6)]), dtype=npd)
            e, r = oracle.row_echelon(a, m, n, pc)
            reg.append({"source": "oracle (regression net, parity unpinned)", "dtype": dname, "rows": m, "cols": n,
                        "pivot_cols": pc, "in": a.ravel().tolist(), "out": e.ravel().tolist(), "rank": r.tolist()})
    with open(os.path.join(HERE, "row_echelon_regression.json"), "w") as f:
        json.dump(reg, f)
    print("wrote", len(reg), "rowEchelonForm cases")


if __name__ == "__main__":
    main()
