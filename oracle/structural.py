"""ORACLE — test infrastructure only.

Restatement of the flat-vector backend's STRUCTURAL operations
(src/Data/Tensor/Vector.hs:260-354, :400-429) as index tables: for every
operation `table[i]` is the input linear offset that output linear offset `i`
reads (pure permutation / gather, so one table per shape serves any element type
and any batch).  Each function follows the cited reference lines step by step,
including the 1-based multi-indices of linearize/unlinearize (:150-168).

Parity: this code exists in the reference, so parity is PINNED — tests/ check
these tables against oracle/gadt_model.py (the GADT backend the reference's own
tests run on, tests/Tensor.hs:25-80) and against the reference's properties
(split . append == id, append (t1 x) y == x |: y, slice x allS == x, rev/unRev
versus concat/unConcat).

One deliberate deviation: `ta` follows the GADT semantics (src/Data/Tensor.hs:245-247),
not Vector.hs:284-288 whose slice offset is off by one.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def product(xs: Sequence[int]) -> int:
    p = 1
    for x in xs:
        p *= int(x)
    return p


def linearize(ds: Sequence[int], ix: Sequence[int]) -> int:
    # Vector.hs:150-158 — go es js acc = acc + (head js - 1) * product (tail es)
    acc = 0
    es, js = list(ds), list(ix)
    while js:
        t = es[1:]
        acc += (js[0] - 1) * product(t)
        es, js = t, js[1:]
    return acc


def unlinearize(ds: Sequence[int], i: int) -> List[int]:
    # Vector.hs:160-168 — (q,r) = quotRem j (product t); emit q+1
    out, es, j = [], list(ds), int(i)
    while es:
        t = es[1:]
        q, r = divmod(j, product(t))
        out.append(q + 1)
        es, j = t, r
    return out


def transposeV(ds: Sequence[int], i: int) -> int:
    # Vector.hs:306-307
    return linearize(ds, list(reversed(unlinearize(list(reversed(ds)), i))))


# ---- append / split (IsTensor instance, Vector.hs:322-354) -------------------------
def append_table(axis: int, d: Sequence[int], e: Sequence[int]) -> Tuple[List[int], List[Tuple[int, int]]]:
    """`append sh x y` with fromPI sh == axis (1-based).  Returns (out dims, [(src, offset)])
    with src 0 = x, 1 = y.  Vector.hs:322-335."""
    i = axis - 1
    f = [d[i] + e[i]] + list(d[i + 1:])
    out_dims = list(d[:i]) + f
    length = product(out_dims)
    m, n, o = product(d[i:]), product(e[i:]), product(f)
    table = []
    for k in range(length):
        q, r = divmod(k, o)
        table.append((0, q * m + r) if r < m else (1, q * n + r - m))
    return out_dims, table


def split_tables(axis: int, ks: Sequence[int], first_extent: int):
    """`split sh t`: the two halves of `t` cut at `first_extent` along `axis` (1-based).
    Vector.hs:336-354: u1 = unsafeTensorGen f1 (unsafeTensorGet t);
    u2 = unsafeTensorGen f2 (unsafeTensorGet t . shift)."""
    i = axis - 1
    f1 = [first_extent if a == i else x for a, x in enumerate(ks)]
    f2 = [x - first_extent if a == i else x for a, x in enumerate(ks)]
    t1 = [linearize(ks, unlinearize(f1, j)) for j in range(product(f1))]
    t2 = []
    for j in range(product(f2)):
        ix = unlinearize(f2, j)
        ix = [x + first_extent if a == i else x for a, x in enumerate(ix)]
        t2.append(linearize(ks, ix))
    return (f1, t1), (f2, t2)


# ---- at / ta (MultiIndexable instance) ------------------------------------------------
def at_table(ds: Sequence[int], ix: Sequence[int]) -> Tuple[List[int], List[int]]:
    """`t at ix`: fix all but the FIRST index (Vector.hs:277-283)."""
    h, t = ds[0], list(ds[1:])
    l = product(t)
    i = linearize(t, ix)
    return [h], [x * l + i for x in range(h)]


def ta_table(ds: Sequence[int], i: int) -> Tuple[List[int], List[int]]:
    """`[i] ta t`: fix the FIRST index (1-based).  GADT semantics, Tensor.hs:245-247."""
    t = list(ds[1:])
    l = product(t)
    return t, [(i - 1) * l + x for x in range(l)]


# ---- slice (Vector.hs:400-429) -----------------------------------------------------------
def slice_size(sh: Sequence[int], sl: Sequence[int | None]) -> int:
    # Vector.hs:419-423
    acc = 1
    for i in reversed(range(len(sh))):
        if sl[i] is None:
            acc = sh[i] * acc
    return acc


def slice_table(sh: Sequence[int], sl: Sequence[int | None]) -> Tuple[List[int], List[int]]:
    """sliceV / sliceSh.  `sl[a]` is None (Nothing: keep the axis) or a 1-based index (Just)."""
    # ks = prescanr' (*) 1 sh  : strides (Vector.hs:403)
    ks = [product(sh[a + 1:]) for a in range(len(sh))]

    def f(zh, js, zl, acc, x):
        # Vector.hs:404-417
        if not zh:
            return acc
        if zl[0] is not None:
            return f(zh[1:], js[1:], zl[1:], acc + (zl[0] - 1) * js[0], x)
        q, r = divmod(x, slice_size(zh[1:], zl[1:]))
        return f(zh[1:], js[1:], zl[1:], acc + q * js[0], r)

    out_dims = [d for d, s in zip(sh, sl) if s is None]          # sliceSh, Vector.hs:425-429
    return out_dims, [f(list(sh), ks, list(sl), 0, x) for x in range(slice_size(sh, sl))]


# ---- concat / unConcat / rev / unRev on nested tensors (Vector.hs:260-301) -------------
# A nested tensor `t is (t js e)` is flat [outer row-major][inner row-major].
def concat_table(ds: Sequence[int], js: Sequence[int]):
    """concat :: t (i:is) (t js e) -> t is (t (i:js) e).  Vector.hs:260-267."""
    h, t = ds[0], list(ds[1:])
    l, dj = product(t), product(js)
    table = []
    for x in range(l):
        for k in range(h):                       # G.concat [content (u ! (x + k*l)) | k <- [0..h-1]]
            table.extend((x + k * l) * dj + e for e in range(dj))
    return (t, [h] + list(js)), table


def unconcat_table(ds: Sequence[int], es: Sequence[int]):
    """unConcat :: t is (t (i:js) e) -> t (i:is) (t js e).  Vector.hs:268-276."""
    h = es[0]
    t = list(es[1:])
    tl, de, pd = product(t), product(es), product(ds)
    table = []
    for n in range(pd * h):
        i, is_ = divmod(n, pd)
        table.extend(is_ * de + i * tl + e for e in range(tl))
    return ([h] + list(ds), t), table


def rev_table(sh: Sequence[int], js: Sequence[int]):
    """rev :: t is (t js e) -> t (reverse is ++ js) e.  Vector.hs:289-294."""
    s, dj = product(sh), product(js)
    table = []
    for i in range(s):
        src = transposeV(sh, i)
        table.extend(src * dj + e for e in range(dj))
    return list(reversed(sh)) + list(js), table


def unrev_table(sh: Sequence[int], rh: Sequence[int]):
    """unRev :: t js (t is e) -> t (reverse is ++ js) e, outer form sh (= js), inner form rh (= is).
    Vector.hs:295-301."""
    s, dr = product(sh), product(rh)
    zh = list(reversed(rh)) + list(sh)
    table = []
    for i in range(product(zh)):
        q, r = divmod(i, s)
        table.append(r * dr + transposeV(rh, q))
    return zh, table


# ---- applying a table to a batch ------------------------------------------------------------
def gather(a: np.ndarray, d_in: int, table: Sequence[int]) -> np.ndarray:
    a = np.ascontiguousarray(a).reshape(-1, d_in)
    return a[:, np.asarray(table, dtype=np.int64)]


def gather2(a: np.ndarray, da: int, b: np.ndarray, db: int, table: Sequence[Tuple[int, int]]) -> np.ndarray:
    a = np.ascontiguousarray(a).reshape(-1, da)
    b = np.ascontiguousarray(b).reshape(-1, db)
    both = np.concatenate([a, b], axis=1)
    idx = np.asarray([off if src == 0 else da + off for src, off in table], dtype=np.int64)
    return both[:, idx]
