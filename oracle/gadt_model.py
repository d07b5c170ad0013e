"""ORACLE — test infrastructure only.

Independent pure-Python restatement of the reference's *GADT* backend
(src/Data/Tensor.hs:142-147 and the MultiIndexable instance :226-255).  The
reference's own tests run on this backend only (tests/Tensor.hs:16-19, :42-80)
and tie ``unRev`` to iterated ``unConcat`` (:62-79); ``transpose = unRev . t0``
(src/Data/Indexable.hs:101-102).  tests/ use this model as a second derivation
to pin the flat-vector ``transposeV`` restatement in t5_oracle.cpp.

Representation: ``T0 e`` is the class below; ``T1 t`` is ``[t]``; ``t :| ts`` is
``[t] + ts`` — i.e. a rank>=1 tensor is the Python list of its slices along the
first index.
"""
from __future__ import annotations


class T0:
    __slots__ = ("e",)

    def __init__(self, e):
        self.e = e

    def __eq__(self, o):
        return isinstance(o, T0) and self.e == o.e

    def __repr__(self):
        return f"T0({self.e!r})"


def tmap(f, t):
    # Tensor.hs:215-217
    if isinstance(t, T0):
        return T0(f(t.e))
    return [tmap(f, s) for s in t]


def tap(fs, t):
    # Tensor.hs:218-220
    if isinstance(fs, T0):
        return T0(fs.e(t.e))
    assert len(fs) == len(t)
    return [tap(f, s) for f, s in zip(fs, t)]


def unT0(t):
    assert isinstance(t, T0)
    return t.e


def concat(t):
    # Tensor.hs:229-230
    #   concat (T1 t)    = map T1 t
    #   concat (t :| ts) = (:|) `map` t `ap` (concat ts)
    assert isinstance(t, list) and len(t) >= 1
    if len(t) == 1:
        return tmap(lambda inner: [inner], t[0])
    head, tail = t[0], t[1:]
    return tap(tmap(lambda u: (lambda us: [u] + us), head), concat(tail))


def unconcat(t):
    # Tensor.hs:231-242
    if isinstance(t, T0):
        inner = t.e
        assert isinstance(inner, list)
        if len(inner) == 1:                      # unConcat (T0 (T1 t)) = T1 $ T0 t
            return [T0(inner[0])]
        return [T0(inner[0])] + unconcat(T0(inner[1:]))   # T0 t :| unConcat (T0 ts)
    if len(t) == 1:                              # unConcat (T1 t) = putT1 $ unConcat t
        return [[u] for u in unconcat(t[0])]
    a, b = unconcat(t[0]), unconcat(t[1:])       # unConcat t |:: unConcat ts
    assert len(a) == len(b)
    return [[u] + v for u, v in zip(a, b)]


def inner_rank(t):
    """Rank of the element tensors of an outer tensor whose elements are tensors."""
    while not isinstance(t, T0):
        t = t[0]
    r, e = 0, t.e
    while not isinstance(e, T0):
        r, e = r + 1, e[0]
    return r


def outer_rank(t):
    r = 0
    while not isinstance(t, T0):
        r, t = r + 1, t[0]
    return r


def rev(t):
    # Tensor.hs:248-251: rev' R0 = unT0 ; rev' (R1 r) = rev' r . concat
    while outer_rank(t) > 0:
        t = concat(t)
    return unT0(t)


def unrev(t):
    # Tensor.hs:252-255: unRev' R0 = fmap unT0 ; unRev' (R1 r) = unRev' r . unConcat
    while inner_rank(t) > 0:
        t = unconcat(t)
    return tmap(unT0, t)


def transpose(t):
    # Indexable.hs:101-102: transpose = unRev . t0
    return unrev(T0(t))


# ---- row-major list <-> GADT tensor (Tensor.hs:294-300 IsList instance) -----
def from_list(dims, flat):
    dims = list(dims)
    if not dims:
        assert len(flat) == 1
        return T0(flat[0])
    step = 1
    for d in dims[1:]:
        step *= d
    assert len(flat) == dims[0] * step
    return [from_list(dims[1:], flat[i * step:(i + 1) * step]) for i in range(dims[0])]


def to_list(t):
    if isinstance(t, T0):
        return [t.e]
    out = []
    for s in t:
        out.extend(to_list(s))
    return out


def shape(t):
    s = []
    while not isinstance(t, T0):
        s.append(len(t))
        t = t[0]
    return s


# ---- indexing, at / ta (Tensor.hs:203-206, :243-247) ---------------------------------
def index(t, ix):
    """t ! ix with a 1-based multi-index (list)."""
    for i in ix:
        t = t[i - 1]
    return unT0(t)


def at(t, ix):
    # T1 t `at` is = T1 $ T0 (t ! is) ; (t :| ts) `at` is = T0 (t ! is) :| (ts `at` is)
    return [T0(index(s, ix)) for s in t]


def ta(i, t):
    # OneCons _ `ta` (t :| _) = t ; HeadSucc is `ta` (_ :| ts) = is `ta` ts   (1-based i)
    return t[i - 1]


# ---- append / split (Tensor.hs:267-288) -----------------------------------------------
def append(n, x, y):
    if n == 1:          # A1 / A1S: walk down x's first axis, then cons onto y
        return list(x) + list(y)
    assert len(x) == len(y)
    return [append(n - 1, a, b) for a, b in zip(x, y)]   # An: zip along the first axis


def split(n, first_extent, t):
    if n == 1:
        return list(t[:first_extent]), list(t[first_extent:])
    parts = [split(n - 1, first_extent, s) for s in t]
    return [p[0] for p in parts], [p[1] for p in parts]


# ---- slice (Tensor.hs:372-379) -----------------------------------------------------------
def slice_(t, sl):
    """sl: list with None (AllCons) or a 1-based index (OneConsSl / HeadSuccSl chain)."""
    if not sl:
        assert isinstance(t, T0)
        return t
    if sl[0] is None:
        return [slice_(s, sl[1:]) for s in t]
    return slice_(t[sl[0] - 1], sl[1:])


def from_nested(outer_dims, inner_dims, flat):
    """Row-major flat data -> outer tensor whose elements are inner tensors."""
    di = 1
    for d in inner_dims:
        di *= d
    n = 1
    for d in outer_dims:
        n *= d
    inners = [from_list(inner_dims, flat[i * di:(i + 1) * di]) for i in range(n)]
    return from_list(outer_dims, inners)


def nested_to_list(t):
    """Flatten a tensor of tensors (or of scalars) row-major."""
    out = []
    for e in to_list(t):
        out.extend(to_list(e) if isinstance(e, (list, T0)) else [e])
    return out
