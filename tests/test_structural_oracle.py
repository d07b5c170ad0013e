"""CPU tests pinning the structural-operation oracle (oracle/structural.py, a restatement of
src/Data/Tensor/Vector.hs:260-354, :400-429) against the GADT backend restatement
(oracle/gadt_model.py, src/Data/Tensor.hs:226-288, :372-379) and against the reference's own
properties (tests/Tensor.hs:25-80)."""
import itertools

import numpy as np
import pytest

from oracle import gadt_model as G
from oracle import structural as S


def flat(dims, start=0):
    return list(range(start, start + S.product(dims)))


@pytest.mark.parametrize("d,e,axis", [([2], [3], 1), ([2, 3], [3, 3], 1), ([1, 3], [2, 3], 1), ([2, 2], [2, 3], 2),
                                      ([2, 3, 2], [2, 1, 2], 2), ([2, 2, 3], [2, 2, 4], 3), ([4, 4], [4, 4], 1)])
def test_append_and_split_match_gadt_and_roundtrip(d, e, axis):
    x, y = flat(d), flat(e, 1000)
    out_dims, table = S.append_table(axis, d, e)
    got = S.gather2(np.array(x), len(x), np.array(y), len(y), table)[0].tolist()
    want = G.append(axis, G.from_list(d, x), G.from_list(e, y))
    assert G.shape(want) == out_dims and G.to_list(want) == got
    # split sh . append sh == id   (tests/Tensor.hs:25-29)
    (f1, t1), (f2, t2) = S.split_tables(axis, out_dims, d[axis - 1])
    assert f1 == list(d) and f2 == list(e)
    assert [got[i] for i in t1] == x and [got[i] for i in t2] == y
    g1, g2 = G.split(axis, d[axis - 1], want)
    assert G.to_list(g1) == x and G.to_list(g2) == y


def test_append_t1_is_cons():
    # append SOne (t1 x) y == x |: y   (tests/Tensor.hs:31-34), x :: [3], y :: [2,3]
    x, y = flat([3]), flat([2, 3], 100)
    _, table = S.append_table(1, [1, 3], [2, 3])
    got = S.gather2(np.array(x), 3, np.array(y), 6, table)[0].tolist()
    assert got == x + y


@pytest.mark.parametrize("dims", [[2, 3], [3, 2, 2], [4], [2, 1, 3, 2]])
def test_at_ta_match_gadt(dims):
    x = flat(dims)
    t = G.from_list(dims, x)
    for i in range(1, dims[0] + 1):
        od, table = S.ta_table(dims, i)
        assert od == dims[1:] and [x[j] for j in table] == G.to_list(G.ta(i, t))
    for ix in itertools.product(*[range(1, d + 1) for d in dims[1:]]):
        od, table = S.at_table(dims, list(ix))
        assert od == [dims[0]] and [x[j] for j in table] == G.to_list(G.at(t, list(ix)))


@pytest.mark.parametrize("dims,sl", [([], []), ([2, 3], [None, None]), ([2, 3], [2, None]), ([2, 3], [None, 3]),
                                     ([3, 2, 4], [None, 2, None]), ([3, 2, 4], [3, None, 1]), ([2, 2], [1, 2])])
def test_slice_matches_gadt(dims, sl):
    x = flat(dims)
    od, table = S.slice_table(dims, sl)
    want = G.slice_(G.from_list(dims, x), sl)
    assert G.shape(want) == od and G.to_list(want) == [x[j] for j in table]
    if all(s is None for s in sl):          # slice x (allCons ... nilS) == x  (tests/Tensor.hs:36-40)
        assert table == list(range(len(x)))


@pytest.mark.parametrize("outer,inner", [([], []), ([], [2, 3]), ([2], []), ([2], [3]), ([2], [3, 4]),
                                         ([2, 3], []), ([2, 3], [4]), ([2, 3], [4, 5]), ([3, 2, 2], [2])])
def test_rev_unrev_concat_unconcat_match_gadt(outer, inner):
    x = flat(list(outer) + list(inner))
    t = G.from_nested(outer, inner, x)
    # rev (tests/Tensor.hs:44-61)
    od, table = S.rev_table(outer, inner)
    want = G.rev(t)
    assert G.shape(want) == od and G.to_list(want) == [x[j] for j in table]
    # unRev (tests/Tensor.hs:62-79): outer form = js, inner form = is
    od, table = S.unrev_table(outer, inner)
    want = G.unrev(t)
    assert G.shape(want) == od and G.to_list(want) == [x[j] for j in table]
    if outer:
        (o2, i2), table = S.concat_table(outer, inner)
        want = G.concat(t)
        assert G.nested_to_list(want) == [x[j] for j in table]
        assert G.shape(want) == o2        # outer shape of the result
    if inner:
        (o2, i2), table = S.unconcat_table(outer, inner)
        want = G.unconcat(t)
        assert G.nested_to_list(want) == [x[j] for j in table]


def test_transpose_is_unrev_of_t0():
    for dims in ([2, 3], [2, 3, 4], [3]):
        x = flat(dims)
        od, table = S.unrev_table([], dims)        # transpose = unRev . t0  (Indexable.hs:101-102)
        assert od == list(reversed(dims))
        assert [x[j] for j in table] == [x[S.transposeV(dims, i)] for i in range(len(x))]


# ---- the library's own host-side table builder (pure C function, no GPU needed) -----------
def _lib_table(op, dims, params):
    from tensor_b200 import _lib
    return _lib.structural_table(op, dims, params)


def test_library_tables_match_reference_restatement():
    from tensor_b200 import _lib
    for d, e, axis in [([2], [3], 1), ([2, 3], [3, 3], 1), ([2, 2], [2, 3], 2), ([2, 3, 2], [2, 1, 2], 2), ([4, 4], [4, 4], 1)]:
        od, table = S.append_table(axis, d, e)
        dA = S.product(d)
        want = [off if src == 0 else dA + off for src, off in table]
        assert _lib_table(_lib.ST_APPEND, d, [axis] + list(e)) == want
        (f1, t1), (f2, t2) = S.split_tables(axis, od, d[axis - 1])
        assert _lib_table(_lib.ST_SPLIT_FIRST, od, [axis, d[axis - 1]]) == t1
        assert _lib_table(_lib.ST_SPLIT_SECOND, od, [axis, d[axis - 1]]) == t2
    for dims in ([2, 3], [3, 2, 2], [4], [2, 1, 3, 2]):
        for i in range(1, dims[0] + 1):
            assert _lib_table(_lib.ST_TA, dims, [i]) == S.ta_table(dims, i)[1]
        for ix in itertools.product(*[range(1, d + 1) for d in dims[1:]]):
            assert _lib_table(_lib.ST_AT, dims, list(ix)) == S.at_table(dims, list(ix))[1]
    for dims, sl in [([2, 3], [None, None]), ([2, 3], [2, None]), ([3, 2, 4], [None, 2, None]), ([3, 2, 4], [3, None, 1]), ([2, 2], [1, 2])]:
        assert _lib_table(_lib.ST_SLICE, dims, [0 if s is None else s for s in sl]) == S.slice_table(dims, sl)[1]
    # nested reshuffles are axis permutations of the flat [outer][inner] layout
    for outer, inner in [([2], [3]), ([2, 3], [4]), ([2, 3], [4, 5]), ([3, 2, 2], [2]), ([], [2, 3]), ([2, 3], [])]:
        dims = list(outer) + list(inner)
        r, n = len(outer), len(dims)
        assert _lib_table(_lib.ST_PERMUTE, dims, list(reversed(range(r))) + list(range(r, n))) == S.rev_table(outer, inner)[1]
        assert _lib_table(_lib.ST_PERMUTE, dims, list(reversed(range(r, n))) + list(range(r))) == S.unrev_table(outer, inner)[1]
        if outer:
            assert _lib_table(_lib.ST_PERMUTE, dims, list(range(1, r)) + [0] + list(range(r, n))) == S.concat_table(outer, inner)[1]
        if inner:
            assert _lib_table(_lib.ST_PERMUTE, dims, [r] + list(range(r)) + list(range(r + 1, n))) == S.unconcat_table(outer, inner)[1]


def test_library_table_rejects_bad_requests():
    from tensor_b200 import _lib
    with pytest.raises(ValueError):
        _lib_table(_lib.ST_APPEND, [2, 3], [1, 2, 4])       # other axes differ
    with pytest.raises(ValueError):
        _lib_table(_lib.ST_TA, [2, 3], [3])                 # index out of range
    with pytest.raises(ValueError):
        _lib_table(_lib.ST_PERMUTE, [2, 3], [0, 0])         # not a permutation
    with pytest.raises(ValueError):
        _lib_table(_lib.ST_SPLIT_FIRST, [2, 3], [1, 2])     # empty second half
