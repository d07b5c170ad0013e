"""GPU parity of the structural operations (append / split / at / ta / slice / concat /
unConcat / rev / unRev) against the restatement of src/Data/Tensor/Vector.hs:260-354,
:400-429 in oracle/structural.py.  Pure index maps: bit-exact for every dtype."""
import numpy as np
import pytest

import oracle
import tensor_b200 as tb
from oracle import structural as S
from tensor_b200 import structure as st

pytestmark = pytest.mark.gpu
NPD = {oracle.F64: np.float64, oracle.F32: np.float32, oracle.I64: np.int64}
ALL = [oracle.F64, oracle.F32, oracle.I64]


def data(dc, dims, batch, op=21):
    return oracle.fill(dc, oracle.SEED_A, op, batch * S.product(dims))


@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("d,e,axis", [([2], [3], 1), ([2, 3], [3, 3], 1), ([3, 3], [3, 1], 2), ([2, 3, 2], [2, 1, 2], 2), ([4, 4], [4, 4], 1)])
def test_append_split(ctx, dc, d, e, axis):
    for batch in (1, 1000, 4099):
        x, y = data(dc, d, batch), data(dc, e, batch, 22)
        X, Y = ctx.from_numpy(x, d), ctx.from_numpy(y, e)
        od, table = S.append_table(axis, d, e)
        want = S.gather2(x, S.product(d), y, S.product(e), table)
        Z = st.append(axis, X, Y)
        assert list(Z.shape) == od
        assert np.array_equal(Z.to_numpy().reshape(batch, -1), want)
        # split n . append n == id  (tests/Tensor.hs:25-29)
        A, B = st.split(axis, d[axis - 1], Z)
        assert list(A.shape) == list(d) and list(B.shape) == list(e)
        assert np.array_equal(A.to_numpy().ravel(), x) and np.array_equal(B.to_numpy().ravel(), y)


def test_augmented_matrix_workflow(ctx):
    # [A | b] via append at axis 2, the use case SURVEY §8f-2 names
    batch = 2000
    a, b = data(oracle.F64, [3, 3], batch), data(oracle.F64, [3, 1], batch, 23)
    Ab = st.append(2, ctx.from_numpy(a, (3, 3)), ctx.from_numpy(b, (3, 1))).to_numpy()
    assert np.array_equal(Ab, np.concatenate([a.reshape(batch, 3, 3), b.reshape(batch, 3, 1)], axis=2))


@pytest.mark.parametrize("dc", ALL)
@pytest.mark.parametrize("dims", [[3, 3], [4, 4], [2, 3, 4], [5]])
def test_at_ta(ctx, dc, dims):
    batch = 777
    x = data(dc, dims, batch)
    X = ctx.from_numpy(x, dims)
    xr = x.reshape(batch, *dims)
    for i in range(1, dims[0] + 1):
        assert np.array_equal(st.ta(i, X).to_numpy(), xr[:, i - 1])                    # row extractor
    ix = [1 + (k % d) for k, d in enumerate(dims[1:])]
    got = st.at(X, ix).to_numpy()
    assert np.array_equal(got, xr[(slice(None), slice(None)) + tuple(j - 1 for j in ix)])  # column extractor
    assert np.array_equal(got.reshape(batch, -1), S.gather(x, S.product(dims), S.at_table(dims, ix)[1]))


@pytest.mark.parametrize("dims,sl", [([2, 3], [None, None]), ([2, 3], [2, None]), ([3, 2, 4], [None, 2, None]),
                                     ([3, 2, 4], [3, None, 1]), ([2, 2], [1, 2]), ([8, 8], [None, 5])])
def test_slice(ctx, dims, sl):
    for dc in ALL:
        batch = 1025
        x = data(dc, dims, batch)
        od, table = S.slice_table(dims, sl)
        Y = st.slice_(ctx.from_numpy(x, dims), sl)
        assert list(Y.shape) == od
        assert np.array_equal(Y.to_numpy().reshape(batch, -1), S.gather(x, S.product(dims), table))


@pytest.mark.parametrize("outer,inner", [([2], [3]), ([2, 3], [4]), ([2, 3], [4, 5]), ([3, 2, 2], [2]), ([], [2, 3]), ([2, 3], [])])
def test_nested_reshuffles(ctx, outer, inner):
    dims = list(outer) + list(inner)
    batch = 513
    x = data(oracle.I64, dims, batch)
    N = tb.Nested(ctx.from_numpy(x, dims), len(outer))
    d = S.product(dims)
    od, table = S.rev_table(outer, inner)
    R = st.rev(N)
    assert list(R.shape) == od and np.array_equal(R.to_numpy().reshape(batch, -1), S.gather(x, d, table))
    od, table = S.unrev_table(outer, inner)
    U = st.unrev(N)
    assert list(U.shape) == od and np.array_equal(U.to_numpy().reshape(batch, -1), S.gather(x, d, table))
    if outer:
        (o2, i2), table = S.concat_table(outer, inner)
        Cc = st.concat(N)
        assert list(Cc.outer) == o2 and list(Cc.inner) == i2
        assert np.array_equal(Cc.flat.to_numpy().reshape(batch, -1), S.gather(x, d, table))
        # rev == unT0 . concat . ... . concat  (tests/Tensor.hs:50-61)
        M = N
        while M.outer_rank > 0:
            M = st.concat(M)
        assert np.array_equal(M.flat.to_numpy().ravel(), R.to_numpy().ravel())
    if inner:
        (o2, i2), table = S.unconcat_table(outer, inner)
        Uc = st.unconcat(N)
        assert list(Uc.outer) == o2 and list(Uc.inner) == i2
        assert np.array_equal(Uc.flat.to_numpy().reshape(batch, -1), S.gather(x, d, table))
    # transpose == unRev . t0  (Indexable.hs:101-102)
    T0 = tb.Nested(N.flat, 0)
    assert np.array_equal(st.unrev(T0).to_numpy().ravel(), tb.transpose(N.flat).to_numpy().ravel())


@pytest.mark.parametrize("dc", ALL)
def test_structural_ops_on_soa_batches(ctx, dc):
    """SoA batches: every structural operation is a copy of whole rows - the same results as on AoS, in SoA."""
    for batch in (1, 1000, 4099):
        d, e, dims = [3, 3], [3, 2], [2, 3, 4]
        x, y, z = data(dc, d, batch), data(dc, e, batch, 24), data(dc, dims, batch, 25)
        Xa, Ya, Za = ctx.from_numpy(x, d), ctx.from_numpy(y, e), ctx.from_numpy(z, dims)
        Xs, Ys, Zs = ctx.from_numpy(x, d, tb.SOA), ctx.from_numpy(y, e, tb.SOA), ctx.from_numpy(z, dims, tb.SOA)
        J = st.append(2, Xs, Ys)
        assert J.layout == tb.SOA and np.array_equal(J.to_numpy(), st.append(2, Xa, Ya).to_numpy())
        L, R = st.split(2, 3, J)
        assert L.layout == tb.SOA and np.array_equal(L.to_numpy().ravel(), x) and np.array_equal(R.to_numpy().ravel(), y)
        assert np.array_equal(st.ta(2, Zs).to_numpy(), st.ta(2, Za).to_numpy())
        assert np.array_equal(st.at(Zs, [2, 3]).to_numpy(), st.at(Za, [2, 3]).to_numpy())
        assert np.array_equal(st.slice_(Zs, [None, 2, None]).to_numpy(), st.slice_(Za, [None, 2, None]).to_numpy())
        assert np.array_equal(st.permute_axes(Zs, [2, 0, 1]).to_numpy(), st.permute_axes(Za, [2, 0, 1]).to_numpy())
        T = tb.transpose(Zs)
        assert T.layout == tb.SOA and np.array_equal(T.to_numpy(), oracle.transpose(z, dims).reshape(batch, 4, 3, 2))


def test_structural_errors(ctx):
    X = ctx.empty((2, 3), np.float64, 4)
    with pytest.raises(tb.WrongListLength):
        st.append(1, X, ctx.empty((2, 4), np.float64, 4))
    with pytest.raises(tb.WrongListLength):
        st.ta(3, X)
    with pytest.raises(tb.WrongListLength):
        st.slice_(X, [None])
