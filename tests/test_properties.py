"""Property tests in the style of the reference's QuickCheck suites (tests/Tensor.hs:25-80,
tests/MultiIndex.hs), restated with hypothesis over RANDOM shapes:

* CPU: the library's own host-side index-map builder (t5_structural_table, a pure C function of
  libt5b200.so) against the GADT-backend restatement for append/split, at/ta, slice, rev/unRev -
  the same equalities the reference checks, but on shapes hypothesis chooses;
* GPU (marked): the same round trips and the linear-algebra metamorphic identities through the C ABI
  on random small shapes, every result compared with the oracle bit for bit."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
from oracle import gadt_model as G
from oracle import structural as S
from tensor_b200 import _lib

dims_st = st.lists(st.integers(1, 4), min_size=1, max_size=4)
CPU = settings(max_examples=60, deadline=None)
GPU = settings(max_examples=int(os.environ.get("T5_HYP_EXAMPLES", "40")), deadline=None,
               suppress_health_check=[HealthCheck.function_scoped_fixture])   # (T5_HYP_EXAMPLES=400 for a soak run)


def flat(dims, start=0):
    return list(range(start, start + S.product(dims)))


# ---- CPU: product host logic vs the GADT model --------------------------------------------------
@CPU
@given(dims_st, st.data())
def test_split_of_append_is_identity_on_library_tables(dims, data):
    """split sh (append sh x y) == (x, y)   (tests/Tensor.hs:25-29), any rank, any axis."""
    axis = data.draw(st.integers(1, len(dims)))
    other = data.draw(st.integers(1, 4))
    e = list(dims)
    e[axis - 1] = other
    x, y = flat(dims), flat(e, 10_000)
    table = _lib.structural_table(_lib.ST_APPEND, dims, [axis] + e)
    dA = len(x)
    joined = [x[t] if t < dA else y[t - dA] for t in table]
    want = G.append(axis, G.from_list(dims, x), G.from_list(e, y))
    assert G.to_list(want) == joined
    out_dims = list(dims)
    out_dims[axis - 1] += other
    t1 = _lib.structural_table(_lib.ST_SPLIT_FIRST, out_dims, [axis, dims[axis - 1]])
    t2 = _lib.structural_table(_lib.ST_SPLIT_SECOND, out_dims, [axis, dims[axis - 1]])
    assert [joined[i] for i in t1] == x and [joined[i] for i in t2] == y


@CPU
@given(dims_st, st.data())
def test_slice_at_ta_on_library_tables(dims, data):
    x = flat(dims)
    t = G.from_list(dims, x)
    sl = [data.draw(st.one_of(st.none(), st.integers(1, d))) for d in dims]
    table = _lib.structural_table(_lib.ST_SLICE, dims, [0 if s is None else s for s in sl])
    assert [x[j] for j in table] == G.to_list(G.slice_(t, sl))
    # slice x (allCons .. nilS) == x   (tests/Tensor.hs:36-40)
    assert _lib.structural_table(_lib.ST_SLICE, dims, [0] * len(dims)) == list(range(len(x)))
    i = data.draw(st.integers(1, dims[0]))
    assert [x[j] for j in _lib.structural_table(_lib.ST_TA, dims, [i])] == G.to_list(G.ta(i, t))
    if len(dims) > 1:
        ix = [data.draw(st.integers(1, d)) for d in dims[1:]]
        assert [x[j] for j in _lib.structural_table(_lib.ST_AT, dims, ix)] == G.to_list(G.at(t, ix))


@CPU
@given(st.lists(st.integers(1, 3), min_size=0, max_size=3), st.lists(st.integers(1, 3), min_size=0, max_size=3))
def test_rev_unrev_are_axis_permutations(outer, inner):
    """rev / unRev of nested tensors (tests/Tensor.hs:42-80) as permutations of the flat [outer][inner] layout,
    and unRev . rev == id."""
    dims = list(outer) + list(inner)
    if not dims:
        return
    r, n = len(outer), len(dims)
    rev_perm = list(reversed(range(r))) + list(range(r, n))
    assert _lib.structural_table(_lib.ST_PERMUTE, dims, rev_perm) == S.rev_table(outer, inner)[1]
    rev_dims = [dims[p] for p in rev_perm]
    back = _lib.structural_table(_lib.ST_PERMUTE, rev_dims, rev_perm)       # reversing the outer axes twice
    fwd = _lib.structural_table(_lib.ST_PERMUTE, dims, rev_perm)
    assert [fwd[j] for j in back] == list(range(S.product(dims)))
    # transpose == full reversal == transposeV (Vector.hs:306-307), and it is an involution
    tr = _lib.structural_table(_lib.ST_PERMUTE, dims, list(reversed(range(n))))
    assert tr == [S.transposeV(dims, i) for i in range(S.product(dims))]
    tr_back = _lib.structural_table(_lib.ST_PERMUTE, list(reversed(dims)), list(reversed(range(n))))
    assert [tr[j] for j in tr_back] == list(range(S.product(dims)))


@CPU
@given(st.integers(1, 6), st.integers(1, 6), st.integers(1, 6), st.integers(0, 2 ** 32 - 1))
def test_oracle_product_identities(m, k, n, seed):
    """(A.B)^T == B^T.A^T exactly for Int64 (wrapping) AND for Double (same products, same order); A.I == A."""
    rng = np.random.default_rng(seed)
    for dc, mk in ((oracle.I64, lambda s: rng.integers(-2 ** 62, 2 ** 62, s, dtype=np.int64)),
                   (oracle.F64, lambda s: rng.standard_normal(s))):
        a, b = mk(3 * m * k), mk(3 * k * n)
        ab_t = oracle.transpose(oracle.matmul(a, b, m, k, n), [m, n])
        bt_at = oracle.matmul(oracle.transpose(b, [k, n]), oracle.transpose(a, [m, k]), n, k, m)
        assert np.array_equal(ab_t.ravel().view(np.uint64), bt_at.ravel().view(np.uint64))
        eye = np.broadcast_to(np.eye(k, dtype=a.dtype), (3, k, k)).copy()
        assert np.array_equal(oracle.matmul(a, eye, m, k, k).ravel().view(np.uint64), np.asarray(a).view(np.uint64))


# ---- GPU: the same properties through the C ABI ---------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    import torch
    import tensor_b200 as tb
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    c = tb.Context([0])
    yield c
    c.close()


@pytest.mark.gpu
@GPU
@given(st.integers(1, 24), st.integers(1, 24), st.integers(1, 24), st.sampled_from([oracle.F64, oracle.F32, oracle.I64]),
       st.integers(1, 700), st.integers(0, 2 ** 16))
def test_gpu_matmul_random_shapes_bit_exact(ctx, m, k, n, dc, batch, op_id):
    import tensor_b200 as tb
    a = oracle.fill(dc, oracle.SEED_A, op_id, batch * m * k)
    b = oracle.fill(dc, oracle.SEED_B, op_id, batch * k * n)
    A, B = ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))
    from _parity import assert_matmul_matches, matmul_is_exact
    C = tb.matmul(A, B).to_numpy()
    assert_matmul_matches(C, a, b, m, k, n, dc, f"{m}x{k}x{n}")       # bit-exact unless a tensor-core shape (17..24 x 8k x 24 Double)
    if matmul_is_exact(dc, m, k, n) and matmul_is_exact(dc, n, k, m):
        # (A.B)^T == B^T.A^T, computed entirely on the device, bit for bit
        lhs = tb.transpose(tb.matmul(A, B)).to_numpy()
        rhs = tb.matmul(tb.transpose(B), tb.transpose(A)).to_numpy()
        assert np.array_equal(lhs.view(np.uint8), rhs.view(np.uint8))


@pytest.mark.gpu
@GPU
@given(dims_st, st.data())
def test_gpu_split_of_append_and_transpose_involution(ctx, dims, data):
    import tensor_b200 as tb
    from tensor_b200 import structure as T
    axis = data.draw(st.integers(1, len(dims)))
    e = list(dims)
    e[axis - 1] = data.draw(st.integers(1, 4))
    batch = data.draw(st.integers(1, 300))
    x = oracle.fill(oracle.I64, oracle.SEED_A, 31, batch * S.product(dims))
    y = oracle.fill(oracle.I64, oracle.SEED_B, 31, batch * S.product(e))
    X, Y = ctx.from_numpy(x, tuple(dims)), ctx.from_numpy(y, tuple(e))
    Z = T.append(axis, X, Y)
    X2, Y2 = T.split(axis, dims[axis - 1], Z)
    assert np.array_equal(X2.to_numpy().ravel(), x) and np.array_equal(Y2.to_numpy().ravel(), y)
    assert np.array_equal(tb.transpose(tb.transpose(Z)).to_numpy(), Z.to_numpy())
    assert np.array_equal(tb.transpose(X).to_numpy().reshape(batch, -1),
                          oracle.transpose(x, list(dims)).reshape(batch, -1))


@pytest.mark.gpu
@GPU
@given(st.integers(1, 8), st.integers(1, 16), st.data(), st.sampled_from([oracle.F64, oracle.F32]), st.integers(1, 500),
       st.integers(0, 2 ** 16))
def test_gpu_row_echelon_random_shapes_bit_exact(ctx, m, n, data, dc, batch, op_id):
    """rowEchelonForm on random shapes / pivot-column counts with small-integer entries (so that exact rank
    defects occur naturally): bit-identical to the oracle, idempotent, rank == rank of the leading block."""
    import tensor_b200 as tb
    pc = data.draw(st.integers(0, n))
    raw = oracle.fill(oracle.I64, oracle.SEED_A, op_id, batch * m * n)
    a = ((raw >> 7) % 5 - 2).astype(oracle.NP_DTYPE[dc]).reshape(batch, m, n)
    got, rank = tb.row_echelon_form(ctx.from_numpy(a, (m, n)), pc)
    want, wrank = oracle.row_echelon(a, m, n, pc)
    g = got.to_numpy()
    assert np.array_equal(rank.to_numpy(), wrank)
    assert np.array_equal(g.view(np.uint8), want.view(np.uint8))
    again, rank2 = tb.row_echelon_form(got, pc)
    assert np.array_equal(again.to_numpy(), g) and np.array_equal(rank2.to_numpy(), wrank)
    assert np.all(wrank <= min(m, pc))


@pytest.mark.gpu
@GPU
@given(st.integers(2, 8), st.sampled_from([oracle.F64, oracle.F32]), st.integers(1, 300), st.integers(0, 2 ** 16))
def test_gpu_inverse_solve_echelon_agree_on_rank(ctx, n, dc, batch, op_id):
    """inverse status, det == 0 and rowEchelonForm rank < n single out the same items (they share the pivot rule and
    the multipliers below the diagonal), and the three agree with the oracle bit for bit."""
    import tensor_b200 as tb
    raw = oracle.fill(oracle.I64, oracle.SEED_B, op_id, batch * n * n)
    a = ((raw >> 9) % 3 - 1).astype(oracle.NP_DTYPE[dc]).reshape(batch, n, n)     # entries in {-1, 0, 1}: many singular items
    A = ctx.from_numpy(a, (n, n))
    inv, st_ = tb.inverse(A)
    d = tb.det(A).to_numpy()
    _, rank = tb.row_echelon_form(A)
    singular = st_.to_numpy() == 1
    assert np.array_equal(singular, rank.to_numpy() < n)
    assert np.array_equal(singular, d == 0)
    oinv, ost = oracle.inverse(a, n)
    assert np.array_equal(st_.to_numpy(), ost)
    assert np.array_equal(inv.to_numpy().view(np.uint8), oinv.view(np.uint8))
    assert np.array_equal(d.view(np.uint8), oracle.det(a, n).view(np.uint8))


@pytest.mark.gpu
@GPU
@given(st.integers(17, 140), st.integers(4, 48), st.integers(12, 80), st.integers(1, 40), st.integers(0, 2 ** 16))
def test_gpu_double_tensor_core_random_shapes_within_tolerance(ctx, m, k2, n2, batch, op_id):
    """Double products of random shapes in the FP64 tensor-core class (k, n even): every tile geometry - four
    32 x 32 or two 48 x 48 / 64 x 64 items per tile, the 64 / 96 / 128 tiles - with ragged rows, columns, K tails and
    part-filled item tiles; whatever the library declares exact must be exact, the rest within 1e-12 sum|a||b|."""
    import tensor_b200 as tb
    from _parity import assert_matmul_matches
    k, n = 2 * k2, 2 * n2
    a = oracle.fill(oracle.F64, oracle.SEED_A, op_id, batch * m * k)
    b = oracle.fill(oracle.F64, oracle.SEED_B, op_id, batch * k * n)
    C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
    assert_matmul_matches(C, a, b, m, k, n, oracle.F64, f"{m}x{k}x{n} batch {batch}")


@pytest.mark.gpu
@GPU
@given(st.integers(17, 70), st.integers(2, 24), st.integers(5, 18), st.integers(1, 300), st.integers(0, 2 ** 16))
def test_gpu_float_packed_items_within_tolerance(ctx, m, k4, n4, batch, op_id):
    """Float items up to 70 x 72 with column counts that are multiples of 4: four (<= 32 x 32) or two (<= 64 x 64) items
    to a tcgen05 tile, one item per tile beyond; random row counts, depths and batch sizes (part-filled last tiles,
    part-stored column blocks).  Whatever the library declares exact must be exact, the rest within 1e-5 sum|a||b|."""
    import tensor_b200 as tb
    from _parity import assert_matmul_matches
    k, n = 4 * k4, 4 * n4
    a = oracle.fill(oracle.F32, oracle.SEED_A, op_id, batch * m * k)
    b = oracle.fill(oracle.F32, oracle.SEED_B, op_id, batch * k * n)
    C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, n))).to_numpy()
    assert_matmul_matches(C, a, b, m, k, n, oracle.F32, f"{m}x{k}x{n} batch {batch}")


@pytest.mark.gpu
@GPU
@given(st.integers(5, 48), st.sampled_from(["square", "matvec", "vecmat"]), st.sampled_from([oracle.F64, oracle.F32, oracle.I64]),
       st.integers(1, 400), st.integers(0, 2 ** 16))
def test_gpu_square_and_matvec_sizes_bit_exact(ctx, n, kind, dc, batch, op_id):
    """Every square size 5..48 and its matrix.vector / vector.matrix forms - the sizes served by the row-per-thread
    kernel (aligned and padded-row variants, packed results), the staged kernel and the blocked product."""
    import tensor_b200 as tb
    m, k, q = {"square": (n, n, n), "matvec": (n, n, 1), "vecmat": (1, n, n)}[kind]
    a = oracle.fill(dc, oracle.SEED_A, op_id, batch * m * k)
    b = oracle.fill(dc, oracle.SEED_B, op_id, batch * k * q)
    from _parity import assert_matmul_matches
    C = tb.matmul(ctx.from_numpy(a, (m, k)), ctx.from_numpy(b, (k, q))).to_numpy()
    assert_matmul_matches(C, a, b, m, k, q, dc, f"{m}x{k}x{q}")      # Double squares 24 / 32 / 40 / 48: FP64 tensor pipe, 1e-12
