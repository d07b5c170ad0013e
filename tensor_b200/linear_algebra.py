"""Python mirror of the re-introduced ``Data.Tensor.LinearAlgebra`` classes
(the reference removed them in 0.2.0 — CHANGELOG.md:24-29; SURVEY.md §8a
specifies them): ``MatrixProduct`` ``(.*.)``, ``DotProduct`` ``dot``,
``TensorProduct`` ``(⊗)`` / ``prod n``, ``Transpose`` (src/Data/Indexable.hs:101-102),
``SquareMatrix`` ``det`` / ``inverse`` and ``LinearSystem`` ``solve``.

Every function takes and returns device ``Batch`` values and makes exactly one
C-ABI call; shapes are checked the way the Haskell type-checker would check the
class instances, and a mismatch raises ``WrongListLength``.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np

from . import _lib
from ._lib import AOS, U8
from .device import Batch, Context, NP_OF, WrongListLength, check


def _same_ctx(*bs: Batch) -> Context:
    ctx = bs[0].ctx
    for b in bs[1:]:
        if b.ctx is not ctx:
            raise ValueError("operands live in different contexts")
        if b.batch != bs[0].batch:
            raise WrongListLength(f"batch counts differ: {b.batch} vs {bs[0].batch}")
        if b.dtype != bs[0].dtype:
            raise ValueError("element types differ")
        if b.layout != bs[0].layout:
            raise ValueError("layouts differ")
    return ctx


def matmul(a: Batch, b: Batch) -> Batch:
    """``(.*.)``: contract the last index of ``a`` with the first of ``b`` (§8a a6).

    matrix.matrix, matrix.vector and vector.matrix are the same operation.
    """
    ctx = _same_ctx(a, b)
    if len(a.shape) < 1 or len(b.shape) < 1 or a.shape[-1] != b.shape[0]:
        raise WrongListLength(f"(.*.) of shapes {a.shape} and {b.shape}")
    if len(a.shape) > 2 or len(b.shape) > 2:
        return prod(a, b, 1)
    m = a.shape[0] if len(a.shape) == 2 else 1
    k = a.shape[-1]
    n = b.shape[1] if len(b.shape) == 2 else 1
    out_shape = a.shape[:-1] + b.shape[1:]
    out = Batch(ctx, out_shape, a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_matmul(ctx._h, a.dtype, m, k, n, a._h, b._h, out._h), ctx)
    return out


def matmul_is_exact(dtype: int, m: int, k: int, n: int) -> bool:
    """True when ``(m x k) .*. (k x n)`` of this element type is bit-identical to the sequential fold, False when the
    library runs it on the tensor cores (Double: |delta| <= 1e-12 sum|a||b|, Float: 1e-5; include/t5b200.h t5_matmul)."""
    r = _lib.lib().t5_matmul_is_exact(dtype, m, k, n)
    if r < 0:
        check(-r, None)
    return bool(r)


def dot(x: Batch, y: Batch) -> Batch:
    """``dot``: one scalar per pair (§8a a7)."""
    ctx = _same_ctx(x, y)
    if len(x.shape) != 1 or x.shape != y.shape:
        raise WrongListLength(f"dot of shapes {x.shape} and {y.shape}")
    out = Batch(ctx, (), x.dtype, x.batch, x.layout)
    check(_lib.lib().t5_dot(ctx._h, x.dtype, x.shape[0], x._h, y._h, out._h), ctx)
    return out


def dot_sum(x: Batch, y: Batch):
    """sum over the batch of ``dot`` — the only cross-device reduction (§8e)."""
    ctx = _same_ctx(x, y)
    if len(x.shape) != 1 or x.shape != y.shape:
        raise WrongListLength(f"dot of shapes {x.shape} and {y.shape}")
    out = np.zeros((1,), dtype=NP_OF[x.dtype])
    check(_lib.lib().t5_dot_sum(ctx._h, x.dtype, x.shape[0], x._h, y._h, out.ctypes.data_as(ctypes.c_void_p)), ctx)
    return out[0]


def prod(a: Batch, b: Batch, n: int) -> Batch:
    """``prod n``: contract the last ``n`` indices of ``a`` with the first ``n`` of ``b`` (§8a a8)."""
    ctx = _same_ctx(a, b)
    if n < 0 or n > len(a.shape) or n > len(b.shape) or tuple(a.shape[len(a.shape) - n:]) != tuple(b.shape[:n]):
        raise WrongListLength(f"prod {n} of shapes {a.shape} and {b.shape}")
    out_shape = a.shape[: len(a.shape) - n] + b.shape[n:]
    out = Batch(ctx, out_shape, a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_prod(ctx._h, a.dtype, len(a.shape), _lib.u64_array(a.shape), len(b.shape),
                             _lib.u64_array(b.shape), n, a._h, b._h, out._h), ctx)
    return out


def tensor_product(a: Batch, b: Batch) -> Batch:
    """``(⊗)`` = ``prod 0``: rank-r ⊗ rank-s -> rank-(r+s), one multiply per component."""
    return prod(a, b, 0)


def transpose(a: Batch) -> Batch:
    """``transpose = unRev . t0`` (Indexable.hs:101-102): full index reversal, shape reversed."""
    ctx = a.ctx
    out = Batch(ctx, tuple(reversed(a.shape)), a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_transpose(ctx._h, a.dtype, len(a.shape), _lib.u64_array(a.shape), a._h, out._h), ctx)
    return out


def _square(a: Batch, what: str) -> int:
    if len(a.shape) != 2 or a.shape[0] != a.shape[1]:
        raise WrongListLength(f"{what} needs a square matrix, got shape {a.shape}")
    return a.shape[0]


def det(a: Batch) -> Batch:
    """``det`` by Gaussian elimination, first-non-zero pivot; singular -> exact 0 (§8a a9)."""
    n = _square(a, "det")
    out = Batch(a.ctx, (), a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_det(a.ctx._h, a.dtype, n, a._h, out._h), a.ctx)
    return out


def inverse(a: Batch) -> Tuple[Batch, Batch]:
    """``inverse :: … -> Maybe``: (inverses, status) with status 0 = Just, 1 = Nothing (§8a a10)."""
    n = _square(a, "inverse")
    out = Batch(a.ctx, a.shape, a.dtype, a.batch, a.layout)
    status = Batch(a.ctx, (), U8, a.batch, a.layout)
    check(_lib.lib().t5_inverse(a.ctx._h, a.dtype, n, a._h, out._h, status._h), a.ctx)
    return out, status


def solve(a: Batch, b: Batch) -> Tuple[Batch, Batch]:
    """``solve``: unique-solution square systems ``a x = b``; ``b`` is an n-vector or n x m (§8a a11)."""
    ctx = _same_ctx(a, b)
    n = _square(a, "solve")
    if len(b.shape) not in (1, 2) or b.shape[0] != n:
        raise WrongListLength(f"solve of shapes {a.shape} and {b.shape}")
    nrhs = b.shape[1] if len(b.shape) == 2 else 1
    out = Batch(ctx, b.shape, a.dtype, a.batch, a.layout)
    status = Batch(ctx, (), U8, a.batch, a.layout)
    check(_lib.lib().t5_solve(ctx._h, a.dtype, n, nrhs, a._h, b._h, out._h, status._h), ctx)
    return out, status


def tr(a: Batch) -> Batch:
    """``tr``: the trace, ``fold (+) 0`` of the diagonal in index order."""
    n = _square(a, "tr")
    out = Batch(a.ctx, (), a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_trace(a.ctx._h, a.dtype, n, a._h, out._h), a.ctx)
    return out


def poly_eval(a: Batch, coeffs) -> Batch:
    """``polyEval a [c0, c1, ..]`` = c0 I + c1 a + c2 a^2 + .. by Horner's rule (CHANGELOG.md:28-29
    names the method; semantics SURVEY §8f-4).  One coefficient list for the whole batch."""
    n = _square(a, "polyEval")
    c = np.ascontiguousarray(coeffs, dtype=NP_OF[a.dtype]).reshape(-1)
    out = Batch(a.ctx, a.shape, a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_poly_eval(a.ctx._h, a.dtype, n, c.ctypes.data_as(ctypes.c_void_p), int(c.size), a._h, out._h), a.ctx)
    return out


def char_poly(a: Batch) -> Batch:
    """``charPoly``: the n+1 coefficients of det(xI - a), low order first, leading 1 (Faddeev-LeVerrier)."""
    n = _square(a, "charPoly")
    out = Batch(a.ctx, (n + 1,), a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_char_poly(a.ctx._h, a.dtype, n, a._h, out._h), a.ctx)
    return out


def min_poly(a: Batch) -> Tuple[Batch, Batch]:
    """``minPoly``: (coefficients low order first, zero above the degree; degree per item)."""
    n = _square(a, "minPoly")
    out = Batch(a.ctx, (n + 1,), a.dtype, a.batch, a.layout)
    deg = Batch(a.ctx, (), U8, a.batch, a.layout)
    check(_lib.lib().t5_min_poly(a.ctx._h, a.dtype, n, a._h, out._h, deg._h), a.ctx)
    return out, deg


def row_echelon_form(a: Batch, pivot_cols: int | None = None) -> Tuple[Batch, Batch]:
    """``rowEchelonForm``: (reduced row echelon form of every rows x cols item, rank per item).  Pivots are
    looked for in the first ``pivot_cols`` columns only (default: all of them)."""
    if len(a.shape) != 2:
        raise WrongListLength(f"rowEchelonForm needs a matrix, got shape {a.shape}")
    rows, cols = a.shape
    pc = cols if pivot_cols is None else int(pivot_cols)
    out = Batch(a.ctx, a.shape, a.dtype, a.batch, a.layout)
    rank = Batch(a.ctx, (), U8, a.batch, a.layout)
    check(_lib.lib().t5_row_echelon(a.ctx._h, a.dtype, rows, cols, pc, a._h, out._h, rank._h), a.ctx)
    return out, rank


def triangular_solve(a: Batch, b: Batch) -> Tuple[Batch, Batch, Batch]:
    """``triangularSolve a b``: row operations on the augmented [a | b] until a is in reduced row echelon form;
    returns (reduced a, transformed b, rank of a).  b is a matrix with a's row count, or a vector of that length."""
    from . import structure
    if len(a.shape) != 2 or len(b.shape) not in (1, 2) or b.shape[0] != a.shape[0]:
        raise WrongListLength(f"triangularSolve: shapes {a.shape} and {b.shape} do not form an augmented matrix")
    vec = len(b.shape) == 1
    saved = b.shape
    try:
        if vec:
            b.shape = (saved[0], 1)      # an n-vector and an n x 1 matrix are the same record (Vector.hs:139-144)
        aug = structure.append(2, a, b)
    finally:
        b.shape = saved
    ech, rank = row_echelon_form(aug, a.shape[1])
    ra, rb = structure.split(2, a.shape[1], ech)
    aug.free()
    ech.free()
    if vec:
        rb.shape = saved
    return ra, rb, rank


_ZIP = {"+": 0, "-": 1, "*": 2}


def zip_with(op: str, a: Batch, b: Batch) -> Batch:
    """``liftA2 (+) / (-) / (*)`` (Vector.hs:187-190, :251)."""
    ctx = _same_ctx(a, b)
    if a.shape != b.shape:
        raise WrongListLength(f"liftA2 of shapes {a.shape} and {b.shape}")
    out = Batch(ctx, a.shape, a.dtype, a.batch, a.layout)
    check(_lib.lib().t5_zip(ctx._h, a.dtype, _ZIP[op], a._h, b._h, out._h), ctx)
    return out


def scale(k, a: Batch) -> Batch:
    """``fmap (k *)`` (Vector.hs:180-181, :250)."""
    out = Batch(a.ctx, a.shape, a.dtype, a.batch, a.layout)
    kk = np.asarray([k], dtype=NP_OF[a.dtype])
    check(_lib.lib().t5_scale(a.ctx._h, a.dtype, kk.ctypes.data_as(ctypes.c_void_p), a._h, out._h), a.ctx)
    return out


def replicate(ctx: Context, record, shape, batch: int, dtype=None, layout: int = AOS) -> Batch:
    """Every item a copy of ``record`` (one tensor of ``shape``, row-major): the record form of ``pure``."""
    rec = np.ascontiguousarray(record, dtype=dtype)
    shape = tuple(int(x) for x in shape)
    out = ctx.empty(shape, rec.dtype, batch, layout)
    if rec.size != out.elems:
        out.free()
        raise WrongListLength(f"replicate: record of {rec.size} elements for shape {shape}")
    check(_lib.lib().t5_replicate(ctx._h, out.dtype, rec.ctypes.data_as(ctypes.c_void_p), out._h), ctx)
    return out


def pure(ctx: Context, value, shape, dtype, batch: int, layout: int = AOS) -> Batch:
    """``pure a`` (Vector.hs:185-188): all components equal to ``value``."""
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    return replicate(ctx, np.full((n,), value, dtype=dtype), shape, batch, dtype, layout)


def unit(ctx: Context, n: int, dtype, batch: int, layout: int = AOS) -> Batch:
    """``unit``: a batch of n x n identity matrices (the old SquareMatrix method)."""
    return replicate(ctx, np.eye(n, dtype=dtype), (n, n), batch, dtype, layout)


def matmul_host(ctx: Context, a: np.ndarray, b: np.ndarray, m: int, k: int, n: int, out: np.ndarray) -> np.ndarray:
    """``(.*.)`` on HOST arrays (the end-to-end entry): chunked upload/compute/download pipeline."""
    if a.dtype != b.dtype or a.dtype != out.dtype:
        raise ValueError("element types differ")
    batch = a.size // (m * k)
    if a.size != batch * m * k or b.size != batch * k * n or out.size != batch * m * n:
        raise WrongListLength("host operand sizes do not match (m, k, n)")
    from .device import DTYPE_OF
    check(_lib.lib().t5_matmul_host(ctx._h, DTYPE_OF[a.dtype], m, k, n, batch,
                                    a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                                    out.ctypes.data_as(ctypes.c_void_p)), ctx)
    return out
