"""Host-array forms of the hot-path operations (the ``t5_*_host`` entries of include/t5b200.h).

This is the call a user of the reference makes: tensors live in ordinary host memory
(``Data.Tensor.Vector``'s ``content`` vector, src/Data/Tensor/Vector.hs:121-124; here a
C-contiguous numpy array ``[batch, *shape]`` in the reference's row-major wire order,
Vector.hs:360-367) and one call returns the result in host memory.  Inside the library
the batch is streamed through the GPU in chunks — upload, kernel, download overlapped on
three streams.  Pass arrays backed by :class:`tensor_b200.PinnedArray` (and ``out=``) for
full PCIe speed.  There is no CPU arithmetic here: every function is one C-ABI call.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .device import Context, DTYPE_OF, WrongListLength, check


def _vp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _in(a, what: str) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype not in DTYPE_OF:
        raise TypeError(f"{what}: element type {a.dtype} has no device kernels")
    return np.ascontiguousarray(a)


def _items(a: np.ndarray, shape: Tuple[int, ...], what: str) -> int:
    d = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if d == 0 or a.size % d:
        raise WrongListLength(f"{what}: {a.size} elements are not a whole number of {shape} tensors")
    return a.size // d


def _out(out: Optional[np.ndarray], shape, dtype, what: str) -> np.ndarray:
    if out is None:
        return np.empty(shape, dtype=dtype)
    if out.dtype != dtype or not out.flags.c_contiguous:
        raise ValueError(f"{what}: out= must be a C-contiguous {np.dtype(dtype)} array")
    if out.size != int(np.prod(shape, dtype=np.int64)):
        raise WrongListLength(f"{what}: out= holds {out.size} elements, need {int(np.prod(shape, dtype=np.int64))}")
    return out


def matmul(ctx: Context, a, b, m: int, k: int, n: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    """``(.*.)`` of ``batch`` m x k by k x n matrices (n == 1: matrix . vector, m == 1: vector . matrix)."""
    a, b = _in(a, "(.*.)"), _in(b, "(.*.)")
    if a.dtype != b.dtype:
        raise ValueError("element types differ")
    batch = _items(a, (m, k), "(.*.)")
    if _items(b, (k, n), "(.*.)") != batch:
        raise WrongListLength("(.*.): batch counts differ")
    out = _out(out, (batch, m, n), a.dtype, "(.*.)")
    check(_lib.lib().t5_matmul_host(ctx._h, DTYPE_OF[a.dtype], m, k, n, batch, _vp(a), _vp(b), _vp(out)), ctx)
    return out


def dot(ctx: Context, x, y, n: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    """``dot`` of ``batch`` pairs of n-vectors: one scalar per pair."""
    x, y = _in(x, "dot"), _in(y, "dot")
    if x.dtype != y.dtype:
        raise ValueError("element types differ")
    batch = _items(x, (n,), "dot")
    if _items(y, (n,), "dot") != batch:
        raise WrongListLength("dot: batch counts differ")
    out = _out(out, (batch,), x.dtype, "dot")
    check(_lib.lib().t5_dot_host(ctx._h, DTYPE_OF[x.dtype], n, batch, _vp(x), _vp(y), _vp(out)), ctx)
    return out


def prod(ctx: Context, a, dims_a, b, dims_b, ncontract: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    """``prod n`` / ``(⊗)`` (ncontract == 0) of ``batch`` tensors of shapes dims_a and dims_b."""
    a, b = _in(a, "prod"), _in(b, "prod")
    dims_a, dims_b = tuple(int(x) for x in dims_a), tuple(int(x) for x in dims_b)
    if a.dtype != b.dtype:
        raise ValueError("element types differ")
    if ncontract < 0 or ncontract > len(dims_a) or ncontract > len(dims_b) or dims_a[len(dims_a) - ncontract:] != dims_b[:ncontract]:
        raise WrongListLength(f"prod {ncontract} of shapes {dims_a} and {dims_b}")
    batch = _items(a, dims_a, "prod")
    if _items(b, dims_b, "prod") != batch:
        raise WrongListLength("prod: batch counts differ")
    out_dims = dims_a[: len(dims_a) - ncontract] + dims_b[ncontract:]
    out = _out(out, (batch, *out_dims), a.dtype, "prod")
    check(_lib.lib().t5_prod_host(ctx._h, DTYPE_OF[a.dtype], len(dims_a), _lib.u64_array(dims_a), len(dims_b),
                                  _lib.u64_array(dims_b), ncontract, batch, _vp(a), _vp(b), _vp(out)), ctx)
    return out


def transpose(ctx: Context, a, dims, out: Optional[np.ndarray] = None) -> np.ndarray:
    """``transpose`` (full index reversal, Indexable.hs:101-102) of ``batch`` tensors of shape dims."""
    a = _in(a, "transpose")
    dims = tuple(int(x) for x in dims)
    batch = _items(a, dims, "transpose")
    out = _out(out, (batch, *dims[::-1]), a.dtype, "transpose")
    check(_lib.lib().t5_transpose_host(ctx._h, DTYPE_OF[a.dtype], len(dims), _lib.u64_array(dims), batch, _vp(a), _vp(out)), ctx)
    return out


def det(ctx: Context, a, n: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    a = _in(a, "det")
    batch = _items(a, (n, n), "det")
    out = _out(out, (batch,), a.dtype, "det")
    check(_lib.lib().t5_det_host(ctx._h, DTYPE_OF[a.dtype], n, batch, _vp(a), _vp(out)), ctx)
    return out


def inverse(ctx: Context, a, n: int, out: Optional[np.ndarray] = None,
            status: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(inverses, status): status 0 = Just, 1 = Nothing (singular; the item is zero-filled)."""
    a = _in(a, "inverse")
    batch = _items(a, (n, n), "inverse")
    out = _out(out, (batch, n, n), a.dtype, "inverse")
    status = _out(status, (batch,), np.uint8, "inverse status")
    check(_lib.lib().t5_inverse_host(ctx._h, DTYPE_OF[a.dtype], n, batch, _vp(a), _vp(out), _vp(status)), ctx)
    return out, status


def solve(ctx: Context, a, b, n: int, nrhs: int = 1, out: Optional[np.ndarray] = None,
          status: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Unique-solution square systems a x = b; b holds ``batch`` n x nrhs right-hand sides."""
    a, b = _in(a, "solve"), _in(b, "solve")
    if a.dtype != b.dtype:
        raise ValueError("element types differ")
    batch = _items(a, (n, n), "solve")
    if _items(b, (n, nrhs), "solve") != batch:
        raise WrongListLength("solve: batch counts differ")
    out = _out(out, (batch, n, nrhs), a.dtype, "solve")
    status = _out(status, (batch,), np.uint8, "solve status")
    check(_lib.lib().t5_solve_host(ctx._h, DTYPE_OF[a.dtype], n, nrhs, batch, _vp(a), _vp(b), _vp(out), _vp(status)), ctx)
    return out, status
