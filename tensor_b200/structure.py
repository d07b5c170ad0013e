"""Structural operations on device batches — the Python mirror of the IsTensor /
MultiIndexable / Sliceable methods of the flat-vector backend
(src/Data/Tensor/Vector.hs:260-354, :400-463).  Pure index maps: one gather kernel each,
bit-exact for every element type.  Indices and axes are 1-based like the reference's
MultiIndex.

Nested tensors (`t is (t js e)`, the argument of concat / unConcat / rev / unRev) are a
`Nested` view: a flat batch of shape `is ++ js` plus the rank of the outer part.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

from . import _lib
from .device import Batch, WrongListLength, check


def append(n: int, x: Batch, y: Batch) -> Batch:
    """``append n x y`` (Vector.hs:322-335): concatenate along axis n (1-based)."""
    if x.ctx is not y.ctx or x.dtype != y.dtype or x.batch != y.batch:
        raise ValueError("operands differ in context, element type or batch count")
    r = len(x.shape)
    if r != len(y.shape) or not (1 <= n <= r) or any(a != b for k, (a, b) in enumerate(zip(x.shape, y.shape)) if k != n - 1):
        raise WrongListLength(f"append {n} of shapes {x.shape} and {y.shape}")
    shape = tuple(a + b if k == n - 1 else a for k, (a, b) in enumerate(zip(x.shape, y.shape)))
    out = Batch(x.ctx, shape, x.dtype, x.batch, x.layout)
    check(_lib.lib().t5_append(x.ctx._h, x.dtype, n, r, _lib.u64_array(x.shape), _lib.u64_array(y.shape),
                               x._h, y._h, out._h), x.ctx)
    return out


def split(n: int, first_extent: int, t: Batch) -> Tuple[Batch, Batch]:
    """``split n t`` (Vector.hs:336-354): inverse of append, cutting axis n after `first_extent`."""
    r = len(t.shape)
    if not (1 <= n <= r) or not (0 < first_extent < t.shape[n - 1]):
        raise WrongListLength(f"split {n} at {first_extent} of shape {t.shape}")
    s1 = tuple(first_extent if k == n - 1 else a for k, a in enumerate(t.shape))
    s2 = tuple(a - first_extent if k == n - 1 else a for k, a in enumerate(t.shape))
    o1, o2 = Batch(t.ctx, s1, t.dtype, t.batch, t.layout), Batch(t.ctx, s2, t.dtype, t.batch, t.layout)
    check(_lib.lib().t5_split(t.ctx._h, t.dtype, n, r, _lib.u64_array(t.shape), first_extent, t._h, o1._h, o2._h), t.ctx)
    return o1, o2


def at(t: Batch, ix: Sequence[int]) -> Batch:
    """``t `at` ix`` (Vector.hs:277-283): fix all but the first index; the column extractor."""
    if len(t.shape) < 1 or len(ix) != len(t.shape) - 1:
        raise WrongListLength(f"at {list(ix)} of shape {t.shape}")
    out = Batch(t.ctx, (t.shape[0],), t.dtype, t.batch, t.layout)
    check(_lib.lib().t5_at(t.ctx._h, t.dtype, len(t.shape), _lib.u64_array(t.shape), _lib.u64_array(ix), t._h, out._h), t.ctx)
    return out


def ta(i: int, t: Batch) -> Batch:
    """``[i] `ta` t``: fix the first index; the row extractor (GADT semantics, Tensor.hs:245-247)."""
    if len(t.shape) < 1:
        raise WrongListLength("ta of a rank-0 tensor")
    out = Batch(t.ctx, tuple(t.shape[1:]), t.dtype, t.batch, t.layout)
    check(_lib.lib().t5_ta(t.ctx._h, t.dtype, len(t.shape), _lib.u64_array(t.shape), i, t._h, out._h), t.ctx)
    return out


def slice_(t: Batch, sl: Sequence[Optional[int]]) -> Batch:
    """``slice t sl`` (Vector.hs:400-429): ``None`` keeps an axis (Nothing), ``k`` fixes it (Just k)."""
    if len(sl) != len(t.shape):
        raise WrongListLength(f"slicer {list(sl)} for shape {t.shape}")
    shape = tuple(d for d, s in zip(t.shape, sl) if s is None)
    out = Batch(t.ctx, shape, t.dtype, t.batch, t.layout)
    check(_lib.lib().t5_slice(t.ctx._h, t.dtype, len(t.shape), _lib.u64_array(t.shape),
                              _lib.u64_array([0 if s is None else s for s in sl]), t._h, out._h), t.ctx)
    return out


def permute_axes(t: Batch, perm: Sequence[int]) -> Batch:
    """Output axis a is input axis perm[a] (0-based)."""
    shape = tuple(t.shape[p] for p in perm)
    out = Batch(t.ctx, shape, t.dtype, t.batch, t.layout)
    arr = (ctypes.c_int * max(1, len(perm)))(*perm)
    check(_lib.lib().t5_permute_axes(t.ctx._h, t.dtype, len(t.shape), _lib.u64_array(t.shape), arr, t._h, out._h), t.ctx)
    return out


@dataclass
class Nested:
    """``t is (t js e)``: `flat` has shape is ++ js, `outer_rank` = length of `is`."""
    flat: Batch
    outer_rank: int

    @property
    def outer(self):
        return tuple(self.flat.shape[: self.outer_rank])

    @property
    def inner(self):
        return tuple(self.flat.shape[self.outer_rank:])


def concat(t: Nested) -> Nested:
    """concat :: t (i:is) (t js e) -> t is (t (i:js) e)  (Vector.hs:260-267)."""
    r = t.outer_rank
    if r < 1:
        raise WrongListLength("concat needs an outer tensor of rank >= 1")
    perm = list(range(1, r)) + [0] + list(range(r, len(t.flat.shape)))
    return Nested(permute_axes(t.flat, perm), r - 1)


def unconcat(t: Nested) -> Nested:
    """unConcat :: t is (t (i:js) e) -> t (i:is) (t js e)  (Vector.hs:268-276)."""
    r = t.outer_rank
    if len(t.flat.shape) - r < 1:
        raise WrongListLength("unConcat needs inner tensors of rank >= 1")
    perm = [r] + list(range(0, r)) + list(range(r + 1, len(t.flat.shape)))
    return Nested(permute_axes(t.flat, perm), r + 1)


def rev(t: Nested) -> Batch:
    """rev :: t is (t js e) -> t (reverse is ++ js) e  (Vector.hs:289-294)."""
    r = t.outer_rank
    perm = list(reversed(range(r))) + list(range(r, len(t.flat.shape)))
    return permute_axes(t.flat, perm)


def unrev(t: Nested) -> Batch:
    """unRev :: t js (t is e) -> t (reverse is ++ js) e  (Vector.hs:295-301); transpose = unRev . t0."""
    r, n = t.outer_rank, len(t.flat.shape)
    perm = list(reversed(range(r, n))) + list(range(r))
    return permute_axes(t.flat, perm)
