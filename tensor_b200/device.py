"""Device-resident batches — the Python mirror of the ``Data.Tensor.Device`` module
that INTEGRATION.md adds beside ``Data.Tensor.Vector`` (reference storage type:
src/Data/Tensor/Vector.hs:121-124 ``Tensor {form, content}``).

``Batch`` carries the fixed per-item shape (the type-level ``is :: [PI]`` of the
reference, MultiIndex.hs:46-61, kept here as a tuple), the element type and a
run-time batch count.  Marshalling follows the reference's
``IsList`` instance (Vector.hs:360-367): row-major, last index fastest, and a
wrong length raises ``WrongListLength`` (Tensor.hs:96-100).
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import AOS, F32, F64, I64, SOA, U8

NP_OF = {F64: np.dtype(np.float64), F32: np.dtype(np.float32), I64: np.dtype(np.int64), U8: np.dtype(np.uint8)}
DTYPE_OF = {v: k for k, v in NP_OF.items()}


class TensorException(Exception):
    """src/Data/Tensor.hs:96-100."""


class WrongListLength(TensorException):
    """Raised where the reference throws ``WrongListLength`` (Vector.hs:364-366)."""


class Unsupported(TensorException):
    """T5_ERR_UNSUPPORTED: no device kernel for the request (there is no CPU fallback)."""


class DeviceError(RuntimeError):
    pass


def check(rc: int, ctx: Optional["Context"] = None):
    if rc == _lib.OK:
        return
    msg = _lib.lib().t5_strerror(rc).decode()
    if ctx is not None and ctx._h:
        extra = _lib.lib().t5_last_error(ctx._h).decode()
        if extra:
            msg += f" [{extra}]"
    if rc == _lib.ERR_SHAPE:
        raise WrongListLength(msg)
    if rc == _lib.ERR_UNSUPPORTED:
        raise Unsupported(msg)
    if rc == _lib.ERR_INVALID:
        raise ValueError(msg)
    raise DeviceError(msg)


def _prod(xs: Iterable[int]) -> int:
    p = 1
    for x in xs:
        p *= int(x)
    return p


class Context:
    """``t5_ctx``: a set of devices, one stream each, batch axis sharded across them."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        L = _lib.lib()
        h = ctypes.c_void_p()
        if devices is None:
            rc = L.t5_init(None, 0, ctypes.byref(h))
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = L.t5_init(arr, len(devices), ctypes.byref(h))
        self._h = h if rc == _lib.OK else None
        check(rc)
        self._bufs = weakref.WeakSet()

    @property
    def ndev(self) -> int:
        return _lib.lib().t5_ndev(self._h)

    def sync(self):
        check(_lib.lib().t5_sync(self._h), self)

    def launch_count(self) -> int:
        return int(_lib.lib().t5_launch_count(self._h))

    def timer_begin(self):
        check(_lib.lib().t5_timer_begin(self._h), self)

    def timer_end(self) -> float:
        ms = ctypes.c_float()
        check(_lib.lib().t5_timer_end(self._h, ctypes.byref(ms)), self)
        return float(ms.value)

    def timer_per_device(self):
        """After timer_end(): milliseconds each device's stream took (one entry per device)."""
        n = self.ndev
        ms = (ctypes.c_float * n)()
        check(_lib.lib().t5_timer_per_device(self._h, ms, n), self)
        return [float(x) for x in ms]

    def live_buffers(self) -> int:
        return int(_lib.lib().t5_live_buffers(self._h))

    def flush_l2(self):
        check(_lib.lib().t5_flush_l2(self._h), self)

    def close(self):
        if self._h:
            for b in list(self._bufs):
                b.free()
            _lib.lib().t5_shutdown(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- construction ---------------------------------------------------------
    def empty(self, shape: Sequence[int], dtype, batch: int, layout: int = AOS) -> "Batch":
        return Batch(self, tuple(int(s) for s in shape), _code(dtype), int(batch), layout)

    def from_numpy(self, arr: np.ndarray, shape: Sequence[int], layout: int = AOS) -> "Batch":
        """arr is [batch, *shape] (or flat); element order = the reference's toList order."""
        shape = tuple(int(s) for s in shape)
        d = _prod(shape)
        arr = np.ascontiguousarray(arr)
        if arr.size % d != 0:
            raise WrongListLength(f"{arr.size} elements is not a multiple of the item size {d}")
        b = self.empty(shape, arr.dtype, arr.size // d, layout)
        b.upload(arr)
        return b

    def from_list(self, xs: Sequence, shape: Sequence[int], dtype, batch: int, layout: int = AOS) -> "Batch":
        """``fromList`` for a whole batch (Vector.hs:362-366): length must be batch * prod(shape)."""
        if len(xs) != batch * _prod(shape):
            raise WrongListLength(f"list of {len(xs)} elements for batch {batch} of shape {tuple(shape)}")
        return self.from_numpy(np.asarray(xs, dtype=NP_OF[_code(dtype)]), shape, layout)

    def synthetic(self, shape: Sequence[int], dtype, batch: int, seed: int, op_id: int, layout: int = AOS) -> "Batch":
        b = self.empty(shape, dtype, batch, layout)
        check(_lib.lib().t5_fill_synthetic(b._h, seed, op_id), self)
        return b


def _code(dtype) -> int:
    if isinstance(dtype, int) and dtype in NP_OF:
        return dtype
    return DTYPE_OF[np.dtype(dtype)]


class Batch:
    """``batch`` tensors of fixed ``shape`` and element type, resident on the context's devices."""

    def __init__(self, ctx: Context, shape: Tuple[int, ...], dtype: int, batch: int, layout: int = AOS):
        self.ctx, self.shape, self.dtype, self.batch, self.layout = ctx, shape, dtype, batch, layout
        self.elems = _prod(shape)
        h = ctypes.c_void_p()
        check(_lib.lib().t5_alloc(ctx._h, dtype, self.elems, batch, layout, ctypes.byref(h)), ctx)
        self._h = h
        ctx._bufs.add(self)

    @property
    def np_dtype(self):
        return NP_OF[self.dtype]

    @property
    def nbytes(self) -> int:
        return self.batch * self.elems * self.np_dtype.itemsize

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=self.np_dtype)
        check(_lib.lib().t5_upload(self._h, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes), self.ctx)

    def to_numpy(self) -> np.ndarray:
        """``toList`` (Vector.hs:367) as an array of shape [batch, *shape]."""
        out = np.empty((self.batch, *self.shape), dtype=self.np_dtype)
        check(_lib.lib().t5_download(self._h, out.ctypes.data_as(ctypes.c_void_p), out.nbytes), self.ctx)
        return out

    def to_list(self) -> list:
        return self.to_numpy().ravel().tolist()

    def relayout(self, layout: int) -> "Batch":
        out = Batch(self.ctx, self.shape, self.dtype, self.batch, layout)
        check(_lib.lib().t5_relayout(out._h, self._h), self.ctx)
        return out

    def plant_specials(self, swap_every: int = 1024, sing_every: int = 65536):
        assert len(self.shape) == 2 and self.shape[0] == self.shape[1]
        check(_lib.lib().t5_plant_specials(self._h, self.shape[0], swap_every, sing_every), self.ctx)
        return self

    def free(self):
        if self._h:
            _lib.lib().t5_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if self._h and self.ctx._h:
                self.free()
        except Exception:
            pass

    # operator sugar mirroring the reference's infix names
    def __matmul__(self, other: "Batch") -> "Batch":  # (.*.)
        from .linear_algebra import matmul
        return matmul(self, other)

    def __add__(self, other: "Batch") -> "Batch":  # liftA2 (+)
        from .linear_algebra import zip_with
        return zip_with("+", self, other)

    def __sub__(self, other: "Batch") -> "Batch":
        from .linear_algebra import zip_with
        return zip_with("-", self, other)

    def __repr__(self):
        lay = "AoS" if self.layout == AOS else "SoA"
        return f"Batch(shape={self.shape}, dtype={self.np_dtype}, batch={self.batch}, {lay})"


class PinnedArray:
    """numpy view over ``t5_pinned_alloc`` memory (page-locked; full-speed PCIe copies)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape, dtype=np.int64)) * self.dtype.itemsize
        self._p = _lib.lib().t5_pinned_alloc(self.nbytes)
        if not self._p:
            raise DeviceError("t5_pinned_alloc failed")
        buf = (ctypes.c_char * max(1, self.nbytes)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=self.nbytes // self.dtype.itemsize).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            _lib.lib().t5_pinned_free(self._p)
            self._p = None
