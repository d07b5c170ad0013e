"""ctypes binding of ``libt5b200.so`` (the C ABI declared in include/t5b200.h).

There is no CPU fallback anywhere in this package: if the CUDA library has not
been built (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C tensor_b200/csrc``) importing the binding raises, and every operation
on a box without a GPU fails in ``t5_init``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libt5b200.so")

F64, F32, I64, U8 = 0, 1, 2, 3
AOS, SOA = 0, 1
OK, ERR_SHAPE, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_OOM, ERR_INVALID = range(7)

_vp, _sz, _i, _u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
_u64p = ctypes.POINTER(ctypes.c_uint64)
_szp = ctypes.POINTER(ctypes.c_size_t)
_vpp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); must list every symbol of include/t5b200.h
SIGNATURES = {
    "t5_strerror": (ctypes.c_char_p, [_i]),
    "t5_last_error": (ctypes.c_char_p, [_vp]),
    "t5_version": (ctypes.c_char_p, []),
    "t5_init": (_i, [ctypes.POINTER(ctypes.c_int), _i, _vpp]),
    "t5_shutdown": (_i, [_vp]),
    "t5_live_buffers": (_i, [_vp]),
    "t5_ndev": (_i, [_vp]),
    "t5_sync": (_i, [_vp]),
    "t5_partition": (_i, [_sz, _i, _i, _szp, _szp]),
    "t5_alloc": (_i, [_vp, _i, _sz, _sz, _i, _vpp]),
    "t5_free": (_i, [_vp]),
    "t5_upload": (_i, [_vp, _vp, _sz]),
    "t5_download": (_i, [_vp, _vp, _sz]),
    "t5_relayout": (_i, [_vp, _vp]),
    "t5_pinned_alloc": (_vp, [_sz]),
    "t5_pinned_free": (None, [_vp]),
    "t5_buf_batch": (_sz, [_vp]),
    "t5_buf_elems": (_sz, [_vp]),
    "t5_buf_dtype": (_i, [_vp]),
    "t5_buf_layout": (_i, [_vp]),
    "t5_matmul": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "t5_dot": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "t5_dot_sum": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "t5_dot_sum_used_nccl": (_i, [_vp]),
    "t5_matmul_is_exact": (_i, [_i, _i, _i, _i]),
    "t5_prod": (_i, [_vp, _i, _i, _u64p, _i, _u64p, _i, _vp, _vp, _vp]),
    "t5_transpose": (_i, [_vp, _i, _i, _u64p, _vp, _vp]),
    "t5_det": (_i, [_vp, _i, _i, _vp, _vp]),
    "t5_inverse": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "t5_solve": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "t5_trace": (_i, [_vp, _i, _i, _vp, _vp]),
    "t5_poly_eval": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "t5_char_poly": (_i, [_vp, _i, _i, _vp, _vp]),
    "t5_min_poly": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "t5_row_echelon": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "t5_zip": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "t5_scale": (_i, [_vp, _i, _vp, _vp, _vp]),
    "t5_replicate": (_i, [_vp, _i, _vp, _vp]),
    "t5_append": (_i, [_vp, _i, _i, _i, _u64p, _u64p, _vp, _vp, _vp]),
    "t5_split": (_i, [_vp, _i, _i, _i, _u64p, _u64, _vp, _vp, _vp]),
    "t5_at": (_i, [_vp, _i, _i, _u64p, _u64p, _vp, _vp]),
    "t5_ta": (_i, [_vp, _i, _i, _u64p, _u64, _vp, _vp]),
    "t5_slice": (_i, [_vp, _i, _i, _u64p, _u64p, _vp, _vp]),
    "t5_permute_axes": (_i, [_vp, _i, _i, _u64p, ctypes.POINTER(ctypes.c_int), _vp, _vp]),
    "t5_structural_table": (_i, [_i, _i, _u64p, _u64p, _i, ctypes.POINTER(ctypes.c_uint32), _sz, _szp]),
    "t5_matmul_host": (_i, [_vp, _i, _i, _i, _i, _sz, _vp, _vp, _vp]),
    "t5_dot_host": (_i, [_vp, _i, _i, _sz, _vp, _vp, _vp]),
    "t5_prod_host": (_i, [_vp, _i, _i, _u64p, _i, _u64p, _i, _sz, _vp, _vp, _vp]),
    "t5_transpose_host": (_i, [_vp, _i, _i, _u64p, _sz, _vp, _vp]),
    "t5_det_host": (_i, [_vp, _i, _i, _sz, _vp, _vp]),
    "t5_inverse_host": (_i, [_vp, _i, _i, _sz, _vp, _vp, _vp]),
    "t5_solve_host": (_i, [_vp, _i, _i, _i, _sz, _vp, _vp, _vp, _vp]),
    "t5_timer_begin": (_i, [_vp]),
    "t5_timer_end": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "t5_timer_per_device": (_i, [_vp, ctypes.POINTER(ctypes.c_float), _i]),
    "t5_flush_l2": (_i, [_vp]),
    "t5_fill_synthetic": (_i, [_vp, _u64, _u64]),
    "t5_plant_specials": (_i, [_vp, _i, _u64, _u64]),
    "t5_launch_count": (_u64, [_vp]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load libt5b200.so; raises LibraryMissing (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise LibraryMissing(
                f"{SO_PATH} not built — run __graft_entry__.build() (nvcc, sm_100a). "
                "tensor_b200 has no CPU implementation."
            )
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def u64_array(xs):
    xs = [int(x) for x in xs]
    return (ctypes.c_uint64 * max(1, len(xs)))(*xs)


def partition(batch: int, ndev: int, g: int):
    """Pure host function of the library: shard g's (first, count)."""
    f, c = ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib().t5_partition(batch, ndev, g, ctypes.byref(f), ctypes.byref(c))
    if rc != OK:
        raise ValueError(lib().t5_strerror(rc).decode())
    return f.value, c.value


ST_APPEND, ST_SPLIT_FIRST, ST_SPLIT_SECOND, ST_AT, ST_TA, ST_SLICE, ST_PERMUTE = range(7)


def structural_table(op: int, dims, params):
    """Pure host function of the library: the gather table of a structural operation."""
    n = ctypes.c_size_t()
    L = lib()
    rc = L.t5_structural_table(op, len(dims), u64_array(dims), u64_array(params), len(params), None, 0, ctypes.byref(n))
    if rc != OK:
        raise ValueError(L.t5_strerror(rc).decode())
    buf = (ctypes.c_uint32 * max(1, n.value))()
    rc = L.t5_structural_table(op, len(dims), u64_array(dims), u64_array(params), len(params), buf, n.value, ctypes.byref(n))
    if rc != OK:
        raise ValueError(L.t5_strerror(rc).decode())
    return [int(buf[i]) for i in range(n.value)]
