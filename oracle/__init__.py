"""ORACLE — test infrastructure only (see oracle/t5_oracle.cpp header).

ctypes front-end of ``oracle/libt5oracle.so``.  Importable ONLY from ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs.  ``tensor_b200`` never imports this package.

Parity status: transpose / index maps follow existing reference code
(src/Data/Tensor/Vector.hs:150-168, :295-307); every other operation is the
specification of SURVEY.md §8a because the reference checkout has no
Data.Tensor.LinearAlgebra — **parity unpinned** for those.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libt5oracle.so")

F64, F32, I64 = 0, 1, 2
NP_DTYPE = {F64: np.float64, F32: np.float32, I64: np.int64}
DTYPE_OF = {np.dtype(np.float64): F64, np.dtype(np.float32): F32, np.dtype(np.int64): I64}


def build() -> str:
    """Compile the oracle with g++ (idempotent)."""
    src = os.path.join(_HERE, "t5_oracle.cpp")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["bash", os.path.join(_HERE, "build.sh")])
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
        L.t5o_linearize.restype = u64
        L.t5o_linearize.argtypes = [u64p, u64p, i32]
        L.t5o_unlinearize.restype = None
        L.t5o_unlinearize.argtypes = [u64p, i32, u64, u64p]
        L.t5o_transposeV.restype = u64
        L.t5o_transposeV.argtypes = [u64p, i32, u64]
        L.t5o_transpose.argtypes = [i32, i32, u64p, sz, vp, vp, i32]
        L.t5o_matmul.argtypes = [i32, i32, i32, i32, sz, vp, vp, vp, i32]
        L.t5o_dot.argtypes = [i32, i32, sz, vp, vp, vp, i32]
        L.t5o_dot_sum.argtypes = [i32, i32, sz, vp, vp, vp]
        L.t5o_prod.argtypes = [i32, i32, u64p, i32, u64p, i32, sz, vp, vp, vp, i32]
        L.t5o_det.argtypes = [i32, i32, sz, vp, vp, i32]
        L.t5o_inverse.argtypes = [i32, i32, sz, vp, vp, vp, i32]
        L.t5o_solve.argtypes = [i32, i32, i32, sz, vp, vp, vp, vp, i32]
        L.t5o_trace.argtypes = [i32, i32, sz, vp, vp, i32]
        L.t5o_poly_eval.argtypes = [i32, i32, vp, i32, sz, vp, vp, i32]
        L.t5o_char_poly.argtypes = [i32, i32, sz, vp, vp, i32]
        L.t5o_min_poly.argtypes = [i32, i32, sz, vp, vp, vp, i32]
        L.t5o_row_echelon.argtypes = [i32, i32, i32, i32, sz, vp, vp, vp, i32]
        L.t5o_fill.argtypes = [i32, u64, u64, u64, sz, vp]
        L.t5o_plant_specials.argtypes = [i32, i32, u64, sz, vp, u64, u64]
        _lib = L
    return _lib


def _u64(xs):
    arr = (ctypes.c_uint64 * max(1, len(xs)))(*xs)
    return arr


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt=None):
    a = np.ascontiguousarray(a) if dt is None else np.ascontiguousarray(a, dtype=dt)
    return a


def _chk(rc):
    if rc != 0:
        raise ValueError(f"oracle error code {rc}")


# ---- index conventions -----------------------------------------------------
def linearize(ds, ix) -> int:
    return int(lib().t5o_linearize(_u64(ds), _u64(ix), len(ds)))


def unlinearize(ds, i) -> list:
    out = (ctypes.c_uint64 * max(1, len(ds)))()
    lib().t5o_unlinearize(_u64(ds), len(ds), i, out)
    return [int(out[a]) for a in range(len(ds))]


def transposeV(ds, i) -> int:
    return int(lib().t5o_transposeV(_u64(ds), len(ds), i))


# ---- operations; every array is [batch, *dims] C-contiguous -----------------
def transpose(a: np.ndarray, dims, threads: int = 1) -> np.ndarray:
    dims = list(dims)
    a = _c(a)
    batch = a.size // max(1, int(np.prod(dims, dtype=np.int64)))
    out = np.empty((batch, *dims[::-1]), dtype=a.dtype)
    _chk(lib().t5o_transpose(a.dtype.itemsize, len(dims), _u64(dims), batch, _ptr(a), _ptr(out), threads))
    return out


def matmul(a: np.ndarray, b: np.ndarray, m: int, k: int, n: int, threads: int = 1) -> np.ndarray:
    a, b = _c(a), _c(b)
    batch = a.size // (m * k)
    out = np.empty((batch, m, n), dtype=a.dtype)
    _chk(lib().t5o_matmul(DTYPE_OF[a.dtype], m, k, n, batch, _ptr(a), _ptr(b), _ptr(out), threads))
    return out


def dot(x: np.ndarray, y: np.ndarray, n: int, threads: int = 1) -> np.ndarray:
    x, y = _c(x), _c(y)
    batch = x.size // n
    out = np.empty((batch,), dtype=x.dtype)
    _chk(lib().t5o_dot(DTYPE_OF[x.dtype], n, batch, _ptr(x), _ptr(y), _ptr(out), threads))
    return out


def dot_sum(x: np.ndarray, y: np.ndarray, n: int):
    x, y = _c(x), _c(y)
    batch = x.size // n
    out = np.zeros((1,), dtype=x.dtype)
    _chk(lib().t5o_dot_sum(DTYPE_OF[x.dtype], n, batch, _ptr(x), _ptr(y), _ptr(out)))
    return out[0]


def prod(a: np.ndarray, dims_a, b: np.ndarray, dims_b, ncontract: int, threads: int = 1) -> np.ndarray:
    a, b = _c(a), _c(b)
    dims_a, dims_b = list(dims_a), list(dims_b)
    da = int(np.prod(dims_a, dtype=np.int64)) if dims_a else 1
    batch = a.size // da
    out_dims = dims_a[: len(dims_a) - ncontract] + dims_b[ncontract:]
    out = np.empty((batch, *out_dims), dtype=a.dtype)
    _chk(lib().t5o_prod(DTYPE_OF[a.dtype], len(dims_a), _u64(dims_a), len(dims_b), _u64(dims_b), ncontract,
                        batch, _ptr(a), _ptr(b), _ptr(out), threads))
    return out


def det(a: np.ndarray, n: int, threads: int = 1) -> np.ndarray:
    a = _c(a)
    batch = a.size // (n * n)
    out = np.empty((batch,), dtype=a.dtype)
    _chk(lib().t5o_det(DTYPE_OF[a.dtype], n, batch, _ptr(a), _ptr(out), threads))
    return out


def inverse(a: np.ndarray, n: int, threads: int = 1):
    a = _c(a)
    batch = a.size // (n * n)
    out = np.empty((batch, n, n), dtype=a.dtype)
    status = np.empty((batch,), dtype=np.uint8)
    _chk(lib().t5o_inverse(DTYPE_OF[a.dtype], n, batch, _ptr(a), _ptr(out), _ptr(status), threads))
    return out, status


def solve(a: np.ndarray, b: np.ndarray, n: int, nrhs: int, threads: int = 1):
    a, b = _c(a), _c(b)
    batch = a.size // (n * n)
    out = np.empty((batch, n, nrhs), dtype=a.dtype)
    status = np.empty((batch,), dtype=np.uint8)
    _chk(lib().t5o_solve(DTYPE_OF[a.dtype], n, nrhs, batch, _ptr(a), _ptr(b), _ptr(out), _ptr(status), threads))
    return out, status


# ---- tr / polyEval / charPoly / minPoly (parity unpinned) ------
def trace(a: np.ndarray, n: int, threads: int = 1) -> np.ndarray:
    a = _c(a)
    batch = a.size // (n * n)
    out = np.empty((batch,), dtype=a.dtype)
    _chk(lib().t5o_trace(DTYPE_OF[a.dtype], n, batch, _ptr(a), _ptr(out), threads))
    return out


def poly_eval(a: np.ndarray, coeffs, n: int, threads: int = 1) -> np.ndarray:
    """coeffs are low order first: c0 I + c1 A + ..."""
    a = _c(a)
    c = np.ascontiguousarray(coeffs, dtype=a.dtype)
    batch = a.size // (n * n)
    out = np.empty((batch, n, n), dtype=a.dtype)
    _chk(lib().t5o_poly_eval(DTYPE_OF[a.dtype], n, _ptr(c), int(c.size), batch, _ptr(a), _ptr(out), threads))
    return out


def char_poly(a: np.ndarray, n: int, threads: int = 1) -> np.ndarray:
    """[batch, n+1] coefficients of det(xI - A), low order first, last one = 1."""
    a = _c(a)
    batch = a.size // (n * n)
    out = np.empty((batch, n + 1), dtype=a.dtype)
    _chk(lib().t5o_char_poly(DTYPE_OF[a.dtype], n, batch, _ptr(a), _ptr(out), threads))
    return out


def min_poly(a: np.ndarray, n: int, threads: int = 1):
    """([batch, n+1] coefficients low order first, zero above the degree; [batch] degrees)."""
    a = _c(a)
    batch = a.size // (n * n)
    out = np.empty((batch, n + 1), dtype=a.dtype)
    deg = np.empty((batch,), dtype=np.uint8)
    _chk(lib().t5o_min_poly(DTYPE_OF[a.dtype], n, batch, _ptr(a), _ptr(out), _ptr(deg), threads))
    return out, deg


def row_echelon(a: np.ndarray, m: int, n: int, pivot_cols: int | None = None, threads: int = 1):
    """([batch, m, n] reduced row echelon forms, [batch] ranks); pivots are searched in the first
    pivot_cols columns only (default: all n)."""
    a = _c(a)
    batch = a.size // (m * n)
    out = np.empty((batch, m, n), dtype=a.dtype)
    rank = np.empty((batch,), dtype=np.uint8)
    pc = n if pivot_cols is None else int(pivot_cols)
    _chk(lib().t5o_row_echelon(DTYPE_OF[a.dtype], m, n, pc, batch, _ptr(a), _ptr(out), _ptr(rank), threads))
    return out, rank


# ---- synthetic inputs ------------------------------------------
SEED_A = 0x5EED000000000001
SEED_B = 0x5EED000000000002


def fill(dtype: int, seed: int, op_id: int, count: int, first_elem: int = 0) -> np.ndarray:
    out = np.empty((count,), dtype=NP_DTYPE[dtype])
    _chk(lib().t5o_fill(dtype, seed, op_id, first_elem, count, _ptr(out)))
    return out


def plant_specials(a: np.ndarray, n: int, first_item: int = 0, swap_every: int = 1024, sing_every: int = 65536):
    assert a.flags.c_contiguous
    batch = a.size // (n * n)
    _chk(lib().t5o_plant_specials(DTYPE_OF[a.dtype], n, first_item, batch, _ptr(a), swap_every, sing_every))
    return a
