"""Counter-based synthetic inputs on the host, vectorised with numpy.

Same stream as the device generator (``t5_fill_synthetic``) and the oracle's
``t5o_fill``: element i of the global batch = f(splitmix64((seed ^ (op_id << 56)) + i)).
Used by bench.py to build the HOST inputs of the end-to-end leg.
"""
from __future__ import annotations

import numpy as np

SEED_A = 0x5EED000000000001
SEED_B = 0x5EED000000000002
_M = (1 << 64) - 1


def splitmix64(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        z += np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def fill(dtype, seed: int, op_id: int, count: int, first_elem: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    dtype = np.dtype(dtype)
    base = (seed ^ ((op_id << 56) & _M)) & _M
    res = out if out is not None else np.empty((count,), dtype=dtype)
    flat = res.reshape(-1)
    step = 1 << 22
    for s in range(0, count, step):
        n = min(step, count - s)
        with np.errstate(over="ignore"):
            ctr = np.arange(n, dtype=np.uint64) + np.uint64((base + first_elem + s) & _M)
        x = splitmix64(ctr)
        if dtype == np.float64:
            flat[s:s + n] = ((x >> np.uint64(11)).astype(np.float64) * 2.0 ** -53) * 2.0 - 1.0
        elif dtype == np.float32:
            flat[s:s + n] = ((x >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)) * np.float32(2.0) - np.float32(1.0)
        elif dtype == np.int64:
            flat[s:s + n] = x.view(np.int64)
        else:
            raise ValueError(dtype)
    return res
