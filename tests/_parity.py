"""Shared comparison for `.*.` results: bit-exact against the oracle where the library promises the sequential fold
(`t5_matmul_is_exact`), the stated tolerance where the shape runs on the tensor cores (include/t5b200.h, t5_matmul:
Double |delta| <= 1e-12 * sum_j |a_ij||b_jc|, Float 1e-5)."""
import numpy as np

import oracle

TOL = {oracle.F64: 1e-12, oracle.F32: 1e-5}


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize])


def matmul_is_exact(dc, m, k, n):
    import tensor_b200 as tb
    return tb.matmul_is_exact({oracle.F64: tb.F64, oracle.F32: tb.F32, oracle.I64: tb.I64}[dc], m, k, n)


def assert_matmul_matches(C, a, b, m, k, n, dc, what=""):
    """C: the library's (batch, m, n) result for operands a (batch*m*k) and b (batch*k*n) of oracle dtype code dc."""
    batch = a.size // (m * k)
    C = np.asarray(C).reshape(batch, m, n)
    want = oracle.matmul(a, b, m, k, n, threads=8).reshape(batch, m, n)
    if matmul_is_exact(dc, m, k, n):
        if not np.array_equal(bits(C), bits(want)):
            bad = np.argwhere(bits(C).ravel() != bits(want).ravel()).ravel()
            raise AssertionError(f"{what}: {bad.size} mismatching elements, first at {bad[:5]}: "
                                 f"got {C.ravel()[bad[:5]]} want {want.ravel()[bad[:5]]}")
        return "exact"
    a3 = np.abs(a.reshape(batch, m, k)).astype(np.float64)
    b3 = np.abs(b.reshape(batch, k, n)).astype(np.float64)
    bound = TOL[dc] * np.einsum("bij,bjk->bik", a3, b3)
    if dc == oracle.F32:        # the Float bar is against the real product, not against the Float fold's own rounding
        want = np.einsum("bij,bjk->bik", a.reshape(batch, m, k).astype(np.float64), b.reshape(batch, k, n).astype(np.float64))
    err = np.abs(C.astype(np.float64) - want)
    assert np.all(err <= bound), f"{what}: max excess {np.max(err / np.maximum(bound, 1e-300))}"
    return "tolerance"
