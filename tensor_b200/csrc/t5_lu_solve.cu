// solve (any number of right-hand sides) for 9 <= n <= 128, and small n with unusual column counts (t5_lu.cuh, MODE 1).
#include "t5_lu.cuh"

namespace t5 {
int lu_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (nrhs < 1) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) return lu::dispatch<double, 1>(s, n, nrhs, A, B, X, status);
    if (dtype == T5D_F32) return lu::dispatch<float, 1>(s, n, nrhs, A, B, X, status);
    return T5I_NOT_SPECIALISED;
}
}  // namespace t5
