// solve Float, 9 <= n <= 16 and 30 <= n <= 32, a matrix row per lane (t5_warp_lu.cuh, MODE 1).  One operation per TU so that they compile in parallel.
#include "t5_warp_lu.cuh"

namespace t5 {
int warp_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (dtype == T5D_F32) return wlu::dispatch<float, 1>(s, n, nrhs, A, B, X, status);
    return T5I_NOT_SPECIALISED;
}
}  // namespace t5
