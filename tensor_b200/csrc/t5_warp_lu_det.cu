// det Float, 9 <= n <= 16 and 30 <= n <= 32, a matrix row per lane (t5_warp_lu.cuh, MODE 0).  One operation per TU so that they compile in parallel.
#include "t5_warp_lu.cuh"

namespace t5 {
int warp_det(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (dtype == T5D_F32) return wlu::dispatch<float, 0>(s, n, 0, A, nullptr, O, nullptr);
    return T5I_NOT_SPECIALISED;
}
}  // namespace t5
