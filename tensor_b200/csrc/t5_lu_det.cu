// det for 9 <= n <= 128 (t5_lu.cuh, MODE 0).  One operation per TU: the three instantiate ~20 kernels each.
#include "t5_lu.cuh"

namespace t5 {
int lu_det(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (dtype == T5D_F64) return lu::dispatch<double, 0>(s, n, 0, A, nullptr, O, nullptr);
    if (dtype == T5D_F32) return lu::dispatch<float, 0>(s, n, 0, A, nullptr, O, nullptr);
    return T5I_NOT_SPECIALISED;
}
}  // namespace t5
