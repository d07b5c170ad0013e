// inverse for 9 <= n <= 128 (t5_lu.cuh, MODE 2: right-hand sides = the permuted identity).
#include "t5_lu.cuh"

namespace t5 {
int lu_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status) {
    if (dtype == T5D_F64) return lu::dispatch<double, 2>(s, n, n, A, nullptr, O, status);
    if (dtype == T5D_F32) return lu::dispatch<float, 2>(s, n, n, A, nullptr, O, status);
    return T5I_NOT_SPECIALISED;
}
}  // namespace t5
