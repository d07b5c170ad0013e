// rowEchelonForm in registers for n = 5..8: n x n and n x (n+1) (up to 72 elements, up to 44 step
// instances per shape); see t5_echelon_op.cuh.  The Float half of t5_echelon_large.cu.  Bit-exact TU: -fmad=false.
#include "t5_echelon_op.cuh"

namespace t5 {

template <class T>
static int echelon_large_t(const Shard& s, int m, int n, const OpParams& prm, const void* A, void* O, void* rank) {
#define T5_RE(M_, N_) \
    if (m == M_ && n == N_) return launch_op<RowEchelonOp<T, M_, N_>>(s, A, nullptr, O, rank, prm);
    T5_RE(5, 5) T5_RE(6, 6) T5_RE(7, 7) T5_RE(8, 8)
    T5_RE(5, 6) T5_RE(6, 7) T5_RE(7, 8) T5_RE(8, 9)
#undef T5_RE
    return T5I_NOT_SPECIALISED;
}

int large_row_echelon_f32(const Shard& s, int m, int n, const OpParams& prm, const void* A, void* O, void* rank) {
    return echelon_large_t<float>(s, m, n, prm, A, O, rank);
}

}  // namespace t5
