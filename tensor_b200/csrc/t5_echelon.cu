// rowEchelonForm in registers, n = 2..4 (n x n, n x (n+1), n x 2n); the op itself: t5_echelon_op.cuh.
// n = 5..8 live in t5_echelon_large.cu (their 20-44 step instances each compile for a while; two TUs build
// in parallel).  Bit-exact TU: -fmad=false.
#include "t5_echelon_op.cuh"

namespace t5 {

int large_row_echelon(const Shard& s, int dtype, int m, int n, const OpParams& prm, const void* A, void* O, void* rank);

template <class T>
static int echelon_t(const Shard& s, int m, int n, const OpParams& prm, const void* A, void* O, void* rank) {
#define T5_RE(M_, N_) \
    if (m == M_ && n == N_) return launch_op<RowEchelonOp<T, M_, N_>>(s, A, nullptr, O, rank, prm);
    T5_RE(2, 2) T5_RE(3, 3) T5_RE(4, 4)       // rowEchelonForm of the small square matrices
    T5_RE(2, 3) T5_RE(3, 4) T5_RE(4, 5)       // [A | b]
    T5_RE(2, 4) T5_RE(3, 6) T5_RE(4, 8)       // [A | I]
#undef T5_RE
    return T5I_NOT_SPECIALISED;
}

int small_row_echelon(const Shard& s, int dtype, int m, int n, int pivot_cols, const void* A, void* O, void* rank) {
    OpParams prm{};
    prm.n = pivot_cols;
    if (m >= 5) return large_row_echelon(s, dtype, m, n, prm, A, O, rank);
    if (dtype == T5D_F64) return echelon_t<double>(s, m, n, prm, A, O, rank);
    if (dtype == T5D_F32) return echelon_t<float>(s, m, n, prm, A, O, rank);
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
