// solve A X = B with a matrix right-hand side for 5 <= n <= 8 (2, 3, 4 or n columns), record per thread
// in the tile pipeline (SolveManyOp, t5_small_ops.cuh: LU with kept multipliers, right-hand side read
// lazily from its shared-memory slot two columns at a time, results streamed back in place).  Other
// column counts stay with the row-per-lane kernel (t5_rowelim.cu).  Bit-exact TU: -fmad=false.
#include "t5_tile_launch.cuh"

namespace t5 {

// Two operand records per item: from ~450 B on (6x6 Double with 4+ columns, 7x7, 8x8) a two-stage ring leaves one or two CTAs per SM;
// the one-stage pipeline (t5_tile.cuh) doubles that.
template <class T, int N, int R> struct Geo<SolveManyOp<T, N, R>> {
    using type = std::conditional_t<(sizeof(T) * (N * N + N * R) >= 448), G<64, 1, 1>, G<64, 1, 2>>;
};

template <class T, int N>
static int many_n(const Shard& s, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (nrhs == 2) return launch_op<SolveManyOp<T, N, 2>>(s, A, B, X, status);
    if (nrhs == 3) return launch_op<SolveManyOp<T, N, 3>>(s, A, B, X, status);
    if (nrhs == 4) return launch_op<SolveManyOp<T, N, 4>>(s, A, B, X, status);
    if (nrhs == N) return launch_op<SolveManyOp<T, N, N>>(s, A, B, X, status);
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int many_t(const Shard& s, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    switch (n) {
        case 5: return many_n<T, 5>(s, nrhs, A, B, X, status);
        case 6: return many_n<T, 6>(s, nrhs, A, B, X, status);
        case 7: return many_n<T, 7>(s, nrhs, A, B, X, status);
        case 8: return many_n<T, 8>(s, nrhs, A, B, X, status);
    }
    return T5I_NOT_SPECIALISED;
}

int many_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (dtype == T5D_F64) return many_t<double>(s, n, nrhs, A, B, X, status);
    if (dtype == T5D_F32) return many_t<float>(s, n, nrhs, A, B, X, status);
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
