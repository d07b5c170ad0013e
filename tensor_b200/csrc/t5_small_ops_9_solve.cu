// The `solve` (one right-hand side) instances of t5_small_ops_9.cu (see there): one translation unit per operation, so that the three build in parallel.
#include "t5_tile_launch.cuh"

namespace t5 {

template <class Op>
static int launch_reg(const Shard& s, const void* in0, const void* in1, void* out0, void* out1) {
    return launch_tile<Op, 64, 1, 1, 1>(s, in0, in1, out0, out1, OpParams{});
}

int reg9_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (s.layout != T5L_AOS || nrhs != 1) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) {
        if (n == 9) return launch_reg<SolveOp<double, 9, 1>>(s, A, B, X, status);
        if (n == 10) return launch_reg<SolveOp<double, 10, 1>>(s, A, B, X, status);
        if (n == 11) return launch_reg<SolveOp<double, 11, 1>>(s, A, B, X, status);
        if (n == 12) return launch_reg<SolveOp<double, 12, 1>>(s, A, B, X, status);
        if (n == 13) return launch_reg<SolveOp<double, 13, 1>>(s, A, B, X, status);
    } else if (dtype == T5D_F32) {
        if (n == 9) return launch_reg<SolveOp<float, 9, 1>>(s, A, B, X, status);
        if (n == 10) return launch_reg<SolveOp<float, 10, 1>>(s, A, B, X, status);
        if (n == 11) return launch_reg<SolveOp<float, 11, 1>>(s, A, B, X, status);
        if (n == 12) return launch_reg<SolveOp<float, 12, 1>>(s, A, B, X, status);
        if (n == 13) return launch_reg<SolveOp<float, 13, 1>>(s, A, B, X, status);
        if (n == 14) return launch_reg<SolveOp<float, 14, 1>>(s, A, B, X, status);
        if (n == 15) return launch_reg<SolveOp<float, 15, 1>>(s, A, B, X, status);
        if (n == 16) return launch_reg<SolveOp<float, 16, 1>>(s, A, B, X, status);
        if (n == 17) return launch_reg<SolveOp<float, 17, 1>>(s, A, B, X, status);
        if (n == 18) return launch_reg<SolveOp<float, 18, 1>>(s, A, B, X, status);
        if (n == 19) return launch_reg<SolveOp<float, 19, 1>>(s, A, B, X, status);
        if (n == 20) return launch_reg<SolveOp<float, 20, 1>>(s, A, B, X, status);
    }
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
