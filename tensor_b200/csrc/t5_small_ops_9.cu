// det / solve / inverse of 9x9 and 10x10 (Float: up to 12x12) as a record per thread in the one-stage tile pipeline.
//
// The record-per-thread eliminations of t5_small_ops.cu stop at 8x8 and the CTA-per-matrix kernels of t5_lu.cuh take over:
// 1Mi 9x9 Double `det` 0.82 ms against 0.096 ms for 8x8 - a factor of 8.5 for 27 % more data (`solve` 2.4 ms against
// 0.12).  An 81- or 100-element Double record still fits a thread's 255 registers for `det` and `solve` (the LU is in
// place; `inverse` parks the factors in shared memory and replays all columns, t5_small_ops.cuh), so these sizes keep the
// thread-per-item form: compile-time indices, no shuffles, no barriers, the oracle's operation order (bit-identical).
// AoS only (SoA operands arrive through the host layer's relayout detour).  Compiled with -fmad=false like every exact TU.
#include "t5_tile_launch.cuh"
#include <cstdlib>

namespace t5 {

static bool reg9_off() { static const bool off = getenv("T5_NO_REG9") != nullptr; return off; }   // tuning hook

template <class Op>
static int launch_reg(const Shard& s, const void* in0, const void* in1, void* out0, void* out1) {
    return launch_tile<Op, 64, 1, 1, 1>(s, in0, in1, out0, out1, OpParams{});
}

int reg9_det(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (s.layout != T5L_AOS || reg9_off()) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) {
        if (n == 9) return launch_reg<DetOp<double, 9>>(s, A, nullptr, O, nullptr);
        if (n == 10) return launch_reg<DetOp<double, 10>>(s, A, nullptr, O, nullptr);
    } else if (dtype == T5D_F32) {
        if (n == 9) return launch_reg<DetOp<float, 9>>(s, A, nullptr, O, nullptr);
        if (n == 10) return launch_reg<DetOp<float, 10>>(s, A, nullptr, O, nullptr);
        if (n == 11) return launch_reg<DetOp<float, 11>>(s, A, nullptr, O, nullptr);
        if (n == 12) return launch_reg<DetOp<float, 12>>(s, A, nullptr, O, nullptr);
        if (n == 13) return launch_reg<DetOp<float, 13>>(s, A, nullptr, O, nullptr);
        if (n == 14) return launch_reg<DetOp<float, 14>>(s, A, nullptr, O, nullptr);
    }
    return T5I_NOT_SPECIALISED;
}

int reg9_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (s.layout != T5L_AOS || nrhs != 1 || reg9_off()) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) {
        if (n == 9) return launch_reg<SolveOp<double, 9, 1>>(s, A, B, X, status);
        if (n == 10) return launch_reg<SolveOp<double, 10, 1>>(s, A, B, X, status);
    } else if (dtype == T5D_F32) {
        if (n == 9) return launch_reg<SolveOp<float, 9, 1>>(s, A, B, X, status);
        if (n == 10) return launch_reg<SolveOp<float, 10, 1>>(s, A, B, X, status);
        if (n == 11) return launch_reg<SolveOp<float, 11, 1>>(s, A, B, X, status);
        if (n == 12) return launch_reg<SolveOp<float, 12, 1>>(s, A, B, X, status);
        if (n == 13) return launch_reg<SolveOp<float, 13, 1>>(s, A, B, X, status);
        if (n == 14) return launch_reg<SolveOp<float, 14, 1>>(s, A, B, X, status);
    }
    return T5I_NOT_SPECIALISED;
}

int reg9_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status) {
    if (s.layout != T5L_AOS || reg9_off()) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) {
        if (n == 9) return launch_reg<InverseOp<double, 9>>(s, A, nullptr, O, status);
        if (n == 10) return launch_reg<InverseOp<double, 10>>(s, A, nullptr, O, status);
    } else if (dtype == T5D_F32) {
        if (n == 9) return launch_reg<InverseOp<float, 9>>(s, A, nullptr, O, status);
        if (n == 10) return launch_reg<InverseOp<float, 10>>(s, A, nullptr, O, status);
        if (n == 11) return launch_reg<InverseOp<float, 11>>(s, A, nullptr, O, status);
        if (n == 12) return launch_reg<InverseOp<float, 12>>(s, A, nullptr, O, status);
    }
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
