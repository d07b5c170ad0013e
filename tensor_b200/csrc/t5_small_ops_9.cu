// det / solve / inverse just beyond 8x8 as a record per thread in the one-stage tile pipeline.
//
// The record-per-thread eliminations of t5_small_ops.cu stop at 8x8 and the CTA-per-matrix kernels of t5_lu.cuh take over:
// 1Mi 9x9 Double `det` took 0.82 ms against 0.096 ms for 8x8 - a factor of 8.5 for 27 % more data (`solve` 1.7 ms against
// 0.12).  An 81-element Double record still fits a thread's 255 registers for `det` and `solve` (the LU is in place;
// `inverse` parks the factors in shared memory and replays all columns, t5_small_ops.cuh), and a thread that spills a few
// hundred bytes to local memory (L1-resident: 64-thread CTAs) is still far ahead of a CTA that synchronises on every
// elimination step.  So these sizes keep the thread-per-item form - compile-time indices, no shuffles, no barriers, the
// oracle's operation order (bit-identical) - up to the size where the other kernels win, measured per operation and type
// (1Mi items, ms, before -> after; profiles/r02o.. r02q_time_lu_reg9.txt):
//   Double det   9: 0.816 -> 0.118 (0.90 of the HBM peak)  10: 0.878 -> 0.219  11: 0.946 -> 0.331  12: 0.893 -> 0.615  (13: 0.842 -> 0.931, not taken)
//   Double solve 9: 1.734 -> 0.156   10: 1.911 -> 0.371   11: 2.102 -> 0.559   12: 2.021 -> 1.166   13: 1.900 -> 1.375
//   Double inverse 9: 2.475 -> 0.441   10: 2.714 -> 0.906   11: 2.971 -> 1.628   (12: 2.854 -> 3.125, not taken)
//   Float det   9: 0.517 -> 0.059  10: 0.069  11: 0.095  12: 0.686 -> 0.148  14: 0.821 -> 0.188  16: 0.934 -> 0.407  17: 1.709 -> 0.554  18: 1.632 -> 1.083
//   Float solve 9: 1.186 -> 0.071  12: 1.301 -> 0.176  14: 1.467 -> 0.226  16: 1.462 -> 0.573  17: 4.071 -> 1.069  18: 3.963 -> 1.433
//               19: 1.680 and 20: 2.240 (743Ki / 671Ki items; the CTA-per-matrix kernel needs ~3.9 ms at these sizes)
//   Float inverse 9: 1.717 -> 0.265  10: 1.769 -> 0.324  11: 1.950 -> 0.597  12: 2.040 -> 0.965  13: 2.133 -> 1.216  14: 2.326 -> 1.586  (15, 16 slower)
// Registers / spills (ptxas): Double det 9 254 / 0, 10 255 / 348 B, 12 255 / 2.6 KB; Float det 14 226 / 0, 16 255 / 256 B.
// AoS only (SoA operands arrive through the host layer's relayout detour).  Compiled with -fmad=false like every exact TU.
// This file holds the `det` instances; `solve` and `inverse` are t5_small_ops_9_solve.cu / _inverse.cu (parallel builds).
#include "t5_tile_launch.cuh"

namespace t5 {

template <class Op>
static int launch_reg(const Shard& s, const void* in0, const void* in1, void* out0, void* out1) {
    return launch_tile<Op, 64, 1, 1, 1>(s, in0, in1, out0, out1, OpParams{});
}

int reg9_det(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) {
        if (n == 9) return launch_reg<DetOp<double, 9>>(s, A, nullptr, O, nullptr);
        if (n == 10) return launch_reg<DetOp<double, 10>>(s, A, nullptr, O, nullptr);
        if (n == 11) return launch_reg<DetOp<double, 11>>(s, A, nullptr, O, nullptr);
        if (n == 12) return launch_reg<DetOp<double, 12>>(s, A, nullptr, O, nullptr);
    } else if (dtype == T5D_F32) {
        if (n == 9) return launch_reg<DetOp<float, 9>>(s, A, nullptr, O, nullptr);
        if (n == 10) return launch_reg<DetOp<float, 10>>(s, A, nullptr, O, nullptr);
        if (n == 11) return launch_reg<DetOp<float, 11>>(s, A, nullptr, O, nullptr);
        if (n == 12) return launch_reg<DetOp<float, 12>>(s, A, nullptr, O, nullptr);
        if (n == 13) return launch_reg<DetOp<float, 13>>(s, A, nullptr, O, nullptr);
        if (n == 14) return launch_reg<DetOp<float, 14>>(s, A, nullptr, O, nullptr);
        if (n == 15) return launch_reg<DetOp<float, 15>>(s, A, nullptr, O, nullptr);
        if (n == 16) return launch_reg<DetOp<float, 16>>(s, A, nullptr, O, nullptr);
        if (n == 17) return launch_reg<DetOp<float, 17>>(s, A, nullptr, O, nullptr);
        if (n == 18) return launch_reg<DetOp<float, 18>>(s, A, nullptr, O, nullptr);
    }
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
