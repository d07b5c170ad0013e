// The `inverse` instances of t5_small_ops_9.cu (see there): one translation unit per operation, so that the three build in parallel.
#include "t5_tile_launch.cuh"

namespace t5 {

template <class Op>
static int launch_reg(const Shard& s, const void* in0, const void* in1, void* out0, void* out1) {
    return launch_tile<Op, 64, 1, 1, 1>(s, in0, in1, out0, out1, OpParams{});
}

int reg9_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) {
        if (n == 9) return launch_reg<InverseOp<double, 9>>(s, A, nullptr, O, status);
        if (n == 10) return launch_reg<InverseOp<double, 10>>(s, A, nullptr, O, status);
        if (n == 11) return launch_reg<InverseOp<double, 11>>(s, A, nullptr, O, status);
    } else if (dtype == T5D_F32) {
        if (n == 9) return launch_reg<InverseOp<float, 9>>(s, A, nullptr, O, status);
        if (n == 10) return launch_reg<InverseOp<float, 10>>(s, A, nullptr, O, status);
        if (n == 11) return launch_reg<InverseOp<float, 11>>(s, A, nullptr, O, status);
        if (n == 12) return launch_reg<InverseOp<float, 12>>(s, A, nullptr, O, status);
        if (n == 13) return launch_reg<InverseOp<float, 13>>(s, A, nullptr, O, status);
        if (n == 14) return launch_reg<InverseOp<float, 14>>(s, A, nullptr, O, status);
    }
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
