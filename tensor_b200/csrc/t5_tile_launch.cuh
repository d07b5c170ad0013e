// Launchers of the tile pipeline (t5_tile.cuh), shared by the translation units that instantiate
// thread-per-item ops (t5_small_ops.cu, t5_echelon.cu): grid sizing from the occupancy of the
// instantiated kernel, tile claims per scheduler atomic, PDL launch, and the per-op tile geometry.
// A TU specialises Geo<Op> for its ops BEFORE the first launch_op<Op> it instantiates.
#pragma once
#include "t5_small_ops.cuh"
#include "t5_internal.h"
#include <type_traits>

namespace t5 {

// Tiles claimed per scheduler atomic.  One L2 address sustains only ~300 M atomics/s chip-wide, and at
// the HBM rate a tile of `tile_bytes` (operands + results) lasts tile_bytes / 6.5 TB/s: tiles under
// ~48 KB (64-thread tiles of small Float records: 10 KB) would be claimed faster than that, so they
// are claimed several at a time - as long as every CTA still gets >= 4 claims (the tail is one claim
// long).  Large batches claim several tiles anyway (>= 16 claims per CTA).
static inline size_t tile_claim(size_t tiles, int grid, size_t tile_bytes) {
    size_t claim = tiles / ((size_t)grid * 16);
    const size_t want = (48 * 1024 + tile_bytes - 1) / (tile_bytes ? tile_bytes : 1);
    const size_t room = tiles / ((size_t)grid * 4);
    const size_t small = want < room ? want : room;
    if (claim < small) claim = small;
    return claim < 1 ? 1 : (claim > 8 ? 8 : claim);
}

template <class Op, int THREADS, int IPT, int STAGES, int MIN_BLOCKS>
static int launch_tile(const Shard& s, const void* in0, const void* in1, void* out0, void* out1, const OpParams& prm) {
    using C = TileCfg<Op, THREADS, IPT, STAGES>;
    static_assert(C::SMEM <= 227 * 1024, "tile does not fit in shared memory");
    void (*kern)(TileArgs);
    if constexpr (STAGES == 1) kern = tile_kernel_1s<Op, THREADS, IPT, MIN_BLOCKS>;
    else kern = tile_kernel<Op, THREADS, IPT, STAGES, MIN_BLOCKS>;
    static int grid_cap[T5_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!grid_cap[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) return T5I_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, C::SMEM) != cudaSuccess) return T5I_CUDA;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (occ < 1) occ = 1;
        if (sms < 1) sms = T5_SM_COUNT_FALLBACK;
        grid_cap[dev] = occ * sms;
    }
    if (s.count == 0) return T5I_OK;
    if (!s.sched) return T5I_CUDA;
    size_t tiles = (s.count + C::TILE - 1) / C::TILE;
    int grid = (int)(tiles < (size_t)grid_cap[dev] ? tiles : (size_t)grid_cap[dev]);
    const size_t claim = tile_claim(tiles, grid, (size_t)C::TILE * (C::I0::bytes + C::I1::bytes + C::O0::bytes + C::O1::bytes));
    TileArgs a{in0, in1, out0, out1, (unsigned long long)s.count, s.sched, (unsigned)claim, prm};
    if (pdl_launch(kern, grid, THREADS, C::SMEM, s.stream, a) != cudaSuccess) { cudaGetLastError(); return T5I_CUDA; }
    return T5I_OK;
}

template <class Op, int THREADS, int IPT, int STAGES, int MIN_BLOCKS>
static int launch_soa(const Shard& s, const void* in0, const void* in1, void* out0, void* out1, const OpParams& prm) {
    constexpr int TILE = THREADS * IPT;
    constexpr int SMEM = STAGES * TILE * (Op::In0::bytes + Op::In1::bytes);
    static_assert(SMEM <= 227 * 1024, "tile does not fit in shared memory");
    auto kern = soa_tile_kernel<Op, THREADS, IPT, STAGES, MIN_BLOCKS>;
    static int grid_cap[T5_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!grid_cap[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return T5I_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, SMEM) != cudaSuccess) return T5I_CUDA;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        grid_cap[dev] = (occ < 1 ? 1 : occ) * (sms < 1 ? T5_SM_COUNT_FALLBACK : sms);
    }
    if (s.count == 0) return T5I_OK;
    if (!s.sched || s.soa_stride % 32) return T5I_CUDA;
    const size_t tiles = (s.count + TILE - 1) / TILE;
    const int grid = (int)(tiles < (size_t)grid_cap[dev] ? tiles : (size_t)grid_cap[dev]);
    const size_t claim = tile_claim(tiles, grid, (size_t)TILE * (Op::In0::bytes + Op::In1::bytes + Op::Out0::bytes + Op::Out1::bytes));
    SoaArgs a{in0, in1, out0, out1, (unsigned long long)s.count, (unsigned long long)s.soa_stride, s.sched, (unsigned)claim, prm};
    if (pdl_launch(kern, grid, THREADS, SMEM, s.stream, a) != cudaSuccess) { cudaGetLastError(); return T5I_CUDA; }
    return T5I_OK;
}

// Tile geometry {THREADS, records per thread, ring depth, min CTAs/SM}, picked per op from the sweep
// tools/tune_tile.cu -> profiles/r01_tune_tile_v3.csv (PDL launches, rotating operand sets).  With
// PDL the footprint matters: 3+ CTAs per SM let the next launch's CTAs move in while the previous
// launch's stragglers drain (1Mi 3x3 .*.: 41.9 us with one 147-KB CTA per SM, 36.9 us with three 74-KB
// ones).  The elimination ops are latency-bound on FP64 division chains and want more, smaller CTAs.
template <int T, int I, int S, int B = 1> struct G { static constexpr int threads = T, ipt = I, stages = S, min_blocks = B; };
template <class Op> struct Geo {
    static constexpr int in_bytes = Op::In0::bytes + Op::In1::bytes;
    using type = std::conditional_t<(in_bytes <= 32), G<256, 2, 3>,
                 std::conditional_t<(in_bytes <= 64), G<256, 1, 2>,
                 std::conditional_t<(in_bytes <= 160), G<128, 2, 2>, G<128, 2, 2>>>>;
};
template <class Op>
static int launch_op(const Shard& s, const void* in0, const void* in1, void* out0, void* out1, const OpParams& prm = OpParams{}) {
    using g = typename Geo<Op>::type;
    // (the one-stage geometry exists for the AoS pipeline only: SoA batches keep a two-stage ring)
    if (s.layout == T5L_SOA) return launch_soa<Op, g::threads, g::ipt, (g::stages < 2 ? 2 : g::stages), g::min_blocks>(s, in0, in1, out0, out1, prm);
    return launch_tile<Op, g::threads, g::ipt, g::stages, g::min_blocks>(s, in0, in1, out0, out1, prm);
}

}  // namespace t5
