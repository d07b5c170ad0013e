// AoS tile pipeline: the skeleton every thread-per-item kernel runs in.
//
// A batch is `n_items` records per stream, records concatenated (the reference's
// wire order, src/Data/Tensor/Vector.hs:360-367).  Records are 12..128 bytes, so
// a thread reading "its" record straight from HBM would issue 32 scattered
// sectors per warp instruction.  Instead a persistent CTA walks tiles of
// TILE = THREADS*IPT records:
//
//   global --cp.async 16 B (LDGSTS, coalesced, STAGES-deep ring)--> shared
//   shared --LDS (record-per-thread, conflict-free by construction)--> registers
//   registers: Op::run  (pure register arithmetic, shared with the SoA skeleton)
//   registers --STS--> shared out-stage --LDS.128 + STG.128 (coalesced, streaming)--> global
//
// Conflict-free record access: a record of B bytes is stored with stride
// B (+16 when B/16 is even).  Then B/16 odd -> LDS.128 hits 8 distinct 16-B bank
// groups per quarter-warp; B = 8*odd -> LDS.64 hits 16 distinct bank pairs per
// half-warp; B = 4*odd -> LDS.32 hits 32 distinct banks.
//
// Two __syncthreads per tile; the ring keeps (STAGES-1) tiles of loads in flight
// per CTA, and several CTAs are resident per SM, which is what covers the HBM
// latency (bytes in flight per SM = CTAs/SM * (STAGES-1) * tile bytes).
#pragma once
#include "t5_device.cuh"

namespace t5 {

template <class T, int N>
struct Rec {
    using type = T;
    static constexpr int n = N;
    static constexpr int bytes = N * (int)sizeof(T);
    static constexpr int pad = (bytes > 0 && bytes % 16 == 0 && ((bytes / 16) % 2 == 0)) ? 16 : 0;
    static constexpr int stride = bytes + pad;
};
using NoRec = Rec<unsigned char, 0>;

template <class R>
struct RegArr {
    alignas(16) typename R::type v[R::n > 0 ? R::n : 1];
};

struct TileArgs {
    const void* in0;
    const void* in1;
    void* out0;
    void* out1;
    unsigned long long n_items;
};

// ---- cooperative tile movement ------------------------------------------------
template <class R, int TILE, int THREADS>
__device__ __forceinline__ void tile_load_async(unsigned char* stage, const unsigned char* gsrc, int n_valid, int tid) {
    if constexpr (R::bytes > 0) {
        constexpr int NCH = TILE * R::bytes / 16;  // TILE % 32 == 0 makes this exact
        constexpr int CPI = R::bytes / 16;         // only used when padded (bytes % 16 == 0)
        if (n_valid == TILE) {
            #pragma unroll
            for (int i = 0; i < (NCH + THREADS - 1) / THREADS; ++i) {
                int ch = tid + i * THREADS;
                if ((NCH % THREADS == 0) || ch < NCH) {
                    unsigned char* d = R::pad ? stage + (ch / (CPI > 0 ? CPI : 1)) * R::stride + (ch % (CPI > 0 ? CPI : 1)) * 16
                                              : stage + ch * 16;
                    cp_async16(d, gsrc + (size_t)ch * 16);
                }
            }
        } else {
            const int vbytes = n_valid * R::bytes;
            const int nch = (vbytes + 15) / 16;
            for (int ch = tid; ch < nch; ch += THREADS) {
                unsigned char* d = R::pad ? stage + (ch / (CPI > 0 ? CPI : 1)) * R::stride + (ch % (CPI > 0 ? CPI : 1)) * 16
                                          : stage + ch * 16;
                int rem = vbytes - ch * 16;
                cp_async16(d, gsrc + (size_t)ch * 16, rem < 16 ? rem : 16);
            }
        }
    }
}

template <class R, int TILE, int THREADS>
__device__ __forceinline__ void tile_store(const unsigned char* stage, unsigned char* gdst, int n_valid, int tid) {
    if constexpr (R::bytes > 0) {
        constexpr int NCH = TILE * R::bytes / 16;
        constexpr int CPI = R::bytes / 16;
        if (n_valid == TILE) {
            #pragma unroll
            for (int i = 0; i < (NCH + THREADS - 1) / THREADS; ++i) {
                int ch = tid + i * THREADS;
                if ((NCH % THREADS == 0) || ch < NCH) {
                    const unsigned char* s = R::pad ? stage + (ch / (CPI > 0 ? CPI : 1)) * R::stride + (ch % (CPI > 0 ? CPI : 1)) * 16
                                                    : stage + ch * 16;
                    st_global_stream16(gdst + (size_t)ch * 16, *reinterpret_cast<const uint4*>(s));
                }
            }
        } else {
            const int vbytes = n_valid * R::bytes;
            const int nfull = vbytes / 16;
            for (int ch = tid; ch < nfull; ch += THREADS) {
                const unsigned char* s = R::pad ? stage + (ch / (CPI > 0 ? CPI : 1)) * R::stride + (ch % (CPI > 0 ? CPI : 1)) * 16
                                                : stage + ch * 16;
                st_global_stream16(gdst + (size_t)ch * 16, *reinterpret_cast<const uint4*>(s));
            }
            // ragged last chunk (only possible when the record size is not a multiple of 16)
            for (int b = nfull * 16 + tid; b < vbytes; b += THREADS) gdst[b] = stage[b];
        }
    }
}

// ---- the kernel ------------------------------------------------------------------
// Op provides: In0, In1, Out0, Out1 (Rec types; NoRec when absent) and
//   __device__ static void run(const In0::type*, const In1::type*, Out0::type*, Out1::type*)
// operating on register arrays.
template <class Op, int THREADS, int IPT, int STAGES>
struct TileCfg {
    static constexpr int TILE = THREADS * IPT;
    static constexpr int IN_STAGE = TILE * (Op::In0::stride + Op::In1::stride);
    static constexpr int OUT_STAGE = TILE * (Op::Out0::stride + Op::Out1::stride);
    static constexpr int SMEM = STAGES * IN_STAGE + OUT_STAGE;
};

template <class Op, int THREADS, int IPT, int STAGES>
__global__ void __launch_bounds__(THREADS) tile_kernel(TileArgs a) {
    using C = TileCfg<Op, THREADS, IPT, STAGES>;
    using I0 = typename Op::In0; using I1 = typename Op::In1;
    using O0 = typename Op::Out0; using O1 = typename Op::Out1;
    constexpr int TILE = C::TILE;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* out_stage0 = smem + STAGES * C::IN_STAGE;
    unsigned char* out_stage1 = out_stage0 + TILE * O0::stride;

    const int tid = threadIdx.x;
    const unsigned long long n_tiles = (a.n_items + TILE - 1) / TILE;
    const unsigned char* g_in0 = (const unsigned char*)a.in0;
    const unsigned char* g_in1 = (const unsigned char*)a.in1;
    unsigned char* g_out0 = (unsigned char*)a.out0;
    unsigned char* g_out1 = (unsigned char*)a.out1;

    auto issue = [&](unsigned long long tile, int stage) {
        if (tile < n_tiles) {
            unsigned long long first = tile * TILE;
            unsigned long long left = a.n_items - first;
            int nv = left < (unsigned long long)TILE ? (int)left : TILE;
            unsigned char* s0 = smem + stage * C::IN_STAGE;
            tile_load_async<I0, TILE, THREADS>(s0, g_in0 + first * I0::bytes, nv, tid);
            tile_load_async<I1, TILE, THREADS>(s0 + TILE * I0::stride, g_in1 + first * I1::bytes, nv, tid);
        }
        cp_async_commit();  // always commit: keeps the group count uniform
    };

    unsigned long long tile = blockIdx.x;
    #pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(tile + (unsigned long long)s * gridDim.x, s);

    int stage = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
        int pf = stage + STAGES - 1; if (pf >= STAGES) pf -= STAGES;
        issue(tile + (unsigned long long)(STAGES - 1) * gridDim.x, pf);
        cp_async_wait<STAGES - 1>();
        __syncthreads();  // (A) tile landed for every thread; previous copy-out finished

        unsigned long long first = tile * TILE;
        unsigned long long left = a.n_items - first;
        const int nv = left < (unsigned long long)TILE ? (int)left : TILE;
        const unsigned char* s0 = smem + stage * C::IN_STAGE;
        const unsigned char* s1 = s0 + TILE * I0::stride;
        #pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * THREADS;
            if (it < nv) {
                RegArr<I0> r0; RegArr<I1> r1; RegArr<O0> w0; RegArr<O1> w1;
                if constexpr (I0::bytes > 0) lds_item<I0::bytes>(s0 + it * I0::stride, r0.v);
                if constexpr (I1::bytes > 0) lds_item<I1::bytes>(s1 + it * I1::stride, r1.v);
                Op::run(r0.v, r1.v, w0.v, w1.v);
                if constexpr (O0::bytes > 0) sts_item<O0::bytes>(out_stage0 + it * O0::stride, w0.v);
                if constexpr (O1::bytes > 0) sts_item<O1::bytes>(out_stage1 + it * O1::stride, w1.v);
            }
        }
        __syncthreads();  // (B) out-stage complete; in-stage free for the next prefetch
        tile_store<O0, TILE, THREADS>(out_stage0, g_out0 + first * O0::bytes, nv, tid);
        tile_store<O1, TILE, THREADS>(out_stage1, g_out1 + first * O1::bytes, nv, tid);
        if (++stage == STAGES) stage = 0;
    }
    cp_async_wait<0>();
}

// ---- SoA skeleton: element e of item b at base[e * stride + b]; perfectly
// coalesced thread-per-item access, no staging needed. -----------------------------
struct SoaArgs {
    const void* in0;
    const void* in1;
    void* out0;
    void* out1;
    unsigned long long n_items;
    unsigned long long stride;  // shard capacity in items
};

template <class Op, int THREADS>
__global__ void __launch_bounds__(THREADS) soa_kernel(SoaArgs a) {
    using I0 = typename Op::In0; using I1 = typename Op::In1;
    using O0 = typename Op::Out0; using O1 = typename Op::Out1;
    const typename I0::type* in0 = (const typename I0::type*)a.in0;
    const typename I1::type* in1 = (const typename I1::type*)a.in1;
    typename O0::type* out0 = (typename O0::type*)a.out0;
    typename O1::type* out1 = (typename O1::type*)a.out1;
    for (unsigned long long b = (unsigned long long)blockIdx.x * THREADS + threadIdx.x; b < a.n_items;
         b += (unsigned long long)gridDim.x * THREADS) {
        RegArr<I0> r0; RegArr<I1> r1; RegArr<O0> w0; RegArr<O1> w1;
        #pragma unroll
        for (int e = 0; e < I0::n; ++e) r0.v[e] = __ldg(in0 + (size_t)e * a.stride + b);
        #pragma unroll
        for (int e = 0; e < I1::n; ++e) r1.v[e] = __ldg(in1 + (size_t)e * a.stride + b);
        Op::run(r0.v, r1.v, w0.v, w1.v);
        #pragma unroll
        for (int e = 0; e < O0::n; ++e) out0[(size_t)e * a.stride + b] = w0.v[e];
        #pragma unroll
        for (int e = 0; e < O1::n; ++e) out1[(size_t)e * a.stride + b] = w1.v[e];
    }
}

}  // namespace t5
