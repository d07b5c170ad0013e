// Mid-size contractions (records of 0.25-4 KiB, e.g. the 8x8 . 8x8 `prod 1` of BASELINE
// configs[3]): too big for one thread per item, far too small for tensor-core tiles.
//
// One thread per ROW of the result: P threads share an item.  A tile of G = THREADS/P items
// is staged with the same cp.async ring + dynamic tile scheduler as t5_tile.cuh; a thread
// keeps its A row and its C row in registers and reads B row by row from shared memory
// (all P threads of an item read the same address: a broadcast, one wavefront per
// quarter-warp).  A rows are padded to an odd number of 16-B chunks and B items likewise,
// so neither the per-row nor the per-item accesses collide.  Results leave straight from
// registers (a C row is >= 16 contiguous bytes per thread).
//
// Arithmetic: acc = 0; for k ascending: acc = acc + a[k]*b[k][q] - the oracle's order;
// this TU is compiled with -fmad=false, so results are bit-identical to the oracle.
#include "t5_tile.cuh"
#include "t5_internal.h"

namespace t5 {
namespace mid {

struct Args {
    const void* A;
    const void* B;
    void* C;
    unsigned long long n_items;
    unsigned long long* sched;
    unsigned claim;
};

constexpr int up16(int x) { return (x + 15) / 16 * 16; }

template <class T, int P, int K, int Q, int THREADS, int STAGES>
struct Cfg {
    static constexpr int G = THREADS / P;                       // items per tile
    static constexpr int A_ROW = K * (int)sizeof(T);
    static constexpr int B_ROW = Q * (int)sizeof(T);
    // rows that are whole 16-byte chunks are copied chunk-wise and B items stay packed; other extents (odd n for
    // 8-byte elements, n % 4 != 0 for Float) are copied element-wise into rows padded to the next chunk
    static constexpr bool ALIGNED = A_ROW % 16 == 0 && B_ROW % 16 == 0;
    static constexpr int A_CPR = up16(A_ROW) / 16;              // chunks per (padded) A row
    static constexpr int A_RS = up16(A_ROW) + ((A_CPR % 2 == 0) ? 16 : 0);
    static constexpr int A_STAGE = G * P * A_RS;
    static constexpr int B_RS = ALIGNED ? B_ROW : up16(B_ROW);  // B row pitch inside an item
    static constexpr int B_BYTES = K * B_RS;
    static constexpr int B_CPI = B_BYTES / 16;
    static constexpr int B_ITEM = B_BYTES + ((B_CPI % 2 == 0) ? 16 : 0);
    static constexpr int B_STAGE = G * B_ITEM;
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int SMEM = STAGES * STAGE;
    static_assert(THREADS % 32 == 0 && G >= 1, "bad thread count");
};

template <int BYTES>
__device__ __forceinline__ void cp_async_unit(void* smem_dst, const void* gsrc) {
    if constexpr (BYTES == 16) cp_async16(smem_dst, gsrc);
    else cp_async_small<BYTES>(smem_dst, gsrc);
}

template <class T, int P, int K, int Q, int THREADS, int STAGES>
__global__ void __launch_bounds__(THREADS) rowprod_kernel(Args a) {
    using C = Cfg<T, P, K, Q, THREADS, STAGES>;
    using R = Arith<T>;
    constexpr int G = C::G;
    constexpr int VPR = 16 / (int)sizeof(T);   // elements per 16-B chunk
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long tile_q[STAGES + 1];

    const int tid = threadIdx.x;
    const unsigned long long n_tiles = (a.n_items + G - 1) / G;
    const unsigned char* gA = (const unsigned char*)a.A;
    const unsigned char* gB = (const unsigned char*)a.B;
    T* gC = (T*)a.C;

    auto issue = [&](unsigned long long tile, int stage) {
        if (tile < n_tiles) {
            const unsigned long long first = tile * G;
            const unsigned long long left = a.n_items - first;
            const int nv = left < (unsigned long long)G ? (int)left : G;
            unsigned char* sA = smem + stage * C::STAGE;
            unsigned char* sB = sA + C::A_STAGE;
            const unsigned char* srcA = gA + first * (size_t)(P * C::A_ROW);
            const unsigned char* srcB = gB + first * (size_t)(K * C::B_ROW);
            if constexpr (C::ALIGNED) {
                const int nchA = nv * P * C::A_CPR, nchB = nv * C::B_CPI;
                for (int ch = tid; ch < nchA; ch += THREADS)
                    cp_async16(sA + (ch / C::A_CPR) * C::A_RS + (ch % C::A_CPR) * 16, srcA + (size_t)ch * 16);
                for (int ch = tid; ch < nchB; ch += THREADS)
                    cp_async16(sB + (ch / C::B_CPI) * C::B_ITEM + (ch % C::B_CPI) * 16, srcB + (size_t)ch * 16);
            } else {
                // Rows that are not whole chunks: each operand moves in the largest unit its OWN row length allows - 16 bytes
                // (matrix . vector: A's rows may be whole chunks although a one-element B row never is), 8 bytes (Float rows
                // of an even length), else one element per asynchronous copy.
                // (Tried and dropped, profiles/r02r_sweep_matmul_vector_staging.txt: reading these contiguous runs as 16-byte
                // vectors into registers and scattering the elements to their padded places from there - the copy is no longer
                // asynchronous and every size lost: 9x9 Double 3.8 -> 3.2 TB/s, 19x19 2.8 -> 2.0, 16x16 . 16 Float 4.3 -> 3.6.)
                constexpr int ES = (int)sizeof(T);
                constexpr int AU = C::A_ROW % 16 == 0 ? 16 : (C::A_ROW % 8 == 0 ? 8 : ES), UA = C::A_ROW / AU;
                constexpr int BU = C::B_ROW % 16 == 0 ? 16 : (C::B_ROW % 8 == 0 ? 8 : ES), UB = C::B_ROW / BU;
                const int nuA = nv * P * UA, nuB = nv * K * UB;
                for (int u = tid; u < nuA; u += THREADS)
                    cp_async_unit<AU>(sA + (u / UA) * C::A_RS + (u % UA) * AU, srcA + (size_t)u * AU);
                for (int u = tid; u < nuB; u += THREADS) {
                    const int item = u / (K * UB), rem = u - item * (K * UB);
                    cp_async_unit<BU>(sB + item * C::B_ITEM + (rem / UB) * C::B_RS + (rem % UB) * BU, srcB + (size_t)u * BU);
                }
            }
        }
        cp_async_commit();
    };

    const int il = tid / P, r = tid % P;
    // Rows that are not whole chunks would leave as 8- or 4-byte stores one row pitch apart (ncu: 8 of every 32 bytes
    // of a sector used, half the L2 sectors excessive): those results are packed into the (consumed) A region of the
    // slot first and the CTA writes the tile as one contiguous run.  One-stage kernels only (the slot is not a ring).
    constexpr bool PACK_OUT = !C::ALIGNED && STAGES == 1 && Q <= K;
    auto compute = [&](unsigned long long tile, int stage) {
        const unsigned long long first = tile * G;
        const unsigned long long left = a.n_items - first;
        const int nv = left < (unsigned long long)G ? (int)left : G;
        constexpr int KP = C::A_CPR * VPR, QC = (Q + VPR - 1) / VPR, QP = QC * VPR;   // extents rounded up to whole chunks
        alignas(16) T acc[QP];
        const bool active = il < nv && il < G;          // (threads beyond G * P idle when P does not divide the CTA)
        if (active) {
            const unsigned char* sA = smem + stage * C::STAGE + (il * P + r) * C::A_RS;
            const unsigned char* sB = smem + stage * C::STAGE + C::A_STAGE + il * C::B_ITEM;
            alignas(16) T av[KP];
            #pragma unroll
            for (int c = 0; c < C::A_CPR; ++c) reinterpret_cast<uint4*>(av)[c] = reinterpret_cast<const uint4*>(sA)[c];
            #pragma unroll
            for (int q = 0; q < Q; ++q) acc[q] = R::zero();
            #pragma unroll
            for (int k = 0; k < K; ++k) {
                alignas(16) T bv[QP];            // (the padding lanes of a partial last chunk are never read)
                #pragma unroll
                for (int c = 0; c < QC; ++c) reinterpret_cast<uint4*>(bv)[c] = reinterpret_cast<const uint4*>(sB + k * C::B_RS)[c];
                #pragma unroll
                for (int q = 0; q < Q; ++q) acc[q] = R::add(acc[q], R::mul(av[k], bv[q]));
            }
            if constexpr (!PACK_OUT) {
                T* dst = gC + ((first + il) * P + r) * (size_t)Q;
                if constexpr (C::ALIGNED) {
                    #pragma unroll
                    for (int c = 0; c < Q / VPR; ++c) st_global_stream16(dst + c * VPR, reinterpret_cast<const uint4*>(acc)[c]);
                } else {
                    #pragma unroll
                    for (int q = 0; q < Q; ++q) dst[q] = acc[q];
                }
            }
        }
        if constexpr (PACK_OUT) {
            T* sC = reinterpret_cast<T*>(smem + stage * C::STAGE);          // G * P * Q elements <= the A region (Q <= K)
            __syncthreads();                                                 // every A row is in registers
            if (active) {
                #pragma unroll
                for (int q = 0; q < Q; ++q) sC[(il * P + r) * Q + q] = acc[q];
            }
            __syncthreads();
            T* dst = gC + first * (size_t)(P * Q);
            const int total = nv * P * Q;
            #pragma unroll 4
            for (int e = tid; e < total; e += THREADS) dst[e] = sC[e];
        }
    };

    pdl_trigger();
    pdl_wait();
    TileClaimer claimer;
    if constexpr (STAGES == 1) {
        // one slot per CTA: residency is bounded by threads instead of by the ring (16x16 Double: 3 -> 6 CTAs per SM);
        // the other resident CTAs cover this one's copy, the next claim travels while the tile lands
        if (tid == 0) tile_q[0] = claimer.next(a.sched, a.claim);
        __syncthreads();
        int q = 0;
        while (true) {
            const unsigned long long tile = tile_q[q];
            if (tile >= n_tiles) break;
            issue(tile, 0);
            if (tid == 0) tile_q[q ^ 1] = claimer.next(a.sched, a.claim);
            cp_async_wait<0>();
            __syncthreads();
            compute(tile, 0);
            __syncthreads();
            q ^= 1;
        }
    } else {
        unsigned long long next_claim = 0;
        if (tid == 0) {
            #pragma unroll
            for (int s = 0; s < STAGES - 1; ++s) tile_q[s] = claimer.next(a.sched, a.claim);
            next_claim = claimer.next(a.sched, a.claim);
        }
        __syncthreads();
        #pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) issue(tile_q[s], s);

        int stage = 0;
        while (true) {
            const unsigned long long tile = tile_q[stage];
            if (tile >= n_tiles) break;
            int pf = stage + STAGES - 1; if (pf >= STAGES) pf -= STAGES;
            if (tid == 0) tile_q[pf] = next_claim;
            cp_async_wait<(STAGES >= 2 ? STAGES - 2 : 0)>();
            __syncthreads();
            issue(tile_q[pf], pf);
            if (tid == 0) next_claim = claimer.next(a.sched, a.claim);
            compute(tile, stage);
            if (++stage == STAGES) stage = 0;
        }
    }
    cp_async_wait<0>();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.sched + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
            a.sched[0] = 0;
            a.sched[1] = 0;
            __threadfence();
        }
    }
}

template <class T, int P, int K, int Q, int THREADS, int STAGES>
static int launch(const Shard& s, const void* A, const void* B, void* Cc) {
    using C = Cfg<T, P, K, Q, THREADS, STAGES>;
    static_assert(C::SMEM <= 227 * 1024, "ring does not fit");
    auto kern = rowprod_kernel<T, P, K, Q, THREADS, STAGES>;
    static int grid_cap[T5_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!grid_cap[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) return T5I_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, C::SMEM) != cudaSuccess) return T5I_CUDA;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        grid_cap[dev] = (occ < 1 ? 1 : occ) * (sms < 1 ? T5_SM_COUNT_FALLBACK : sms);
    }
    if (s.count == 0) return T5I_OK;
    if (!s.sched) return T5I_CUDA;
    size_t tiles = (s.count + C::G - 1) / C::G;
    int grid = (int)(tiles < (size_t)grid_cap[dev] ? tiles : (size_t)grid_cap[dev]);
    size_t claim = tiles / ((size_t)grid * 16);
    claim = claim < 1 ? 1 : (claim > 8 ? 8 : claim);
    Args a{A, B, Cc, (unsigned long long)s.count, s.sched, (unsigned)claim};
    if (pdl_launch(kern, grid, THREADS, C::SMEM, s.stream, a) != cudaSuccess) { cudaGetLastError(); return T5I_CUDA; }
    return T5I_OK;
}

template <class T>
static int dispatch(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
    // geometry from tools/tune_mid.cu (profiles/r01_tune_mid.csv): small CTAs, many per SM
    // 8x8: 256 threads x 2 stages = 3 CTAs (768 threads) per SM.  64 x 4 is as fast in isolation but, with
    // only 8 warps per SM on the FP64 chains, lost 25% inside a long bench run (clock-sensitive).
    if (m == 8 && k == 8 && n == 8) return launch<T, 8, 8, 8, 256, 2>(s, A, B, C);
    if (m == 16 && k == 16 && n == 16) return launch<T, 16, 16, 16, 128, 1>(s, A, B, C);
    if (m == 12 && k == 12 && n == 12) return launch<T, 12, 12, 12, 192, 1>(s, A, B, C);
    if (m == 20 && k == 20 && n == 20) return launch<T, 20, 20, 20, 160, 1>(s, A, B, C);
    if (m == 24 && k == 24 && n == 24) return launch<T, 24, 24, 24, 192, 1>(s, A, B, C);
    if constexpr (sizeof(T) == 4) {   // a warp per matrix; 8-byte 32x32 measured slower than the blocked product (FP64 / IMAD bound)
        if (m == 32 && k == 32 && n == 32) return launch<T, 32, 32, 32, 128, 1>(s, A, B, C);
    }
    if (m == 48 && k == 48 && n == 48) return launch<T, 48, 48, 48, 96, 1>(s, A, B, C);
    // matrix . vector beyond the tile pipeline's n <= 8: a thread per row, the vector broadcast from shared memory
    if (n == 1 && m == k) {
        if (m == 9) return launch<T, 9, 9, 1, 256, 1>(s, A, B, C);
        if (m == 10) return launch<T, 10, 10, 1, 256, 1>(s, A, B, C);
        if (m == 12) return launch<T, 12, 12, 1, 256, 1>(s, A, B, C);
        if (m == 16) return launch<T, 16, 16, 1, 256, 1>(s, A, B, C);
        if (m == 20) return launch<T, 20, 20, 1, 256, 1>(s, A, B, C);
        if (m == 24) return launch<T, 24, 24, 1, 256, 1>(s, A, B, C);
        if (m == 32) return launch<T, 32, 32, 1, 256, 1>(s, A, B, C);
    }
    // odd extents (element-wise staging into padded rows): one-stage, CTAs of whole warps, a few idle threads
    if constexpr (sizeof(T) == 8) {
        if (m == 7 && k == 7 && n == 7) return launch<T, 7, 7, 7, 256, 1>(s, A, B, C);
        if (m == 9 && k == 9 && n == 9) return launch<T, 9, 9, 9, 256, 1>(s, A, B, C);
        if (m == 11 && k == 11 && n == 11) return launch<T, 11, 11, 11, 256, 1>(s, A, B, C);
        if (m == 13 && k == 13 && n == 13) return launch<T, 13, 13, 13, 256, 1>(s, A, B, C);
        if (m == 15 && k == 15 && n == 15) return launch<T, 15, 15, 15, 256, 1>(s, A, B, C);
        if (m == 17 && k == 17 && n == 17) return launch<T, 17, 17, 17, 256, 1>(s, A, B, C);
        if (m == 19 && k == 19 && n == 19) return launch<T, 19, 19, 19, 256, 1>(s, A, B, C);
    } else {
        if (m == 9 && k == 9 && n == 9) return launch<T, 9, 9, 9, 256, 1>(s, A, B, C);
        if (m == 10 && k == 10 && n == 10) return launch<T, 10, 10, 10, 256, 1>(s, A, B, C);
        if (m == 11 && k == 11 && n == 11) return launch<T, 11, 11, 11, 256, 1>(s, A, B, C);
        if (m == 13 && k == 13 && n == 13) return launch<T, 13, 13, 13, 256, 1>(s, A, B, C);
        if (m == 14 && k == 14 && n == 14) return launch<T, 14, 14, 14, 256, 1>(s, A, B, C);
        if (m == 15 && k == 15 && n == 15) return launch<T, 15, 15, 15, 256, 1>(s, A, B, C);
        if (m == 17 && k == 17 && n == 17) return launch<T, 17, 17, 17, 256, 1>(s, A, B, C);
        if (m == 18 && k == 18 && n == 18) return launch<T, 18, 18, 18, 256, 1>(s, A, B, C);
        if (m == 19 && k == 19 && n == 19) return launch<T, 19, 19, 19, 256, 1>(s, A, B, C);
    }
    if constexpr (sizeof(T) == 8) {   // rows must be whole 16-byte chunks
        if (m == 14 && k == 14 && n == 14) return launch<T, 14, 14, 14, 224, 1>(s, A, B, C);
        if (m == 18 && k == 18 && n == 18) return launch<T, 18, 18, 18, 288, 1>(s, A, B, C);
        if (m == 6 && k == 6 && n == 6) return launch<T, 6, 6, 6, 192, 3>(s, A, B, C);
        if (m == 10 && k == 10 && n == 10) return launch<T, 10, 10, 10, 160, 1>(s, A, B, C);
    }
    return T5I_NOT_SPECIALISED;
}

}  // namespace mid

int mid_matmul(const Shard& s, int dtype, int m, int k, int n, const void* A, const void* B, void* C) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    switch (dtype) {
        case T5D_F64: return mid::dispatch<double>(s, m, k, n, A, B, C);
        case T5D_F32: return mid::dispatch<float>(s, m, k, n, A, B, C);
        case T5D_I64: return mid::dispatch<long long>(s, m, k, n, A, B, C);
    }
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
