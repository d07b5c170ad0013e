// Batched Float `.*.` for large matrices on the 5th-generation tensor cores: tcgen05.mma kind::tf32
// with the operand split a = a_hi + a_lo ("3xTF32"), accumulators in TMEM.
//
// A single TF32 pass keeps 10 mantissa bits (~1e-3): not acceptable.  With
//      A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi        (a_hi = a truncated to TF32, a_lo = a - a_hi, exact)
// the dropped term is ~2^-20 |a||b|, well inside the |delta| <= 1e-5 * sum_j |a_ij||b_jk| bar that
// north_star states for Float (tests hold this kernel to it; shapes off the 128/32 tile grid, and
// Int64, stay on the bit-exact SIMT product of t5_xgemm.cu).  128x128x128 items are then HBM-bound
// (21 flop/B) instead of FP32-issue-bound.
//
//   warp 0 (1 lane)   TMA producer: A chunk 128 x 32 (K-major, SWIZZLE_128B), B chunk 32 x 128
//                     (MN-major: [4 n-blocks][32 k-rows][32 floats], SWIZZLE_128B_ATOM_32B - the
//                     only layout tcgen05 accepts for MN-major 32-bit operands; one box per
//                     128-byte column block, so n need not be a multiple of 32)       --full[s]-->
//   warps 4-7         split: x_lo = x - trunc_tf32(x) for both chunks into twin tiles of the same
//                     (swizzled) layout; fence.proxy.async                            --conv[s]-->
//   warp 1 (1 lane)   tcgen05.mma cta_group::1 kind::tf32, M=128 N=128 K=8: per K step
//                     (A_lo,B) (A,B_lo) (A,B) into one of two 128-column TMEM accumulators;
//                     tcgen05.commit frees the stage (--empty[s]-->) and, after the last
//                     chunk, hands the accumulator over                               --tmem_full[a]-->
//   warps 8-11        epilogue: tcgen05.ld 32 lanes x 32 columns, rows straight to HBM --tmem_empty[a]-->
//   warp 2            TMEM allocation (256 columns) and release
//
// The hardware reads the FP32 bit patterns of the "hi" tiles and ignores the 13 low mantissa bits,
// so the raw operands serve as a_hi / b_hi and only the lo tiles are computed.
#include <atomic>

#include "t5_async.cuh"
#include "t5_device.cuh"
#include "t5_internal.h"

namespace t5 {
namespace tf {

constexpr int BK = 32;
constexpr int THREADS = 384;
constexpr int EPI_PITCH = 36 * 4;                                   // staging row pitch: 32 floats + 16 B (conflict-free v4 access)
constexpr int EPI_WARP = 32 * EPI_PITCH;                            // one (<= 32) x 32 block per epilogue warp

// Square tiles of 128 (3 stages of 64 KB) or 64 (6 stages of 16 KB) - the UMMA shapes M = N = 128 / 64.
// ITEMS = 4 (round 2): items of at most 32 rows and exactly 32 columns, FOUR to a 128 x 128 tile.  The tensor maps carry
// the item as a dimension, so one box {32 k, 32 rows, 4 items} lands as the A image of a 128-row tile (item q = rows
// 32q..32q+31) and one box {32, 32 k-rows, 1 block, 4 items} as the B image of a 128-column tile (item q = column block
// q): the MMA stream is the 128 tile's, and only the four diagonal 32 x 32 blocks of the accumulator are read back.
// Per byte of HBM traffic that is exactly the tensor / split / epilogue work of a 128^3 item (one K chunk per 48 KB).
template <int TILE_, int STAGES_, int ITEMS_ = 1>
struct Geo {
    static constexpr int BM = TILE_, BN = TILE_, STAGES = STAGES_, ITEMS = ITEMS_;
    static_assert(ITEMS == 1 || ((ITEMS == 4 || ITEMS == 2) && TILE_ == 128), "item packing: four 32-row or two 64-row items on the 128 tile");
    static constexpr int ITEM_ROWS = BM / ITEMS;                     // rows (and columns) of the tile one packed item owns
    static constexpr int ITEM_BLOCKS = BN / 32 / ITEMS;              // 128-byte column blocks per packed item
    static constexpr int TILE_A = BM * BK * 4, TILE_B = BK * BN * 4;
    static constexpr int STAGE = 2 * (TILE_A + TILE_B);             // hi + lo tiles
    static constexpr int SMEM = STAGES * STAGE + 1024 + 4 * EPI_WARP;
    static constexpr int TMEM_COLS = 2 * BN;                        // two FP32 accumulators (power of two >= 32)
    static constexpr int ROWS_PER_WARP = BM / 4;                    // M = 64 keeps 16 rows in lanes 0-15 of every 32-lane quarter
    // Instruction descriptor: D = F32, A = B = TF32, A K-major, B MN-major, N, M.
    static constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) |
                                      ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
    static_assert(SMEM <= 227 * 1024, "ring does not fit");
};
using Geo128 = Geo<128, 3>;
using Geo64 = Geo<64, 6>;
using Geo32x4 = Geo<128, 3, 4>;
using Geo64x2 = Geo<128, 3, 2>;                                       // two items of at most 64 x 64 per tile (rows 64q.., column blocks 2q, 2q + 1)

constexpr float FLT_MAX_F = 3.4028234664e38f;

struct Args {
    const float* A;
    const float* B;
    float* C;
    int M, N, K;
    unsigned long long count;
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b,
                                            unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// Shared-memory matrix descriptor (SM100 UMMA): start address, leading / stride byte offsets in
// 16-byte units, version 1, SWIZZLE_128B.
// layout: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks; the only layout the
// hardware accepts for MN-major 32-bit operands).
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, unsigned lbo16, unsigned sbo16, unsigned layout) {
    return (unsigned long long)((smem_addr >> 4) & 0x3fffu) | ((unsigned long long)(lbo16 & 0x3fffu) << 16) |
           ((unsigned long long)(sbo16 & 0x3fffu) << 32) | (1ull << 46) | ((unsigned long long)layout << 61);
}

__device__ __forceinline__ float lo_part(float x) {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const float lo = x - hi;                                   // exact
    return (fabsf(x) <= 3.4028234664e38f) ? lo : 0.0f;         // inf / nan: such tiles are recomputed exactly (see special_tile)
}

template <class G>
__global__ void __launch_bounds__(THREADS, 1)
tf32x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Args p) {
    constexpr int BM = G::BM, BN = G::BN, STAGES = G::STAGES, TILE_A = G::TILE_A, TILE_B = G::TILE_B, STAGE = G::STAGE;
    constexpr int TMEM_COLS = G::TMEM_COLS, RPW = G::ROWS_PER_WARP, ITEMS = G::ITEMS;
    constexpr unsigned IDESC = G::IDESC;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[3 * STAGES + 4];   // full, conv, empty, tmem_full[2], tmem_empty[2]
    __shared__ unsigned tmem_base_smem;
    // Non-finite operands: inf * (lo part) would turn the reference's inf into NaN.  The split warps
    // record the sequence number of any tile that contains one; its epilogue recomputes that tile
    // with the sequential exact fold instead of reading TMEM (rare path, same results as the oracle).
    __shared__ unsigned special_tile[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned bar0 = smem_u32(bars);
    auto full_bar = [&](unsigned s) { return bar0 + 8 * s; };
    auto conv_bar = [&](unsigned s) { return bar0 + 8 * (STAGES + s); };
    auto empty_bar = [&](unsigned s) { return bar0 + 8 * (2 * STAGES + s); };
    auto tfull_bar = [&](unsigned a) { return bar0 + 8 * (3 * STAGES + a); };
    auto tempty_bar = [&](unsigned a) { return bar0 + 8 * (3 * STAGES + 2 + a); };

    // ragged shapes: the tensor maps carry the item as their own (outermost) dimension, so the part of a box that lies
    // beyond an item's M rows, N columns or K depth is out of bounds for TMA and arrives as ZEROS (never the next item's
    // data); zero rows / columns / k-slices add nothing, and the epilogue stores only what exists
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN, kch = (p.K + BK - 1) / BK;
    const unsigned long long tiles_per_item = (unsigned long long)tiles_m * tiles_n;      // 1 when ITEMS > 1 (host checks)
    const unsigned long long n_tiles = ITEMS > 1 ? (p.count + ITEMS - 1) / ITEMS : p.count * tiles_per_item;

    if (tid < 8) special_tile[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(conv_bar(s), 4);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        mbar_init_fence();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     ::"r"(smem_u32(&tmem_base_smem)), "r"((unsigned)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ---- TMA producer ----
        if (lane == 0) {
            unsigned it = 0;
            for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const unsigned long long item = ITEMS > 1 ? tile * ITEMS : tile / tiles_per_item;   // (items past the batch: zeros)
                const int rem = ITEMS > 1 ? 0 : (int)(tile - item * tiles_per_item);
                const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
                for (int kc = 0; kc < kch; ++kc, ++it) {
                    const unsigned s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), TILE_A + TILE_B);                 // out-of-bounds parts are zero-filled and counted
                    tma_load_3d(ring + s * STAGE, &mapA, full_bar(s), kc * BK, tm * BM, (int)item);
                    // B: one box of 32 columns x 32 k-rows per 128-byte column block (columns / items beyond the operand: zeros)
                    #pragma unroll
                    for (int blk = 0; blk < BN / 32; ++blk)
                        tma_load_3d(ring + s * STAGE + TILE_A + blk * 4096, &mapB, full_bar(s),
                                    ITEMS > 1 ? (blk % G::ITEM_BLOCKS) * 32 : tn * BN + blk * 32, kc * BK,
                                    ITEMS > 1 ? (int)item + blk / G::ITEM_BLOCKS : (int)item);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        if (lane == 0) {
            unsigned it = 0, tl = 0;
            for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
                const unsigned acc = tl & 1;
                mbar_wait(tempty_bar(acc), ((tl >> 1) & 1) ^ 1);            // epilogue has drained this accumulator
                tc_fence_after();
                const unsigned tmem_d = tmem_base + acc * BN;
                for (int kc = 0; kc < kch; ++kc, ++it) {
                    const unsigned s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(conv_bar(s), ph);                             // operands and their lo parts are in place
                    tc_fence_after();
                    const unsigned a_hi = ring + s * STAGE, b_hi = a_hi + TILE_A;
                    const unsigned a_lo = b_hi + TILE_B, b_lo = a_lo + TILE_A;
                    #pragma unroll
                    for (int k8 = 0; k8 < BK / 8; ++k8) {
                        if (k8 && kc * BK + k8 * 8 >= p.K) break;           // K tail: nothing but zeros beyond
                        // A (K-major): 32 bytes further along the 128-byte swizzled rows; 8-row groups 1024 B apart
                        const unsigned long long da_hi = umma_desc(a_hi + k8 * 32, 1, 64, 2);
                        const unsigned long long da_lo = umma_desc(a_lo + k8 * 32, 1, 64, 2);
                        // B (MN-major, 32-byte swizzle atoms of 4 k-rows x 128 B): 8 k-rows = 1024 B further per step,
                        // 4-row atoms 512 B apart (SBO), 32-column blocks 4096 B apart (LBO)
                        const unsigned long long db_hi = umma_desc(b_hi + k8 * 1024, 256, 32, 1);
                        const unsigned long long db_lo = umma_desc(b_lo + k8 * 1024, 256, 32, 1);
                        tc_mma_tf32(tmem_d, da_lo, db_hi, IDESC, (kc | k8) ? 1u : 0u);   // small terms first
                        tc_mma_tf32(tmem_d, da_hi, db_lo, IDESC, 1u);
                        tc_mma_tf32(tmem_d, da_hi, db_hi, IDESC, 1u);
                    }
                    tc_commit(empty_bar(s));                                // stage free once these MMAs have read it
                }
                tc_commit(tfull_bar(acc));                                  // accumulator complete
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ---- split warps: lo tiles ----
        const int t = tid - 128;
        unsigned it = 0, tl = 0;
        for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            for (int kc = 0; kc < kch; ++kc, ++it) {
                const unsigned s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                const unsigned hi = ring + s * STAGE, lo = hi + TILE_A + TILE_B;
                bool special = false;
                #pragma unroll 4
                for (int c = 0; c < (TILE_A + TILE_B) / 16 / 128; ++c) {
                    const unsigned off = (unsigned)(t + c * 128) * 16u;
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(hi + off));
                    special = special || !(fabsf(v.x) <= FLT_MAX_F && fabsf(v.y) <= FLT_MAX_F && fabsf(v.z) <= FLT_MAX_F && fabsf(v.w) <= FLT_MAX_F);
                    v.x = lo_part(v.x); v.y = lo_part(v.y); v.z = lo_part(v.z); v.w = lo_part(v.w);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(lo + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
                if (special) *(volatile unsigned*)&special_tile[tl & 7] = tl + 1;
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(conv_bar(s));
            }
        }
    } else if (warp >= 8) {
        // ---- epilogue: TMEM -> registers -> HBM; warp q owns TMEM lanes 32q .. 32q+31 = rows of the tile ----
        const int q = warp & 3;
        const unsigned stage_w = ring + STAGES * STAGE + q * EPI_WARP;
        unsigned tl = 0;
        for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            const unsigned acc = tl & 1;
            constexpr int WPI = 4 / ITEMS;                                    // epilogue warps per packed item
            const int qi = q / WPI, qr = q % WPI;                             // packed: item of the tile, 32-row block within the item
            const unsigned long long item = ITEMS > 1 ? tile * ITEMS + qi : tile / tiles_per_item;
            const int rem = ITEMS > 1 ? 0 : (int)(tile - item * tiles_per_item);
            const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
            mbar_wait(tfull_bar(acc), (tl >> 1) & 1);
            tc_fence_after();
            const int row_base = ITEMS > 1 ? qr * 32 : tm * BM + q * RPW;     // first row (within the item) of this warp's block
            float* cblk = p.C + item * (size_t)p.M * p.N + (size_t)row_base * p.N + tn * BN;   // this warp's rows
            float* crow = cblk + (size_t)lane * p.N;
            const int rows_here = (ITEMS > 1 && item >= p.count) ? 0 : p.M - row_base;   // rows of this warp's block that exist (may be <= 0)
            const int cols_here = p.N - tn * BN;                              // columns of the tile that exist (multiple of 4)
            if (*(volatile unsigned*)&special_tile[tl & 7] == tl + 1) {
                // exact sequential fold for this thread's row (multiply, round, add, round; k ascending)
                const float* arow = p.A + item * (size_t)p.M * p.K + (size_t)(row_base + lane) * p.K;
                const float* bcol = p.B + item * (size_t)p.K * p.N + tn * BN;
                for (int j = 0; j < BN && j < cols_here && lane < RPW && lane < rows_here; ++j) {
                    float acc_x = 0.0f;
                    for (int kk = 0; kk < p.K; ++kk) acc_x = __fadd_rn(acc_x, __fmul_rn(__ldg(arow + kk), __ldg(bcol + (size_t)kk * p.N + j)));
                    crow[j] = acc_x;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                continue;
            }
            #pragma unroll
            for (int cb = 0; cb < (ITEMS > 1 ? G::ITEM_BLOCKS : BN / 32); ++cb) {
                if (cb * 32 >= cols_here) break;                              // (warp-uniform)
                const int cc = cb;                                            // column block of C; packed: the item's blocks are
                unsigned r[32];                                               // found in the DIAGONAL part of the accumulator
                const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + acc * BN + (ITEMS > 1 ? qi * G::ITEM_BLOCKS + cb : cb) * 32;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                               "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                // a thread holds 32 columns of ONE row: transpose the warp's 32 x 32 block through shared memory so
                // that every store instruction writes four full 128-byte row segments instead of 32 half sectors
                __syncwarp();
                #pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(stage_w + lane * EPI_PITCH + j * 16),
                                 "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
                __syncwarp();
                #pragma unroll
                for (int i = 0; i < RPW / 4; ++i) {
                    const int rr = i * 4 + (lane >> 3), c4 = lane & 7;
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                 : "r"(stage_w + rr * EPI_PITCH + c4 * 16));
                    if (rr < rows_here && cb * 32 + c4 * 4 < cols_here) st_global_stream16(cblk + (size_t)rr * p.N + cc * 32 + c4 * 4, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((unsigned)TMEM_COLS) : "memory");
    }
}

template <class G>
static int launch(const Shard& s, int m, int k, int n, const float* A, const float* B, float* C, int sms) {
    constexpr int BM = G::BM, BN = G::BN;
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return T5I_CUDA;
    if (s.count > 0x7fffffffull) return T5I_UNSUPPORTED;                  // TMA coordinates are int32
    CUtensorMap ma, mb;
    {   // A: {k, m, count} floats, box 32 floats x BM rows x 1 item, K-major.  Rows >= m / depth >= k: zero fill.
        const cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)m, (cuuint64_t)s.count};
        const cuuint64_t strides[2] = {(cuuint64_t)k * 4, (cuuint64_t)m * k * 4};
        const cuuint32_t box[3] = {BK, (cuuint32_t)(BM / G::ITEMS), (cuuint32_t)G::ITEMS}, estr[3] = {1, 1, 1};
        if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(A), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return T5I_CUDA;
    }
    {   // B: [k rows][n] floats per item, {n, k, count}; box = one 128-byte column block: 32 columns x 32 k-rows x 1 item
        // (one copy per block: n only has to keep the row pitch a multiple of 16 bytes, columns past n arrive as zeros)
        const cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)k, (cuuint64_t)s.count};
        const cuuint64_t strides[2] = {(cuuint64_t)n * 4, (cuuint64_t)k * n * 4};
        const cuuint32_t box[3] = {32, BK, 1}, estr[3] = {1, 1, 1};
        if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(B), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return T5I_CUDA;
    }
    const unsigned long long tiles = G::ITEMS > 1 ? ((unsigned long long)s.count + G::ITEMS - 1) / G::ITEMS
                                                  : (unsigned long long)s.count * ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    const int grid = (int)(tiles < (unsigned long long)sms ? tiles : (unsigned long long)sms);
    Args a{A, B, C, m, n, k, (unsigned long long)s.count};
    tf32x3_kernel<G><<<grid, THREADS, G::SMEM, s.stream>>>(ma, mb, a);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

}  // namespace tf

// padded tile area over useful area for a square tiling of edge T
static inline double tf_waste(int m, int n, int T) {
    return (double)((m + T - 1) / T * T) * (double)((n + T - 1) / T * T) / ((double)m * n);
}

// Shape rule of the tcgen05 Float product (also what t5_matmul_is_exact reports); *geo = tile edge (32: four items per tile).
// TMA: row pitches must be whole 16-byte units (k % 4, n % 4); m is free.
// Items of at least 33 x 36 x 32 (smaller ones: the packed geometry below, or the exact kernels).
// The tensor pipe is ~1/3 busy on full tiles (the kernel is HBM-bound), so up to 2x padded tile area is affordable;
// between the two geometries the one with less padding wins, the larger tile on a tie (operands read once).
// Items of at most 32 x 32 go four to a 128 tile (*geo = 32) once they carry ~8 KB (below that the per-tile pipeline cost,
// not HBM, bounds the kernel and the exact kernels tie: 24^3 3.9 TB/s either way, 20^3 2.6 against 4.7), items of at most
// 64 x 64 two to a tile (*geo = 2; 48^3: 2.0 TB/s exact, 3.1 one per 64 tile, 3.7 two per 128 tile; 64^3: 4.8 -> 5.7).
bool tf32x3_takes(int m, int k, int n, int* geo_out) {
    if (n % 4 || k % 4 || k < 8) return false;
    if (m <= 32 && n <= 32) {
        if (m < 17 || n < 20 || m * k + k * n + m * n < 2048) return false;
        if (geo_out) *geo_out = 32;
        return true;
    }
    if (k < tf::BK || m < 33 || n < 36) return false;
    if (m <= 64 && n <= 64) { if (geo_out) *geo_out = 2; return true; }
    const double w128 = tf_waste(m, n, 128), w64 = tf_waste(m, n, 64);
    if (w128 > 3.5 && w64 > 3.5) return false;
    if (geo_out) *geo_out = w128 <= w64 * 1.3 ? 128 : 64;
    return true;
}

int tf32x3_matmul(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    int geo = 0;
    if (!tf32x3_takes(m, k, n, &geo)) return T5I_NOT_SPECIALISED;
    static std::atomic<bool> configured[T5_MAX_DEVICES];   // zero-initialised; entries of different contexts / pipeline threads may race here (idempotent work)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!configured[dev]) {
        if (cudaFuncSetAttribute(tf::tf32x3_kernel<tf::Geo128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tf::Geo128::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(tf::tf32x3_kernel<tf::Geo64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tf::Geo64::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(tf::tf32x3_kernel<tf::Geo32x4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tf::Geo32x4::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(tf::tf32x3_kernel<tf::Geo64x2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tf::Geo64x2::SMEM) != cudaSuccess) return T5I_CUDA;
        configured[dev] = true;
    }
    if (s.count == 0) return T5I_OK;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = T5_SM_COUNT_FALLBACK;
    if (geo == 32) return tf::launch<tf::Geo32x4>(s, m, k, n, (const float*)A, (const float*)B, (float*)C, sms);
    if (geo == 2) return tf::launch<tf::Geo64x2>(s, m, k, n, (const float*)A, (const float*)B, (float*)C, sms);
    return geo == 128 ? tf::launch<tf::Geo128>(s, m, k, n, (const float*)A, (const float*)B, (float*)C, sms)
               : tf::launch<tf::Geo64>(s, m, k, n, (const float*)A, (const float*)B, (float*)C, sms);
}

}  // namespace t5
