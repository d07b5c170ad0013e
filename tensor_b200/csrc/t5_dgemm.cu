// Batched Double `.*.` for large matrices on the FP64 tensor pipe (SURVEY.md §8a a6,
// BASELINE configs[4]: 64Ki products of 128x128; north_star: ">= 64x64 products use DMMA").
//
// tcgen05.mma has no f64 kind: on sm_100a the FP64 tensor instruction is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), fed from registers - TMEM does not apply.  What does
// apply from the Blackwell playbook is the data movement: TMA tensor maps, an mbarrier ring and
// warp specialisation (below).  One persistent CTA owns a BM x BN tile of C; 8 DMMA warps, 2 x 4
// over the tile, hold the accumulators in registers and store them straight to HBM.
//
// Rounding differs from the sequential fold of the oracle (the tensor pipe fuses and
// re-associates inside a k4 step): tests hold this kernel to
// |delta| <= 1e-12 * sum_j |a_ij||b_jk|, the tolerance north_star states for Double.
// History (numbers in DESIGN.md §5.9): a cp.async ring issued by the DMMA warps with a CTA barrier
// per K chunk kept the DMMA pipe 74 % busy (9.49 ms for 64Ki x 128^3); this kernel 96 % (7.73 ms).
#include <atomic>

#include "t5_async.cuh"
#include "t5_device.cuh"
#include "t5_internal.h"

namespace t5 {

// One FP64 tensor instruction: D(8x8) += A(8x4) . B(4x8); lane (g = lane/4, t = lane%4) holds
// A[g][t], B[t][g] and the accumulator pair C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------
// The operand ring is filled by ONE producer lane with cp.async.bulk.tensor (3-D tensor maps,
// SWIZZLE_64B) and handed over through full/empty mbarriers, so the eight DMMA warps never
// meet at a CTA barrier: they drift apart, one warp of a scheduler's pair stores its
// accumulators while the other keeps the FP64 tensor pipe busy, and no DMMA-warp issue slot
// or register goes to address generation for the copies.
//
//   A map: dims {K, M, count}, one box {8, BM, ITEMS} per group of 8 k    -> shared [4 k-groups][item][BM rows][8 doubles]
//   B map: dims {N, K, count}, one box {8, 32, ITEMS} per group of 8 n    -> shared [BN/8 n-groups][item][32 k-rows][8 doubles]
//   (64-byte rows, 16-B chunks XOR-ed with (row >> 1) & 3 by the TMA unit; stage bases 1024-B aligned)
// The ITEM is its own outermost dimension (round 2): the part of a box beyond an item's M rows, N columns or K depth is
// out of bounds for TMA and arrives as ZEROS, never as the next item's data, so shapes that are not multiples of the
// tile run too - zero rows / columns / k-slices add nothing, and the epilogue stores only what exists.  One box per
// 64-byte group (not one 4-D box over a {8, rows, cols/8} view): K and N only have to be EVEN (16-byte row pitch).
//
// The four k of a DMMA step are free to be any four k of the chunk as long as A and B agree.
// Lane (g, t) of step s uses k = (t&1) + 2(s&1) + 4(t>>1) + 8((s>>1)&1) + 16(s>>2): with the
// 64-B swizzle this makes the 16 lanes of each half-warp (LDS.64 is served per half-warp) hit 16
// distinct bank pairs for both operands, without padding.  (A first version on SWIZZLE_128B was
// conflict-free only across the full warp: ncu showed 2x the shared-memory wavefronts.)
namespace dt {
constexpr int BK = 32;
// 8 DMMA warps (two warpgroups) + one producer warpgroup of which a single lane works.  Registers
// are a per-scheduler pool (3 warps x 168 at launch); setmaxnreg moves the producer group's share
// to the DMMA warps (2 x 232 + 40 <= 512 per scheduler lane), which need 64 accumulators + fragments.
constexpr int CONSUMER_WARPS = 8, THREADS = (CONSUMER_WARPS + 4) * 32;
// Tile geometry: 2 x 4 DMMA warps over BM x BN; 128 x 128 (64 accumulators per thread, 3 stages of
// 64 KB) for the large squares, 64 x 64 (16 accumulators, 6 stages of 32 KB) for 64-multiples.
// Items of at most 32 x 32 (round 2): ONE tile is ITEMS = 4 consecutive items - the item is a tensor-map dimension, so
// a box {8, 32, 4, 4} brings four items' operands in one bulk copy, [item][k group | n group][row][8 doubles], and two
// warps share an item (16 x 32 each).  K-steps beyond an item's K are skipped, not multiplied by zeros.
template <int BM_, int BN_, int STAGES_, int ITEMS_ = 1, int WM_ = 2, int WN_ = 4>
struct Geo {
    static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_, ITEMS = ITEMS_, WM = WM_, WN = WN_;
    static constexpr int WTM = BM / WM, WTN = BN / WN, FI = WTM / 8, FJ = WTN / 8;
    static constexpr int A_GROUP = ITEMS * BM * 64, B_GROUP = ITEMS * BK * 64;       // one TMA box: 8 doubles x rows x items
    static constexpr int A_STAGE = (BK / 8) * A_GROUP, B_STAGE = (BN / 8) * B_GROUP, STAGE = A_STAGE + B_STAGE;
    static constexpr int SMEM = STAGES * STAGE + 1024;     // + slack to align the ring to 1024 B
    static_assert(SMEM <= 227 * 1024, "ring does not fit");
    static_assert(WTN % 8 == 0 && WTM % 8 == 0, "warp tile must be whole 8x8 DMMA fragments");
    static_assert(ITEMS * WM * WN == CONSUMER_WARPS, "eight DMMA warps");
    static_assert(A_GROUP % 512 == 0 && B_GROUP % 512 == 0 && (BM * 64) % 512 == 0, "box and item bases keep the swizzle phase");
};
using Geo128 = Geo<128, 128, 3>;
using Geo64 = Geo<64, 64, 6>;
using Geo64x2 = Geo<64, 64, 3, 2, 2, 2>;        // items of at most 64 x 64, two to a tile: 32 x 32 warp tiles (the 128 tile's operand reuse)
using Geo96 = Geo<96, 96, 4>;                   // 72 / 80 / 96 / 160 / 192: less padding than the 128 tile, more reuse than 64
using Geo48x2 = Geo<48, 48, 4, 2, 2, 2>;        // two items of at most 48 x 48, four warps (24 x 24 each) per item
using Geo32x4 = Geo<32, 32, 3, 4, 2, 1>;        // four items of at most 32 x 32, two warps (16 x 32 each) per item

struct Args {
    double* C;
    int M, N, K;
    unsigned long long count;
};

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}

template <class G>
__global__ void __launch_bounds__(THREADS, 1)
dmma_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Args p) {
    constexpr int BM = G::BM, BN = G::BN, STAGES = G::STAGES, A_STAGE = G::A_STAGE, STAGE = G::STAGE, ITEMS = G::ITEMS;
    constexpr int WTM = G::WTM, WTN = G::WTN, FI = G::FI, FJ = G::FJ;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * STAGES];       // full[0..S), empty[S..2S)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned bar0 = smem_u32(bars);
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN, kch = (p.K + BK - 1) / BK;
    const unsigned long long tiles_per_item = (unsigned long long)tiles_m * tiles_n;      // 1 when ITEMS > 1 (host checks)
    const unsigned long long n_tiles = ITEMS > 1 ? (p.count + ITEMS - 1) / ITEMS : p.count * tiles_per_item;

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar0 + 8 * s, 1);                                 // full: the producer's expect_tx arrive
            mbar_init(bar0 + 8 * (STAGES + s), CONSUMER_WARPS);         // empty: one arrive per DMMA warp
        }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp >= CONSUMER_WARPS) {
        // ---- producer: one lane walks (tile, K chunk) in the consumers' order ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp != CONSUMER_WARPS || lane != 0) return;
        unsigned it = 0;
        for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const unsigned long long item = ITEMS > 1 ? tile * ITEMS : tile / tiles_per_item;   // (items past the batch: zeros)
            const int rem = ITEMS > 1 ? 0 : (int)(tile - item * tiles_per_item);
            const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
            for (int kc = 0; kc < kch; ++kc, ++it) {
                const unsigned s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(bar0 + 8 * (STAGES + s), ph ^ 1);              // stage released by all 8 warps
                const unsigned full = bar0 + 8 * s;
                mbar_expect_tx(full, STAGE);                              // (zero-filled parts of a box count too)
                #pragma unroll
                for (int g8 = 0; g8 < BK / 8; ++g8)
                    tma_load_3d(ring + s * STAGE + g8 * G::A_GROUP, &mapA, full, kc * BK + g8 * 8, tm * BM, (int)item);
                #pragma unroll
                for (int g8 = 0; g8 < BN / 8; ++g8)
                    tma_load_3d(ring + s * STAGE + A_STAGE + g8 * G::B_GROUP, &mapB, full, tn * BN + g8 * 8, kc * BK, (int)item);
            }
        }
        return;
    }

    // ---- consumers: 8 DMMA warps, 2 x 4 over the tile ----
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = lane >> 2, t = lane & 3, t0 = t & 1, t1 = t >> 1;
    const int wq = warp / (G::WM * G::WN), wm = (warp / G::WN) % G::WM, wn = warp % G::WN;   // item of the tile, warp row, warp column
    // runtime parts of the swizzled fragment addresses (see the header comment)
    const unsigned a_thr = (unsigned)wq * (BM * 64) + (unsigned)(wm * WTM + g) * 64u + (unsigned)t0 * 8u;
    const unsigned b_thr = (unsigned)wq * 2048u + (unsigned)(wn * (WTN / 8)) * G::B_GROUP + (unsigned)(t0 + 4 * t1) * 64u + (unsigned)(g & 1) * 8u;
    const unsigned x_thr = (unsigned)((g >> 1) ^ (2 * t1));                      // 16-B chunk = x_thr ^ (s & 1), both operands

    double acc[FI][FJ][2];
    unsigned it = 0;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        #pragma unroll
        for (int i = 0; i < FI; ++i)
            #pragma unroll
            for (int j = 0; j < FJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kc = 0; kc < kch; ++kc, ++it) {
            const unsigned s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(bar0 + 8 * s, ph);                                 // TMA bytes have landed
            const unsigned As = ring + s * STAGE + a_thr, Bs = ring + s * STAGE + A_STAGE + b_thr;
            const int ksteps = ((p.K - kc * BK + 7) >> 3) << 1;          // steps 0..s-1 (s even) hold exactly k < 4s (see the k order above)
            #pragma unroll
            for (int st = 0; st < BK / 4; ++st) {
                if (st >= 2 && st >= ksteps) break;                       // K tail: nothing but zeros beyond
                const unsigned chunk = (x_thr ^ (unsigned)(st & 1)) << 4;
                const unsigned a_off = (unsigned)(st >> 1) * (unsigned)G::A_GROUP + chunk;
                const unsigned b_off = (unsigned)(2 * (st & 1) + 8 * ((st >> 1) & 1) + 16 * (st >> 2)) * 64u + chunk;
                double a[FI], b[FJ];
                #pragma unroll
                for (int i = 0; i < FI; ++i) a[i] = lds_f64(As + a_off + (unsigned)i * 512u);
                #pragma unroll
                for (int j = 0; j < FJ; ++j) b[j] = lds_f64(Bs + b_off + (unsigned)j * (unsigned)G::B_GROUP);
                #pragma unroll
                for (int i = 0; i < FI; ++i)
                    #pragma unroll
                    for (int j = 0; j < FJ; ++j) dmma(acc[i][j], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (STAGES + s));         // this warp is done with the stage
        }

        const unsigned long long item = ITEMS > 1 ? tile * ITEMS + wq : tile / tiles_per_item;
        if (ITEMS > 1 && item >= p.count) continue;                      // the last tile's missing items
        const int rem = ITEMS > 1 ? 0 : (int)(tile - item * tiles_per_item);
        const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
        const int row0 = tm * BM + wm * WTM + g, col0 = tn * BN + wn * WTN + 2 * t;
        double* Cg = p.C + item * (size_t)p.M * p.N + (size_t)row0 * p.N + col0;
        #pragma unroll
        for (int i = 0; i < FI; ++i)
            #pragma unroll
            for (int j = 0; j < FJ; ++j) {
                if (row0 + i * 8 >= p.M || col0 + j * 8 >= p.N) continue;        // beyond a ragged item (N is even: pairs are whole)
                uint4 v;
                reinterpret_cast<double*>(&v)[0] = acc[i][j][0];
                reinterpret_cast<double*>(&v)[1] = acc[i][j][1];
                st_global_stream16(Cg + (size_t)(i * 8) * p.N + j * 8, v);
            }
    }
}

// `count` items of rows x cols doubles, row-major: {cols, rows, count}; box {8 doubles, box_rows, box_items}
static bool make_map(CUtensorMap* map, const void* base, unsigned long long count, int rows, int cols, int box_rows, int box_items) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, count};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 8, (cuuint64_t)rows * cols * 8};
    const cuuint32_t box[3] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_items};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class G>
static int launch(const Shard& s, int m, int k, int n, const double* A, const double* B, double* C, int sms) {
    constexpr int BM = G::BM, BN = G::BN;
    if (s.count > 0x7fffffffull) return T5I_UNSUPPORTED;                  // TMA coordinates are int32
    CUtensorMap ma, mb;
    if (!make_map(&ma, A, (unsigned long long)s.count, m, k, BM, G::ITEMS)) return T5I_CUDA;
    if (!make_map(&mb, B, (unsigned long long)s.count, k, n, BK, G::ITEMS)) return T5I_CUDA;
    const unsigned long long tiles = G::ITEMS > 1 ? ((unsigned long long)s.count + G::ITEMS - 1) / G::ITEMS
                                                  : (unsigned long long)s.count * ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    const int grid = (int)(tiles < (unsigned long long)sms ? tiles : (unsigned long long)sms);
    Args a{C, m, n, k, (unsigned long long)s.count};
    dmma_tma_kernel<G><<<grid, THREADS, G::SMEM, s.stream>>>(ma, mb, a);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}
}  // namespace dt


// Shape rule of the FP64 tensor-core product (also what t5_matmul_is_exact reports); *geo = tile edge chosen.
// TMA needs 16-byte row pitches: k and n even; m is free.
// Geometry = the least padded tile area (ties and near-ties go to the larger tile: more operand reuse per byte of
// shared-memory traffic).  Measured against the exact kernels (profiles/r02r_dmma_small_items.txt): the tensor pipe wins
// from 17 x 24 items on (24^3: 4.2 -> 6.7 TB/s, 32^3: 3.6 -> 6.0, 48^3: 1.9 -> 4.1) and still by 1.8x when two thirds of
// the tile are padding (72^3 on the 128 tile), so only thin items stay on the SIMT kernels.
bool dmma_takes(int m, int k, int n, int* geo_out) {
    if (n % 2 || k % 2 || k < 8 || m < 17 || n < 24) return false;
    const double area = (double)m * n;
    auto padded = [&](int T) { return (double)((m + T - 1) / T * T) * (double)((n + T - 1) / T * T) / area; };
    int geo = 128; double best = padded(128);
    for (int T : {96, 64}) if (padded(T) < best * 0.95) { geo = T; best = padded(T); }
    if (m <= 48 && n <= 48 && padded(48) < best * 0.95) { geo = 48; best = padded(48); }
    if (m <= 32 && n <= 32) { geo = 32; best = padded(32); }
    if (best > 4.0) return false;
    if (geo_out) *geo_out = geo;
    return true;
}

int dmma_matmul(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    int geo = 0;
    if (!dmma_takes(m, k, n, &geo)) return T5I_NOT_SPECIALISED;
    static std::atomic<bool> configured[T5_MAX_DEVICES];   // zero-initialised; entries of different contexts / pipeline threads may race here (idempotent work)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!configured[dev]) {
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo128>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo128::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo64>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo64::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo32x4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo32x4::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo48x2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo48x2::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo64x2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo64x2::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo96>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo96::SMEM) != cudaSuccess) return T5I_CUDA;
        configured[dev] = true;
    }
    if (s.count == 0) return T5I_OK;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = T5_SM_COUNT_FALLBACK;
    const double *a = (const double*)A, *b = (const double*)B;
    double* c = (double*)C;
    switch (geo) {
        case 32: return dt::launch<dt::Geo32x4>(s, m, k, n, a, b, c, sms);
        case 48: return dt::launch<dt::Geo48x2>(s, m, k, n, a, b, c, sms);
        // items that fit one 64 tile go two to a CTA (64^3: 5.38 -> 5.89 TB/s, 0.90 of the HBM peak); a single K chunk is
        // better served by the deeper ring of the one-item geometry (64 x 32 . 32 x 64: 6.3 vs 5.8 TB/s)
        case 64: return (m <= 64 && n <= 64 && k >= 2 * dt::BK) ? dt::launch<dt::Geo64x2>(s, m, k, n, a, b, c, sms)
                                                               : dt::launch<dt::Geo64>(s, m, k, n, a, b, c, sms);
        case 96: return dt::launch<dt::Geo96>(s, m, k, n, a, b, c, sms);
        default: return dt::launch<dt::Geo128>(s, m, k, n, a, b, c, sms);
    }
}

}  // namespace t5
