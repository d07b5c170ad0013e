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
//   A map: dims {8, count*M, K/8}, box {8, BM, 4}     -> shared [4][BM rows][8 doubles]
//   B map: dims {8, count*K, N/8}, box {8, 32, BN/8}  -> shared [BN/8][32 k-rows][8 doubles]
//   (64-byte rows, 16-B chunks XOR-ed with (row >> 1) & 3 by the TMA unit; stage bases 1024-B aligned)
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
template <int BM_, int BN_, int STAGES_>
struct Geo {
    static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_;
    static constexpr int WTM = BM / 2, WTN = BN / 4, FI = WTM / 8, FJ = WTN / 8;
    static constexpr int A_STAGE = BM * BK * 8, B_STAGE = BK * BN * 8, STAGE = A_STAGE + B_STAGE;
    static constexpr int SMEM = STAGES * STAGE + 1024;     // + slack to align the ring to 1024 B
    static_assert(SMEM <= 227 * 1024, "ring does not fit");
    static_assert(WTN % 8 == 0 && WTM % 8 == 0, "warp tile must be whole 8x8 DMMA fragments");
};
using Geo128 = Geo<128, 128, 3>;
using Geo64 = Geo<64, 64, 6>;

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
    constexpr int BM = G::BM, BN = G::BN, STAGES = G::STAGES, A_STAGE = G::A_STAGE, STAGE = G::STAGE;
    constexpr int WTM = G::WTM, WTN = G::WTN, FI = G::FI, FJ = G::FJ;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * STAGES];       // full[0..S), empty[S..2S)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned bar0 = smem_u32(bars);
    const int tiles_m = p.M / BM, tiles_n = p.N / BN, kch = p.K / BK;
    const unsigned long long tiles_per_item = (unsigned long long)tiles_m * tiles_n;
    const unsigned long long n_tiles = p.count * tiles_per_item;

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
            const unsigned long long item = tile / tiles_per_item;
            const int rem = (int)(tile - item * tiles_per_item);
            const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
            const int a_row = (int)(item * p.M) + tm * BM;
            const int b_row = (int)(item * p.K);
            for (int kc = 0; kc < kch; ++kc, ++it) {
                const unsigned s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(bar0 + 8 * (STAGES + s), ph ^ 1);              // stage released by all 8 warps
                const unsigned full = bar0 + 8 * s;
                mbar_expect_tx(full, STAGE);
                tma_load_3d(ring + s * STAGE, &mapA, full, 0, a_row, kc * (BK / 8));
                tma_load_3d(ring + s * STAGE + A_STAGE, &mapB, full, 0, b_row + kc * BK, tn * (BN / 8));
            }
        }
        return;
    }

    // ---- consumers: 8 DMMA warps, 2 x 4 over the tile ----
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = lane >> 2, t = lane & 3, t0 = t & 1, t1 = t >> 1;
    const int wm = warp >> 2, wn = warp & 3;
    // runtime parts of the swizzled fragment addresses (see the header comment)
    const unsigned a_thr = (unsigned)(wm * WTM + g) * 64u + (unsigned)t0 * 8u;
    const unsigned b_thr = (unsigned)(wn * (WTN / 8)) * 2048u + (unsigned)(t0 + 4 * t1) * 64u + (unsigned)(g & 1) * 8u;
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
            #pragma unroll
            for (int st = 0; st < BK / 4; ++st) {
                const unsigned chunk = (x_thr ^ (unsigned)(st & 1)) << 4;
                const unsigned a_off = (unsigned)(st >> 1) * (unsigned)(BM * 64) + chunk;
                const unsigned b_off = (unsigned)(2 * (st & 1) + 8 * ((st >> 1) & 1) + 16 * (st >> 2)) * 64u + chunk;
                double a[FI], b[FJ];
                #pragma unroll
                for (int i = 0; i < FI; ++i) a[i] = lds_f64(As + a_off + (unsigned)i * 512u);
                #pragma unroll
                for (int j = 0; j < FJ; ++j) b[j] = lds_f64(Bs + b_off + (unsigned)j * 2048u);
                #pragma unroll
                for (int i = 0; i < FI; ++i)
                    #pragma unroll
                    for (int j = 0; j < FJ; ++j) dmma(acc[i][j], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (STAGES + s));         // this warp is done with the stage
        }

        const unsigned long long item = tile / tiles_per_item;
        const int rem = (int)(tile - item * tiles_per_item);
        const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
        double* Cg = p.C + item * (size_t)p.M * p.N + (size_t)(tm * BM + wm * WTM + g) * p.N + tn * BN + wn * WTN + 2 * t;
        #pragma unroll
        for (int i = 0; i < FI; ++i)
            #pragma unroll
            for (int j = 0; j < FJ; ++j) {
                uint4 v;
                reinterpret_cast<double*>(&v)[0] = acc[i][j][0];
                reinterpret_cast<double*>(&v)[1] = acc[i][j][1];
                st_global_stream16(Cg + (size_t)(i * 8) * p.N + j * 8, v);
            }
    }
}

// rows x cols doubles, row-major, viewed as {8, rows, cols/8}; box {8, box_rows, box_grp}
static bool make_map(CUtensorMap* map, const void* base, unsigned long long rows, int cols, int box_rows, int box_grp) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {8, rows, (cuuint64_t)(cols / 8)};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 8, 64};
    const cuuint32_t box[3] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_grp};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class G>
static int launch(const Shard& s, int m, int k, int n, const double* A, const double* B, double* C, int sms) {
    constexpr int BM = G::BM, BN = G::BN;
    // TMA coordinates are int32: walk the shard in pieces whose row counts fit
    const size_t per = (size_t)0x7fffffff / (size_t)(m > k ? m : k);
    for (size_t off = 0; off < s.count; off += per) {
        const size_t cnt = s.count - off < per ? s.count - off : per;
        CUtensorMap ma, mb;
        if (!make_map(&ma, A + off * (size_t)m * k, (unsigned long long)cnt * m, k, BM, BK / 8)) return T5I_CUDA;
        if (!make_map(&mb, B + off * (size_t)k * n, (unsigned long long)cnt * k, n, BK, BN / 8)) return T5I_CUDA;
        const unsigned long long tiles = (unsigned long long)cnt * (m / BM) * (n / BN);
        const int grid = (int)(tiles < (unsigned long long)sms ? tiles : (unsigned long long)sms);
        Args a{C + off * (size_t)m * n, m, n, k, (unsigned long long)cnt};
        dmma_tma_kernel<G><<<grid, THREADS, G::SMEM, s.stream>>>(ma, mb, a);
        if (cudaGetLastError() != cudaSuccess) return T5I_CUDA;
    }
    return T5I_OK;
}
}  // namespace dt


int dmma_matmul(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (m % 64 || n % 64 || k % dt::BK || k < dt::BK) return T5I_NOT_SPECIALISED;
    const bool big = m % 128 == 0 && n % 128 == 0;
    static std::atomic<bool> configured[T5_MAX_DEVICES];   // zero-initialised; entries of different contexts / pipeline threads may race here (idempotent work)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!configured[dev]) {
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo128>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo128::SMEM) != cudaSuccess) return T5I_CUDA;
        if (cudaFuncSetAttribute(dt::dmma_tma_kernel<dt::Geo64>, cudaFuncAttributeMaxDynamicSharedMemorySize, dt::Geo64::SMEM) != cudaSuccess) return T5I_CUDA;
        configured[dev] = true;
    }
    if (s.count == 0) return T5I_OK;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = T5_SM_COUNT_FALLBACK;
    return big ? dt::launch<dt::Geo128>(s, m, k, n, (const double*)A, (const double*)B, (double*)C, sms)
               : dt::launch<dt::Geo64>(s, m, k, n, (const double*)A, (const double*)B, (double*)C, sms);
}

}  // namespace t5
