// Batched Double `.*.` for large matrices on the FP64 tensor pipe (SURVEY.md §8a a6,
// BASELINE configs[4]: 64Ki products of 128x128).
//
// tcgen05.mma has no f64 kind: on sm_100a the FP64 tensor instruction is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), fed from registers.  One persistent CTA
// (8 warps) owns a 128x128 tile of C; a warp owns 64x32 of it (8x4 DMMA fragments,
// 64 accumulator doubles per thread).  K is walked in chunks of 32:
//
//   global --cp.async 16 B, 3-deep ring, running ACROSS tiles/items--> shared
//          A chunk 128 x 32 (row stride 36 doubles), B chunk 32 x 128 (row stride 132)
//   shared --LDS.64--> fragments: the strides are = 4 (mod 16) eight-byte bank pairs, so
//          the 8x4 (A) and 4x8 (B) fragment reads of a half-warp hit 16 distinct pairs
//   DMMA, accumulate in registers; epilogue: STG.128 straight from the accumulators
//          (4 lanes cover 64 contiguous bytes of a C row).
//
// Tiles are claimed from a global atomic counter two tiles ahead, so the ring never
// drains between items.  Rounding differs from the sequential fold of the oracle (the
// tensor pipe fuses and re-associates inside a k4 step): tests hold this kernel to
// |delta| <= 1e-12 * sum_j |a_ij||b_jk|, the tolerance north_star states for Double.
#include "t5_device.cuh"
#include "t5_internal.h"

namespace t5 {

namespace dg {
constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3, THREADS = 256;
constexpr int A_LD = BK + 4;    // doubles per A row in shared memory
constexpr int B_LD = BN + 4;    // doubles per B row in shared memory
constexpr int A_STAGE = BM * A_LD * 8;
constexpr int B_STAGE = BK * B_LD * 8;
constexpr int STAGE = A_STAGE + B_STAGE;
constexpr int SMEM = STAGES * STAGE;
static_assert(SMEM <= 227 * 1024, "ring does not fit");

struct Args {
    const double* A;
    const double* B;
    double* C;
    int M, N, K;                 // per item; M % 128 == 0, N % 128 == 0, K % 32 == 0
    unsigned long long count;    // items
    unsigned long long* sched;   // {next tile, finished CTAs}
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(THREADS, 1) dmma_gemm_kernel(Args p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long tile_q[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps over the 128 x 128 tile
    const int tiles_m = p.M / BM, tiles_n = p.N / BN, kch = p.K / BK;
    const unsigned long long tiles_per_item = (unsigned long long)tiles_m * tiles_n;
    const unsigned long long n_tiles = p.count * tiles_per_item;

    if (tid == 0) {
        tile_q[0] = atomicAdd(p.sched, 1ull);
        tile_q[1] = atomicAdd(p.sched, 1ull);
    }
    __syncthreads();

    // producer cursor: (sequence number of the tile inside this CTA, K chunk) and the global
    // addresses of that chunk; the tile -> (item, tm, tn) division runs once per tile only
    unsigned pf_n = 0; int pf_kc = 0;
    bool pf_valid = false;
    const double* pf_A = nullptr;
    const double* pf_B = nullptr;
    auto pf_locate = [&]() {
        const unsigned long long tile = tile_q[pf_n & 3];
        pf_valid = tile < n_tiles;
        if (pf_valid) {
            const unsigned long long item = tile / tiles_per_item;
            const int rem = (int)(tile - item * tiles_per_item);
            const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
            // this thread's first 16-B chunk of the A block (row = tid/16, col chunk = tid%16) and of the B block
            pf_A = p.A + item * (size_t)p.M * p.K + (size_t)(tm * BM + (tid >> 4)) * p.K + (tid & 15) * 2;
            pf_B = p.B + item * (size_t)p.K * p.N + (size_t)(tid >> 6) * p.N + tn * BN + (tid & 63) * 2;
        }
    };
    const unsigned a_dst = (tid >> 4) * (A_LD * 8) + (tid & 15) * 16;
    const unsigned b_dst = (tid >> 6) * (B_LD * 8) + (tid & 63) * 16;
    const size_t a_step = (size_t)(THREADS / 16) * p.K;     // 16 rows further down per pass
    const size_t b_step = (size_t)(THREADS / 64) * p.N;     // 4 rows further down per pass
    auto issue = [&](int stage) {
        if (pf_kc == 0) pf_locate();
        if (pf_valid) {
            unsigned char* As = smem + stage * STAGE;
            unsigned char* Bs = As + A_STAGE;
            #pragma unroll
            for (int i = 0; i < (BM * BK / 2) / THREADS; ++i)       // A block: 128 rows x 16 chunks, 16 rows per pass
                cp_async16(As + a_dst + i * (THREADS / 16) * (A_LD * 8), pf_A + i * a_step);
            #pragma unroll
            for (int i = 0; i < (BK * BN / 2) / THREADS; ++i)       // B block: 32 rows x 64 chunks, 4 rows per pass
                cp_async16(Bs + b_dst + i * (THREADS / 64) * (B_LD * 8), pf_B + i * b_step);
            pf_A += BK;                      // next K chunk: 32 columns right in A,
            pf_B += (size_t)BK * p.N;        // 32 rows down in B
        }
        cp_async_commit();
        if (++pf_kc == kch) { pf_kc = 0; ++pf_n; }
    };

    #pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    double acc[8][4][2];
    int stage = 0;
    for (unsigned n = 0;; ++n) {
        const unsigned long long tile = tile_q[n & 3];
        if (tile >= n_tiles) break;
        if (tid == 0) tile_q[(n + 2) & 3] = atomicAdd(p.sched, 1ull);   // published by the syncs below
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            #pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kc = 0; kc < kch; ++kc) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();                       // chunk landed for all; the stage consumed last is free
            int pf = stage + STAGES - 1; if (pf >= STAGES) pf -= STAGES;
            issue(pf);
            const double* As = reinterpret_cast<const double*>(smem + stage * STAGE) + (wm * 64 + g) * A_LD + t;
            const double* Bs = reinterpret_cast<const double*>(smem + stage * STAGE + A_STAGE) + t * B_LD + wn * 32 + g;
            #pragma unroll
            for (int k4 = 0; k4 < BK / 4; ++k4) {
                double a[8], b[4];
                #pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = As[i * 8 * A_LD + k4 * 4];
                #pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[k4 * 4 * B_LD + j * 8];
                #pragma unroll
                for (int i = 0; i < 8; ++i)
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
            }
            if (++stage == STAGES) stage = 0;
        }

        // epilogue: accumulators -> C, 16 B per lane, 64 contiguous bytes per 4 lanes
        const unsigned long long item = tile / tiles_per_item;
        const int rem = (int)(tile - item * tiles_per_item);
        const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
        double* Cg = p.C + item * (size_t)p.M * p.N + (size_t)(tm * BM + wm * 64 + g) * p.N + tn * BN + wn * 32 + 2 * t;
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 v;
                reinterpret_cast<double*>(&v)[0] = acc[i][j][0];
                reinterpret_cast<double*>(&v)[1] = acc[i][j][1];
                st_global_stream16(Cg + (size_t)(i * 8) * p.N + j * 8, v);
            }
    }
    cp_async_wait<0>();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(p.sched + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
        }
    }
}
}  // namespace dg

int dmma_matmul(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
    using namespace dg;
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (m % BM || n % BN || k % BK || k < BK) return T5I_NOT_SPECIALISED;
    static bool configured[T5_MAX_DEVICES] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!configured[dev]) {
        if (cudaFuncSetAttribute(dmma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return T5I_CUDA;
        configured[dev] = true;
    }
    if (s.count == 0) return T5I_OK;
    if (!s.sched) return T5I_CUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = T5_SM_COUNT_FALLBACK;
    const unsigned long long tiles = (unsigned long long)s.count * (m / BM) * (n / BN);
    const int grid = (int)(tiles < (unsigned long long)sms ? tiles : (unsigned long long)sms);
    Args a{(const double*)A, (const double*)B, (double*)C, m, n, k, (unsigned long long)s.count, s.sched};
    dmma_gemm_kernel<<<grid, THREADS, SMEM, s.stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

}  // namespace t5
