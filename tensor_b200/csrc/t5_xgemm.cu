// Exact blocked product: per item C[P x Q] = A[P x K] . B[K x Q] for every shape and element type
// that has no faster specialisation - large Float / Int64 matrices (no TF32: the 1e-5 and bit-exact
// bars of north_star forbid it), Double shapes that are not multiples of the DMMA tiles.
//
// Classic register-tiled SIMT kernel, but with the ORACLE's arithmetic: every output element is
// acc = 0; for k ascending: acc = acc + a[i,k] * b[k,j]  - one rounding after the multiply, one after
// the add (this TU is compiled with -fmad=false), k strictly sequential per element.  Blocking
// only changes which thread computes an element and where its operands are staged, never the
// order of its operations, so results are bit-identical to the oracle for Double, Float and Int64.
//
//   CTA = 8 x 8 threads, thread tile TM x TN, CTA tile (8 TM) x (8 TN), K walked in chunks of 16
//   global --LDG (coalesced along rows; next chunk prefetched into registers)--> STS
//          A chunk stored TRANSPOSED As[k][row] (+4 pad), B chunk Bs[k][col]
//   shared --LDS.128: a thread's rows/cols are groups of 16 B at stride 8 groups, so every quarter-
//          warp reads 128 contiguous bytes (B) or one broadcast address (A)--> registers
//   TM x TN multiply+add per k; epilogue: 16-B stores where the row pitch allows, scalar otherwise
//
// Rows/columns beyond P/Q are loaded as zeros and never stored; the K tail is handled by the loop
// bound (never by zero padding: 0 * inf would poison a sum the oracle never forms).
#include "t5_device.cuh"
#include "t5_internal.h"
#include <mutex>
#include <type_traits>

namespace t5 {
namespace xg {

constexpr int BK = 16, THREADS = 64;

// The thread's TM x TN accumulators and the one operation on them: acc[i][j] = acc[i][j] + a[i] * b[j], the multiply
// rounded, then the add rounded (never fused).
template <class T, int TM, int TN>
struct Acc {
    using R = Arith<T>;
    T v[TM][TN];
    __device__ __forceinline__ void zero() {
        #pragma unroll
        for (int i = 0; i < TM; ++i)
            #pragma unroll
            for (int j = 0; j < TN; ++j) v[i][j] = R::zero();
    }
    __device__ __forceinline__ void step(const T (&a)[TM], const T (&b)[TN], unsigned long long, unsigned long long) {
        #pragma unroll
        for (int i = 0; i < TM; ++i)
            #pragma unroll
            for (int j = 0; j < TN; ++j) v[i][j] = R::add(v[i][j], R::mul(a[i], b[j]));
    }
    __device__ __forceinline__ uint4 vec16(int i, int h) const {      // elements [i][h * VE .. h * VE + VE)
        constexpr int VE = 16 / (int)sizeof(T);
        uint4 r;
        #pragma unroll
        for (int e = 0; e < VE; ++e) reinterpret_cast<T*>(&r)[e] = v[i][h * VE + e];
        return r;
    }
    __device__ __forceinline__ T get(int i, int j) const { return v[i][j]; }
};

// Float: Blackwell's packed FP32 pairs (SASS FFMA2: two FP32 lanes per register pair, one issue slot).  ptxas CONTRACTS
// a mul.rn.f32x2 feeding an add.rn.f32x2 into one fused FFMA2 even under -fmad=false (seen in the SASS: a * b + acc with
// a single rounding - not the oracle's arithmetic), so the two roundings are spelled as two packed FMAs that cannot be
// merged:  t = fma(a, b, -0.0)  is the correctly rounded product (x + -0.0 = x for every x, signed zeros included),
// acc = fma(t, 1.0, acc)  the correctly rounded sum.  With literal constants ptxas folds the pair back into ONE fused
// FFMA2 (x * 1 -> x, y + -0 -> y, then the same contraction), so 1.0 and -0.0 arrive as kernel parameters whose values
// it cannot know.  Bit-identical to FMUL + FADD (tested against the oracle, specials included), but one issue slot now
// carries two multiplies (or two adds): the scalar version was ISSUE-bound (ncu: issue active 72 %, FMUL + FADD + LDS +
// addressing sharing the slots, 0.43 of the FP32 lanes).
template <int TM, int TN>
struct Acc<float, TM, TN> {
    static_assert(TN % 2 == 0, "pairs along the row");
    unsigned long long v[TM][TN / 2];               // {acc[i][2j] (low half), acc[i][2j+1] (high half)}
    __device__ __forceinline__ void zero() {
        #pragma unroll
        for (int i = 0; i < TM; ++i)
            #pragma unroll
            for (int j = 0; j < TN / 2; ++j) v[i][j] = 0ull;        // (+0.0f, +0.0f)
    }
    __device__ __forceinline__ void step(const float (&a)[TM], const float (&b)[TN], unsigned long long one2, unsigned long long negzero2) {
        unsigned long long bp[TN / 2];
        #pragma unroll
        for (int j = 0; j < TN / 2; ++j)
            asm("mov.b64 %0, {%1, %2};" : "=l"(bp[j]) : "r"(__float_as_uint(b[2 * j])), "r"(__float_as_uint(b[2 * j + 1])));
        #pragma unroll
        for (int i = 0; i < TM; ++i) {
            unsigned long long aa, t;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "r"(__float_as_uint(a[i])));
            #pragma unroll
            for (int j = 0; j < TN / 2; ++j) {
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(aa), "l"(bp[j]), "l"(negzero2));        // a * b, rounded
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v[i][j]) : "l"(t), "l"(one2), "l"(v[i][j]));   // + acc, rounded
            }
        }
    }
    __device__ __forceinline__ uint4 vec16(int i, int h) const {
        uint4 r;
        r.x = (unsigned)v[i][2 * h]; r.y = (unsigned)(v[i][2 * h] >> 32);
        r.z = (unsigned)v[i][2 * h + 1]; r.w = (unsigned)(v[i][2 * h + 1] >> 32);
        return r;
    }
    __device__ __forceinline__ float get(int i, int j) const {
        const unsigned long long x = v[i][j / 2];
        return __uint_as_float((j & 1) ? (unsigned)(x >> 32) : (unsigned)x);
    }
};

struct Args {
    const void* A;
    const void* B;
    void* C;
    int P, K, Q;
    unsigned long long count;
    unsigned long long one2, negzero2;   // packed (1.0f, 1.0f) and (-0.0f, -0.0f): kernel PARAMETERS, so that ptxas cannot fold them (below)
};

template <class T, int TM, int TN>
struct Cfg {
    static constexpr int VE = 16 / (int)sizeof(T);            // elements per 16-B vector
    static constexpr int BM = 8 * TM, BN = 8 * TN;
    static constexpr int A_LD = BM + 4;                       // As[k][row], padded
    static constexpr int A_ELEMS = BK * A_LD, B_ELEMS = BK * BN;
    static constexpr int SMEM = 2 * (A_ELEMS + B_ELEMS) * (int)sizeof(T);
    static constexpr int A_PER_THREAD = BM * BK / THREADS, B_PER_THREAD = BK * BN / THREADS;
    static_assert(TM % VE == 0 && TN % VE == 0, "thread tile must be whole 16-byte vectors");
    static_assert(SMEM <= 48 * 1024, "static shared memory limit");
};

template <class T, int TM, int TN>
__global__ void __launch_bounds__(THREADS, 6) xgemm_kernel(Args p) {
    using C = Cfg<T, TM, TN>;
    constexpr int VE = C::VE, BM = C::BM, BN = C::BN, A_LD = C::A_LD;
    __shared__ __align__(16) T smem[2 * (C::A_ELEMS + C::B_ELEMS)];

    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int tiles_m = (p.P + BM - 1) / BM, tiles_n = (p.Q + BN - 1) / BN;
    const unsigned long long tiles_per_item = (unsigned long long)tiles_m * tiles_n;
    const unsigned long long n_tiles = p.count * tiles_per_item;
    const int kchunks = (p.K + BK - 1) / BK;
    const bool vec_store = ((size_t)p.Q * sizeof(T)) % 16 == 0;

    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long item = tile / tiles_per_item;
        const int rem = (int)(tile - item * tiles_per_item);
        const int tm = rem / tiles_n, tn = rem - tm * tiles_n;
        const int r0 = tm * BM, c0 = tn * BN;
        const T* gA = (const T*)p.A + item * (size_t)p.P * p.K;
        const T* gB = (const T*)p.B + item * (size_t)p.K * p.Q;
        T* gC = (T*)p.C + item * (size_t)p.P * p.Q;

        T pa[C::A_PER_THREAD], pb[C::B_PER_THREAD];
        auto fetch = [&](int k0) {          // global -> registers (zeros outside the matrix)
            #pragma unroll
            for (int i = 0; i < C::A_PER_THREAD; ++i) {
                const int e = tid + i * THREADS, r = e / BK, k = e % BK;
                pa[i] = (r0 + r < p.P && k0 + k < p.K) ? __ldg(gA + (size_t)(r0 + r) * p.K + k0 + k) : T(0);
            }
            #pragma unroll
            for (int i = 0; i < C::B_PER_THREAD; ++i) {
                const int e = tid + i * THREADS, k = e / BN, c = e % BN;
                pb[i] = (k0 + k < p.K && c0 + c < p.Q) ? __ldg(gB + (size_t)(k0 + k) * p.Q + c0 + c) : T(0);
            }
        };
        auto stash = [&](int buf) {         // registers -> shared
            T* As = smem + buf * (C::A_ELEMS + C::B_ELEMS);
            T* Bs = As + C::A_ELEMS;
            #pragma unroll
            for (int i = 0; i < C::A_PER_THREAD; ++i) {
                const int e = tid + i * THREADS, r = e / BK, k = e % BK;
                As[k * A_LD + r] = pa[i];
            }
            #pragma unroll
            for (int i = 0; i < C::B_PER_THREAD; ++i) {
                const int e = tid + i * THREADS;
                Bs[e] = pb[i];              // e = k * BN + c
            }
        };

        Acc<T, TM, TN> acc;
        acc.zero();

        __syncthreads();                    // the previous tile's readers are done with both buffers
        fetch(0);
        stash(0);
        __syncthreads();
        for (int kc = 0; kc < kchunks; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < kchunks) fetch((kc + 1) * BK);       // in flight during the math below
            const T* As = smem + buf * (C::A_ELEMS + C::B_ELEMS);
            const T* Bs = As + C::A_ELEMS;
            const int kleft = p.K - kc * BK;
            auto step = [&](int k) {
                alignas(16) T a[TM], b[TN];
                #pragma unroll
                for (int g = 0; g < TM / VE; ++g)
                    *reinterpret_cast<uint4*>(a + g * VE) = *reinterpret_cast<const uint4*>(As + k * A_LD + g * (8 * VE) + ty * VE);
                #pragma unroll
                for (int g = 0; g < TN / VE; ++g)
                    *reinterpret_cast<uint4*>(b + g * VE) = *reinterpret_cast<const uint4*>(Bs + k * BN + g * (8 * VE) + tx * VE);
                acc.step(a, b, p.one2, p.negzero2);
            };
            if (kleft >= BK) {
                #pragma unroll 4
                for (int k = 0; k < BK; ++k) step(k);
            } else {
                for (int k = 0; k < kleft; ++k) step(k);
            }
            if (kc + 1 < kchunks) {
                stash(buf ^ 1);             // that buffer was last read in iteration kc - 1, before the sync below
                __syncthreads();
            }
        }

        // epilogue: thread rows  r0 + g*(8 VE) + ty*VE + v,  columns  c0 + h*(8 VE) + tx*VE + (0..VE-1)
        #pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = r0 + (i / VE) * (8 * VE) + ty * VE + (i % VE);
            if (r >= p.P) continue;
            #pragma unroll
            for (int h = 0; h < TN / VE; ++h) {
                const int c = c0 + h * (8 * VE) + tx * VE;
                T* dst = gC + (size_t)r * p.Q + c;
                if (vec_store && c + VE <= p.Q) {
                    st_global_stream16(dst, acc.vec16(i, h));
                } else {
                    #pragma unroll
                    for (int v = 0; v < VE; ++v)
                        if (c + v < p.Q) dst[v] = acc.get(i, h * VE + v);
                }
            }
        }
    }
}

template <class T, int TM, int TN>
static int launch(const Shard& s, int P, int K, int Q, const void* A, const void* B, void* Cc) {
    using C = Cfg<T, TM, TN>;
    auto kern = xgemm_kernel<T, TM, TN>;
    static int cap[T5_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!cap[dev]) {
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, 0) != cudaSuccess) return T5I_CUDA;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cap[dev] = (occ < 1 ? 1 : occ) * (sms < 1 ? T5_SM_COUNT_FALLBACK : sms);
    }
    const unsigned long long tiles = (unsigned long long)s.count * ((P + C::BM - 1) / C::BM) * ((Q + C::BN - 1) / C::BN);
    const int grid = (int)(tiles < (unsigned long long)cap[dev] ? tiles : (unsigned long long)cap[dev]);
    Args a{A, B, Cc, P, K, Q, (unsigned long long)s.count, 0x3f8000003f800000ull, 0x8000000080000000ull};
    kern<<<grid, THREADS, 0, s.stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

// padded work of a tiling relative to the useful work
static inline double waste(int P, int Q, int BM, int BN) {
    const double padded = (double)((P + BM - 1) / BM * BM) * (double)((Q + BN - 1) / BN * BN);
    return padded / ((double)P * Q);
}

template <class T, int TM_BIG, int TN_BIG, int TM_SMALL, int TN_SMALL>
static int pick(const Shard& s, int P, int K, int Q, const void* A, const void* B, void* C) {
    const double wb = waste(P, Q, 8 * TM_BIG, 8 * TN_BIG), ws = waste(P, Q, 8 * TM_SMALL, 8 * TN_SMALL);
    if (ws > 2.0) return T5I_NOT_SPECIALISED;           // tiny outputs: the any-shape kernel wastes less
    if (wb <= 1.15 * ws) return launch<T, TM_BIG, TN_BIG>(s, P, K, Q, A, B, C);
    return launch<T, TM_SMALL, TN_SMALL>(s, P, K, Q, A, B, C);
}

}  // namespace xg

int exact_blocked_matmul(const Shard& s, int dtype, int P, int K, int Q, const void* A, const void* B, void* C) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (K < 8 || (long long)P * Q < 256) return T5I_NOT_SPECIALISED;
    if (s.count == 0) return T5I_OK;
    switch (dtype) {
        case T5D_F32: return xg::pick<float, 8, 8, 4, 4>(s, P, K, Q, A, B, C);
        case T5D_F64: return xg::pick<double, 8, 4, 4, 4>(s, P, K, Q, A, B, C);
        case T5D_I64: return xg::pick<long long, 8, 4, 4, 4>(s, P, K, Q, A, B, C);
    }
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
