// Any-shape kernels: the general form of every operation (so that the C ABI is a
// drop-in for arbitrary fixed dimensions, not just the benchmarked ones), plus
// the utility kernels (synthetic fill, re-layout, reductions, element-wise ops).
// Compiled with -fmad=false (bit-exact against the oracle's multiply-round-add).
#include "t5_device.cuh"
#include "t5_internal.h"
#include "t5_recip.cuh"
#include <cstdlib>
#include <type_traits>
#include <mutex>
#include <unordered_map>

namespace t5 {

static inline int sm_count() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : T5_SM_COUNT_FALLBACK;
}
static inline int ok() { return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA; }

// Persistent grid-stride kernels are launched with exactly as many CTAs as can be RESIDENT
// (occupancy x SMs).  A fixed "8 per SM" put 1184 CTAs on 888 slots for the 40-register gather
// kernel: a second, one-third-full wave that cost 30% (ncu, profiles/r01c).
static int resident_blocks(const void* kern, int threads) {
    static std::mutex mu;
    static std::unordered_map<const void*, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(kern);
    if (it != cache.end()) return it->second;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 2048 / threads / 2;
    }
    cache[kern] = occ;
    return occ;
}
static inline int resident_grid(const void* kern, int threads, size_t units) {
    const size_t cap = (size_t)sm_count() * resident_blocks(kern, threads);
    return (int)(units < cap ? (units ? units : 1) : cap);
}

// =====================================================================================
// prod n: per item C[P x Q] = A[P x K] . B[K x Q]; consecutive threads
// own consecutive output vectors (coalesced stores, B rows coalesced, A broadcast).
// =====================================================================================
template <class T, int VEC> struct VecT;
template <> struct VecT<double, 2> { using type = double2; };
template <> struct VecT<long long, 2> { using type = longlong2; };
template <> struct VecT<float, 4> { using type = float4; };
template <class T> struct VecT<T, 1> { using type = T; };

template <class T, int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS)
prod_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int P, int K, int Q, int nc,
            unsigned long long count, int G) {
    using R = Arith<T>;
    using V = typename VecT<T, VEC>::type;
    const int PQ = P * Q;
    const int PQv = PQ / VEC;
    const size_t strideA = (size_t)P * K, strideB = (size_t)K * Q;
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const int gc = left < (unsigned long long)G ? (int)left : G;
        const int total = gc * PQv;
        for (int l = threadIdx.x; l < total; l += THREADS) {
            const int il = l / PQv;
            const int e = (l - il * PQv) * VEC;
            const int p = e / Q;
            const int q = e - p * Q;
            const T* a = A + (g0 + il) * strideA + (size_t)p * K;
            const T* b = B + (g0 + il) * strideB + q;
            T acc[VEC];
            if (nc == 0) {
                const T av = __ldg(a);
                V bv = __ldg(reinterpret_cast<const V*>(b));
                #pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] = R::mul(av, reinterpret_cast<const T*>(&bv)[j]);
            } else {
                #pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] = R::zero();
                for (int k = 0; k < K; ++k) {
                    const T av = __ldg(a + k);
                    V bv = __ldg(reinterpret_cast<const V*>(b + (size_t)k * Q));
                    #pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j] = R::add(acc[j], R::mul(av, reinterpret_cast<const T*>(&bv)[j]));
                }
            }
            V out;
            #pragma unroll
            for (int j = 0; j < VEC; ++j) reinterpret_cast<T*>(&out)[j] = acc[j];
            *reinterpret_cast<V*>(C + (g0 + il) * (size_t)PQ + e) = out;
        }
    }
}

template <class T, int VEC>
static int prod_launch(const Shard& s, int P, int K, int Q, int nc, const void* A, const void* B, void* C) {
    constexpr int THREADS = 256;
    if (s.count == 0) return T5I_OK;
    const long long PQv = (long long)P * Q / VEC;
    int G = (int)((THREADS * 8 + PQv - 1) / PQv);
    if (G < 1) G = 1;
    if ((long long)G * PQv > 0x3fffffff) return T5I_UNSUPPORTED;
    const int grid = resident_grid((const void*)prod_kernel<T, VEC, THREADS>, THREADS, (s.count + G - 1) / G);
    prod_kernel<T, VEC, THREADS><<<grid, THREADS, 0, s.stream>>>((const T*)A, (const T*)B, (T*)C, P, K, Q, nc,
                                                                 (unsigned long long)s.count, G);
    return ok();
}

// ---- products of small records with ODD extents (7x7, 9x9 .. 19x19; any n for 4-byte elements) ----
// Between the specialised kernels (n <= 8, and the row-per-thread kernel for 16-byte-multiple rows) the per-output
// kernel above re-reads A and B through L1 with one thread per result element, unvectorised for odd n, and two
// run-time integer divisions each (tools/sweep_matmul.py: 0.13-0.3 of the HBM peak for n = 9..19).  Here a CTA
// copies G whole items of A and of B into shared memory as two contiguous, coalesced runs; a thread owns four
// adjacent results of one row: per k one A element (a broadcast for its neighbours) and four B elements (stride-1
// across the warp), accumulated in the oracle's order - k ascending, multiply then add - so results are
// bit-identical for all three element types; (item, row, column block) come from two multiply-highs.  1.5-1.9x the
// per-output kernel on those shapes (a 2 x 4 register block measured slower: the kernel is latency-, not
// LDS-bound); even extents of 8-byte elements and rectangular shapes measured no better and stay where they were.
template <class T, int JT, int THREADS>
__global__ void __launch_bounds__(THREADS)
stagedprod_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int m, int k, int n, int nb, unsigned magic_units,
                  unsigned magic_nb, unsigned long long count, int G) {
    using R = Arith<T>;
    extern __shared__ __align__(16) unsigned char sp_smem[];
    const unsigned dA = m * k, dB = k * n, dC = m * n, units = m * nb;
    constexpr unsigned V = 16 / sizeof(T);
    T* sA = reinterpret_cast<T*>(sp_smem);
    T* sB = sA + ((size_t)G * dA + V - 1) / V * V;              // 16-byte aligned like sA
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const unsigned gc = left < (unsigned long long)G ? (unsigned)left : (unsigned)G;
        const T* ga = A + (size_t)g0 * dA;
        const T* gb = B + (size_t)g0 * dB;
        // the two runs are contiguous and land as they are: 16-byte asynchronous copies when the group starts on a 16-byte
        // boundary (the host picks G as a multiple of the elements per chunk), one element per lane otherwise (21^3 Float
        // 1.20 -> 1.36 TB/s, 5x24x9 1.88 -> 2.10, 9x9x5 Double 3.24 -> 3.42; profiles/r02u_time_stagedprod.txt)
        if (((reinterpret_cast<uintptr_t>(ga) | reinterpret_cast<uintptr_t>(gb)) & 15) == 0) {
            const unsigned nA = gc * dA, nB = gc * dB, vA = nA / V, vB = nB / V;
            for (unsigned v = threadIdx.x; v < vA; v += THREADS) cp_async16(sA + v * V, ga + v * V);
            for (unsigned v = threadIdx.x; v < vB; v += THREADS) cp_async16(sB + v * V, gb + v * V);
            for (unsigned l = vA * V + threadIdx.x; l < nA; l += THREADS) sA[l] = __ldg(ga + l);
            for (unsigned l = vB * V + threadIdx.x; l < nB; l += THREADS) sB[l] = __ldg(gb + l);
            cp_async_commit();
            cp_async_wait<0>();
        } else {
            #pragma unroll 4
            for (unsigned l = threadIdx.x; l < gc * dA; l += THREADS) sA[l] = __ldg(ga + l);
            #pragma unroll 4
            for (unsigned l = threadIdx.x; l < gc * dB; l += THREADS) sB[l] = __ldg(gb + l);
        }
        __syncthreads();
        for (unsigned u = threadIdx.x; u < gc * units; u += THREADS) {
            const unsigned item = units == 1 ? u : __umulhi(u, magic_units), r = u - item * units;   // (a divisor of 1 has no 32-bit magic)
            const unsigned i = nb == 1 ? r : __umulhi(r, magic_nb), j0 = (r - i * nb) * JT;
            const T* a = sA + item * dA + i * k;
            const T* b = sB + item * dB + j0;
            T* c = C + (size_t)(g0 + item) * dC + (size_t)i * n + j0;
            T acc[JT];
            #pragma unroll
            for (int t = 0; t < JT; ++t) acc[t] = R::zero();
            if (j0 + JT <= (unsigned)n) {
                for (int kk = 0; kk < k; ++kk) {
                    const T av = a[kk];
                    #pragma unroll
                    for (int t = 0; t < JT; ++t) acc[t] = R::add(acc[t], R::mul(av, b[kk * n + t]));
                }
                #pragma unroll
                for (int t = 0; t < JT; ++t) c[t] = acc[t];
            } else {
                const int jt = n - (int)j0;
                for (int kk = 0; kk < k; ++kk) {
                    const T av = a[kk];
                    #pragma unroll
                    for (int t = 0; t < JT; ++t)
                        if (t < jt) acc[t] = R::add(acc[t], R::mul(av, b[kk * n + t]));
                }
                #pragma unroll
                for (int t = 0; t < JT; ++t)
                    if (t < jt) c[t] = acc[t];
            }
        }
        __syncthreads();
    }
}

template <class T>
static int stagedprod_launch(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
    constexpr int THREADS = 256, JT = 4;
    auto kern = stagedprod_kernel<T, JT, THREADS>;
    const size_t item_bytes = ((size_t)m * k + (size_t)k * n) * sizeof(T);
    const int nb = (n + JT - 1) / JT;
    const unsigned units = (unsigned)m * nb;
    long long G = (long long)((32 * 1024) / item_bytes);
    if (G * units > 65535) G = 65535 / units;           // the multiply-high divisions are exact below 2^16
    if (G < 2) return T5I_NOT_SPECIALISED;
    constexpr long long V = 16 / sizeof(T);
    if (G >= V) G = G / V * V;                          // every group starts on a 16-byte boundary: chunk-wise staging
    const int smem = (int)(G * item_bytes) + 16;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const size_t groups = (s.count + G - 1) / G, cap = (size_t)occ * sm_count();
    const int grid = (int)(groups < cap ? groups : cap);
    const unsigned mu = (unsigned)((0x100000000ull + units - 1) / units), mn = (unsigned)((0x100000000ull + (unsigned)nb - 1) / (unsigned)nb);
    kern<<<grid, THREADS, smem, s.stream>>>((const T*)A, (const T*)B, (T*)C, m, k, n, nb, mu, mn, (unsigned long long)s.count, (int)G);
    return ok();
}

int generic_stagedprod(const Shard& s, int dtype, int m, int k, int n, const void* A, const void* B, void* C) {
    if (s.layout != T5L_AOS || m < 5 || k < 5 || n < 5 || m > 24 || k > 24 || n > 24) return T5I_NOT_SPECIALISED;
    if (dtype != T5D_F32 && n % 2 == 0) return T5I_NOT_SPECIALISED;     // even rows of 8-byte elements: the vectorised per-output kernel ties
    if (s.count == 0) return T5I_OK;
    switch (dtype) {
        case T5D_F64: return stagedprod_launch<double>(s, m, k, n, A, B, C);
        case T5D_F32: return stagedprod_launch<float>(s, m, k, n, A, B, C);
        case T5D_I64: return stagedprod_launch<long long>(s, m, k, n, A, B, C);
    }
    return T5I_UNSUPPORTED;
}

// ---- long dot products: `dot` of n-vectors and `prod n` down to a scalar (BASELINE configs[3] with
// n = 2: 8x8 : 8x8 -> scalar), n > 4 ----
// The fold `acc = acc + x_j*y_j`, j ascending, is one sequential chain per item (the specification
// fixes the order, so lanes cannot share an item), but a thread that reads its own 2 x n elements
// from HBM touches 32 scattered sectors per warp load (the any-shape kernel: 0.30 / 0.15 of the HBM
// peak for 64-element Double / Float vectors).  A CTA therefore stages chunks of KC elements of TPB
// consecutive items into shared memory TRANSPOSED (coalesced global reads in runs of KC elements;
// element e of item t at s[e][t], pitch TPB+1: conflict-free both ways) and every thread folds its own
// column.  Same operations, same order: bit-identical for all three element types.
template <class T, int TPB, int KC>
__global__ void __launch_bounds__(TPB)
longdot_kernel(const T* __restrict__ X, const T* __restrict__ Y, T* __restrict__ out, int K, unsigned long long count) {
    using R = Arith<T>;
    __shared__ T sx[KC][TPB + 1];
    __shared__ T sy[KC][TPB + 1];
    const int tid = threadIdx.x;
    const unsigned long long n_tiles = (count + TPB - 1) / TPB;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long first = tile * TPB;
        const int nv = (count - first < (unsigned long long)TPB) ? (int)(count - first) : TPB;
        const T* gx = X + first * (size_t)K;
        const T* gy = Y + first * (size_t)K;
        T acc = R::zero();
        for (int k0 = 0; k0 < K; k0 += KC) {
            const int kc = (K - k0 < KC) ? K - k0 : KC;
            const int sq = TPB / kc, sr = TPB - sq * kc;
            int item = tid / kc, e = tid - item * kc;
            const int total = nv * kc;
            __syncthreads();                                   // the previous chunk has been folded
            #pragma unroll 8
            for (int idx = tid; idx < total; idx += TPB) {
                const size_t g = (size_t)item * K + k0 + e;
                sx[e][item] = __ldg(gx + g);
                sy[e][item] = __ldg(gy + g);
                item += sq; e += sr;
                if (e >= kc) { e -= kc; ++item; }
            }
            __syncthreads();
            if (tid < nv) {
                #pragma unroll 8
                for (int k = 0; k < kc; ++k) acc = R::add(acc, R::mul(sx[k][tid], sy[k][tid]));
            }
        }
        if (tid < nv) out[first + tid] = acc;
    }
}

// The same fold for vectors of whole 16-byte groups (the 64-element `dot`, `prod 2` of 8x8 : 8x8).  The
// kernel above moves one element per lane and instruction, visits every row once per chunk (a 256-byte Float row in two
// halves, microseconds apart) and its CTA alternates between loading and folding: 0.68 (Float) / 0.75 (Double) of the HBM
// peak on 64-element vectors.  Here the rows of TPB consecutive items arrive WHOLE - one contiguous run per operand, 16-byte
// cp.async straight into shared memory, no registers in between - at a pitch of an odd number of 16-byte units, so that each
// thread folds its own row with conflict-free 16-byte shared loads (a quarter-warp hits eight different bank groups).  One
// stage per CTA; the copies of the other resident CTAs overlap the fold.  Same operations in the same order: bit-identical.
// (Tried first: transposed staging fed from registers holding the next chunk, 16 vectors per thread - 126 registers, 4 CTAs
// per SM, 3.7 TB/s against 4.5 for the kernel above.)
template <class T, int TPB>
__global__ void __launch_bounds__(TPB)
longdot_rows_kernel(const T* __restrict__ X, const T* __restrict__ Y, T* __restrict__ out, int K, unsigned kvc_max, unsigned long long count) {
    using R = Arith<T>;
    constexpr unsigned V = 16 / sizeof(T);
    extern __shared__ __align__(16) unsigned char ldr_smem[];
    const unsigned kv = (unsigned)K / V, pitch = kvc_max | 1u;   // 16-byte units per row; staged pitch
    uint4* sx = reinterpret_cast<uint4*>(ldr_smem);
    uint4* sy = sx + TPB * pitch;
    const unsigned tid = threadIdx.x;
    const unsigned long long n_tiles = (count + TPB - 1) / TPB;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long first = tile * TPB;
        const unsigned nv = (count - first < (unsigned long long)TPB) ? (unsigned)(count - first) : (unsigned)TPB;
        const uint4* gx = reinterpret_cast<const uint4*>(X + first * (size_t)K);
        const uint4* gy = reinterpret_cast<const uint4*>(Y + first * (size_t)K);
        T acc = R::zero();
        for (unsigned c0 = 0; c0 < kv; c0 += kvc_max) {          // rows of up to 512 bytes: one pass
            const unsigned kvc = kv - c0 < kvc_max ? kv - c0 : kvc_max;
            const unsigned step_q = TPB / kvc, step_r = TPB - step_q * kvc;
            const unsigned total = nv * kvc;
            __syncthreads();                                   // the previous chunk has been folded
            unsigned item = tid / kvc, j = tid - item * kvc;
            for (unsigned v = tid; v < total; v += TPB) {
                const size_t g = (size_t)item * kv + c0 + j;
                cp_async16(sx + item * pitch + j, gx + g);
                cp_async16(sy + item * pitch + j, gy + g);
                item += step_q; j += step_r;
                if (j >= kvc) { j -= kvc; ++item; }
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            if (tid < nv) {
                const uint4* rx = sx + tid * pitch;
                const uint4* ry = sy + tid * pitch;
                #pragma unroll 4
                for (unsigned q = 0; q < kvc; ++q) {
                    const uint4 a = rx[q], b = ry[q];
                    #pragma unroll
                    for (unsigned u = 0; u < V; ++u) acc = R::add(acc, R::mul(reinterpret_cast<const T*>(&a)[u], reinterpret_cast<const T*>(&b)[u]));
                }
            }
        }
        if (tid < nv) out[first + tid] = acc;
    }
}

// (rows of more than 512 bytes go through the same staging in 512-byte pieces)
template <class T>
static int longdot_rows_launch(const Shard& s, int K, const void* X, const void* Y, void* O) {
    const size_t kv = (size_t)K * sizeof(T) / 16, kvc = kv <= 32 ? kv : 32, row = (kvc | 1) * 16;
    const bool wide = 2 * 128 * row <= 24 * 1024;              // short rows: 128 items per tile
    const int tpb = wide ? 128 : 64;
    const size_t smem = 2 * (size_t)tpb * row;
    const void* kern = wide ? (const void*)longdot_rows_kernel<T, 128> : (const void*)longdot_rows_kernel<T, 64>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return T5I_CUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, tpb, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const size_t tiles = (s.count + tpb - 1) / tpb, cap = (size_t)occ * sm_count();
    const int grid = (int)(tiles < cap ? tiles : cap);
    if (wide) longdot_rows_kernel<T, 128><<<grid, 128, smem, s.stream>>>((const T*)X, (const T*)Y, (T*)O, K, (unsigned)kvc, s.count);
    else longdot_rows_kernel<T, 64><<<grid, 64, smem, s.stream>>>((const T*)X, (const T*)Y, (T*)O, K, (unsigned)kvc, s.count);
    return ok();
}

int generic_longdot(const Shard& s, int dtype, int K, const void* X, const void* Y, void* O) {
    constexpr int TPB = 128, KC8 = 16, KC4 = 32;    // 128-byte runs per item and chunk; 33 KB of static shared memory
    if (s.layout != T5L_AOS || K < 1) return T5I_NOT_SPECIALISED;
    if (s.count == 0) return T5I_OK;
    const size_t tiles = (s.count + TPB - 1) / TPB;
    const size_t row_bytes = (size_t)K * (dtype == T5D_F32 ? 4 : 8);
    if (row_bytes % 16 == 0 && (((uintptr_t)X | (uintptr_t)Y) & 15) == 0) {
        switch (dtype) {
            case T5D_F64: return longdot_rows_launch<double>(s, K, X, Y, O);
            case T5D_F32: return longdot_rows_launch<float>(s, K, X, Y, O);
            case T5D_I64: return longdot_rows_launch<long long>(s, K, X, Y, O);
            default: return T5I_UNSUPPORTED;
        }
    }
    switch (dtype) {
        case T5D_F64: longdot_kernel<double, TPB, KC8><<<resident_grid((const void*)longdot_kernel<double, TPB, KC8>, TPB, tiles), TPB, 0, s.stream>>>((const double*)X, (const double*)Y, (double*)O, K, s.count); break;
        case T5D_F32: longdot_kernel<float, TPB, KC4><<<resident_grid((const void*)longdot_kernel<float, TPB, KC4>, TPB, tiles), TPB, 0, s.stream>>>((const float*)X, (const float*)Y, (float*)O, K, s.count); break;
        case T5D_I64: longdot_kernel<long long, TPB, KC8><<<resident_grid((const void*)longdot_kernel<long long, TPB, KC8>, TPB, tiles), TPB, 0, s.stream>>>((const long long*)X, (const long long*)Y, (long long*)O, K, s.count); break;
        default: return T5I_UNSUPPORTED;
    }
    return ok();
}

// ---- outer product (prod 0, the rank-r (x) rank-s tensor product of BASELINE configs[3]) ----
// Write-bound: 2 x 512 B in, 32 KiB out per 8x8 (x) 8x8 item.  A CTA walks items; thread
// (tp, tq) keeps its B vector in registers and streams rows p = tp, tp+R, ... of the
// P x Q result: one broadcast load of A[p], VEC multiplies, one 16-B streaming store.  No
// index division in the loop; every warp store instruction covers 512 contiguous bytes.
template <class T, int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS)
outer_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int P, int Q, unsigned long long count,
             int items_per_cta) {
    using R = Arith<T>;
    using V = typename VecT<T, VEC>::type;
    const int Qv = Q / VEC;
    const int rows = THREADS / Qv;          // host guarantees THREADS % Qv == 0
    const int tq = threadIdx.x % Qv, tp = threadIdx.x / Qv;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * items_per_cta;
    const unsigned long long i1 = i0 + items_per_cta < count ? i0 + items_per_cta : count;
    for (unsigned long long item = i0; item < i1; ++item) {
        const V bv = __ldg(reinterpret_cast<const V*>(B + item * (size_t)Q) + tq);
        const T* a = A + item * (size_t)P;
        T* c = C + item * (size_t)P * Q + (size_t)tq * VEC;
        int p = tp;
        for (; p + 3 * rows < P; p += 4 * rows) {
            T av[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) av[u] = __ldg(a + p + u * rows);
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint4 o;
                #pragma unroll
                for (int j = 0; j < VEC; ++j) reinterpret_cast<T*>(&o)[j] = R::mul(av[u], reinterpret_cast<const T*>(&bv)[j]);
                st_global_stream16(c + (size_t)(p + u * rows) * Q, o);
            }
        }
        for (; p < P; p += rows) {
            const T av = __ldg(a + p);
            uint4 o;
            #pragma unroll
            for (int j = 0; j < VEC; ++j) reinterpret_cast<T*>(&o)[j] = R::mul(av, reinterpret_cast<const T*>(&bv)[j]);
            st_global_stream16(c + (size_t)p * Q, o);
        }
    }
}

template <class T, int VEC>
static int outer_launch(const Shard& s, int P, int Q, const void* A, const void* B, void* C) {
    constexpr int THREADS = 256;
    if (s.count == 0) return T5I_OK;
    // a CTA streams ~512 KiB of results: long enough to amortise its launch, short enough that
    // the hardware block scheduler evens out the SMs
    long long per = (512ll << 10) / ((long long)P * Q * (long long)sizeof(T));
    if (per < 1) per = 1;
    if (per > 64) per = 64;
    size_t blocks = (s.count + per - 1) / per;
    if (blocks > 0x7fffffffull) return T5I_UNSUPPORTED;
    outer_kernel<T, VEC, THREADS><<<(int)blocks, THREADS, 0, s.stream>>>((const T*)A, (const T*)B, (T*)C, P, Q,
                                                                        (unsigned long long)s.count, (int)per);
    return ok();
}

static int gather_geometry(const void* kern, size_t count, size_t d_out, int threads, int* G);

// Small outer products (3 (x) 3, 4 (x) 4, 3x3 (x) 3x3 -> 3^4, ...): the results of G consecutive items form
// one contiguous run, thread l of the run writes output l (coalesced for any P, Q, odd ones included) and
// finds its (item, p, q) by constant steps and one multiply-high - the any-shape kernel spent two run-time
// integer divisions per output (0.29 / 0.18 of the HBM peak for 4Mi 3x3 (x) 3x3 Double / Float).  The
// operands are re-read through L1 (they are 1/P + 1/Q of the output bytes).
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
outer_small_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, unsigned P, unsigned Q, unsigned magicQ,
                   unsigned long long count, int G) {
    using R = Arith<T>;
    const unsigned PQ = P * Q;
    const unsigned step_q = THREADS / PQ, step_r = THREADS - step_q * PQ;
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const unsigned gc = left < (unsigned long long)G ? (unsigned)left : (unsigned)G;
        const unsigned total = gc * PQ;
        unsigned il = threadIdx.x / PQ, e = threadIdx.x - il * PQ;
        const T* a = A + (size_t)g0 * P;
        const T* b = B + (size_t)g0 * Q;
        T* c = C + (size_t)g0 * PQ;
        #pragma unroll 4
        for (unsigned l = threadIdx.x; l < total; l += THREADS) {
            const unsigned p = __umulhi(e, magicQ);            // e / Q, exact for e < 2^16 (host checks P*Q)
            const unsigned q = e - p * Q;
            c[l] = R::mul(__ldg(a + (size_t)il * P + p), __ldg(b + (size_t)il * Q + q));
            il += step_q; e += step_r;
            if (e >= PQ) { e -= PQ; ++il; }
        }
    }
}

template <class T>
static int outer_small_launch(const Shard& s, int P, int Q, const void* A, const void* B, void* C) {
    constexpr int THREADS = 256;
    if (s.count == 0) return T5I_OK;
    int G;
    const int grid = gather_geometry((const void*)outer_small_kernel<T, THREADS>, s.count, (size_t)P * Q, THREADS, &G);
    const unsigned magic = (unsigned)((0x100000000ull + (unsigned)Q - 1) / (unsigned)Q);
    outer_small_kernel<T, THREADS><<<grid, THREADS, 0, s.stream>>>((const T*)A, (const T*)B, (T*)C, (unsigned)P, (unsigned)Q, magic,
                                                                   (unsigned long long)s.count, G);
    return ok();
}

int generic_soa_prod(const Shard& s, int dtype, int P, int K, int Q, int nc, const void* A, const void* B, void* C);

// the streaming outer-product kernel applies when a CTA's threads tile whole rows of the result
static bool outer_ok(int P, int Q, int vec) {
    if (Q % vec) return false;
    const int Qv = Q / vec;
    return Qv <= 256 && 256 % Qv == 0 && (long long)P * Qv >= 256;
}

int generic_prod(const Shard& s, int dtype, int P, int K, int Q, int nc, const void* A, const void* B, void* C) {
    if (s.layout == T5L_SOA) return generic_soa_prod(s, dtype, P, K, Q, nc, A, B, C);
    if (s.layout != T5L_AOS) return T5I_UNSUPPORTED;
    if ((long long)P * Q > (1 << 28) || (long long)P * K > (1 << 28) || (long long)K * Q > (1 << 28)) return T5I_UNSUPPORTED;
    if (nc == 0) {
        if (dtype == T5D_F64 && outer_ok(P, Q, 2)) return outer_launch<double, 2>(s, P, Q, A, B, C);
        if (dtype == T5D_I64 && outer_ok(P, Q, 2)) return outer_launch<long long, 2>(s, P, Q, A, B, C);
        if (dtype == T5D_F32 && outer_ok(P, Q, 4)) return outer_launch<float, 4>(s, P, Q, A, B, C);
        if ((long long)P * Q < 65536 && Q > 1) {
            if (dtype == T5D_F64) return outer_small_launch<double>(s, P, Q, A, B, C);
            if (dtype == T5D_I64) return outer_small_launch<long long>(s, P, Q, A, B, C);
            if (dtype == T5D_F32) return outer_small_launch<float>(s, P, Q, A, B, C);
        }
    }
    switch (dtype) {
        case T5D_F64: return (Q % 2 == 0) ? prod_launch<double, 2>(s, P, K, Q, nc, A, B, C) : prod_launch<double, 1>(s, P, K, Q, nc, A, B, C);
        case T5D_I64: return (Q % 2 == 0) ? prod_launch<long long, 2>(s, P, K, Q, nc, A, B, C) : prod_launch<long long, 1>(s, P, K, Q, nc, A, B, C);
        case T5D_F32: return (Q % 4 == 0) ? prod_launch<float, 4>(s, P, K, Q, nc, A, B, C) : prod_launch<float, 1>(s, P, K, Q, nc, A, B, C);
    }
    return T5I_UNSUPPORTED;
}

// =====================================================================================
// transpose, any rank: out[b][i] = in[b][perm[i]] with perm = transposeV dims
// (src/Data/Tensor/Vector.hs:306-307) tabulated once per shape by the host layer.
// =====================================================================================
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
gather_kernel(const T* __restrict__ in, T* __restrict__ out, const uint32_t* __restrict__ table, unsigned d_in,
              unsigned d_out, unsigned long long count, int G) {
    const unsigned step_q = THREADS / d_out, step_r = THREADS - step_q * d_out;
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const unsigned gc = left < (unsigned long long)G ? (unsigned)left : (unsigned)G;
        const unsigned total = gc * d_out;
        // (item, element) of this thread's l-th output advance by constant steps: no division in the loop
        unsigned il = threadIdx.x / d_out, i = threadIdx.x - il * d_out;
        const T* src = in + (size_t)g0 * d_in;
        T* dst = out + (size_t)g0 * d_out;
        #pragma unroll 4
        for (unsigned l = threadIdx.x; l < total; l += THREADS) {
            dst[l] = __ldg(src + (size_t)il * d_in + __ldg(table + i));
            il += step_q; i += step_r;
            if (i >= d_out) { i -= d_out; ++il; }
        }
    }
}

// two-source gather (append): table value v < dA reads A[v], else B[v - dA]
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
gather2_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ out, const uint32_t* __restrict__ table,
               unsigned dA, unsigned dB, unsigned d_out, unsigned long long count, int G) {
    const unsigned step_q = THREADS / d_out, step_r = THREADS - step_q * d_out;
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const unsigned gc = left < (unsigned long long)G ? (unsigned)left : (unsigned)G;
        const unsigned total = gc * d_out;
        unsigned il = threadIdx.x / d_out, i = threadIdx.x - il * d_out;
        const T* srcA = A + (size_t)g0 * dA;
        const T* srcB = B + (size_t)g0 * dB;
        T* dst = out + (size_t)g0 * d_out;
        #pragma unroll 4
        for (unsigned l = threadIdx.x; l < total; l += THREADS) {
            const unsigned v = __ldg(table + i);
            dst[l] = v < dA ? __ldg(srcA + (size_t)il * dA + v) : __ldg(srcB + (size_t)il * dB + (v - dA));
            il += step_q; i += step_r;
            if (i >= d_out) { i -= d_out; ++il; }
        }
    }
}

// Staged form of the gathers above for records of up to a few KB (append / split / at / ta / slice / axis permutations of
// tiny tensors): the direct kernels move one element per lane and instruction - 128-byte warp accesses for Float, and source
// reads that follow the table - and reached 0.55-0.83 of the HBM peak.  Here G consecutive items arrive as ONE contiguous
// run per operand (16-byte cp.async), the index tables sit in shared memory, and every result leaves as 16-byte vectors of
// its own contiguous run, each lane collecting its 2 / 4 elements from the staged records.  Up to two sources (table value
// v < dA: A[v], else B[v - dA]) and two results (split: both halves from one read of the input).
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
staged_gather_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ O1, T* __restrict__ O2,
                     const uint32_t* __restrict__ t1, const uint32_t* __restrict__ t2, unsigned dA, unsigned dB, unsigned d1, unsigned d2,
                     unsigned long long count, unsigned G, unsigned offB, unsigned offT) {
    constexpr unsigned V = 16 / sizeof(T);
    struct alignas(16) Vec { T v[V]; };
    extern __shared__ __align__(16) unsigned char sg_smem[];
    T* sA = reinterpret_cast<T*>(sg_smem);
    T* sB = sA + offB;
    uint32_t* sT1 = reinterpret_cast<uint32_t*>(sA + offT);
    uint32_t* sT2 = sT1 + d1;
    for (unsigned i = threadIdx.x; i < d1; i += THREADS) sT1[i] = __ldg(t1 + i);
    for (unsigned i = threadIdx.x; i < d2; i += THREADS) sT2[i] = __ldg(t2 + i);
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const unsigned gc = left < (unsigned long long)G ? (unsigned)left : G;
        __syncthreads();                                     // the previous group has left shared memory (and the tables are in)
        #pragma unroll
        for (int op = 0; op < 2; ++op) {
            const unsigned dd = op ? dB : dA;
            if (dd == 0) continue;
            const T* src = (op ? B : A) + (size_t)g0 * dd;   // g0 % V == 0: 16-byte aligned
            T* dst = op ? sB : sA;
            const unsigned total = gc * dd, nvec = total / V;
            for (unsigned l = threadIdx.x; l < nvec; l += THREADS) cp_async16(dst + l * V, src + l * V);
            for (unsigned l = nvec * V + threadIdx.x; l < total; l += THREADS) dst[l] = __ldg(src + l);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        #pragma unroll
        for (int res = 0; res < 2; ++res) {
            const unsigned dd = res ? d2 : d1;
            if (dd == 0) continue;
            const uint32_t* tab = res ? sT2 : sT1;
            T* out = (res ? O2 : O1) + (size_t)g0 * dd;
            const unsigned total = gc * dd, nvec = total / V;
            const unsigned step_q = (THREADS * V) / dd, step_r = THREADS * V - step_q * dd;
            unsigned il = (threadIdx.x * V) / dd, i = threadIdx.x * V - il * dd;
            for (unsigned lv = threadIdx.x; lv < nvec; lv += THREADS) {
                Vec o;
                unsigned il2 = il, i2 = i;
                #pragma unroll
                for (unsigned u = 0; u < V; ++u) {
                    const unsigned v = tab[i2];
                    o.v[u] = v < dA ? sA[il2 * dA + v] : sB[il2 * dB + (v - dA)];
                    if (++i2 == dd) { i2 = 0; ++il2; }
                }
                *reinterpret_cast<Vec*>(out + (size_t)lv * V) = o;
                il += step_q; i += step_r;
                if (i >= dd) { i -= dd; ++il; }
            }
            for (unsigned l = nvec * V + threadIdx.x; l < total; l += THREADS) {      // (< V elements: the last group of an odd batch)
                const unsigned il2 = l / dd, v = tab[l - il2 * dd];
                out[l] = v < dA ? sA[il2 * dA + v] : sB[il2 * dB + (v - dA)];
            }
        }
    }
}

// Records that can be staged: V of them in 24 KB, tables of at most 2048 entries.  Measured on 4Mi items (TB/s, direct ->
// staged): Float append 3x3|3 4.4 -> 5.4, split 4x4 3.7 -> 5.2, split 8x8 3.8 -> 5.0, permutations of 2x3x4 5.1 -> 5.3; Double
// split 4x4 3.9 -> 5.4, split 8x8 3.7 -> 5.3 - but Double append 5.4 -> 5.3, permutations 5.9 -> 5.7, transpose 8x8 6.4 -> 5.5, and
// picks of a third or a quarter of the record (at / ta / slice) 2-6 % slower in both types: the direct gather reads only the
// sectors it needs.  So: 4-byte elements when the results are at least half the record, 8-byte elements for the split only.
bool staged_gather_fits(int elem_bytes, size_t d_in_total, size_t d_out_total, bool two_results) {
    if (elem_bytes != 4 && elem_bytes != 8) return false;
    if (elem_bytes == 8 && !two_results) return false;
    const size_t V = 16 / elem_bytes;
    return d_in_total > 0 && d_out_total <= 2048 && d_in_total * elem_bytes * V <= 24 * 1024 && d_out_total * 2 >= d_in_total;
}

// T5I_NOT_SPECIALISED when the records are too large to stage (the direct kernels take them)
int staged_gather(const Shard& s, int elem_bytes, size_t dA, size_t dB, size_t d1, size_t d2, const uint32_t* t1, const uint32_t* t2,
                  const void* A, const void* B, void* O1, void* O2) {
    constexpr int THREADS = 256;
    if (s.layout != T5L_AOS || !staged_gather_fits(elem_bytes, dA + dB, d1 + d2, d1 && d2)) return T5I_NOT_SPECIALISED;
    if (s.count == 0 || d1 + d2 == 0) return T5I_OK;
    const size_t V = 16 / elem_bytes, item = (dA + dB) * elem_bytes;
    if ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)O1 | (uintptr_t)O2) & 15) != 0) return T5I_NOT_SPECIALISED;
    size_t G = (24 * 1024) / item / V * V;
    const size_t want = (size_t)THREADS * V * 4;                       // ~4 result vectors per thread and group are plenty
    const size_t dmin = d1 == 0 ? d2 : d2 == 0 ? d1 : (d2 < d1 ? d2 : d1);
    if (G * dmin > want) G = ((want + dmin - 1) / dmin + V - 1) / V * V;
    if (G < V) G = V;
    const size_t offB = (G * dA + V - 1) / V * V, offT = offB + (G * dB + V - 1) / V * V;
    const size_t smem = offT * elem_bytes + (d1 + d2) * 4;
    const void* kern = elem_bytes == 8 ? (const void*)staged_gather_kernel<unsigned long long, THREADS> : (const void*)staged_gather_kernel<unsigned, THREADS>;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const size_t groups = (s.count + G - 1) / G, cap = (size_t)occ * sm_count();
    const int grid = (int)(groups < cap ? groups : cap);
    if (elem_bytes == 8)
        staged_gather_kernel<unsigned long long, THREADS><<<grid, THREADS, smem, s.stream>>>((const unsigned long long*)A, (const unsigned long long*)B,
            (unsigned long long*)O1, (unsigned long long*)O2, t1, t2, (unsigned)dA, (unsigned)dB, (unsigned)d1, (unsigned)d2, s.count, (unsigned)G, (unsigned)offB, (unsigned)offT);
    else
        staged_gather_kernel<unsigned, THREADS><<<grid, THREADS, smem, s.stream>>>((const unsigned*)A, (const unsigned*)B,
            (unsigned*)O1, (unsigned*)O2, t1, t2, (unsigned)dA, (unsigned)dB, (unsigned)d1, (unsigned)d2, s.count, (unsigned)G, (unsigned)offB, (unsigned)offT);
    return ok();
}

static int gather_geometry(const void* kern, size_t count, size_t d_out, int threads, int* G) {
    int g = (int)((threads * 16 + d_out - 1) / d_out);
    if (g < 1) g = 1;
    *G = g;
    return resident_grid(kern, threads, (count + g - 1) / g);
}

int generic_gather(const Shard& s, int elem_bytes, size_t d_in, size_t d_out, const uint32_t* table_dev, const void* A, void* O) {
    constexpr int THREADS = 256;
    if (s.layout != T5L_AOS) return T5I_UNSUPPORTED;
    if (s.count == 0 || d_out == 0) return T5I_OK;
    if (d_in > (1u << 28) || d_out > (1u << 28)) return T5I_UNSUPPORTED;
    {
        const int r = staged_gather(s, elem_bytes, d_in, 0, d_out, 0, table_dev, nullptr, A, nullptr, O, nullptr);
        if (r != T5I_NOT_SPECIALISED) return r;
    }
    int G;
    const int grid = gather_geometry(elem_bytes == 8 ? (const void*)gather_kernel<unsigned long long, THREADS>
                                                     : (const void*)gather_kernel<unsigned, THREADS>, s.count, d_out, THREADS, &G);
    if (elem_bytes == 8)
        gather_kernel<unsigned long long, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned long long*)A, (unsigned long long*)O, table_dev, (unsigned)d_in, (unsigned)d_out, s.count, G);
    else if (elem_bytes == 4)
        gather_kernel<unsigned, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned*)A, (unsigned*)O, table_dev, (unsigned)d_in, (unsigned)d_out, s.count, G);
    else
        return T5I_UNSUPPORTED;
    return ok();
}

int generic_gather2(const Shard& s, int elem_bytes, size_t dA, size_t dB, size_t d_out, const uint32_t* table_dev,
                    const void* A, const void* B, void* O) {
    constexpr int THREADS = 256;
    if (s.layout != T5L_AOS) return T5I_UNSUPPORTED;
    if (s.count == 0 || d_out == 0) return T5I_OK;
    if (dA + dB > (1u << 28) || d_out > (1u << 28)) return T5I_UNSUPPORTED;
    {
        const int r = staged_gather(s, elem_bytes, dA, dB, d_out, 0, table_dev, nullptr, A, B, O, nullptr);
        if (r != T5I_NOT_SPECIALISED) return r;
    }
    int G;
    const int grid = gather_geometry(elem_bytes == 8 ? (const void*)gather2_kernel<unsigned long long, THREADS>
                                                     : (const void*)gather2_kernel<unsigned, THREADS>, s.count, d_out, THREADS, &G);
    if (elem_bytes == 8)
        gather2_kernel<unsigned long long, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned long long*)A, (const unsigned long long*)B, (unsigned long long*)O, table_dev, (unsigned)dA, (unsigned)dB, (unsigned)d_out, s.count, G);
    else if (elem_bytes == 4)
        gather2_kernel<unsigned, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned*)A, (const unsigned*)B, (unsigned*)O, table_dev, (unsigned)dA, (unsigned)dB, (unsigned)d_out, s.count, G);
    else
        return T5I_UNSUPPORTED;
    return ok();
}

// one-source, two-destination scatter (split): every input element belongs to exactly one of the two
// results, so the input is read ONCE, coalesced (two gathers read it twice: 384 B of traffic for the
// 256 algorithmic bytes of a 4x4 -> 4x3 | 4x1 Double split).  table[e] = offset in the first result, or
// 0x80000000 | offset in the second.
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
scatter2_kernel(const T* __restrict__ in, T* __restrict__ out1, T* __restrict__ out2, const uint32_t* __restrict__ table,
                unsigned d_in, unsigned d1, unsigned d2, unsigned long long count, int G) {
    const unsigned step_q = THREADS / d_in, step_r = THREADS - step_q * d_in;
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const unsigned gc = left < (unsigned long long)G ? (unsigned)left : (unsigned)G;
        const unsigned total = gc * d_in;
        unsigned il = threadIdx.x / d_in, i = threadIdx.x - il * d_in;
        const T* src = in + (size_t)g0 * d_in;
        T* dst1 = out1 + (size_t)g0 * d1;
        T* dst2 = out2 + (size_t)g0 * d2;
        #pragma unroll 4
        for (unsigned l = threadIdx.x; l < total; l += THREADS) {
            const unsigned v = __ldg(table + i);
            const T x = __ldg(src + l);
            if (v & 0x80000000u) dst2[(size_t)il * d2 + (v & 0x7fffffffu)] = x;
            else dst1[(size_t)il * d1 + v] = x;
            il += step_q; i += step_r;
            if (i >= d_in) { i -= d_in; ++il; }
        }
    }
}

int generic_scatter2(const Shard& s, int elem_bytes, size_t d_in, size_t d1, size_t d2, const uint32_t* table_dev,
                     const void* A, void* O1, void* O2) {
    constexpr int THREADS = 256;
    if (s.layout != T5L_AOS) return T5I_UNSUPPORTED;
    if (s.count == 0 || d_in == 0) return T5I_OK;
    if (d_in > (1u << 28)) return T5I_UNSUPPORTED;
    int G;
    const int grid = gather_geometry(elem_bytes == 8 ? (const void*)scatter2_kernel<unsigned long long, THREADS>
                                                     : (const void*)scatter2_kernel<unsigned, THREADS>, s.count, d_in, THREADS, &G);
    if (elem_bytes == 8)
        scatter2_kernel<unsigned long long, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned long long*)A, (unsigned long long*)O1, (unsigned long long*)O2, table_dev, (unsigned)d_in, (unsigned)d1, (unsigned)d2, s.count, G);
    else if (elem_bytes == 4)
        scatter2_kernel<unsigned, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned*)A, (unsigned*)O1, (unsigned*)O2, table_dev, (unsigned)d_in, (unsigned)d1, (unsigned)d2, s.count, G);
    else
        return T5I_UNSUPPORTED;
    return ok();
}

// ---- matrix transpose for extents >= 16 (a4: rank-2 index reversal) -------------------------------
// The table-driven gather reads a column of the source per output row: for 16x16 Double records that is
// 128-byte strides (0.59 of the HBM peak), for 128x128 ones 1-KB strides.  Large matrices: a CTA moves
// 32 x 32 sub-blocks through a padded shared tile, row segments contiguous on both sides.
template <class T, int TILE, int ROWS>
__global__ void __launch_bounds__(TILE * ROWS)
transpose2d_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int C, int tr, int tc, unsigned long long total_tiles) {
    __shared__ T tile[TILE][TILE + 1];
    const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
    const unsigned per_item = (unsigned)(tr * tc);
    const size_t d = (size_t)R * C;
    for (unsigned long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const unsigned long long item = t / per_item;
        const unsigned rem = (unsigned)(t - item * per_item);
        const int ti = rem / tc, tj = rem - ti * tc;
        const T* src = in + item * d;
        T* dst = out + item * d;
        {
            const int x = tj * TILE + tx, y0 = ti * TILE;
            #pragma unroll
            for (int j = ty; j < TILE; j += ROWS)
                if (x < C && y0 + j < R) tile[j][tx] = __ldg(src + (size_t)(y0 + j) * C + x);
        }
        __syncthreads();
        {
            const int x = ti * TILE + tx, y0 = tj * TILE;
            #pragma unroll
            for (int j = ty; j < TILE; j += ROWS)
                if (x < R && y0 + j < C) dst[(size_t)(y0 + j) * R + x] = tile[tx][j];
        }
        __syncthreads();
    }
}

template <class T, int TILE, int ROWS>
static int transpose2d_launch(const Shard& s, int R, int C, const void* A, void* O) {
    const int tr = (R + TILE - 1) / TILE, tc = (C + TILE - 1) / TILE;
    const unsigned long long total = (unsigned long long)s.count * tr * tc;
    const int grid = resident_grid((const void*)transpose2d_kernel<T, TILE, ROWS>, TILE * ROWS, total);
    transpose2d_kernel<T, TILE, ROWS><<<grid, TILE * ROWS, 0, s.stream>>>((const T*)A, (T*)O, R, C, tr, tc, total);
    return ok();
}

// 4-byte elements, even extents from 48 up: 64 x 64 sub-blocks and PAIRS of elements per global access on both sides (a warp
// instruction moves 256 contiguous bytes instead of 128: 128x128 Float 4.86 -> 6.26 TB/s).  Thread (ty, tx) loads the pair
// (row y, columns 2tx, 2tx + 1) and stores the pair (output row x, output columns 2tx, 2tx + 1) = source rows 2tx, 2tx + 1 of
// source column x; the tile pitch is odd, so both shared-memory patterns are conflict-free.
template <class T, class T2, int ROWS>
__global__ void __launch_bounds__(32 * ROWS)
transpose2d_pair_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int C, int tr, int tc, unsigned long long total_tiles) {
    constexpr int TILE = 64;
    __shared__ T tile[TILE][TILE + 1];
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const unsigned per_item = (unsigned)(tr * tc);
    const size_t d = (size_t)R * C;
    for (unsigned long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const unsigned long long item = t / per_item;
        const unsigned rem = (unsigned)(t - item * per_item);
        const int ti = rem / tc, tj = rem - ti * tc;
        const T* src = in + item * d;
        T* dst = out + item * d;
        {
            const int x = tj * TILE + 2 * tx, y0 = ti * TILE;
            #pragma unroll
            for (int j = ty; j < TILE; j += ROWS)
                if (x < C && y0 + j < R) {
                    const T2 v = __ldg(reinterpret_cast<const T2*>(src + (size_t)(y0 + j) * C + x));
                    tile[j][2 * tx] = v.x;
                    tile[j][2 * tx + 1] = v.y;
                }
        }
        __syncthreads();
        {
            const int x = ti * TILE + 2 * tx, y0 = tj * TILE;        // output row y0 + j = source column, output columns x, x + 1 = source rows
            #pragma unroll
            for (int j = ty; j < TILE; j += ROWS)
                if (x < R && y0 + j < C) {
                    T2 v;
                    v.x = tile[2 * tx][j];
                    v.y = tile[2 * tx + 1][j];
                    *reinterpret_cast<T2*>(dst + (size_t)(y0 + j) * R + x) = v;
                }
        }
        __syncthreads();
    }
}

template <class T, class T2>
static int transpose2d_pair_launch(const Shard& s, int R, int C, const void* A, void* O) {
    constexpr int ROWS = 8;
    const int tr = (R + 63) / 64, tc = (C + 63) / 64;
    const unsigned long long total = (unsigned long long)s.count * tr * tc;
    const int grid = resident_grid((const void*)transpose2d_pair_kernel<T, T2, ROWS>, 32 * ROWS, total);
    transpose2d_pair_kernel<T, T2, ROWS><<<grid, 32 * ROWS, 0, s.stream>>>((const T*)A, (T*)O, R, C, tr, tc, total);
    return ok();
}

// Matrices of a few KB (16x16 .. ~45x45 Double): whole items are staged instead - G consecutive items
// arrive as one contiguous run (coalesced), rows at an odd pitch; the result leaves as one contiguous
// run too, each thread reading down a source column (stride = the odd pitch: conflict-free).  Index
// arithmetic: constant steps + one multiply-high per element, no division.
// VEC = 2 (4-byte elements, R and C even): global accesses are 8-byte pairs - a pair never straddles a row on
// either side - so a warp instruction still moves 256 contiguous bytes.
template <class T, int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS)
transpose_staged_kernel(const T* __restrict__ in, T* __restrict__ out, unsigned R, unsigned C, unsigned pitch, unsigned magicC,
                        unsigned magicR, unsigned long long count, int G) {
    extern __shared__ __align__(16) unsigned char tsm[];
    T* sm = reinterpret_cast<T*>(tsm);
    struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
    const unsigned d = R * C, item_pitch = R * pitch;
    const unsigned dv = d / VEC;                                   // vectors per item
    const unsigned step_q = THREADS / dv, step_r = THREADS - step_q * dv;
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * G; g0 < count; g0 += (unsigned long long)gridDim.x * G) {
        const unsigned long long left = count - g0;
        const unsigned gc = left < (unsigned long long)G ? (unsigned)left : (unsigned)G;
        const unsigned total = gc * dv;
        const Pack* src = reinterpret_cast<const Pack*>(in + (size_t)g0 * d);
        Pack* dst = reinterpret_cast<Pack*>(out + (size_t)g0 * d);
        unsigned il = threadIdx.x / dv, ev = threadIdx.x - il * dv;
        #pragma unroll 4
        for (unsigned l = threadIdx.x; l < total; l += THREADS) {
            const unsigned e = ev * VEC;
            const unsigned i = __umulhi(e, magicC), j = e - i * C;          // e = i*C + j
            const Pack p = src[l];
            #pragma unroll
            for (int u = 0; u < VEC; ++u) sm[il * item_pitch + i * pitch + j + u] = p.v[u];
            il += step_q; ev += step_r;
            if (ev >= dv) { ev -= dv; ++il; }
        }
        __syncthreads();
        il = threadIdx.x / dv; ev = threadIdx.x - il * dv;
        #pragma unroll 4
        for (unsigned l = threadIdx.x; l < total; l += THREADS) {
            const unsigned e = ev * VEC;
            const unsigned j = __umulhi(e, magicR), i = e - j * R;          // output e = j*R + i
            Pack p;
            #pragma unroll
            for (int u = 0; u < VEC; ++u) p.v[u] = sm[il * item_pitch + (i + u) * pitch + j];
            dst[l] = p;
            il += step_q; ev += step_r;
            if (ev >= dv) { ev -= dv; ++il; }
        }
        __syncthreads();
    }
}

template <class T, int VEC>
static int transpose_staged_launch(const Shard& s, int R, int C, const void* A, void* O) {
    constexpr int THREADS = 256;
    auto kern = transpose_staged_kernel<T, VEC, THREADS>;
    const unsigned pitch = (unsigned)C | 1u;
    const size_t item_bytes = (size_t)R * pitch * sizeof(T);
    int G = (int)((32 * 1024) / item_bytes);
    if (G < 1) G = 1;
    const int smem = (int)(G * item_bytes);
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return T5I_CUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const size_t groups = (s.count + G - 1) / G, cap = (size_t)occ * sm_count();
    const int grid = (int)(groups < cap ? groups : cap);
    const unsigned mC = (unsigned)((0x100000000ull + (unsigned)C - 1) / (unsigned)C), mR = (unsigned)((0x100000000ull + (unsigned)R - 1) / (unsigned)R);
    kern<<<grid, THREADS, smem, s.stream>>>((const T*)A, (T*)O, (unsigned)R, (unsigned)C, pitch, mC, mR, (unsigned long long)s.count, G);
    return ok();
}

int generic_transpose2d(const Shard& s, int elem_bytes, int R, int C, const void* A, void* O) {
    if (s.layout != T5L_AOS || (R < 16 && C < 16)) return T5I_NOT_SPECIALISED;   // smaller: the table gather is at 0.95+
    if (elem_bytes != 4 && elem_bytes != 8) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    if ((size_t)R * ((unsigned)C | 1u) * elem_bytes <= 16 * 1024 && (size_t)R * C < 65536)
        return elem_bytes == 8 ? transpose_staged_launch<unsigned long long, 1>(s, R, C, A, O)
             : (R % 2 == 0 && C % 2 == 0) ? transpose_staged_launch<unsigned, 2>(s, R, C, A, O) : transpose_staged_launch<unsigned, 1>(s, R, C, A, O);
    if (R < 16 || C < 16) return T5I_NOT_SPECIALISED;
    // (8-byte elements in 16-byte pairs measured slower than the 32 x 32 tile: 128x128 Double 5.65 -> 5.29 TB/s)
    if (elem_bytes == 4 && R % 2 == 0 && C % 2 == 0 && R >= 48 && C >= 48) return transpose2d_pair_launch<unsigned, uint2>(s, R, C, A, O);
    return elem_bytes == 8 ? transpose2d_launch<unsigned long long, 32, 8>(s, R, C, A, O) : transpose2d_launch<unsigned, 32, 8>(s, R, C, A, O);
}

// ---- SoA forms of the any-shape operations ---------------------------------------------------------
// In the SoA layout (element e of item b at [e * cap + b], cap a multiple of 32 items) every pure index map -
// transpose of any rank, append / split / at / ta / slice / axis permutation - is a copy of whole ROWS: output
// row e is input row table[e], contiguous on both sides, 16 bytes per thread.  And a thread-per-item contraction
// reads and writes coalesced by construction.  These make every product shape and every structural operation
// available on SoA batches (the specialised tile kernels still take the shapes they cover).
__global__ void __launch_bounds__(256)
soa_rows_kernel(const uint4* __restrict__ A, const uint4* __restrict__ B, uint4* __restrict__ out, const uint32_t* __restrict__ table,
                unsigned dA, unsigned d_out, unsigned long long row_chunks) {
    for (unsigned e = blockIdx.y; e < d_out; e += gridDim.y) {
        const unsigned v = __ldg(table + e);
        const uint4* src = (v < dA ? A + (size_t)v * row_chunks : B + (size_t)(v - dA) * row_chunks);
        uint4* dst = out + (size_t)e * row_chunks;
        // a CTA moves contiguous 16-KB spans (4 x 256 chunks, four loads in flight per thread)
        for (unsigned long long c0 = (unsigned long long)blockIdx.x * 1024; c0 < row_chunks; c0 += (unsigned long long)gridDim.x * 1024) {
            uint4 v4[4];
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long c = c0 + u * 256 + threadIdx.x;
                if (c < row_chunks) v4[u] = ld_global_stream16(src + c);
            }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long c = c0 + u * 256 + threadIdx.x;
                if (c < row_chunks) st_global_stream16(dst + c, v4[u]);
            }
        }
    }
}

// one- or two-source row gather on SoA shards of equal capacity (B may be null when every table entry is < dA)
int generic_soa_gather(const Shard& s, int elem_bytes, size_t dA, size_t d_out, const uint32_t* table_dev, const void* A, const void* B, void* O) {
    if (s.layout != T5L_SOA || (elem_bytes != 4 && elem_bytes != 8)) return T5I_UNSUPPORTED;
    if (s.count == 0 || d_out == 0) return T5I_OK;
    if (s.soa_stride % 32 || d_out > (1u << 28)) return T5I_UNSUPPORTED;
    const unsigned long long row_chunks = s.soa_stride * elem_bytes / 16;       // rows are whole 16-byte chunks (cap % 32 == 0)
    const size_t bx = (row_chunks + 1023) / 1024;
    // ~8 resident CTAs per SM in total, spread over rows first
    const unsigned gy = (unsigned)(d_out < 256 ? d_out : 256);
    unsigned gx = (unsigned)((8 * (size_t)sm_count() + gy - 1) / gy);
    if (gx > bx) gx = (unsigned)bx;
    if (gx < 1) gx = 1;
    dim3 grid(gx, gy);
    soa_rows_kernel<<<grid, 256, 0, s.stream>>>((const uint4*)A, (const uint4*)B, (uint4*)O, table_dev, (unsigned)dA, (unsigned)d_out, row_chunks);
    return ok();
}

template <class T>
__global__ void __launch_bounds__(256)
soa_prod_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int P, int K, int Q, int nc, unsigned long long count,
                unsigned long long stride) {
    using R = Arith<T>;
    for (unsigned long long b = (unsigned long long)blockIdx.x * 256 + threadIdx.x; b < count; b += (unsigned long long)gridDim.x * 256) {
        const T* a = A + b;
        const T* bb = B + b;
        T* c = C + b;
        if (nc == 0) {
            for (int p = 0; p < P; ++p) {
                const T av = a[(size_t)p * stride];
                for (int q = 0; q < Q; ++q) c[((size_t)p * Q + q) * stride] = R::mul(av, __ldg(bb + (size_t)q * stride));
            }
        } else {
            for (int i = 0; i < P; ++i)
                for (int j = 0; j < Q; ++j) {
                    T acc = R::zero();
                    for (int k = 0; k < K; ++k)
                        acc = R::add(acc, R::mul(__ldg(a + ((size_t)i * K + k) * stride), __ldg(bb + ((size_t)k * Q + j) * stride)));
                    c[((size_t)i * Q + j) * stride] = acc;
                }
        }
    }
}

// Contractions whose B operand is a few hundred elements (8x8 . 8x8, the `prod 1` of BASELINE configs[3], in SoA):
// the kernel above re-reads every B element P times through L1 and thrashes it (0.2 of the HBM peak for 8x8).
// Here a CTA first copies its items' B rows into shared memory - [element][item], stride-1 across the warp, so both
// the coalesced fill and the per-thread reads are conflict-free - and each thread then streams its A elements from
// global memory once per four results: global traffic is the algorithmic traffic.  Same fold, k ascending.
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
soa_prod_staged_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int P, int K, int Q, unsigned long long count,
                       unsigned long long stride) {
    using R = Arith<T>;
    extern __shared__ __align__(16) unsigned char spb_smem[];
    T* sB = reinterpret_cast<T*>(spb_smem);                  // [K * Q][THREADS]
    const int KQ = K * Q;
    const unsigned long long n_tiles = (count + THREADS - 1) / THREADS;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long b = tile * THREADS + threadIdx.x;
        const bool live = b < count;
        __syncthreads();                                         // the previous tile's B is no longer read
        if (live)
            for (int e = 0; e < KQ; ++e) cp_async_small<(int)sizeof(T)>(sB + e * THREADS + threadIdx.x, B + (size_t)e * stride + b);
        cp_async_commit();
        cp_async_wait<0>();                                      // (each thread reads only what it copied itself)
        if (live) {
            const T* a = A + b;
            T* c = C + b;
            const T* sb = sB + threadIdx.x;
            constexpr int KMAX = 16;
            for (int i = 0; i < P; ++i) {
                // the A row in registers (compile-time unrolled, predicated): K independent loads in flight, each element read once
                T arow[KMAX];
                if (K <= KMAX) {
                    #pragma unroll
                    for (int k = 0; k < KMAX; ++k)
                        if (k < K) arow[k] = __ldg(a + ((size_t)i * K + k) * stride);
                }
                for (int j0 = 0; j0 < Q; j0 += 4) {
                    T acc[4];
                    #pragma unroll
                    for (int t = 0; t < 4; ++t) acc[t] = R::zero();
                    const int jt = Q - j0 < 4 ? Q - j0 : 4;
                    if (K <= KMAX) {
                        #pragma unroll
                        for (int k = 0; k < KMAX; ++k)
                            if (k < K) {
                                #pragma unroll
                                for (int t = 0; t < 4; ++t)
                                    if (t < jt) acc[t] = R::add(acc[t], R::mul(arow[k], sb[(k * Q + j0 + t) * THREADS]));
                            }
                    } else {
                        for (int k = 0; k < K; ++k) {
                            const T av = __ldg(a + ((size_t)i * K + k) * stride);
                            #pragma unroll
                            for (int t = 0; t < 4; ++t)
                                if (t < jt) acc[t] = R::add(acc[t], R::mul(av, sb[(k * Q + j0 + t) * THREADS]));
                        }
                    }
                    #pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (t < jt) c[((size_t)i * Q + j0 + t) * stride] = acc[t];
                }
            }
        }
    }
}

// Outer products in SoA (the 8x8 (x) 8x8 -> 8^4 of BASELINE configs[3]: 4096 result rows per item).  One item per
// thread writes 256 contiguous bytes per row and warp - every store instruction lands in a different DRAM page -
// and re-reads B through a thrashing L1 (0.5 of the HBM peak).  Here a thread owns 16 bytes' worth of consecutive
// items (2 Double / 4 Float: 512-byte row segments per warp), and the CTA's B vectors are staged once in shared
// memory ([q][thread], stride-1).  One multiply per result, as specified.
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
soa_outer_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int P, int Q, unsigned long long count,
                 unsigned long long stride) {
    using R = Arith<T>;
    constexpr int VEC = 16 / (int)sizeof(T);
    extern __shared__ __align__(16) unsigned char so_smem[];
    uint4* sB = reinterpret_cast<uint4*>(so_smem);           // [Q][THREADS] vectors of VEC items
    const unsigned long long n_vec = (count + VEC - 1) / VEC; // (rows are padded to 32 items: whole vectors are addressable)
    const unsigned long long n_tiles = (n_vec + THREADS - 1) / THREADS;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long v = tile * THREADS + threadIdx.x;
        const bool live = v < n_vec;
        const size_t off = (size_t)v * VEC;
        if (live)
            for (int q = 0; q < Q; ++q) cp_async16(sB + q * THREADS + threadIdx.x, B + (size_t)q * stride + off);
        cp_async_commit();
        cp_async_wait<0>();                                      // (each thread reads only what it copied itself)
        if (live) {
            for (int p = 0; p < P; ++p) {
                const uint4 araw = ld_global_stream16(A + (size_t)p * stride + off);
                const T* av = reinterpret_cast<const T*>(&araw);
                T* crow = C + (size_t)p * Q * stride + off;
                #pragma unroll 4
                for (int q = 0; q < Q; ++q) {
                    const uint4 braw = sB[q * THREADS + threadIdx.x];
                    const T* bv = reinterpret_cast<const T*>(&braw);
                    uint4 o;
                    #pragma unroll
                    for (int t = 0; t < VEC; ++t) reinterpret_cast<T*>(&o)[t] = R::mul(av[t], bv[t]);
                    st_global_stream16(crow + (size_t)q * stride, o);
                }
            }
        }
    }
}

template <class T, int THREADS>
static int soa_outer_launch_t(const Shard& s, int P, int Q, const void* A, const void* B, void* C) {
    auto kern = soa_outer_kernel<T, THREADS>;
    const int smem = Q * THREADS * 16;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return T5I_CUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    constexpr int VEC = 16 / (int)sizeof(T);
    const size_t tiles = ((s.count + VEC - 1) / VEC + THREADS - 1) / THREADS, cap = (size_t)occ * sm_count();
    kern<<<(int)(tiles < cap ? tiles : cap), THREADS, smem, s.stream>>>((const T*)A, (const T*)B, (T*)C, P, Q, (unsigned long long)s.count, s.soa_stride);
    return ok();
}
// (CTA width measured on 256Ki 8x8 (x) 8x8, Double / Float: 32 threads = 512-byte row segments 5.31 / 5.35 TB/s, 64 = 1 KB
// 5.34 / 5.32, 128 = 2 KB 5.60 / 5.28 at one CTA per SM: the segment length is not what holds the 4096-row scatter at 0.82.)
template <class T>
static int soa_outer_launch(const Shard& s, int P, int Q, const void* A, const void* B, void* C) {
    return soa_outer_launch_t<T, 64>(s, P, Q, A, B, C);
}

// SoA contraction with B IN REGISTERS (K x Q <= 64 elements, compile-time): thread per item.  All of B is requested up
// front (K * Q independent coalesced loads in flight per thread), the rows of A stream through a two-deep register
// pipeline (row i + 1 is in flight while row i is multiplied), results leave as coalesced streaming stores - no shared
// memory, no barrier.  The staged-B kernel above waited for its cp.async copies, then paid one exposed global round trip
// per row of A and read every B element from shared memory P times: 1Mi 8x8 Double 0.453 ms (0.54 of the HBM peak).
// Same fold as everywhere: acc = 0; k ascending; multiply, round, add, round.
template <class T, int K, int Q>
__global__ void __launch_bounds__(128)
soa_prod_reg_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int P, unsigned long long count, unsigned long long stride) {
    using R = Arith<T>;
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count; b += step) {
        T breg[K * Q];
        #pragma unroll
        for (int e = 0; e < K * Q; ++e) breg[e] = __ldg(B + (size_t)e * stride + b);
        const T* a = A + b;
        T* c = C + b;
        T nxt[K];
        #pragma unroll
        for (int k = 0; k < K; ++k) nxt[k] = __ldg(a + (size_t)k * stride);
        for (int i = 0; i < P; ++i) {
            T arow[K];
            #pragma unroll
            for (int k = 0; k < K; ++k) arow[k] = nxt[k];
            if (i + 1 < P) {
                #pragma unroll
                for (int k = 0; k < K; ++k) nxt[k] = __ldg(a + ((size_t)(i + 1) * K + k) * stride);
            }
            #pragma unroll
            for (int q = 0; q < Q; ++q) {
                T acc = R::zero();
                #pragma unroll
                for (int k = 0; k < K; ++k) acc = R::add(acc, R::mul(arow[k], breg[k * Q + q]));
                __stcs(c + ((size_t)i * Q + q) * stride, acc);
            }
        }
    }
}
// The same contraction with the result COLUMNS split over QS warps of the CTA (warp w: items of group w / QS, columns
// [h * Q/QS, (h + 1) * Q/QS) with h = w % QS).  A thread keeps K * Q/QS elements of B instead of K * Q - 8x8 Double: 192
// registers and 8 warps per SM become ~128 and 16 - and the rows of A run DEPTH ahead; the QS warps of a group read
// the same rows of A, the second time from L1.  Per result element nothing changes (acc = 0; k ascending).
template <class T, int K, int Q, int QS, int DEPTH>
__global__ void __launch_bounds__(128, 4)
soa_prod_split_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int P, unsigned long long count, unsigned long long stride) {
    using R = Arith<T>;
    constexpr int QH = Q / QS, GROUP = 128 / QS;              // items per CTA
    static_assert(Q % QS == 0 && (128 / 32) % QS == 0, "column blocks are whole warps");
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned h = warp % QS;
    const unsigned long long step = (unsigned long long)gridDim.x * GROUP;
    for (unsigned long long b = (unsigned long long)blockIdx.x * GROUP + (warp / QS) * 32 + lane; b < count; b += step) {
        T breg[K * QH];
        #pragma unroll
        for (int k = 0; k < K; ++k)
            #pragma unroll
            for (int q = 0; q < QH; ++q) breg[k * QH + q] = __ldg(B + (size_t)(k * Q + h * QH + q) * stride + b);
        const T* a = A + b;
        T* c = C + (size_t)(h * QH) * stride + b;
        T pre[DEPTH][K];
        #pragma unroll
        for (int u = 0; u < DEPTH; ++u)
            if (u < P) {
                #pragma unroll
                for (int k = 0; k < K; ++k) pre[u][k] = __ldg(a + ((size_t)u * K + k) * stride);
            }
        for (int i0 = 0; i0 < P; i0 += DEPTH) {
            #pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                const int i = i0 + u;
                if (i < P) {
                    T arow[K];
                    #pragma unroll
                    for (int k = 0; k < K; ++k) arow[k] = pre[u][k];
                    if (i + DEPTH < P) {
                        #pragma unroll
                        for (int k = 0; k < K; ++k) pre[u][k] = __ldg(a + ((size_t)(i + DEPTH) * K + k) * stride);
                    }
                    #pragma unroll
                    for (int q = 0; q < QH; ++q) {
                        T acc = R::zero();
                        #pragma unroll
                        for (int k = 0; k < K; ++k) acc = R::add(acc, R::mul(arow[k], breg[k * QH + q]));
                        __stcs(c + ((size_t)i * Q + q) * stride, acc);
                    }
                }
            }
        }
    }
}
template <class T, int K, int Q, int QS, int DEPTH>
static int soa_prod_split_launch(const Shard& s, int P, const void* A, const void* B, void* C) {
    auto kern = soa_prod_split_kernel<T, K, Q, QS, DEPTH>;
    const size_t tiles = (s.count + 128 / QS - 1) / (128 / QS);
    kern<<<resident_grid((const void*)kern, 128, tiles), 128, 0, s.stream>>>((const T*)A, (const T*)B, (T*)C, P, (unsigned long long)s.count, s.soa_stride);
    return ok();
}

// Float: the same split with TWO consecutive items per thread, so that a warp instruction moves 256 contiguous bytes of a
// row instead of 128 (the one-item kernel: 1Mi 8x8 Float 0.157 ms = 0.78 of the HBM peak).  Rows are padded to 32 items: the
// partner of the last item of an odd count is padding and is written as zero.
template <int K, int Q, int QS, int DEPTH>
__global__ void __launch_bounds__(128, 4)
soa_prod_split2_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int P, unsigned long long count,
                       unsigned long long stride) {
    using R = Arith<float>;
    constexpr int QH = Q / QS, GROUP = 128 / QS;              // item PAIRS per CTA
    static_assert(Q % QS == 0 && (128 / 32) % QS == 0, "column blocks are whole warps");
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned h = warp % QS;
    const unsigned long long step = 2ull * gridDim.x * GROUP;
    for (unsigned long long b = 2 * ((unsigned long long)blockIdx.x * GROUP + (warp / QS) * 32 + lane); b < count; b += step) {
        const bool pad = b + 1 >= count;
        float2 breg[K * QH];
        #pragma unroll
        for (int k = 0; k < K; ++k)
            #pragma unroll
            for (int q = 0; q < QH; ++q) breg[k * QH + q] = __ldg(reinterpret_cast<const float2*>(B + (size_t)(k * Q + h * QH + q) * stride + b));
        const float* a = A + b;
        float* c = C + (size_t)(h * QH) * stride + b;
        float2 pre[DEPTH][K];
        #pragma unroll
        for (int u = 0; u < DEPTH; ++u)
            if (u < P) {
                #pragma unroll
                for (int k = 0; k < K; ++k) pre[u][k] = __ldg(reinterpret_cast<const float2*>(a + ((size_t)u * K + k) * stride));
            }
        for (int i0 = 0; i0 < P; i0 += DEPTH) {
            #pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                const int i = i0 + u;
                if (i < P) {
                    float2 arow[K];
                    #pragma unroll
                    for (int k = 0; k < K; ++k) arow[k] = pre[u][k];
                    if (i + DEPTH < P) {
                        #pragma unroll
                        for (int k = 0; k < K; ++k) pre[u][k] = __ldg(reinterpret_cast<const float2*>(a + ((size_t)(i + DEPTH) * K + k) * stride));
                    }
                    #pragma unroll
                    for (int q = 0; q < QH; ++q) {
                        float2 acc = make_float2(R::zero(), R::zero());
                        #pragma unroll
                        for (int k = 0; k < K; ++k) {
                            acc.x = R::add(acc.x, R::mul(arow[k].x, breg[k * QH + q].x));
                            acc.y = R::add(acc.y, R::mul(arow[k].y, breg[k * QH + q].y));
                        }
                        if (pad) acc.y = 0.0f;
                        __stcs(reinterpret_cast<float2*>(c + ((size_t)i * Q + q) * stride), acc);
                    }
                }
            }
        }
    }
}
template <int K, int Q, int QS, int DEPTH>
static int soa_prod_split2_launch(const Shard& s, int P, const void* A, const void* B, void* C) {
    auto kern = soa_prod_split2_kernel<K, Q, QS, DEPTH>;
    const size_t per_cta = 2 * (128 / QS), tiles = (s.count + per_cta - 1) / per_cta;
    kern<<<resident_grid((const void*)kern, 128, tiles), 128, 0, s.stream>>>((const float*)A, (const float*)B, (float*)C, P, (unsigned long long)s.count, s.soa_stride);
    return ok();
}

template <class T, int K, int Q>
static int soa_prod_reg_launch(const Shard& s, int P, const void* A, const void* B, void* C) {
    auto kern = soa_prod_reg_kernel<T, K, Q>;
    const size_t tiles = (s.count + 127) / 128;
    kern<<<resident_grid((const void*)kern, 128, tiles), 128, 0, s.stream>>>((const T*)A, (const T*)B, (T*)C, P, (unsigned long long)s.count, s.soa_stride);
    return ok();
}
template <class T>
static int soa_prod_reg(const Shard& s, int P, int K, int Q, const void* A, const void* B, void* C) {
    // 8x8 with 8-byte elements: two column blocks, rows of A two ahead (1Mi items, Double: 0.319 ms all of B per thread; 0.308
    // two blocks, one row ahead; 0.290 two ahead; 0.307 three ahead (spills); 0.347 four blocks.  Int64: 0.357 -> 0.301)
    if constexpr (sizeof(T) == 8) if (K == 8 && Q == 8) return soa_prod_split_launch<T, 8, 8, 2, 2>(s, P, A, B, C);
    if constexpr (std::is_same<T, float>::value) {
        if (K == 8 && Q == 8 && s.soa_stride % 2 == 0 && (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 7) == 0) {
            return soa_prod_split2_launch<8, 8, 2, 2>(s, P, A, B, C);   // 0.156 -> 0.150 ms; four column blocks 0.182, one row ahead 0.154
        }
    }
    if (K == 8 && Q == 8) return soa_prod_reg_launch<T, 8, 8>(s, P, A, B, C);
    if (K == 7 && Q == 7) return soa_prod_reg_launch<T, 7, 7>(s, P, A, B, C);
    if (K == 6 && Q == 6) return soa_prod_reg_launch<T, 6, 6>(s, P, A, B, C);
    if (K == 5 && Q == 5) return soa_prod_reg_launch<T, 5, 5>(s, P, A, B, C);
    if (K == 8 && Q == 4) return soa_prod_reg_launch<T, 8, 4>(s, P, A, B, C);
    if (K == 4 && Q == 8) return soa_prod_reg_launch<T, 4, 8>(s, P, A, B, C);
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int soa_prod_staged_launch(const Shard& s, int P, int K, int Q, const void* A, const void* B, void* C) {
    constexpr int THREADS = 128;
    auto kern = soa_prod_staged_kernel<T, THREADS>;
    const int smem = K * Q * THREADS * (int)sizeof(T);
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return T5I_CUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const size_t tiles = (s.count + THREADS - 1) / THREADS, cap = (size_t)occ * sm_count();
    kern<<<(int)(tiles < cap ? tiles : cap), THREADS, smem, s.stream>>>((const T*)A, (const T*)B, (T*)C, P, K, Q, (unsigned long long)s.count, s.soa_stride);
    return ok();
}

int generic_soa_prod(const Shard& s, int dtype, int P, int K, int Q, int nc, const void* A, const void* B, void* C) {
    if (s.layout != T5L_SOA) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    if (nc == 0 && (long long)Q * 64 * 16 <= 200 * 1024 && (long long)P * Q >= 16) {      // outer products with up to 200 B elements
        switch (dtype) {
            case T5D_F64: return soa_outer_launch<double>(s, P, Q, A, B, C);
            case T5D_F32: return soa_outer_launch<float>(s, P, Q, A, B, C);
            case T5D_I64: return soa_outer_launch<long long>(s, P, Q, A, B, C);
        }
    }
    if (nc != 0 && P > 1) {                                  // B of <= 64 elements, common shapes: B in registers
        int r = T5I_NOT_SPECIALISED;
        switch (dtype) {
            case T5D_F64: r = soa_prod_reg<double>(s, P, K, Q, A, B, C); break;
            case T5D_F32: r = soa_prod_reg<float>(s, P, K, Q, A, B, C); break;
            case T5D_I64: r = soa_prod_reg<long long>(s, P, K, Q, A, B, C); break;
        }
        if (r != T5I_NOT_SPECIALISED) return r;
    }
    // B of 16..200 elements and more than one result row: stage B (<= 200 KB of shared memory per 128 items)
    if (nc != 0 && P > 1 && (long long)K * Q >= 16 && (long long)K * Q * 128 * (dtype == T5D_F32 ? 4 : 8) <= 200 * 1024) {
        switch (dtype) {
            case T5D_F64: return soa_prod_staged_launch<double>(s, P, K, Q, A, B, C);
            case T5D_F32: return soa_prod_staged_launch<float>(s, P, K, Q, A, B, C);
            case T5D_I64: return soa_prod_staged_launch<long long>(s, P, K, Q, A, B, C);
        }
    }
    const size_t bx = (s.count + 255) / 256;
    switch (dtype) {
        case T5D_F64: soa_prod_kernel<double><<<resident_grid((const void*)soa_prod_kernel<double>, 256, bx), 256, 0, s.stream>>>((const double*)A, (const double*)B, (double*)C, P, K, Q, nc, s.count, s.soa_stride); break;
        case T5D_F32: soa_prod_kernel<float><<<resident_grid((const void*)soa_prod_kernel<float>, 256, bx), 256, 0, s.stream>>>((const float*)A, (const float*)B, (float*)C, P, K, Q, nc, s.count, s.soa_stride); break;
        case T5D_I64: soa_prod_kernel<long long><<<resident_grid((const void*)soa_prod_kernel<long long>, 256, bx), 256, 0, s.stream>>>((const long long*)A, (const long long*)B, (long long*)C, P, K, Q, nc, s.count, s.soa_stride); break;
        default: return T5I_UNSUPPORTED;
    }
    return ok();
}

int generic_permute(const Shard& s, int elem_bytes, size_t d, const uint32_t* perm_dev, const void* A, void* O) {
    return generic_gather(s, elem_bytes, d, d, perm_dev, A, O);
}

// =====================================================================================
// Gaussian elimination for 5 <= n <= 8 (and any n <= 8 / nrhs <= 8 not specialised in
// t5_small_ops.cu): thread per item, working matrix in local memory.  Same pivot rule
// and operation order as the specialised kernels and the oracle.
// =====================================================================================
constexpr int GEN_MAXN = 8;

template <class T>
__device__ bool gen_eliminate(T* w, int n, int wd, int* swaps) {
    *swaps = 0;
    for (int k = 0; k < n; ++k) {
        int r = k;
        while (r < n && w[r * wd + k] == T(0)) ++r;
        if (r == n) return false;
        if (r != k) {
            for (int j = 0; j < wd; ++j) { T t = w[k * wd + j]; w[k * wd + j] = w[r * wd + j]; w[r * wd + j] = t; }
            ++*swaps;
        }
        for (int i = k + 1; i < n; ++i) {
            const T f = w[i * wd + k] / w[k * wd + k];
            for (int j = k + 1; j < wd; ++j) w[i * wd + j] = w[i * wd + j] - f * w[k * wd + j];
            w[i * wd + k] = T(0);
        }
    }
    return true;
}

template <class T>
__global__ void gen_det_kernel(const T* __restrict__ A, T* __restrict__ out, int n, unsigned long long count) {
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        T w[GEN_MAXN * GEN_MAXN];
        const T* a = A + b * (size_t)(n * n);
        for (int i = 0; i < n * n; ++i) w[i] = a[i];
        int swaps;
        T res = T(0);
        if (gen_eliminate(w, n, n, &swaps)) {
            T acc = T(1);
            for (int k = 0; k < n; ++k) acc = acc * w[k * n + k];
            res = (swaps & 1) ? -acc : acc;
        }
        out[b] = res;
    }
}

template <class T>
__global__ void gen_solve_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ X,
                                 unsigned char* __restrict__ status, int n, int nrhs, bool identity_rhs,
                                 unsigned long long count) {
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        T w[GEN_MAXN * 2 * GEN_MAXN];
        T xs[GEN_MAXN * GEN_MAXN];
        const int wd = n + nrhs;
        const T* a = A + b * (size_t)(n * n);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) w[i * wd + j] = a[i * n + j];
            for (int c = 0; c < nrhs; ++c)
                w[i * wd + n + c] = identity_rhs ? (i == c ? T(1) : T(0)) : B[b * (size_t)(n * nrhs) + i * nrhs + c];
        }
        int swaps;
        T* x = X + b * (size_t)(n * nrhs);
        if (!gen_eliminate(w, n, wd, &swaps)) {
            for (int i = 0; i < n * nrhs; ++i) x[i] = T(0);
            status[b] = 1;
            continue;
        }
        for (int k = n - 1; k >= 0; --k)
            for (int c = 0; c < nrhs; ++c) {
                T s = T(0);
                for (int j = k + 1; j < n; ++j) s = s + w[k * wd + j] * xs[j * nrhs + c];
                xs[k * nrhs + c] = (w[k * wd + n + c] - s) / w[k * wd + k];
            }
        for (int i = 0; i < n * nrhs; ++i) x[i] = xs[i];
        status[b] = 0;
    }
}

// one thread per work unit, grid-stride beyond the resident capacity
template <class... KArgs, class... Args>
static inline void launch_flat(void (*kern)(KArgs...), size_t n, int threads, cudaStream_t st, Args... args) {
    const int grid = resident_grid((const void*)kern, threads, (n + threads - 1) / threads);
    kern<<<grid, threads, 0, st>>>(args...);
}

int generic_det(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (s.layout != T5L_AOS || n < 1 || n > GEN_MAXN) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    if (dtype == T5D_F64) launch_flat(gen_det_kernel<double>, s.count, 128, s.stream, (const double*)A, (double*)O, n, s.count);
    else if (dtype == T5D_F32) launch_flat(gen_det_kernel<float>, s.count, 128, s.stream, (const float*)A, (float*)O, n, s.count);
    else return T5I_UNSUPPORTED;
    return ok();
}

int generic_solve(const Shard& s, int dtype, int n, int nrhs, bool identity_rhs, const void* A, const void* B, void* X, void* status) {
    if (s.layout != T5L_AOS || n < 1 || n > GEN_MAXN || nrhs < 1 || nrhs > GEN_MAXN) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    if (dtype == T5D_F64)
        launch_flat(gen_solve_kernel<double>, s.count, 128, s.stream, (const double*)A, (const double*)B, (double*)X, (unsigned char*)status, n, nrhs, identity_rhs, s.count);
    else if (dtype == T5D_F32)
        launch_flat(gen_solve_kernel<float>, s.count, 128, s.stream, (const float*)A, (const float*)B, (float*)X, (unsigned char*)status, n, nrhs, identity_rhs, s.count);
    else return T5I_UNSUPPORTED;
    return ok();
}

// =====================================================================================
// The rest of the SquareMatrix surface for any n <= 8 (tr, polyEval, charPoly) and
// minPoly (n <= 6): thread per item, working matrices in local memory, the oracle's
// operation order (see oracle/t5_oracle.cpp for the specification).
// =====================================================================================
template <class T>
__device__ __forceinline__ void gen_matmul_nn(const T* a, const T* b, T* c, int n) {
    using R = Arith<T>;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            T acc = R::zero();
            for (int k = 0; k < n; ++k) acc = R::add(acc, R::mul(a[i * n + k], b[k * n + j]));
            c[i * n + j] = acc;
        }
}

template <class T>
__global__ void gen_trace_kernel(const T* __restrict__ A, T* __restrict__ out, int n, unsigned long long count) {
    using R = Arith<T>;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const T* a = A + b * (size_t)(n * n);
        T acc = R::zero();
        for (int i = 0; i < n; ++i) acc = R::add(acc, a[i * n + i]);
        out[b] = acc;
    }
}

template <class T>
__global__ void gen_poly_eval_kernel(const T* __restrict__ A, const T* __restrict__ coef, int ncoef, T* __restrict__ out,
                                     int n, unsigned long long count) {
    using R = Arith<T>;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        T a[GEN_MAXN * GEN_MAXN], r[GEN_MAXN * GEN_MAXN], t[GEN_MAXN * GEN_MAXN];
        const int nn = n * n;
        for (int i = 0; i < nn; ++i) { a[i] = A[b * (size_t)nn + i]; r[i] = R::zero(); }
        if (ncoef > 0)
            for (int i = 0; i < n; ++i) r[i * n + i] = coef[ncoef - 1];
        for (int j = ncoef - 2; j >= 0; --j) {
            gen_matmul_nn(r, a, t, n);
            const T cj = coef[j];
            for (int i = 0; i < n; ++i) t[i * n + i] = R::add(t[i * n + i], cj);
            for (int i = 0; i < nn; ++i) r[i] = t[i];
        }
        for (int i = 0; i < nn; ++i) out[b * (size_t)nn + i] = r[i];
    }
}

template <class T>
__device__ void gen_char_poly(const T* a, T* c, int n) {
    T m[GEN_MAXN * GEN_MAXN], am[GEN_MAXN * GEN_MAXN];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m[i * n + j] = (i == j) ? T(1) : T(0);
    c[n] = T(1);
    for (int k = 1; k <= n; ++k) {
        gen_matmul_nn(a, m, am, n);
        T tr = T(0);
        for (int i = 0; i < n; ++i) tr = tr + am[i * n + i];
        const T ck = -(tr / T(k));
        c[n - k] = ck;
        for (int i = 0; i < n; ++i) am[i * n + i] = am[i * n + i] + ck;
        for (int i = 0; i < n * n; ++i) m[i] = am[i];
    }
}

template <class T>
__global__ void gen_char_poly_kernel(const T* __restrict__ A, T* __restrict__ out, int n, unsigned long long count) {
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        T a[GEN_MAXN * GEN_MAXN], c[GEN_MAXN + 1];
        for (int i = 0; i < n * n; ++i) a[i] = A[b * (size_t)(n * n) + i];
        gen_char_poly(a, c, n);
        for (int j = 0; j <= n; ++j) out[b * (size_t)(n + 1) + j] = c[j];
    }
}

constexpr int MINPOLY_MAXN = 6;

template <class T>
__global__ void gen_min_poly_kernel(const T* __restrict__ A, T* __restrict__ out, unsigned char* __restrict__ degree, int n,
                                    unsigned long long count) {
    constexpr int ML = MINPOLY_MAXN * MINPOLY_MAXN;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        T a[GEN_MAXN * GEN_MAXN], P[ML], Pn[ML], v[ML], g[MINPOLY_MAXN + 1];
        T basis[MINPOLY_MAXN * ML], coef[MINPOLY_MAXN * (MINPOLY_MAXN + 1)];
        int pivot[MINPOLY_MAXN];
        const int L = n * n;
        for (int i = 0; i < L; ++i) a[i] = A[b * (size_t)L + i];
        T* c = out + b * (size_t)(n + 1);
        int nb = 0, deg = -1;
        for (int m = 0; m < n && deg < 0; ++m) {
            if (m == 0) {
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) P[i * n + j] = (i == j) ? T(1) : T(0);
            } else {
                gen_matmul_nn(P, a, Pn, n);
                for (int i = 0; i < L; ++i) P[i] = Pn[i];
            }
            for (int i = 0; i < L; ++i) v[i] = P[i];
            for (int j = 0; j <= n; ++j) g[j] = (j == m) ? T(1) : T(0);
            for (int i = 0; i < nb; ++i) {
                const T* bi = basis + i * L;
                const T* gi = coef + i * (n + 1);
                const T f = v[pivot[i]] / bi[pivot[i]];
                for (int t = 0; t < L; ++t) v[t] = v[t] - f * bi[t];
                v[pivot[i]] = T(0);
                for (int j = 0; j <= n; ++j) g[j] = g[j] - f * gi[j];
            }
            int p = 0;
            while (p < L && v[p] == T(0)) ++p;
            if (p == L) {
                for (int j = 0; j <= n; ++j) c[j] = (j <= m) ? g[j] : T(0);
                deg = m;
            } else {
                for (int t = 0; t < L; ++t) basis[nb * L + t] = v[t];
                for (int j = 0; j <= n; ++j) coef[nb * (n + 1) + j] = g[j];
                pivot[nb++] = p;
            }
        }
        if (deg < 0) {
            T cc[GEN_MAXN + 1];
            gen_char_poly(a, cc, n);
            for (int j = 0; j <= n; ++j) c[j] = cc[j];
            deg = n;
        }
        degree[b] = (unsigned char)deg;
    }
}

int generic_trace(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (s.layout != T5L_AOS || n < 1 || n > (1 << 14)) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    switch (dtype) {
        case T5D_F64: launch_flat(gen_trace_kernel<double>, s.count, 256, s.stream, (const double*)A, (double*)O, n, s.count); break;
        case T5D_F32: launch_flat(gen_trace_kernel<float>, s.count, 256, s.stream, (const float*)A, (float*)O, n, s.count); break;
        case T5D_I64: launch_flat(gen_trace_kernel<long long>, s.count, 256, s.stream, (const long long*)A, (long long*)O, n, s.count); break;
        default: return T5I_UNSUPPORTED;
    }
    return ok();
}

int generic_poly_eval(const Shard& s, int dtype, int n, const void* coef, int ncoef, const void* A, void* O) {
    if (s.layout != T5L_AOS || n < 1 || n > GEN_MAXN || ncoef < 0) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    switch (dtype) {
        case T5D_F64: launch_flat(gen_poly_eval_kernel<double>, s.count, 128, s.stream, (const double*)A, (const double*)coef, ncoef, (double*)O, n, s.count); break;
        case T5D_F32: launch_flat(gen_poly_eval_kernel<float>, s.count, 128, s.stream, (const float*)A, (const float*)coef, ncoef, (float*)O, n, s.count); break;
        case T5D_I64: launch_flat(gen_poly_eval_kernel<long long>, s.count, 128, s.stream, (const long long*)A, (const long long*)coef, ncoef, (long long*)O, n, s.count); break;
        default: return T5I_UNSUPPORTED;
    }
    return ok();
}

int generic_char_poly(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (s.layout != T5L_AOS || n < 1 || n > GEN_MAXN) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    if (dtype == T5D_F64) launch_flat(gen_char_poly_kernel<double>, s.count, 128, s.stream, (const double*)A, (double*)O, n, s.count);
    else if (dtype == T5D_F32) launch_flat(gen_char_poly_kernel<float>, s.count, 128, s.stream, (const float*)A, (float*)O, n, s.count);
    else return T5I_UNSUPPORTED;
    return ok();
}

int generic_min_poly(const Shard& s, int dtype, int n, const void* A, void* O, void* degree) {
    if (s.layout != T5L_AOS || n < 1 || n > MINPOLY_MAXN) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    if (dtype == T5D_F64) launch_flat(gen_min_poly_kernel<double>, s.count, 64, s.stream, (const double*)A, (double*)O, (unsigned char*)degree, n, s.count);
    else if (dtype == T5D_F32) launch_flat(gen_min_poly_kernel<float>, s.count, 64, s.stream, (const float*)A, (float*)O, (unsigned char*)degree, n, s.count);
    else return T5I_UNSUPPORTED;
    return ok();
}

// =====================================================================================
// pure / replicate (Applicative `pure`, src/Data/Tensor/Vector.hs:185-188; also `zero`, `unit`,
// basis vectors): every item of the batch becomes a copy of one d-element record.  Write-only
// stream; the record is read through the read-only cache.  AoS: flat index -> element by constant
// steps (no division in the loop); SoA: row e is the constant rec[e].
// =====================================================================================
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
replicate_kernel(const T* __restrict__ rec, T* __restrict__ out, unsigned d, unsigned long long total) {
    const unsigned long long step = (unsigned long long)gridDim.x * THREADS;
    unsigned long long i = (unsigned long long)blockIdx.x * THREADS + threadIdx.x;
    unsigned e = (unsigned)(i % d);
    const unsigned se = (unsigned)(step % d);
    #pragma unroll 4
    for (; i < total; i += step) {
        out[i] = __ldg(rec + e);
        e += se;
        if (e >= d) e -= d;
    }
}
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
replicate_soa_kernel(const T* __restrict__ rec, T* __restrict__ out, unsigned d, unsigned long long count, unsigned long long stride) {
    for (unsigned e = blockIdx.y; e < d; e += gridDim.y) {
        const T v = __ldg(rec + e);
        T* row = out + (size_t)e * stride;
        for (unsigned long long b = (unsigned long long)blockIdx.x * THREADS + threadIdx.x; b < count; b += (unsigned long long)gridDim.x * THREADS)
            row[b] = v;
    }
}

int generic_replicate(const Shard& s, int elem_bytes, size_t d, const void* rec_dev, void* O) {
    constexpr int THREADS = 256;
    if (s.count == 0 || d == 0) return T5I_OK;
    if (d > (1u << 28)) return T5I_UNSUPPORTED;
    if (elem_bytes != 4 && elem_bytes != 8) return T5I_UNSUPPORTED;
    if (s.layout == T5L_SOA) {
        const size_t bx = (s.count + THREADS - 1) / THREADS;
        dim3 grid((unsigned)(bx < 4096 ? bx : 4096), (unsigned)(d < 64 ? d : 64));
        if (elem_bytes == 8) replicate_soa_kernel<unsigned long long, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned long long*)rec_dev, (unsigned long long*)O, (unsigned)d, s.count, s.soa_stride);
        else replicate_soa_kernel<unsigned, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned*)rec_dev, (unsigned*)O, (unsigned)d, s.count, s.soa_stride);
        return ok();
    }
    const unsigned long long total = (unsigned long long)s.count * d;
    if (elem_bytes == 8) {
        const int grid = resident_grid((const void*)replicate_kernel<unsigned long long, THREADS>, THREADS, (total + THREADS - 1) / THREADS);
        replicate_kernel<unsigned long long, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned long long*)rec_dev, (unsigned long long*)O, (unsigned)d, total);
    } else {
        const int grid = resident_grid((const void*)replicate_kernel<unsigned, THREADS>, THREADS, (total + THREADS - 1) / THREADS);
        replicate_kernel<unsigned, THREADS><<<grid, THREADS, 0, s.stream>>>((const unsigned*)rec_dev, (unsigned*)O, (unsigned)d, total);
    }
    return ok();
}

// =====================================================================================
// Element-wise vector-space operations (Functor/Applicative, Vector.hs:180-190,
// :250-251).  Layout-agnostic: they act on the flat element array of a shard.
// =====================================================================================
// 16-byte streaming accesses, two vectors per operand in flight per thread; `n` elements, buffers
// 256-byte aligned (shards), scalar tail for n % (16 / sizeof(T)).
template <class T, int OP>
__device__ __forceinline__ T zip_op(T x, T y) {
    using R = Arith<T>;
    return OP == 0 ? R::add(x, y) : (OP == 1 ? R::sub(x, y) : R::mul(x, y));
}
template <class T, int OP>
__global__ void __launch_bounds__(256) zip_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ c, unsigned long long n) {
    constexpr int V = 16 / (int)sizeof(T);
    const unsigned long long nv = n / V, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint4* av = reinterpret_cast<const uint4*>(a);
    const uint4* bv = reinterpret_cast<const uint4*>(b);
    uint4* cv = reinterpret_cast<uint4*>(c);
    auto apply = [](const uint4& x, const uint4& y) {
        uint4 r;
        #pragma unroll
        for (int j = 0; j < V; ++j)
            reinterpret_cast<T*>(&r)[j] = zip_op<T, OP>(reinterpret_cast<const T*>(&x)[j], reinterpret_cast<const T*>(&y)[j]);
        return r;
    };
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < nv; i += 2 * stride) {
        const uint4 x0 = ld_global_stream16(av + i), y0 = ld_global_stream16(bv + i);
        const uint4 x1 = ld_global_stream16(av + i + stride), y1 = ld_global_stream16(bv + i + stride);
        st_global_stream16(cv + i, apply(x0, y0));
        st_global_stream16(cv + i + stride, apply(x1, y1));
    }
    for (; i < nv; i += stride) st_global_stream16(cv + i, apply(ld_global_stream16(av + i), ld_global_stream16(bv + i)));
    for (unsigned long long t = nv * V + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
        c[t] = zip_op<T, OP>(a[t], b[t]);
}
template <class T>
__global__ void __launch_bounds__(256) scale_kernel(const T* __restrict__ a, T* __restrict__ c, T k, unsigned long long n) {
    using R = Arith<T>;
    constexpr int V = 16 / (int)sizeof(T);
    const unsigned long long nv = n / V, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint4* av = reinterpret_cast<const uint4*>(a);
    uint4* cv = reinterpret_cast<uint4*>(c);
    auto apply = [k](const uint4& x) {
        uint4 r;
        #pragma unroll
        for (int j = 0; j < V; ++j) reinterpret_cast<T*>(&r)[j] = R::mul(k, reinterpret_cast<const T*>(&x)[j]);
        return r;
    };
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < nv; i += 2 * stride) {
        const uint4 x0 = ld_global_stream16(av + i), x1 = ld_global_stream16(av + i + stride);
        st_global_stream16(cv + i, apply(x0));
        st_global_stream16(cv + i + stride, apply(x1));
    }
    for (; i < nv; i += stride) st_global_stream16(cv + i, apply(ld_global_stream16(av + i)));
    for (unsigned long long t = nv * V + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
        c[t] = R::mul(k, a[t]);
}

template <class T>
static int zip_t(const Shard& s, int op, size_t n, const void* A, const void* B, void* C) {
    const size_t units = n / (16 / sizeof(T)) / 2 + 1;     // one thread per two 16-byte vectors
    if (op == 0) launch_flat(zip_kernel<T, 0>, units, 256, s.stream, (const T*)A, (const T*)B, (T*)C, n);
    else if (op == 1) launch_flat(zip_kernel<T, 1>, units, 256, s.stream, (const T*)A, (const T*)B, (T*)C, n);
    else if (op == 2) launch_flat(zip_kernel<T, 2>, units, 256, s.stream, (const T*)A, (const T*)B, (T*)C, n);
    else return T5I_UNSUPPORTED;
    return ok();
}

// `elems` = number of stored elements of the shard (count*d for AoS, stride*d for SoA)
int generic_zip(const Shard& s, int dtype, int op, size_t elems, const void* A, const void* B, void* C) {
    if (elems == 0) return T5I_OK;
    switch (dtype) {
        case T5D_F64: return zip_t<double>(s, op, elems, A, B, C);
        case T5D_F32: return zip_t<float>(s, op, elems, A, B, C);
        case T5D_I64: return zip_t<long long>(s, op, elems, A, B, C);
    }
    return T5I_UNSUPPORTED;
}

int generic_scale(const Shard& s, int dtype, const void* k, size_t elems, const void* A, void* C) {
    if (elems == 0) return T5I_OK;
    switch (dtype) {
        case T5D_F64: launch_flat(scale_kernel<double>, elems / 4 + 1, 256, s.stream, (const double*)A, (double*)C, *(const double*)k, elems); break;
        case T5D_F32: launch_flat(scale_kernel<float>, elems / 8 + 1, 256, s.stream, (const float*)A, (float*)C, *(const float*)k, elems); break;
        case T5D_I64: launch_flat(scale_kernel<long long>, elems / 4 + 1, 256, s.stream, (const long long*)A, (long long*)C, *(const long long*)k, elems); break;
        default: return T5I_UNSUPPORTED;
    }
    return ok();
}

// =====================================================================================
// sum_b dot(x_b, y_b): a flat reduction of x[i]*y[i] (the product is rounded in the
// element type, the sum is carried in double / wrapping int64).  Deterministic:
// fixed grid, per-block partials reduced by one block in index order.
// =====================================================================================
template <class Acc> __device__ __forceinline__ Acc warp_sum(Acc v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int DS_THREADS = 256;
constexpr int DS_MAX_BLOCKS = 148 * 8;

template <class T, class Acc>
__device__ __forceinline__ void dot_accumulate(Acc& acc, const uint4& a, const uint4& b) {
    using R = Arith<T>;
    constexpr int V = 16 / (int)sizeof(T);
    #pragma unroll
    for (int j = 0; j < V; ++j) acc += (Acc)R::mul(reinterpret_cast<const T*>(&a)[j], reinterpret_cast<const T*>(&b)[j]);
}

// 16-byte streaming loads, two vectors of each operand in flight per thread (the shard buffers
// are 256-B aligned); the assignment of elements to threads is fixed, so the result is reproducible.
template <class T, class Acc>
__global__ void __launch_bounds__(DS_THREADS)
dot_sum_stage1(const T* __restrict__ x, const T* __restrict__ y, Acc* __restrict__ partials, unsigned long long n) {
    using R = Arith<T>;
    constexpr int V = 16 / (int)sizeof(T);
    const unsigned long long nv = n / V;
    const unsigned long long stride = (unsigned long long)gridDim.x * DS_THREADS;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* yv = reinterpret_cast<const uint4*>(y);
    Acc acc = 0;
    unsigned long long i = (unsigned long long)blockIdx.x * DS_THREADS + threadIdx.x;
    for (; i + stride < nv; i += 2 * stride) {
        const uint4 a0 = ld_global_stream16(xv + i), b0 = ld_global_stream16(yv + i);
        const uint4 a1 = ld_global_stream16(xv + i + stride), b1 = ld_global_stream16(yv + i + stride);
        dot_accumulate<T, Acc>(acc, a0, b0);
        dot_accumulate<T, Acc>(acc, a1, b1);
    }
    for (; i < nv; i += stride) dot_accumulate<T, Acc>(acc, ld_global_stream16(xv + i), ld_global_stream16(yv + i));
    for (unsigned long long j = nv * V + (unsigned long long)blockIdx.x * DS_THREADS + threadIdx.x; j < n; j += stride)
        acc += (Acc)R::mul(x[j], y[j]);
    __shared__ Acc sm[DS_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        Acc v = threadIdx.x < DS_THREADS / 32 ? sm[threadIdx.x] : Acc(0);
        v = warp_sum(v);
        if (threadIdx.x == 0) partials[blockIdx.x] = v;
    }
}
template <class Acc>
__global__ void __launch_bounds__(DS_THREADS) dot_sum_stage2(const Acc* __restrict__ partials, int nblocks, Acc* __restrict__ out) {
    Acc acc = 0;
    for (int i = threadIdx.x; i < nblocks; i += DS_THREADS) acc += partials[i];
    __shared__ Acc sm[DS_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        Acc v = threadIdx.x < DS_THREADS / 32 ? sm[threadIdx.x] : Acc(0);
        v = warp_sum(v);
        if (threadIdx.x == 0) *out = v;
    }
}

// SoA shards whose rows carry padding (row_stride > row_valid): only items [0, row_valid) of every row are summed, so
// the result never depends on what an element-wise op left in the pad (0 * inf = NaN there would poison the sum).
// Rows are whole 16-byte vectors (row_stride is a multiple of 32 items); a thread walks (row, vector-in-row) by
// adding the grid stride as (rows, vectors) - one division at entry, none in the loop.
template <class T, class Acc>
__global__ void __launch_bounds__(DS_THREADS)
dot_sum_stage1_rows(const T* __restrict__ x, const T* __restrict__ y, Acc* __restrict__ partials, unsigned long long rows,
                    unsigned long long row_vecs, unsigned long long row_valid) {
    using R = Arith<T>;
    constexpr int V = 16 / (int)sizeof(T);
    const unsigned long long stride = (unsigned long long)gridDim.x * DS_THREADS;
    const unsigned long long srow = stride / row_vecs, svec = stride % row_vecs;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * DS_THREADS + threadIdx.x;
    unsigned long long row = i0 / row_vecs, vec = i0 % row_vecs;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* yv = reinterpret_cast<const uint4*>(y);
    Acc acc = 0;
    while (row < rows) {
        const unsigned long long e0 = vec * V;
        if (e0 < row_valid) {
            const unsigned long long i = row * row_vecs + vec;
            const uint4 a = ld_global_stream16(xv + i), b = ld_global_stream16(yv + i);
            const unsigned long long left = row_valid - e0;
            #pragma unroll
            for (int j = 0; j < V; ++j)
                if ((unsigned long long)j < left) acc += (Acc)R::mul(reinterpret_cast<const T*>(&a)[j], reinterpret_cast<const T*>(&b)[j]);
        }
        row += srow; vec += svec;
        if (vec >= row_vecs) { vec -= row_vecs; ++row; }
    }
    __shared__ Acc sm[DS_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        Acc v = threadIdx.x < DS_THREADS / 32 ? sm[threadIdx.x] : Acc(0);
        v = warp_sum(v);
        if (threadIdx.x == 0) partials[blockIdx.x] = v;
    }
}

template <class T, class Acc>
static void dot_sum_launch(const Shard& s, size_t elems, size_t row_stride, size_t row_valid, const void* X, const void* Y, void* scratch, void* partial) {
    constexpr size_t V = 16 / sizeof(T);
    const size_t blocks = (elems / V + DS_THREADS - 1) / DS_THREADS;   // one 16-B vector per thread and pass
    const int grid = (int)(blocks < (size_t)DS_MAX_BLOCKS ? (blocks ? blocks : 1) : DS_MAX_BLOCKS);
    if (row_stride && row_valid < row_stride)
        dot_sum_stage1_rows<T, Acc><<<grid, DS_THREADS, 0, s.stream>>>((const T*)X, (const T*)Y, (Acc*)scratch, elems / row_stride, row_stride / V, row_valid);
    else
        dot_sum_stage1<T, Acc><<<grid, DS_THREADS, 0, s.stream>>>((const T*)X, (const T*)Y, (Acc*)scratch, elems);
    dot_sum_stage2<Acc><<<1, DS_THREADS, 0, s.stream>>>((const Acc*)scratch, grid, (Acc*)partial);
}

// (A single-kernel version - the last block to finish folds the partials - measured SLOWER: 16Mi 4-vectors Float
// 99.8 -> 104.6 us per call; the extra code costs the streaming kernel more than the second launch does.)
// scratch_dev: >= DS_MAX_BLOCKS * 8 bytes.  AoS: elems = count*d, row_stride = 0.  SoA: elems = stride*d, rows of
// row_stride items of which the first row_valid are data (the rest is padding and is skipped).
int dot_sum_partial(const Shard& s, int dtype, size_t elems, size_t row_stride, size_t row_valid, const void* X, const void* Y, void* scratch, void* partial) {
    switch (dtype) {
        case T5D_F64: dot_sum_launch<double, double>(s, elems, row_stride, row_valid, X, Y, scratch, partial); break;
        case T5D_F32: dot_sum_launch<float, double>(s, elems, row_stride, row_valid, X, Y, scratch, partial); break;
        case T5D_I64: dot_sum_launch<long long, unsigned long long>(s, elems, row_stride, row_valid, X, Y, scratch, partial); break;
        default: return T5I_UNSUPPORTED;
    }
    return ok();
}

// =====================================================================================
// Synthetic inputs, generated where they are consumed.
// =====================================================================================
template <class T> __device__ __forceinline__ T synth(uint64_t x);
template <> __device__ __forceinline__ double synth<double>(uint64_t x) { return ((double)(x >> 11) * 0x1.0p-53) * 2.0 - 1.0; }
template <> __device__ __forceinline__ float synth<float>(uint64_t x) { return ((float)(x >> 40) * 0x1.0p-24f) * 2.0f - 1.0f; }
template <> __device__ __forceinline__ long long synth<long long>(uint64_t x) { return (long long)x; }

template <class T>
__global__ void fill_kernel(T* dst, uint64_t base, unsigned long long first_item, unsigned long long count,
                            unsigned long long d, int soa, unsigned long long stride) {
    const unsigned long long n = count * d;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long gi, loc;
        if (!soa) { gi = first_item * d + i; loc = i; }
        else { unsigned long long e = i / count, b = i - e * count; gi = (first_item + b) * d + e; loc = e * stride + b; }
        dst[loc] = synth<T>(splitmix64(base + gi));
    }
}

int fill_synthetic(const Shard& s, int dtype, size_t d, uint64_t seed, uint64_t op_id, void* dst) {
    if (s.count == 0) return T5I_OK;
    const uint64_t base = seed ^ (op_id << 56);
    const int soa = s.layout == T5L_SOA;
    switch (dtype) {
        case T5D_F64: launch_flat(fill_kernel<double>, s.count * d, 256, s.stream, (double*)dst, base, s.first, s.count, d, soa, s.soa_stride); break;
        case T5D_F32: launch_flat(fill_kernel<float>, s.count * d, 256, s.stream, (float*)dst, base, s.first, s.count, d, soa, s.soa_stride); break;
        case T5D_I64: launch_flat(fill_kernel<long long>, s.count * d, 256, s.stream, (long long*)dst, base, s.first, s.count, d, soa, s.soa_stride); break;
        default: return T5I_UNSUPPORTED;
    }
    return ok();
}

template <class T>
__global__ void plant_kernel(T* A, int n, unsigned long long first_item, unsigned long long count,
                             unsigned long long swap_every, unsigned long long sing_every, int soa, unsigned long long stride) {
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < count;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long g = first_item + b;
        const unsigned long long es = soa ? stride : 1;          // element stride
        T* m = soa ? A + b : A + b * (unsigned long long)(n * n);
        if (swap_every && g % swap_every == 0) m[0] = T(0);
        if (sing_every && g % sing_every == 0 && n >= 2)
            for (int j = 0; j < n; ++j) m[(unsigned long long)(n + j) * es] = m[(unsigned long long)j * es];
    }
}

int plant_specials(const Shard& s, int dtype, int n, uint64_t swap_every, uint64_t sing_every, void* A) {
    if (s.count == 0) return T5I_OK;
    const int soa = s.layout == T5L_SOA;
    if (dtype == T5D_F64) launch_flat(plant_kernel<double>, s.count, 256, s.stream, (double*)A, n, s.first, s.count, swap_every, sing_every, soa, s.soa_stride);
    else if (dtype == T5D_F32) launch_flat(plant_kernel<float>, s.count, 256, s.stream, (float*)A, n, s.first, s.count, swap_every, sing_every, soa, s.soa_stride);
    else return T5I_UNSUPPORTED;
    return ok();
}

// =====================================================================================
// AoS [count][d]  <->  SoA [d][stride]: a 32x32 shared-memory tiled transpose.
// =====================================================================================
template <class T>
__global__ void __launch_bounds__(256)
relayout_kernel(const T* __restrict__ src, T* __restrict__ dst, unsigned long long count, unsigned d,
                unsigned long long stride, int to_soa) {
    __shared__ T tile[32][33];
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const unsigned long long n_bt = (count + 31) / 32;
    const unsigned n_dt = (d + 31) / 32;
    for (unsigned long long t = blockIdx.x; t < n_bt * n_dt; t += gridDim.x) {
        const unsigned long long b0 = (t / n_dt) * 32;
        const unsigned e0 = (unsigned)(t % n_dt) * 32;
        __syncthreads();
        if (to_soa) {
            // read AoS rows: item b0+r, elements e0+tx (contiguous in e)
            for (unsigned r = ty; r < 32; r += 8)
                if (b0 + r < count && e0 + tx < d) tile[r][tx] = src[(b0 + r) * d + e0 + tx];
            __syncthreads();
            for (unsigned r = ty; r < 32; r += 8)
                if (e0 + r < d && b0 + tx < count) dst[(unsigned long long)(e0 + r) * stride + b0 + tx] = tile[tx][r];
        } else {
            for (unsigned r = ty; r < 32; r += 8)
                if (e0 + r < d && b0 + tx < count) tile[r][tx] = src[(unsigned long long)(e0 + r) * stride + b0 + tx];
            __syncthreads();
            for (unsigned r = ty; r < 32; r += 8)
                if (b0 + r < count && e0 + tx < d) dst[(b0 + r) * d + e0 + tx] = tile[tx][r];
        }
    }
}

// Small records (d <= 64 elements: every tiny matrix / vector of the hot path).  The AoS image of
// TI consecutive items is ONE contiguous segment, so it moves with perfectly coalesced linear
// accesses; the SoA side moves as d row segments of TI elements.  In between the tile sits in
// shared memory as [item][element] with an odd pitch, which makes the column reads / writes of the
// SoA phase conflict-free.  No index division in either loop.
constexpr int RL_THREADS = 256;

template <class T, int TI>
__global__ void __launch_bounds__(RL_THREADS)
relayout_small_kernel(const T* __restrict__ src, T* __restrict__ dst, unsigned long long count, unsigned d,
                      unsigned long long stride, int to_soa) {
    extern __shared__ __align__(16) unsigned char rl_smem[];
    T* tile = reinterpret_cast<T*>(rl_smem);
    const unsigned pitch = d | 1u;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned step_q = RL_THREADS / d, step_r = RL_THREADS - step_q * d;
    const unsigned long long n_tiles = (count + TI - 1) / TI;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned long long b0 = t * TI;
        const unsigned nv = count - b0 < (unsigned long long)TI ? (unsigned)(count - b0) : (unsigned)TI;
        const unsigned total = nv * d;
        __syncthreads();                                    // the previous tile has left shared memory
        if (to_soa) {
            const T* a = src + b0 * d;
            unsigned i = tid / d, e = tid - i * d;
            for (unsigned l = tid; l < total; l += RL_THREADS) {
                tile[i * pitch + e] = a[l];
                i += step_q; e += step_r;
                if (e >= d) { e -= d; ++i; }
            }
            __syncthreads();
            for (unsigned r = warp; r < d; r += RL_THREADS / 32) {
                T* row = dst + (unsigned long long)r * stride + b0;
                #pragma unroll
                for (unsigned j = 0; j < TI / 32; ++j) {
                    const unsigned it = lane + 32 * j;
                    if (it < nv) row[it] = tile[it * pitch + r];
                }
            }
        } else {
            for (unsigned r = warp; r < d; r += RL_THREADS / 32) {
                const T* row = src + (unsigned long long)r * stride + b0;
                #pragma unroll
                for (unsigned j = 0; j < TI / 32; ++j) {
                    const unsigned it = lane + 32 * j;
                    if (it < nv) tile[it * pitch + r] = row[it];
                }
            }
            __syncthreads();
            T* a = dst + b0 * d;
            unsigned i = tid / d, e = tid - i * d;
            for (unsigned l = tid; l < total; l += RL_THREADS) {
                a[l] = tile[i * pitch + e];
                i += step_q; e += step_r;
                if (e >= d) { e -= d; ++i; }
            }
        }
    }
}

// 4-byte elements: the same tile, but every global access is an 8-byte PAIR, so that a warp instruction moves 256
// contiguous bytes as the 8-byte kernel's does (1-element accesses left the 4x4 Float relayout at 0.82 of the HBM peak
// against 0.96 for 3x3 Double).  SoA side: two consecutive items of one element row (rows are padded to 32 items: the
// partner of the last item of an odd count is padding, written as zero).  AoS side: two consecutive elements of the
// tile's contiguous AoS image - for even d a pair lies inside one item (pitch d + 1), for odd d the pitch IS d, the tile is
// the linear image and a pair that straddles two items still lands on consecutive words.
template <int TI>
__global__ void __launch_bounds__(RL_THREADS)
relayout_small_pair_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, unsigned long long count, unsigned d,
                           unsigned long long stride, int to_soa) {
    extern __shared__ __align__(16) unsigned char rl_smem[];
    unsigned* tile = reinterpret_cast<unsigned*>(rl_smem);
    const unsigned pitch = d | 1u;
    const bool linear = d & 1u;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned dp = d / 2;                              // pairs per item (even d)
    const unsigned step_q = dp ? RL_THREADS / dp : 0, step_r = dp ? RL_THREADS - step_q * dp : 0;
    const unsigned long long n_tiles = (count + TI - 1) / TI;
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const unsigned long long b0 = t * TI;
        const unsigned nv = count - b0 < (unsigned long long)TI ? (unsigned)(count - b0) : (unsigned)TI;
        const unsigned total = nv * d, totalp = total / 2;
        __syncthreads();                                    // the previous tile has left shared memory
        if (to_soa) {
            const uint2* a = reinterpret_cast<const uint2*>(src + b0 * d);
            if (linear) {
                for (unsigned l = tid; l < totalp; l += RL_THREADS) {
                    const uint2 v = __ldg(a + l);
                    tile[2 * l] = v.x;
                    tile[2 * l + 1] = v.y;
                }
                if ((total & 1u) && tid == 0) tile[total - 1] = __ldg(src + b0 * d + total - 1);
            } else {
                unsigned i = tid / dp, e = tid - i * dp;
                for (unsigned l = tid; l < totalp; l += RL_THREADS) {
                    const uint2 v = __ldg(a + l);
                    tile[i * pitch + 2 * e] = v.x;
                    tile[i * pitch + 2 * e + 1] = v.y;
                    i += step_q; e += step_r;
                    if (e >= dp) { e -= dp; ++i; }
                }
            }
            if (nv < (unsigned)TI && (nv & 1u))             // the partner of the last item: padding, must read as zero
                for (unsigned e2 = tid; e2 < d; e2 += RL_THREADS) tile[nv * pitch + e2] = 0u;
            __syncthreads();
            for (unsigned r = warp; r < d; r += RL_THREADS / 32) {
                uint2* row = reinterpret_cast<uint2*>(dst + (unsigned long long)r * stride + b0);
                #pragma unroll
                for (unsigned j = 0; j < TI / 64; ++j) {
                    const unsigned it = 2 * (lane + 32 * j);
                    if (it < nv) row[lane + 32 * j] = make_uint2(tile[it * pitch + r], tile[(it + 1) * pitch + r]);
                }
            }
        } else {
            for (unsigned r = warp; r < d; r += RL_THREADS / 32) {
                const uint2* row = reinterpret_cast<const uint2*>(src + (unsigned long long)r * stride + b0);
                #pragma unroll
                for (unsigned j = 0; j < TI / 64; ++j) {
                    const unsigned it = 2 * (lane + 32 * j);
                    if (it < nv) {
                        const uint2 v = __ldg(row + lane + 32 * j);
                        tile[it * pitch + r] = v.x;
                        tile[(it + 1) * pitch + r] = v.y;     // (item nv of an odd tail: padding, staged but never stored)
                    }
                }
            }
            __syncthreads();
            uint2* a = reinterpret_cast<uint2*>(dst + b0 * d);
            if (linear) {
                for (unsigned l = tid; l < totalp; l += RL_THREADS) a[l] = make_uint2(tile[2 * l], tile[2 * l + 1]);
                if ((total & 1u) && tid == 0) dst[b0 * d + total - 1] = tile[total - 1];
            } else {
                unsigned i = tid / dp, e = tid - i * dp;
                for (unsigned l = tid; l < totalp; l += RL_THREADS) {
                    a[l] = make_uint2(tile[i * pitch + 2 * e], tile[i * pitch + 2 * e + 1]);
                    i += step_q; e += step_r;
                    if (e >= dp) { e -= dp; ++i; }
                }
            }
        }
    }
}

// Records beyond the staging limit: 64 x 64 (items x elements) tiles with 8-byte accesses on both sides.  4-byte elements
// (d even) move as pairs like the kernel above - the 32 x 32 tile with 1-element accesses ran the 32x32 Float relayout at
// 0.43 of the HBM peak, this one at 0.97 - 8-byte elements one per lane and two column halves per row (0.73 -> 0.9).
template <class T>
__global__ void __launch_bounds__(256)
relayout_tile64_kernel(const T* __restrict__ src, T* __restrict__ dst, unsigned long long count, unsigned d,
                       unsigned long long stride, int to_soa) {
    __shared__ T tile[64][65];
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 lanes x 8 rows
    const unsigned long long n_bt = (count + 63) / 64;
    const unsigned n_dt = (d + 63) / 64;
    for (unsigned long long t = blockIdx.x; t < n_bt * n_dt; t += gridDim.x) {
        const unsigned long long b0 = (t / n_dt) * 64;
        const unsigned e0 = (unsigned)(t % n_dt) * 64;
        __syncthreads();
        if constexpr (sizeof(T) == 4) {
            if (to_soa) {
                // tile[item][element]: AoS rows, pairs of elements
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8) {
                    uint2 v = make_uint2(0u, 0u);                   // items past count: the SoA padding reads as zero
                    if (b0 + r < count && e0 + 2 * tx < d) v = __ldg(reinterpret_cast<const uint2*>(src + (b0 + r) * d + e0 + 2 * tx));
                    tile[r][2 * tx] = v.x;
                    tile[r][2 * tx + 1] = v.y;
                }
                __syncthreads();
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8)
                    if (e0 + r < d && b0 + 2 * tx < count)
                        *reinterpret_cast<uint2*>(dst + (unsigned long long)(e0 + r) * stride + b0 + 2 * tx) = make_uint2(tile[2 * tx][r], tile[2 * tx + 1][r]);
            } else {
                // tile[element][item]: SoA rows, pairs of items
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8)
                    if (e0 + r < d && b0 + 2 * tx < count) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2*>(src + (unsigned long long)(e0 + r) * stride + b0 + 2 * tx));
                        tile[r][2 * tx] = v.x;
                        tile[r][2 * tx + 1] = v.y;
                    }
                __syncthreads();
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8)
                    if (b0 + r < count && e0 + 2 * tx < d)
                        *reinterpret_cast<uint2*>(dst + (b0 + r) * d + e0 + 2 * tx) = make_uint2(tile[2 * tx][r], tile[2 * tx + 1][r]);
            }
        } else {
            if (to_soa) {
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8)
                    #pragma unroll
                    for (unsigned c = tx; c < 64; c += 32)
                        if (b0 + r < count && e0 + c < d) tile[r][c] = __ldg(src + (b0 + r) * d + e0 + c);
                __syncthreads();
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8)
                    #pragma unroll
                    for (unsigned c = tx; c < 64; c += 32)
                        if (e0 + r < d && b0 + c < count) dst[(unsigned long long)(e0 + r) * stride + b0 + c] = tile[c][r];
            } else {
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8)
                    #pragma unroll
                    for (unsigned c = tx; c < 64; c += 32)
                        if (e0 + r < d && b0 + c < count) tile[r][c] = __ldg(src + (unsigned long long)(e0 + r) * stride + b0 + c);
                __syncthreads();
                #pragma unroll
                for (unsigned r = ty; r < 64; r += 8)
                    #pragma unroll
                    for (unsigned c = tx; c < 64; c += 32)
                        if (b0 + r < count && e0 + c < d) dst[(b0 + r) * d + e0 + c] = tile[c][r];
            }
        }
    }
}

// 8-byte elements towards SoA: the same 64 x 64 tile written with one body for both directions (run-time pitches, staged
// elements always written, zero where the source ends) measured 5.76 TB/s on 32x32 Double against 4.94 for the branch of
// the kernel above - and 5.95 against 6.15 the other way, so each direction keeps its faster form.
template <class T>
__global__ void __launch_bounds__(256)
relayout_tile64_uniform_kernel(const T* __restrict__ src, T* __restrict__ dst, unsigned long long count, unsigned d,
                               unsigned long long stride, int to_soa) {
    __shared__ T tile[64][65];
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const unsigned long long n_bt = (count + 63) / 64;
    const unsigned n_dt = (d + 63) / 64;
    for (unsigned long long t = blockIdx.x; t < n_bt * n_dt; t += gridDim.x) {
        const unsigned long long b0 = (t / n_dt) * 64;
        const unsigned e0 = (unsigned)(t % n_dt) * 64;
        // first index of the tile = the side being read; rd_* describe it, wr_* the other one
        const unsigned long long rd_rows = to_soa ? count - b0 : d - e0, rd_cols = to_soa ? d - e0 : count - b0;
        const T* rd = to_soa ? src + b0 * d + e0 : src + (unsigned long long)e0 * stride + b0;
        const unsigned long long rd_pitch = to_soa ? d : stride;
        T* wr = to_soa ? dst + (unsigned long long)e0 * stride + b0 : dst + b0 * d + e0;
        const unsigned long long wr_pitch = to_soa ? stride : d;
        __syncthreads();
        #pragma unroll
        for (unsigned r = ty; r < 64; r += 8)
            #pragma unroll
            for (unsigned c = tx; c < 64; c += 32) {
                T v = T(0);
                if (r < rd_rows && c < rd_cols) v = rd[r * rd_pitch + c];
                tile[r][c] = v;
            }
        __syncthreads();
        #pragma unroll
        for (unsigned r = ty; r < 64; r += 8)
            #pragma unroll
            for (unsigned c = tx; c < 64; c += 32)
                if (r < rd_cols && c < rd_rows) wr[r * wr_pitch + c] = tile[c][r];
    }
}

template <class P>
static int relayout_small_launch(void (*kern)(const P*, P*, unsigned long long, unsigned, unsigned long long, int), cudaStream_t st, int ti,
                                 size_t pitch, size_t d, size_t count, size_t soa_stride, bool to_soa, const void* src, void* dst) {
    const size_t smem = (size_t)ti * pitch * sizeof(P);
    // (an attribute call per launch costs ~1 us on the host; the limit only has to cover this launch)
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return T5I_CUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RL_THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const size_t tiles = (count + ti - 1) / ti;
    const size_t cap = (size_t)sm_count() * occ;
    const int grid = (int)(tiles < cap ? tiles : cap);
    kern<<<grid, RL_THREADS, smem, st>>>((const P*)src, (P*)dst, count, (unsigned)d, soa_stride, to_soa ? 1 : 0);
    return ok();
}

template <class P>
static int relayout_small(cudaStream_t st, int ti, size_t d, size_t count, size_t soa_stride, bool to_soa, const void* src, void* dst) {
    const size_t pitch = d | 1;
    switch (ti) {
        case 64:  return relayout_small_launch<P>(relayout_small_kernel<P, 64>, st, ti, pitch, d, count, soa_stride, to_soa, src, dst);
        case 256: return relayout_small_launch<P>(relayout_small_kernel<P, 256>, st, ti, pitch, d, count, soa_stride, to_soa, src, dst);
        default:  return relayout_small_launch<P>(relayout_small_kernel<P, 128>, st, 128, pitch, d, count, soa_stride, to_soa, src, dst);
    }
}

static int relayout_small_pair(cudaStream_t st, int ti, size_t d, size_t count, size_t soa_stride, bool to_soa, const void* src, void* dst) {
    const size_t pitch = d | 1;
    switch (ti) {
        case 64:  return relayout_small_launch<unsigned>(relayout_small_pair_kernel<64>, st, ti, pitch, d, count, soa_stride, to_soa, src, dst);
        case 256: return relayout_small_launch<unsigned>(relayout_small_pair_kernel<256>, st, ti, pitch, d, count, soa_stride, to_soa, src, dst);
        case 512: return relayout_small_launch<unsigned>(relayout_small_pair_kernel<512>, st, ti, pitch, d, count, soa_stride, to_soa, src, dst);
        default:  return relayout_small_launch<unsigned>(relayout_small_pair_kernel<128>, st, 128, pitch, d, count, soa_stride, to_soa, src, dst);
    }
}

int relayout(cudaStream_t st, int elem_bytes, size_t d, size_t count, size_t soa_stride, bool to_soa, const void* src, void* dst) {
    if (count == 0 || d == 0) return T5I_OK;
    // 8-byte accesses need rows padded to an even number of items and 8-byte aligned bases (cudaMalloc gives 256)
    const bool wide = soa_stride % 2 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 7) == 0;
    // Items per staged tile: row segments of 1 KB (256 four-byte / 128 eight-byte items), twice that while the tile stays
    // under 8 KB (records of a few elements), halved while it exceeds 33 KB (measured: 4x4 Float 5.2 / 4.0 TB/s to / from SoA
    // with 128 items, 6.1 / 6.3 with 256; 2x3 Double 5.4 -> 6.3 with 256; 8x8 Float 5.5 / 6.4 with 128, 5.2 / 5.5 with 256).
    // 8-byte records of more than 32 elements go to the 64 x 64 tiles (8x8 Double 5.3 / 5.6 -> 5.7 / 6.0 TB/s).
    int ti = elem_bytes == 4 ? 256 : 128;
    const size_t item_tile = (d | 1) * elem_bytes;
    if (ti * item_tile < 8 * 1024) ti *= 2;
    while (ti > 64 && ti * item_tile > 33 * 1024) ti /= 2;
    if (elem_bytes == 4 && !wide) ti = 128;
    if (d <= 64 && !(elem_bytes == 8 && d > 32)) {
        if (elem_bytes == 8) return relayout_small<unsigned long long>(st, ti, d, count, soa_stride, to_soa, src, dst);
        if (elem_bytes == 4 && wide) return relayout_small_pair(st, ti, d, count, soa_stride, to_soa, src, dst);
        if (elem_bytes == 4) return relayout_small<unsigned>(st, ti, d, count, soa_stride, to_soa, src, dst);
    }
    const size_t tiles64 = ((count + 63) / 64) * ((d + 63) / 64);
    if (elem_bytes == 8 && to_soa) {
        relayout_tile64_uniform_kernel<unsigned long long><<<resident_grid((const void*)relayout_tile64_uniform_kernel<unsigned long long>, 256, tiles64), 256, 0, st>>>(
            (const unsigned long long*)src, (unsigned long long*)dst, count, (unsigned)d, soa_stride, 1);
        return ok();
    }
    if (elem_bytes == 8) {
        relayout_tile64_kernel<unsigned long long><<<resident_grid((const void*)relayout_tile64_kernel<unsigned long long>, 256, tiles64), 256, 0, st>>>(
            (const unsigned long long*)src, (unsigned long long*)dst, count, (unsigned)d, soa_stride, to_soa);
        return ok();
    }
    if (elem_bytes == 4 && wide && d % 2 == 0) {
        relayout_tile64_kernel<unsigned><<<resident_grid((const void*)relayout_tile64_kernel<unsigned>, 256, tiles64), 256, 0, st>>>(
            (const unsigned*)src, (unsigned*)dst, count, (unsigned)d, soa_stride, to_soa);
        return ok();
    }
    const size_t tiles = ((count + 31) / 32) * ((d + 31) / 32);
    const void* kern = elem_bytes == 8 ? (const void*)relayout_kernel<unsigned long long>
                     : elem_bytes == 4 ? (const void*)relayout_kernel<unsigned> : (const void*)relayout_kernel<unsigned char>;
    const int grid = resident_grid(kern, 256, tiles);
    if (elem_bytes == 8) relayout_kernel<unsigned long long><<<grid, 256, 0, st>>>((const unsigned long long*)src, (unsigned long long*)dst, count, (unsigned)d, soa_stride, to_soa);
    else if (elem_bytes == 4) relayout_kernel<unsigned><<<grid, 256, 0, st>>>((const unsigned*)src, (unsigned*)dst, count, (unsigned)d, soa_stride, to_soa);
    else if (elem_bytes == 1) relayout_kernel<unsigned char><<<grid, 256, 0, st>>>((const unsigned char*)src, (unsigned char*)dst, count, (unsigned)d, soa_stride, to_soa);
    else return T5I_UNSUPPORTED;
    return ok();
}

}  // namespace t5
