// Gaussian elimination with ONE ROW PER LANE for the sizes and element types where it beats the CTA-per-matrix kernel
// of t5_lu.cuh (SquareMatrix det / inverse, LinearSystem solve of SURVEY.md §8a a9-a11): a matrix lives in the
// registers of W = 16 or 32 lanes (two matrices per warp for n <= 16) and everything the lanes tell each other travels
// by warp shuffle - no shared-memory round trip and no barrier inside the elimination.
// (kernel + launch templates; instantiated per operation by t5_warp_lu_det.cu / _solve.cu / _inverse.cu.)
//
// Scope, from measurements (tools/time_lu.py, profiles/r02am_warp_lu.txt).  The kernel is bound by the shuffle /
// shared-memory data path (ncu of det 16x16 Double: 82 % of the LSU wavefront peak, `mio_throttle` the top stall): an
// entry of the pivot row costs one SHFL per 32 bits against the multiply + subtract it feeds.  So it serves
//   Float,  9 <= n <= 16: det, inverse, solve        (16x16: solve 2.39 -> 1.46 ms, inverse 3.48 -> 2.29)
//   Float, 30 <= n <= 32: det, inverse, solve with more than 8 right-hand sides   (32x32 inverse 8.04 -> 6.34)
// and everything else stays on t5_lu.cuh (Double: two SHFL per entry lose to its shared-memory tiles, so the
// kernel is instantiated for Float only; 17 <= n <= 29: too many idle lanes in a
// 32-lane frame; few right-hand sides at n >= 30: t5_lu.cuh eliminates them together with A).  Two variants that were tried and dropped: publishing the
// pivot row through shared memory with 16-byte broadcast loads and a zero row for the lanes above the front (no per-element
// select), and shuffling the next pivot out early to overlap its reciprocal - neither beat the plain version.
//
// Frame.  The n x n matrix sits in the BOTTOM-RIGHT corner of a W x W frame whose top-left corner is the identity:
// lane i holds frame row i in r[0..W) (compile-time indices), real entry (a, b) is frame (a + k0, b + k0), k0 = W - n.
// Steps k < k0 would divide by the identity's ones and subtract zeros: they are skipped (k0 is uniform over the grid),
// and every column j > k of a real step is a real column - no padded work for any n, only idle lanes.
//
// Step k >= k0:  pivot = r[k] of lane k.  Exactly zero -> (warp-wide entry, __any_sync) the first lane p > k with
// r[k] != 0 (ballot + ffs = the oracle's first non-zero row) exchanges its whole row - right-hand sides included - with
// lane k by shuffle; none -> singular.  f = r[k] / pivot in every lane; lanes below k: r[j] -= f * shfl(r[j], k) for
// j > k, and the same for the right-hand-side columns.  Per element these are the oracle's operations in the oracle's
// order (oracle/t5_oracle.cpp forward_eliminate; this TU is compiled -fmad=false -prec-div=true): BIT-IDENTICAL.
// det = (+-) fold_k (acc * pivot_k) from 1, singular -> +0.
// solve / inverse: the eliminated right-hand sides are transposed through shared memory so that lane c owns COLUMN c,
// and back-substitution runs one column per lane with u[k][j] broadcast from lane k by shuffle
// (s = fold_{j>k ascending}(s + u[k,j] x[j]) from 0, x[k] = (y[k] - s) / u[k,k]: the specification's order).
#pragma once
#include <type_traits>

#include "t5_device.cuh"
#include "t5_internal.h"

namespace t5 {
namespace wlu {

struct Args {
    const void* A;
    const void* B;          // n x nrhs per item (solve); unused for det / inverse
    void* X;                // determinant / solution / inverse
    unsigned char* status;  // unused for det
    unsigned long long count;
    int n, nrhs;
};

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

constexpr int WARPS = 4;            // per CTA; warps never talk to each other

// MODE 0: det, 1: solve (NC register columns of right-hand sides, nrhs <= NC), 2: inverse (NC == W)
template <class T, int W, int MODE, int NC>
__global__ void __launch_bounds__(WARPS * 32) wlu_kernel(Args p) {
    static_assert(W == 16 || W == 32, "a matrix per half-warp or per warp");
    constexpr int GPW = 32 / W;                         // matrices per warp
    constexpr int SLOT = W * (W + 1);                   // staging elements per matrix (input rows padded to an odd pitch)
    extern __shared__ __align__(16) unsigned char wlu_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / W, li = lane % W;            // group within the warp, lane within the group = frame row
    T* S = reinterpret_cast<T*>(wlu_smem) + (size_t)(warp * GPW + grp) * SLOT;
    const int n = p.n, k0 = W - n;
    const int pin = n | 1;                              // odd pitch: the lanes of a half-warp read one column conflict-free
    const int ri = li - k0;                             // real row of this lane (< 0: identity rows)
    const int ncol = MODE == 2 ? n : p.nrhs;            // right-hand-side columns
    const int coff = MODE == 2 ? k0 : 0;                // inverse: column c of the frame is real column c - k0
    const T* gA = (const T*)p.A;

    const unsigned long long per_pass = (unsigned long long)gridDim.x * WARPS * GPW;
    for (unsigned long long base = (unsigned long long)blockIdx.x * WARPS * GPW; base < p.count; base += per_pass) {
        // (the groups of a warp stay in lockstep: a group past the end of the batch recomputes the last item, stores nothing)
        const unsigned long long mine = base + (unsigned long long)(warp * GPW + grp);
        const bool live = mine < p.count;
        const unsigned long long item = live ? mine : p.count - 1;
        if (!__any_sync(0xffffffffu, live)) continue;   // (warp-uniform: every group of this warp is past the end)

        // ---- A -> shared (each lane its own row; the L1 absorbs the row-strided pattern) -> frame rows in registers
        const T* a_item = gA + item * (unsigned long long)n * n;
        if (ri >= 0)
            for (int c = 0; c < n; ++c) cp_async_small<(int)sizeof(T)>(S + ri * pin + c, a_item + (size_t)ri * n + c);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        T r[W];
        #pragma unroll
        for (int j = 0; j < W; ++j) {
            T v = (j == li) ? T(1) : T(0);
            if (j >= k0 && ri >= 0) v = S[ri * pin + (j - k0)];
            r[j] = v;
        }
        T y[NC > 0 ? NC : 1];
        if constexpr (MODE == 2) {
            #pragma unroll
            for (int c = 0; c < NC; ++c) y[c] = (c == li) ? T(1) : T(0);
        } else if constexpr (MODE == 1) {
            const T* b_item = (const T*)p.B + item * (unsigned long long)n * ncol;
            #pragma unroll
            for (int c = 0; c < NC; ++c) y[c] = (ri >= 0 && c < ncol) ? __ldg(b_item + (size_t)ri * ncol + c) : T(0);
        }
        __syncwarp();                                   // the staging slot is reused below

        // ---- elimination
        T acc = T(1);
        bool singular = false;
        int swaps = 0;
        static_for<0, W>([&](auto k_) {
            constexpr int k = decltype(k_)::value;
            if (k >= k0) {
                T piv = __shfl_sync(0xffffffffu, r[k], k, W);
                const bool need = piv == T(0);
                if (__any_sync(0xffffffffu, need)) {
                    // rare: planted zeros, exactly singular items.  First row below k with a non-zero entry in column k.
                    const unsigned bal = __ballot_sync(0xffffffffu, li > k && r[k] != T(0));
                    const unsigned mine_bits = W == 32 ? bal : ((bal >> (grp * W)) & 0xffffu);
                    const int pl = mine_bits ? __ffs(mine_bits) - 1 : -1;
                    const bool sw = need && pl >= 0;
                    if (need && pl < 0) singular = true;
                    // lanes k and pl trade rows (everything a row carries: its right-hand sides too); other lanes read themselves
                    const int src = sw ? (li == k ? pl : (li == pl ? k : li)) : li;
                    #pragma unroll
                    for (int j = 0; j < W; ++j) r[j] = __shfl_sync(0xffffffffu, r[j], src, W);
                    if constexpr (MODE != 0) {
                        #pragma unroll
                        for (int c = 0; c < NC; ++c) y[c] = __shfl_sync(0xffffffffu, y[c], src, W);
                    }
                    swaps += sw ? 1 : 0;
                    piv = __shfl_sync(0xffffffffu, r[k], k, W);
                }
                acc = acc * piv;
                const T f = r[k] / piv;
                const bool below = li > k;
                #pragma unroll
                for (int j = k + 1; j < W; ++j) {
                    const T pk = __shfl_sync(0xffffffffu, r[j], k, W);
                    const T upd = r[j] - f * pk;
                    r[j] = below ? upd : r[j];
                }
                if constexpr (MODE != 0) {
                    #pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        if (c >= coff && c - coff < ncol) {                  // (uniform)
                            const T yk = __shfl_sync(0xffffffffu, y[c], k, W);
                            const T upd = y[c] - f * yk;
                            y[c] = below ? upd : y[c];
                        }
                    }
                }
            }
        });

        if constexpr (MODE == 0) {
            if (li == 0 && live) ((T*)p.X)[item] = singular ? T(0) : ((swaps & 1) ? -acc : acc);
        } else {
            // ---- right-hand sides: row per lane -> column per lane (through the staging slot, pitch W + 1)
            #pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c >= coff && c - coff < ncol) S[li * (W + 1) + c] = y[c];
            __syncwarp();
            T x[W];                                     // lane c: column c; x[k] starts as the eliminated right-hand side
            #pragma unroll
            for (int k = 0; k < W; ++k) x[k] = (k >= k0) ? S[k * (W + 1) + li] : T(0);
            __syncwarp();
            static_for<0, W>([&](auto kk_) {
                constexpr int k = W - 1 - decltype(kk_)::value;
                if (k >= k0) {
                    T s = T(0);
                    #pragma unroll
                    for (int j = k + 1; j < W; ++j) {
                        const T u = __shfl_sync(0xffffffffu, r[j], k, W);
                        s = s + u * x[j];
                    }
                    const T ukk = __shfl_sync(0xffffffffu, r[k], k, W);
                    x[k] = (x[k] - s) / ukk;
                }
            });
            // ---- results: X[k - k0][c - coff] from lane c -> staging (pitch ncol) -> one coalesced run
            const int cr = li - coff;
            if (cr >= 0 && cr < ncol) {
                #pragma unroll
                for (int k = 0; k < W; ++k)
                    if (k >= k0) S[(k - k0) * ncol + cr] = singular ? T(0) : x[k];
            }
            __syncwarp();
            if (live) {
                T* gX = (T*)p.X + item * (unsigned long long)n * ncol;
                for (int e = li; e < n * ncol; e += W) gX[e] = S[e];
                if (li == 0) p.status[item] = singular ? 1 : 0;
            }
            __syncwarp();
        }
    }
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else n = T5_SM_COUNT_FALLBACK;
    }
    return n;
}

template <class T, int W, int MODE, int NC>
static int launch(const Shard& s, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    Args p;
    p.A = A; p.B = B; p.X = X; p.status = (unsigned char*)status; p.count = s.count; p.n = n; p.nrhs = nrhs;
    constexpr int GPW = 32 / W;
    const size_t smem = sizeof(T) * (size_t)W * (W + 1) * WARPS * GPW;
    auto kern = wlu_kernel<T, W, MODE, NC>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return T5I_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem) != cudaSuccess || per_sm < 1) return T5I_CUDA;
    const unsigned long long slots = (unsigned long long)per_sm * sm_count();
    const unsigned long long want = (s.count + WARPS * GPW - 1) / (WARPS * GPW);
    const int grid = (int)(want < slots ? want : slots);
    kern<<<grid, WARPS * 32, smem, s.stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

// Sizes and element types of the scope table above; T5I_NOT_SPECIALISED otherwise (the caller goes on to t5_lu.cuh).
template <class T, int MODE>
static int dispatch(const Shard& s, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    static_assert(sizeof(T) == 4, "instantiated for Float only: with two SHFL per entry Double loses to t5_lu.cuh at every size");
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    const bool small = n >= 9 && n <= 16, wide = n >= 30 && n <= 32;
    if (!small && !wide) return T5I_NOT_SPECIALISED;
    if (MODE == 1 && (nrhs < 1 || nrhs > (small ? 16 : 32))) return T5I_NOT_SPECIALISED;
    if (MODE == 1 && wide && nrhs <= 8) return T5I_NOT_SPECIALISED;   // t5_lu.cuh with the right-hand sides riding along: 2.67 vs 2.90 ms
    if (s.count == 0) return T5I_OK;
    if constexpr (MODE == 0) {
        return small ? launch<T, 16, 0, 0>(s, n, 0, A, B, X, status) : launch<T, 32, 0, 0>(s, n, 0, A, B, X, status);
    } else if constexpr (MODE == 2) {
        return small ? launch<T, 16, 2, 16>(s, n, n, A, B, X, status) : launch<T, 32, 2, 32>(s, n, n, A, B, X, status);
    } else {
        if (!small) return launch<T, 32, 1, 32>(s, n, nrhs, A, B, X, status);
        if (nrhs == 1) return launch<T, 16, 1, 1>(s, n, nrhs, A, B, X, status);
        if (nrhs <= 4) return launch<T, 16, 1, 4>(s, n, nrhs, A, B, X, status);
        return launch<T, 16, 1, 16>(s, n, nrhs, A, B, X, status);
    }
}

}  // namespace wlu
}  // namespace t5
