// rowEchelonForm / triangularSolve for the shapes without a register-resident kernel (t5_echelon.cu): any rows x cols with
// rows <= 8, cols <= 16.  AoS only (SoA operands arrive through the host layer's relayout detour).
#include "t5_device.cuh"
#include "t5_internal.h"
#include "t5_recip.cuh"

#include <type_traits>

namespace t5 {

static inline int ech_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else n = T5_SM_COUNT_FALLBACK;
    }
    return n;
}

// =====================================================================================
// rowEchelonForm / triangularSolve (specification: oracle/t5_oracle.cpp
// row_echelon_item): reduced row echelon form of any rows x cols matrix, rows <= 8, cols <= 16.
// The pivot ROW is a run-time quantity here (a column without a pivot does not advance it), so a
// register-resident working matrix is not possible; local memory (3 MB per SM of L1/L2 traffic)
// is what made the first any-n elimination kernels slow.  Instead a CTA stages TPB consecutive
// records into shared memory TRANSPOSED - element e of the CTA's item t at w[e * (TPB+1) + t] -
// with coalesced global reads, every thread reduces its own column of that array in place
// (stride-1 across the warp: conflict-free LDS/STS for any run-time index), and the CTA writes the
// tile back coalesced.  Same operations in the same order as the oracle: bit-identical.
//
// Round 2: the column count NC is a template parameter and the pivot COLUMN c, a run-time loop
// variable, selects code with compile-time column bounds (ech_with_col).  The first version kept
// one body with 16-iteration loops predicated on (j > c && j < n): ncu of 7x9 Double showed
// 19 % ISETP, 17 % branch / reconvergence instructions, 15 % FP64 - 9 800 instructions per
// 32 items of which 1 500 were arithmetic.  The quotients by a pivot use its hoisted exact
// reciprocal (t5_recip.cuh).
//
// Tried and dropped (profiles/r02m_time_echelon_pair_vs_single.txt): two lanes per item (lanes l and l + 16 of a warp own the
// columns of parity 0 and 1, two __syncwarp over the pair per step, multipliers computed by both) to double the warps the
// staged records allow per SM - bit-identical, but the duplicated divisions and boundary predicates cost more than the
// residency bought: 7x9 Double 0.354 -> 0.559 ms, 5x7 0.151 -> 0.323, Float 7x9 0.229 -> 0.362; only 8x16 Double gained
// (0.499 -> 0.422).
// =====================================================================================
constexpr int ECH_MAX_ROWS = 8, ECH_MAX_COLS = 16;

template <int N, int I = 0, class F>
__device__ __forceinline__ void ech_with_col(int c_rt, F&& fn) {
    if constexpr (I < N) {
        if (c_rt == I) fn(std::integral_constant<int, I>{});
        else ech_with_col<N, I + 1>(c_rt, fn);
    }
}

template <class T, int TPB, int NC>
__global__ void __launch_bounds__(TPB)
row_echelon_kernel(const T* __restrict__ A, T* __restrict__ OUT, unsigned char* __restrict__ RANK, int m, int pivot_cols,
                   unsigned long long count) {
    extern __shared__ __align__(16) unsigned char ech_smem[];
    T* w = reinterpret_cast<T*>(ech_smem);
    constexpr int P = TPB + 1;
    constexpr int n = NC;
    const int L = m * n;
    const int tid = threadIdx.x;
    const unsigned long long n_tiles = (count + TPB - 1) / TPB;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long first = tile * TPB;
        const int nv = (count - first < (unsigned long long)TPB) ? (int)(count - first) : TPB;
        const int total = nv * L;
        const T* g = A + first * (size_t)L;
        {
            // asynchronous element copies: a CTA is only 2-4 warps, and with register-staged loads (4 in flight per
            // thread) the kernel ran at the latency, not the bandwidth, of HBM (1.5 TB/s for Float and Double alike)
            int item = tid / L, e = tid - item * L;
            for (int idx = tid; idx < total; idx += TPB) {
                cp_async_small<(int)sizeof(T)>(w + e * P + item, g + idx);
                e += TPB;
                while (e >= L) { e -= L; ++item; }
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        if (tid < nv) {
            T* x = w + tid;                                     // element (i, j) of this item: x[(i * n + j) * P]
            int rank = 0;
            for (int c = 0; c < pivot_cols && rank < m; ++c) {
                int p = rank;
                while (p < m && x[(p * n + c) * P] == T(0)) ++p;
                if (p == m) continue;
                if (p != rank) {
                    #pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        const T t = x[(p * n + j) * P];
                        x[(p * n + j) * P] = x[(rank * n + j) * P];
                        x[(rank * n + j) * P] = t;
                    }
                }
                T* pr = x + (rank * n) * P;
                ech_with_col<NC>(c, [&](auto C_) {
                    constexpr int C = decltype(C_)::value;
                    // m - 1 multipliers and NC - 1 - C normalisations divide by this pivot: exact division, reciprocal hoisted
                    const Recip<T> rc(pr[C * P]);
                    // the pivot row and each target row pass through registers: with both rows addressed through
                    // shared-memory pointers the compiler must keep every load behind the previous store, and the
                    // updates ran one shared-memory latency at a time
                    T prow[NC];
                    #pragma unroll
                    for (int j = C + 1; j < NC; ++j) prow[j] = pr[j * P];
                    // (Double: all multipliers of the column as one batch of independent exact divisions, one fallback test)
                    T fb[ECH_MAX_ROWS];
                    (void)fb;
                    if constexpr (kBatchDiv<T>) {
                        T cv[ECH_MAX_ROWS];
                        bool all_fast = true;
                        #pragma unroll
                        for (int i = 0; i < ECH_MAX_ROWS; ++i) {
                            cv[i] = (i < m && i != rank) ? x[(i * n + C) * P] : T(1);
                            fb[i] = rc.div_try(cv[i], all_fast);
                        }
                        if (!all_fast) {
                            #pragma unroll
                            for (int i = 0; i < ECH_MAX_ROWS; ++i) fb[i] = rc.div_fix(cv[i], fb[i]);
                        }
                    }
                    #pragma unroll
                    for (int i = 0; i < ECH_MAX_ROWS; ++i) {
                        if (i >= m || i == rank) continue;
                        T* ri = x + (i * n) * P;
                        T f;
                        if constexpr (kBatchDiv<T>) f = fb[i]; else f = rc.div(ri[C * P]);
                        T rrow[NC];
                        #pragma unroll
                        for (int j = C + 1; j < NC; ++j) rrow[j] = ri[j * P];
                        #pragma unroll
                        for (int j = C + 1; j < NC; ++j) ri[j * P] = rrow[j] - f * prow[j];
                        ri[C * P] = T(0);
                    }
                    if constexpr (kBatchDiv<T>) {
                        T q[NC];
                        bool all_fast = true;
                        #pragma unroll
                        for (int j = C + 1; j < NC; ++j) q[j] = rc.div_try(prow[j], all_fast);
                        if (!all_fast) {
                            #pragma unroll
                            for (int j = C + 1; j < NC; ++j) q[j] = rc.div_fix(prow[j], q[j]);
                        }
                        #pragma unroll
                        for (int j = C + 1; j < NC; ++j) pr[j * P] = q[j];
                    } else {
                        #pragma unroll
                        for (int j = C + 1; j < NC; ++j) pr[j * P] = rc.div(prow[j]);
                    }
                    pr[C * P] = T(1);
                });
                ++rank;
            }
            RANK[first + tid] = (unsigned char)rank;
        }
        __syncthreads();
        {
            T* o = OUT + first * (size_t)L;
            int item = tid / L, e = tid - item * L;
            #pragma unroll 4
            for (int idx = tid; idx < total; idx += TPB) {
                o[idx] = w[e * P + item];
                e += TPB;
                while (e >= L) { e -= L; ++item; }
            }
        }
        __syncthreads();
    }
}

template <class T, int TPB, int NC>
static int launch_row_echelon(const Shard& s, int m, int pivot_cols, const void* A, void* O, void* rank) {
    auto kern = row_echelon_kernel<T, TPB, NC>;
    const int smem = m * NC * (TPB + 1) * (int)sizeof(T);
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return T5I_CUDA;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
    const size_t tiles = (s.count + TPB - 1) / TPB, cap = (size_t)occ * ech_sm_count();
    const int grid = (int)(tiles < cap ? tiles : cap);
    kern<<<grid, TPB, smem, s.stream>>>((const T*)A, (T*)O, (unsigned char*)rank, m, pivot_cols, (unsigned long long)s.count);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

template <class T, int NC = 1>
static int dispatch_cols(const Shard& s, int m, int n, int pivot_cols, const void* A, void* O, void* rank) {
    if constexpr (NC > ECH_MAX_COLS) {
        return T5I_UNSUPPORTED;
    } else {
        if (n != NC) return dispatch_cols<T, NC + 1>(s, m, n, pivot_cols, A, O, rank);
        // small records: 128 items per CTA (more bytes in flight per CTA); large ones: 64, so that several CTAs fit an SM
        const bool small = (size_t)m * n * sizeof(T) <= 192;
        return small ? launch_row_echelon<T, 128, NC>(s, m, pivot_cols, A, O, rank) : launch_row_echelon<T, 64, NC>(s, m, pivot_cols, A, O, rank);
    }
}

int generic_row_echelon(const Shard& s, int dtype, int m, int n, int pivot_cols, const void* A, void* O, void* rank) {
    if (s.layout != T5L_AOS || m < 1 || n < 1 || m > ECH_MAX_ROWS || n > ECH_MAX_COLS || pivot_cols < 0 || pivot_cols > n) return T5I_UNSUPPORTED;
    if (s.count == 0) return T5I_OK;
    if (dtype == T5D_F64) return dispatch_cols<double>(s, m, n, pivot_cols, A, O, rank);
    if (dtype == T5D_F32) return dispatch_cols<float>(s, m, n, pivot_cols, A, O, rank);
    return T5I_UNSUPPORTED;
}

}  // namespace t5
