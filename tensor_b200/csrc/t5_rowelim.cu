// Gaussian elimination for 5 <= n <= 8: ROW PER LANE, eight lanes per matrix, four matrices per
// warp (north_star: "eliminations are one-thread- or one-warp-per-matrix ... warp shuffles").
//
// A thread-per-item kernel would need the whole n x 2n working matrix per thread: 128 doubles for
// an 8x8 inverse - local memory, 3 MB of it per SM, i.e. L2/HBM traffic on every row operation
// (the any-shape kernel this replaces ran at 0.07-0.22 of the HBM peak).  Here lane r of a group
// keeps ROW r of [A | rhs] in registers (<= 16 values); per column k the pivot row is found with
// one ballot, broadcast with width-8 shuffles, and every lane updates its own row.  Loads are
// coalesced by construction (a group reads one contiguous n*n record, a warp four of them).
//
// Arithmetic is the oracle's, element for element (this TU is compiled with -fmad=false,
// -prec-div=true): pivot = FIRST row >= k with a non-zero entry (exact compare), swap,
// f = w[i,k] / w[k,k], w[i,j] = w[i,j] - f * w[k,j] for j > k including the rhs columns;
// det = (+-) fold_k (acc * pivot_k) from 1; back-substitution s = fold_{j>k ascending}(s + u[k,j] x[j])
// from 0, x[k] = (y[k] - s) / u[k,k].  For the back-substitution the rhs is transposed across the
// group so that lane c owns rhs COLUMN c and runs the sequential recurrence for it, with the
// u[k,j] broadcast from lane k - the order of operations per element is unchanged, so results are
// bit-identical to the oracle.  Singular (no non-zero pivot in some column): status 1, item
// zero-filled, det = +0.
#include "t5_device.cuh"
#include "t5_internal.h"

namespace t5 {
namespace re {

constexpr int THREADS = 128;
constexpr unsigned FULL = 0xffffffffu;

struct Args {
    const void* A;
    const void* B;          // rhs (n x nrhs per item); unused for det / inverse
    void* X;                // solution / inverse / determinant
    unsigned char* status;  // unused for det
    unsigned long long count;
};

template <class T>
__device__ __forceinline__ T bcast(T v, int src) { return __shfl_sync(FULL, v, src, 8); }

// MODE 0: det, 1: solve (NR right-hand sides from B), 2: inverse (rhs = identity, NR == N)
template <class T, int N, int NR, int MODE>
__global__ void __launch_bounds__(THREADS) rowelim_kernel(Args p) {
    constexpr int W = N + NR;
    const int lane = threadIdx.x & 31, g = lane >> 3, r = lane & 7;
    const unsigned long long warp0 = ((unsigned long long)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * THREADS) >> 5;
    const T* gA = (const T*)p.A;
    const T* gB = (const T*)p.B;
    T* gX = (T*)p.X;
    const unsigned long long n_wt = (p.count + 3) / 4;

    for (unsigned long long wt = warp0; wt < n_wt; wt += nwarps) {
        const unsigned long long b = wt * 4 + g;
        const bool item = b < p.count;
        const bool row = item && r < N;

        T w[W];
        #pragma unroll
        for (int j = 0; j < W; ++j) w[j] = T(0);
        if (row) {
            const T* a = gA + (b * N + r) * (size_t)N;
            #pragma unroll
            for (int j = 0; j < N; ++j) w[j] = __ldg(a + j);
            if constexpr (MODE == 1) {
                const T* y = gB + (b * N + r) * (size_t)NR;
                #pragma unroll
                for (int c = 0; c < NR; ++c) w[N + c] = __ldg(y + c);
            } else if constexpr (MODE == 2) {
                #pragma unroll
                for (int c = 0; c < NR; ++c) w[N + c] = (r == c) ? T(1) : T(0);
            }
        }

        bool singular = false;
        int swaps = 0;
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            // first row >= k of this group with a non-zero entry in column k
            const bool nz = row && r >= k && w[k] != T(0);
            const unsigned grp = (__ballot_sync(FULL, nz) >> (g * 8)) & 0xffu;
            singular = singular || grp == 0;
            const int piv = grp ? __ffs((int)grp) - 1 : k;
            if (__any_sync(FULL, piv != k)) {            // rare: some group of the warp swaps rows k and piv
                const int src = r == k ? piv : (r == piv ? k : r);
                #pragma unroll
                for (int j = k; j < W; ++j) w[j] = bcast(w[j], src);   // columns < k of rows >= k are zero already
                swaps += piv != k ? 1 : 0;
            }
            // pivot row to every lane of the group, then each row below it eliminates its entry
            T pk[W];
            #pragma unroll
            for (int j = k; j < W; ++j) pk[j] = bcast(w[j], k);
            if (r > k) {
                const T f = w[k] / pk[k];
                #pragma unroll
                for (int j = k + 1; j < W; ++j) w[j] = w[j] - f * pk[j];
                w[k] = T(0);
            }
        }

        if constexpr (MODE == 0) {
            T acc = T(1);
            #pragma unroll
            for (int k = 0; k < N; ++k) acc = acc * bcast(w[k], k);    // diagonal entry k lives in lane k
            acc = (swaps & 1) ? -acc : acc;
            if (item && r == 0) gX[b] = singular ? T(0) : acc;
        } else {
        // lane c takes rhs column c: ycol[k] = Y[k][c]
        T ycol[N];
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            ycol[k] = T(0);
            #pragma unroll
            for (int c = 0; c < NR; ++c) {
                const T v = bcast(w[N + c], k);
                if (r == c) ycol[k] = v;
            }
        }
        T x[N];
        #pragma unroll
        for (int k = N - 1; k >= 0; --k) {
            T s = T(0);
            #pragma unroll
            for (int j = k + 1; j < N; ++j) s = s + bcast(w[j], k) * x[j];   // u[k][j] from lane k
            x[k] = (ycol[k] - s) / bcast(w[k], k);
        }
        if (item && r < NR) {
            #pragma unroll
            for (int k = 0; k < N; ++k) gX[(b * N + k) * (size_t)NR + r] = singular ? T(0) : x[k];
        }
        if (item && r == 0) p.status[b] = singular ? 1 : 0;
        }
    }
}

template <class T, int N, int NR, int MODE>
static int launch(const Shard& s, const void* A, const void* B, void* X, void* status) {
    auto kern = rowelim_kernel<T, N, NR, MODE>;
    static int cap[T5_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!cap[dev]) {
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, 0) != cudaSuccess) return T5I_CUDA;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cap[dev] = (occ < 1 ? 1 : occ) * (sms < 1 ? T5_SM_COUNT_FALLBACK : sms);
    }
    if (s.count == 0) return T5I_OK;
    const unsigned long long blocks = ((s.count + 3) / 4 + THREADS / 32 - 1) / (THREADS / 32);
    const int grid = (int)(blocks < (unsigned long long)cap[dev] ? blocks : (unsigned long long)cap[dev]);
    Args a{A, B, X, (unsigned char*)status, (unsigned long long)s.count};
    kern<<<grid, THREADS, 0, s.stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

template <class T, int MODE>
static int by_n(const Shard& s, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    // Measured against the thread-per-item any-shape kernel (tools/survey_ops.py, 1Mi items): with 2n
    // columns per row (inverse, matrix right-hand sides) the row-per-lane form wins for every n here
    // (8x8 Double inverse 2.29 -> 0.63 ms, 6x6 0.82 -> 0.58 ms); with n or n+1 columns it wins from
    // n = 7 (8x8 det 0.60 -> 0.26 ms) and loses slightly at n = 5, 6, which stay on the other kernel.
    const bool wide = MODE == 2 || (MODE == 1 && nrhs == n);
    if (!wide && n < 7) return T5I_NOT_SPECIALISED;
#define T5_RE(N_) \
    if (n == N_) { \
        if (MODE == 0) return launch<T, N_, 0, 0>(s, A, nullptr, X, nullptr); \
        if (MODE == 2) return launch<T, N_, N_, 2>(s, A, nullptr, X, status); \
        if (nrhs == 1) return launch<T, N_, 1, 1>(s, A, B, X, status); \
        if (nrhs == N_) return launch<T, N_, N_, 1>(s, A, B, X, status); \
        return T5I_NOT_SPECIALISED; \
    }
    T5_RE(5) T5_RE(6) T5_RE(7) T5_RE(8)
#undef T5_RE
    return T5I_NOT_SPECIALISED;
}

}  // namespace re

int rowlane_det(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) return re::by_n<double, 0>(s, n, 0, A, nullptr, O, nullptr);
    if (dtype == T5D_F32) return re::by_n<float, 0>(s, n, 0, A, nullptr, O, nullptr);
    return T5I_NOT_SPECIALISED;
}

int rowlane_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) return re::by_n<double, 2>(s, n, n, A, nullptr, O, status);
    if (dtype == T5D_F32) return re::by_n<float, 2>(s, n, n, A, nullptr, O, status);
    return T5I_NOT_SPECIALISED;
}

int rowlane_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (dtype == T5D_F64) return re::by_n<double, 1>(s, n, nrhs, A, B, X, status);
    if (dtype == T5D_F32) return re::by_n<float, 1>(s, n, nrhs, A, B, X, status);
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
