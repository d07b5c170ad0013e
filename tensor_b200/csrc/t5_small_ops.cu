// Thread-per-item kernels for the tiny fixed shapes (2x2 .. 4x4 matrices, 2..4
// vectors): the HBM-bound part of the hot path (SURVEY.md §8a a4, a6, a7, a9-a11,
// BASELINE configs 1-3).  Arithmetic is written once as register-level `Op::run`
// and instantiated in the AoS tile pipeline and the SoA skeleton (t5_tile.cuh).
//
// This translation unit MUST be compiled with -fmad=false: the specification is
// multiply-round-add-round in index order, which makes Double/Float results
// bit-identical to the CPU oracle, not merely within tolerance.
#include "t5_tile.cuh"
#include "t5_internal.h"

namespace t5 {

// ---- .*. : C[i,c] = fold_{j asc} (acc + A[i,j]*B[j,c]), acc0 = 0  (§8a a6) -------
template <class T, int M, int K, int N>
struct MatmulOp {
    using In0 = Rec<T, M * K>; using In1 = Rec<T, K * N>;
    using Out0 = Rec<T, M * N>; using Out1 = NoRec;
    __device__ __forceinline__ static void run(const T* a, const T* b, T* c, unsigned char*) {
        using R = Arith<T>;
        #pragma unroll
        for (int i = 0; i < M; ++i)
            #pragma unroll
            for (int col = 0; col < N; ++col) {
                T acc = R::zero();
                #pragma unroll
                for (int j = 0; j < K; ++j) acc = R::add(acc, R::mul(a[i * K + j], b[j * N + col]));
                c[i * N + col] = acc;
            }
    }
};

// ---- transpose of an R x C matrix: pure register renaming (§8a a4) -----------------
template <class T, int R_, int C_>
struct TransposeOp {
    using In0 = Rec<T, R_ * C_>; using In1 = NoRec;
    using Out0 = Rec<T, R_ * C_>; using Out1 = NoRec;
    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* o, unsigned char*) {
        #pragma unroll
        for (int i = 0; i < R_; ++i)
            #pragma unroll
            for (int j = 0; j < C_; ++j) o[j * R_ + i] = a[i * C_ + j];
    }
};

// ---- Gaussian elimination, first-non-zero pivot (§8a pivot rule) --------------------
// w is N x W in registers; every index is a compile-time constant after unrolling,
// the pivot search is a chain of predicated row swaps (no divergence, no local mem).
template <class T, int N, int W>
__device__ __forceinline__ bool eliminate(T (&w)[N][W], int& swaps) {
    bool singular = false;
    swaps = 0;
    #pragma unroll
    for (int k = 0; k < N; ++k) {
        #pragma unroll
        for (int r = k + 1; r < N; ++r) {
            const bool sw = (w[k][k] == T(0)) && (w[r][k] != T(0));
            // columns < k of rows >= k are exact zeros already: swapping them is a no-op
            #pragma unroll
            for (int j = k; j < W; ++j) {
                T x = w[k][j], y = w[r][j];
                w[k][j] = sw ? y : x;
                w[r][j] = sw ? x : y;
            }
            swaps += sw ? 1 : 0;
        }
        singular = singular || (w[k][k] == T(0));
        #pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const T f = w[i][k] / w[k][k];
            #pragma unroll
            for (int j = k + 1; j < W; ++j) w[i][j] = w[i][j] - f * w[k][j];
            w[i][k] = T(0);
        }
    }
    return !singular;
}

template <class T, int N>
struct DetOp {
    using In0 = Rec<T, N * N>; using In1 = NoRec;
    using Out0 = Rec<T, 1>; using Out1 = NoRec;
    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* o, unsigned char*) {
        T w[N][N];
        #pragma unroll
        for (int i = 0; i < N; ++i)
            #pragma unroll
            for (int j = 0; j < N; ++j) w[i][j] = a[i * N + j];
        int swaps;
        const bool ok = eliminate<T, N, N>(w, swaps);
        T acc = T(1);
        #pragma unroll
        for (int k = 0; k < N; ++k) acc = acc * w[k][k];
        acc = (swaps & 1) ? -acc : acc;
        o[0] = ok ? acc : T(0);
    }
};

// solve A X = B, B is N x NRHS; also the body of inverse (B = I).
template <class T, int N, int NRHS>
__device__ __forceinline__ void solve_regs(T (&w)[N][N + NRHS], T* x, unsigned char* status) {
    int swaps;
    const bool ok = eliminate<T, N, N + NRHS>(w, swaps);
    T xs[N][NRHS];
    #pragma unroll
    for (int k = N - 1; k >= 0; --k)
        #pragma unroll
        for (int c = 0; c < NRHS; ++c) {
            T s = T(0);
            #pragma unroll
            for (int j = k + 1; j < N; ++j) s = s + w[k][j] * xs[j][c];
            xs[k][c] = (w[k][N + c] - s) / w[k][k];
        }
    #pragma unroll
    for (int k = 0; k < N; ++k)
        #pragma unroll
        for (int c = 0; c < NRHS; ++c) x[k * NRHS + c] = ok ? xs[k][c] : T(0);
    status[0] = ok ? 0 : 1;
}

template <class T, int N, int NRHS>
struct SolveOp {
    using In0 = Rec<T, N * N>; using In1 = Rec<T, N * NRHS>;
    using Out0 = Rec<T, N * NRHS>; using Out1 = Rec<unsigned char, 1>;
    __device__ __forceinline__ static void run(const T* a, const T* b, T* x, unsigned char* status) {
        T w[N][N + NRHS];
        #pragma unroll
        for (int i = 0; i < N; ++i) {
            #pragma unroll
            for (int j = 0; j < N; ++j) w[i][j] = a[i * N + j];
            #pragma unroll
            for (int c = 0; c < NRHS; ++c) w[i][N + c] = b[i * NRHS + c];
        }
        solve_regs<T, N, NRHS>(w, x, status);
    }
};

template <class T, int N>
struct InverseOp {
    using In0 = Rec<T, N * N>; using In1 = NoRec;
    using Out0 = Rec<T, N * N>; using Out1 = Rec<unsigned char, 1>;
    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* x, unsigned char* status) {
        T w[N][2 * N];
        #pragma unroll
        for (int i = 0; i < N; ++i) {
            #pragma unroll
            for (int j = 0; j < N; ++j) w[i][j] = a[i * N + j];
            #pragma unroll
            for (int c = 0; c < N; ++c) w[i][N + c] = (i == c) ? T(1) : T(0);
        }
        solve_regs<T, N, N>(w, x, status);
    }
};

// ---- launchers ---------------------------------------------------------------------
template <class Op, int THREADS, int IPT, int STAGES>
static int launch_tile(cudaStream_t st, const void* in0, const void* in1, void* out0, void* out1, size_t n) {
    using C = TileCfg<Op, THREADS, IPT, STAGES>;
    static_assert(C::SMEM <= 227 * 1024, "tile does not fit in shared memory");
    auto kern = tile_kernel<Op, THREADS, IPT, STAGES>;
    static int grid_cap[T5_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= T5_MAX_DEVICES) return T5I_CUDA;
    if (!grid_cap[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) return T5I_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int occ = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, C::SMEM) != cudaSuccess) return T5I_CUDA;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (occ < 1) occ = 1;
        if (sms < 1) sms = T5_SM_COUNT_FALLBACK;
        grid_cap[dev] = occ * sms;
    }
    if (n == 0) return T5I_OK;
    size_t tiles = (n + C::TILE - 1) / C::TILE;
    int grid = (int)(tiles < (size_t)grid_cap[dev] ? tiles : (size_t)grid_cap[dev]);
    TileArgs a{in0, in1, out0, out1, (unsigned long long)n};
    kern<<<grid, THREADS, C::SMEM, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

template <class Op>
static int launch_soa(cudaStream_t st, const void* in0, const void* in1, void* out0, void* out1, size_t n, size_t stride) {
    constexpr int THREADS = 256;
    if (n == 0) return T5I_OK;
    size_t blocks = (n + THREADS - 1) / THREADS;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = T5_SM_COUNT_FALLBACK;
    size_t cap = (size_t)sms * 8 * 4;  // 4 waves of full occupancy, grid-stride beyond
    int grid = (int)(blocks < cap ? blocks : cap);
    SoaArgs a{in0, in1, out0, out1, (unsigned long long)n, (unsigned long long)stride};
    soa_kernel<Op, THREADS><<<grid, THREADS, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

// Tile geometry: 128 threads, 1 record per thread; ring depth chosen so a CTA
// stays near 48-64 KB and 3+ CTAs share an SM.
template <class Op>
static int launch_op(const Shard& s, const void* in0, const void* in1, void* out0, void* out1) {
    if (s.layout == T5L_SOA) return launch_soa<Op>(s.stream, in0, in1, out0, out1, s.count, s.soa_stride);
    constexpr int in_bytes = Op::In0::stride + Op::In1::stride;
    if constexpr (in_bytes <= 48) return launch_tile<Op, 256, 2, 4>(s.stream, in0, in1, out0, out1, s.count);
    else if constexpr (in_bytes <= 160) return launch_tile<Op, 128, 1, 4>(s.stream, in0, in1, out0, out1, s.count);
    else return launch_tile<Op, 128, 1, 3>(s.stream, in0, in1, out0, out1, s.count);
}

template <class T>
static int matmul_t(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
#define T5_MM(M_, K_, N_) \
    if (m == M_ && k == K_ && n == N_) return launch_op<MatmulOp<T, M_, K_, N_>>(s, A, B, C, nullptr);
    T5_MM(2, 2, 2) T5_MM(3, 3, 3) T5_MM(4, 4, 4)
    T5_MM(2, 2, 1) T5_MM(3, 3, 1) T5_MM(4, 4, 1)
    T5_MM(1, 2, 2) T5_MM(1, 3, 3) T5_MM(1, 4, 4)
    T5_MM(1, 2, 1) T5_MM(1, 3, 1) T5_MM(1, 4, 1)
#undef T5_MM
    return T5I_NOT_SPECIALISED;
}

int small_matmul(const Shard& s, int dtype, int m, int k, int n, const void* A, const void* B, void* C) {
    switch (dtype) {
        case T5D_F64: return matmul_t<double>(s, m, k, n, A, B, C);
        case T5D_F32: return matmul_t<float>(s, m, k, n, A, B, C);
        case T5D_I64: return matmul_t<long long>(s, m, k, n, A, B, C);
    }
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int transpose_t(const Shard& s, int r, int c, const void* A, void* O) {
#define T5_TR(R_, C_) \
    if (r == R_ && c == C_) return launch_op<TransposeOp<T, R_, C_>>(s, A, nullptr, O, nullptr);
    T5_TR(2, 2) T5_TR(3, 3) T5_TR(4, 4) T5_TR(2, 3) T5_TR(3, 2) T5_TR(3, 4) T5_TR(4, 3) T5_TR(2, 4) T5_TR(4, 2)
#undef T5_TR
    return T5I_NOT_SPECIALISED;
}

int small_transpose(const Shard& s, int elem_bytes, int r, int c, const void* A, void* O) {
    if (elem_bytes == 8) return transpose_t<unsigned long long>(s, r, c, A, O);
    if (elem_bytes == 4) return transpose_t<unsigned int>(s, r, c, A, O);
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int det_t(const Shard& s, int n, const void* A, void* O) {
    if (n == 2) return launch_op<DetOp<T, 2>>(s, A, nullptr, O, nullptr);
    if (n == 3) return launch_op<DetOp<T, 3>>(s, A, nullptr, O, nullptr);
    if (n == 4) return launch_op<DetOp<T, 4>>(s, A, nullptr, O, nullptr);
    return T5I_NOT_SPECIALISED;
}
int small_det(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (dtype == T5D_F64) return det_t<double>(s, n, A, O);
    if (dtype == T5D_F32) return det_t<float>(s, n, A, O);
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int inverse_t(const Shard& s, int n, const void* A, void* O, void* status) {
    if (n == 2) return launch_op<InverseOp<T, 2>>(s, A, nullptr, O, status);
    if (n == 3) return launch_op<InverseOp<T, 3>>(s, A, nullptr, O, status);
    if (n == 4) return launch_op<InverseOp<T, 4>>(s, A, nullptr, O, status);
    return T5I_NOT_SPECIALISED;
}
int small_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status) {
    if (dtype == T5D_F64) return inverse_t<double>(s, n, A, O, status);
    if (dtype == T5D_F32) return inverse_t<float>(s, n, A, O, status);
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int solve_t(const Shard& s, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
#define T5_SV(N_, R_) \
    if (n == N_ && nrhs == R_) return launch_op<SolveOp<T, N_, R_>>(s, A, B, X, status);
    T5_SV(2, 1) T5_SV(3, 1) T5_SV(4, 1) T5_SV(2, 2) T5_SV(3, 3) T5_SV(4, 4)
#undef T5_SV
    return T5I_NOT_SPECIALISED;
}
int small_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (dtype == T5D_F64) return solve_t<double>(s, n, nrhs, A, B, X, status);
    if (dtype == T5D_F32) return solve_t<float>(s, n, nrhs, A, B, X, status);
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
