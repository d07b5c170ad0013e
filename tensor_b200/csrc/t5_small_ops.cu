// Thread-per-item kernels for the tiny fixed shapes (2x2 .. 4x4 matrices, 2..4
// vectors): the HBM-bound part of the hot path (SURVEY.md §8a a4, a6, a7, a9-a11,
// BASELINE configs 1-3).  Arithmetic is written once as register-level `Op::run`
// and instantiated in the AoS tile pipeline and the SoA skeleton (t5_tile.cuh).
//
// This translation unit MUST be compiled with -fmad=false: the specification is
// multiply-round-add-round in index order, which makes Double/Float results
// bit-identical to the CPU oracle, not merely within tolerance.
#include "t5_tile_launch.cuh"
#include <cstring>
#include <type_traits>

namespace t5 {

#define T5_MATVEC_GEO(N_) \
    template <class T> struct Geo<MatmulOp<T, N_, N_, 1>> { using type = G<64, 1, 2>; }; \
    template <class T> struct Geo<MatmulOp<T, 1, N_, N_>> { using type = G<64, 1, 2>; };
T5_MATVEC_GEO(5) T5_MATVEC_GEO(6) T5_MATVEC_GEO(7) T5_MATVEC_GEO(8)
#undef T5_MATVEC_GEO
template <class T> struct Geo<MatmulOp<T, 5, 5, 5>> { using type = G<128, 1, 2>; };
template <class T> struct Geo<MatmulOp<T, 6, 6, 6>> { using type = G<128, 1, 2>; };
template <class T> struct Geo<MatmulOp<T, 7, 7, 7>> { using type = G<128, 1, 2>; };
template <class T> struct Geo<DetOp<T, 3>> { using type = G<256, 1, 2>; };
template <class T> struct Geo<DetOp<T, 4>> { using type = G<128, 1, 2, 6>; };
template <class T> struct Geo<InverseOp<T, 2>> { using type = G<128, 1, 2>; };
template <class T> struct Geo<InverseOp<T, 3>> { using type = G<128, 1, 2, 6>; };   // register cap 80: 6 CTAs/SM
template <class T> struct Geo<InverseOp<T, 4>> { using type = G<64, 1, 2, 8>; };    // register cap 128: 8 CTAs/SM (r01_tune_elim_v3.csv)
// 5 <= n <= 8 elimination: one record per thread, 64-thread CTAs; the one-stage pipeline (t5_tile.cuh) wherever it
// measured faster (profiles/r01f_op_survey.txt: every 7x7 / 8x8 op, the 5x5 / 6x6 det and solve except det 6x6)
template <class T> struct Geo<DetOp<T, 5>> { using type = G<64, 1, 1>; };
template <class T> struct Geo<DetOp<T, 6>> { using type = G<64, 1, 2>; };
template <class T> struct Geo<InverseOp<T, 5>> { using type = G<64, 1, 2>; };
template <class T> struct Geo<InverseOp<T, 6>> { using type = G<64, 1, 2>; };
template <class T> struct Geo<SolveOp<T, 5, 1>> { using type = G<64, 1, 1>; };
template <class T> struct Geo<SolveOp<T, 6, 1>> { using type = G<64, 1, 1>; };
template <class T> struct Geo<DetOp<T, 7>> { using type = G<64, 1, 1>; };
template <class T> struct Geo<DetOp<T, 8>> { using type = G<64, 1, 1>; };
template <class T> struct Geo<InverseOp<T, 7>> { using type = G<64, 1, 1>; };
template <class T> struct Geo<InverseOp<T, 8>> { using type = std::conditional_t<(sizeof(T) == 8), G<64, 1, 1>, G<64, 1, 1, 6>>; };   // Float: a 168-register cap costs no spill and admits 6 CTAs (0.202 -> 0.165 ms)
template <class T> struct Geo<SolveOp<T, 7, 1>> { using type = G<64, 1, 1>; };
template <class T> struct Geo<SolveOp<T, 8, 1>> { using type = G<64, 1, 1>; };
template <class T, int N> struct Geo<MinPolyOp<T, N>> { using type = G<64, 1, 2>; };              // up to 248 registers (4x4 Double)
template <class T, int N, int R> struct Geo<SolveOp<T, N, R>> {
    using type = std::conditional_t<(R == 1), std::conditional_t<(N <= 3), G<256, 1, 3>, G<128, 1, 2, 5>>,
                                    std::conditional_t<(N <= 3), G<128, 1, 2, 6>, G<64, 1, 2, 8>>>;
};

template <class T, int P, int Q> struct Geo<OuterOp<T, P, Q>> {     // by the size of the result record
    using type = std::conditional_t<(P * Q * sizeof(T) >= 512), G<64, 1, 1>,
                 std::conditional_t<(P * Q * sizeof(T) >= 128), G<64, 1, 2>, G<128, 1, 2>>>;
};

template <class T>
static int matmul_t(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C) {
#define T5_MM(M_, K_, N_) \
    if (m == M_ && k == K_ && n == N_) return launch_op<MatmulOp<T, M_, K_, N_>>(s, A, B, C, nullptr);
    T5_MM(2, 2, 2) T5_MM(3, 3, 3) T5_MM(4, 4, 4)
    T5_MM(2, 2, 1) T5_MM(3, 3, 1) T5_MM(4, 4, 1)
    T5_MM(1, 2, 2) T5_MM(1, 3, 3) T5_MM(1, 4, 4)
    T5_MM(1, 2, 1) T5_MM(1, 3, 1) T5_MM(1, 4, 1)
    T5_MM(5, 5, 5)                                             // 75 elements in registers: fine for every type
    T5_MM(5, 5, 1) T5_MM(6, 6, 1) T5_MM(7, 7, 1) T5_MM(8, 8, 1)  // matrix . vector: the matrix record in registers (<= 64 elements)
    T5_MM(1, 5, 5) T5_MM(1, 6, 6) T5_MM(1, 7, 7) T5_MM(1, 8, 8)  // vector . matrix
    if constexpr (sizeof(T) == 4) { T5_MM(6, 6, 6) T5_MM(7, 7, 7) }   // 8-byte 6x6 is the row-per-thread kernel's
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

// Tiny outer products (rank-1 (x) rank-1 up to 4 (x) 4, and 3x3 (x) 3x3 / 4x4 (x) 4x4 seen as 9 (x) 9 / 16 (x) 16) in the tile
// pipeline: a record per thread, the P * Q results leave as whole 16-byte chunks of the tile (the coalesced-run kernel
// of t5_generic.cu writes one element per thread: 3x3 (x) 3x3 Float 3.2 -> 5.6 TB/s, 4x4 (x) 4x4 Float 2.9 -> 6.2,
// 4 (x) 4 Double 4.3 -> 5.6; tools/survey_ops.py).
template <class T>
static int outer_t(const Shard& s, int p, int q, const void* A, const void* B, void* C) {
#define T5_OUT(P_, Q_) \
    if (p == P_ && q == Q_) return launch_op<OuterOp<T, P_, Q_>>(s, A, B, C, nullptr);
    T5_OUT(2, 2) T5_OUT(3, 3) T5_OUT(4, 4) T5_OUT(2, 3) T5_OUT(3, 2) T5_OUT(2, 4) T5_OUT(4, 2) T5_OUT(3, 4) T5_OUT(4, 3)
    T5_OUT(9, 9) T5_OUT(4, 9) T5_OUT(9, 4) T5_OUT(3, 9) T5_OUT(9, 3)
    if constexpr (sizeof(T) == 4) { T5_OUT(16, 16) T5_OUT(4, 16) T5_OUT(16, 4) }     // (8-byte 16 (x) 16 = 2 KB of results per thread: the streaming kernel's)
#undef T5_OUT
    return T5I_NOT_SPECIALISED;
}
int small_outer(const Shard& s, int dtype, int p, int q, const void* A, const void* B, void* C) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;      // SoA: the vectorised SoA outer product of t5_generic.cu
    switch (dtype) {
        case T5D_F64: return outer_t<double>(s, p, q, A, B, C);
        case T5D_F32: return outer_t<float>(s, p, q, A, B, C);
        case T5D_I64: return outer_t<long long>(s, p, q, A, B, C);
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
    if (n == 5) return launch_op<DetOp<T, 5>>(s, A, nullptr, O, nullptr);
    if (n == 6) return launch_op<DetOp<T, 6>>(s, A, nullptr, O, nullptr);
    if (n == 7) return launch_op<DetOp<T, 7>>(s, A, nullptr, O, nullptr);
    if (n == 8) return launch_op<DetOp<T, 8>>(s, A, nullptr, O, nullptr);
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
    if (n == 5) return launch_op<InverseOp<T, 5>>(s, A, nullptr, O, status);
    if (n == 6) return launch_op<InverseOp<T, 6>>(s, A, nullptr, O, status);
    if (n == 7) return launch_op<InverseOp<T, 7>>(s, A, nullptr, O, status);
    if (n == 8) return launch_op<InverseOp<T, 8>>(s, A, nullptr, O, status);
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
    T5_SV(2, 1) T5_SV(3, 1) T5_SV(4, 1) T5_SV(2, 2) T5_SV(3, 3) T5_SV(4, 4) T5_SV(5, 1) T5_SV(6, 1) T5_SV(7, 1) T5_SV(8, 1)
#undef T5_SV
    return T5I_NOT_SPECIALISED;
}
int small_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (dtype == T5D_F64) return solve_t<double>(s, n, nrhs, A, B, X, status);
    if (dtype == T5D_F32) return solve_t<float>(s, n, nrhs, A, B, X, status);
    return T5I_NOT_SPECIALISED;
}

// ---- tr / polyEval / charPoly ------------------------------------------------
template <class T>
static int trace_t(const Shard& s, int n, const void* A, void* O) {
    if (n == 2) return launch_op<TraceOp<T, 2>>(s, A, nullptr, O, nullptr);
    if (n == 3) return launch_op<TraceOp<T, 3>>(s, A, nullptr, O, nullptr);
    if (n == 4) return launch_op<TraceOp<T, 4>>(s, A, nullptr, O, nullptr);
    return T5I_NOT_SPECIALISED;
}
int small_trace(const Shard& s, int dtype, int n, const void* A, void* O) {
    switch (dtype) {
        case T5D_F64: return trace_t<double>(s, n, A, O);
        case T5D_F32: return trace_t<float>(s, n, A, O);
        case T5D_I64: return trace_t<long long>(s, n, A, O);
    }
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int poly_eval_t(const Shard& s, int n, const OpParams& prm, const void* A, void* O) {
    if (n == 2) return launch_op<PolyEvalOp<T, 2>>(s, A, nullptr, O, nullptr, prm);
    if (n == 3) return launch_op<PolyEvalOp<T, 3>>(s, A, nullptr, O, nullptr, prm);
    if (n == 4) return launch_op<PolyEvalOp<T, 4>>(s, A, nullptr, O, nullptr, prm);
    return T5I_NOT_SPECIALISED;
}
// coeffs_host: ncoef elements of dtype, low order first; at most 16 ride in the parameter block
int small_poly_eval(const Shard& s, int dtype, int n, const void* coeffs_host, int ncoef, const void* A, void* O) {
    if (ncoef < 0 || ncoef > 16) return T5I_NOT_SPECIALISED;
    OpParams prm{};
    prm.n = ncoef;
    const int es = dtype_size(dtype);
    for (int j = 0; j < ncoef; ++j) memcpy(&prm.v[j], (const char*)coeffs_host + (size_t)j * es, es);
    switch (dtype) {
        case T5D_F64: return poly_eval_t<double>(s, n, prm, A, O);
        case T5D_F32: return poly_eval_t<float>(s, n, prm, A, O);
        case T5D_I64: return poly_eval_t<long long>(s, n, prm, A, O);
    }
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int char_poly_t(const Shard& s, int n, const void* A, void* O) {
    if (n == 2) return launch_op<CharPolyOp<T, 2>>(s, A, nullptr, O, nullptr);
    if (n == 3) return launch_op<CharPolyOp<T, 3>>(s, A, nullptr, O, nullptr);
    if (n == 4) return launch_op<CharPolyOp<T, 4>>(s, A, nullptr, O, nullptr);
    return T5I_NOT_SPECIALISED;
}
int small_char_poly(const Shard& s, int dtype, int n, const void* A, void* O) {
    if (dtype == T5D_F64) return char_poly_t<double>(s, n, A, O);
    if (dtype == T5D_F32) return char_poly_t<float>(s, n, A, O);
    return T5I_NOT_SPECIALISED;
}

template <class T>
static int min_poly_t(const Shard& s, int n, const void* A, void* O, void* degree) {
    if (n == 2) return launch_op<MinPolyOp<T, 2>>(s, A, nullptr, O, degree);
    if (n == 3) return launch_op<MinPolyOp<T, 3>>(s, A, nullptr, O, degree);
    if (n == 4) return launch_op<MinPolyOp<T, 4>>(s, A, nullptr, O, degree);
    return T5I_NOT_SPECIALISED;
}
int small_min_poly(const Shard& s, int dtype, int n, const void* A, void* O, void* degree) {
    if (dtype == T5D_F64) return min_poly_t<double>(s, n, A, O, degree);
    if (dtype == T5D_F32) return min_poly_t<float>(s, n, A, O, degree);
    return T5I_NOT_SPECIALISED;
}

}  // namespace t5
