// rowEchelonForm / triangularSolve, record per thread in the tile
// pipeline (specification: oracle/t5_oracle.cpp row_echelon_item; other shapes:
// row_echelon_kernel in t5_generic.cu).  Compiled with -fmad=false like every bit-exact TU.
//
// The pivot ROW of column c is a run-time quantity (a column without a pivot does not advance
// the rank), and a run-time row index would push the working matrix into local memory.  So the
// elimination step is instantiated for every (column c, rank r <= c) pair with static indices and
// the run-time rank only SELECTS the instance: `if (rank == r) step<c, r>()`.  A warp whose items
// all have full-rank leading columns (the common case) runs one uniform instance per column; a
// batch of rank-deficient items runs a different but equally uniform chain.  Nothing is indexed
// dynamically, the matrix stays in registers, and the operations of each element are the oracle's
// in the oracle's order (the divisions by a pivot share one exact reciprocal, t5_small_ops.cuh).
#pragma once
#include "t5_tile_launch.cuh"

namespace t5 {

template <class T, int M, int N>
struct RowEchelonOp {
    static constexpr bool uses_params = true;     // prm.n = pivot_cols
    using In0 = Rec<T, M * N>; using In1 = NoRec;
    using Out0 = Rec<T, M * N>; using Out1 = Rec<unsigned char, 1>;

    template <int C, int R>
    __device__ __forceinline__ static void step(T (&w)[M][N], int& rank) {
        if (R + 1 < M && w[R][C] == T(0)) {
            // first row below with a non-zero entry moves up (whole rows: the zeros left of C may differ in sign)
            static_for<R + 1, M>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const bool sw = (w[R][C] == T(0)) && (w[r][C] != T(0));
                #pragma unroll
                for (int j = 0; j < N; ++j) {
                    const T x = w[R][j], y = w[r][j];
                    w[R][j] = sw ? y : x;
                    w[r][j] = sw ? x : y;
                }
            });
        }
        if (w[R][C] != T(0)) {
            const Recip<T> rc(w[R][C]);
            if constexpr (kBatchDiv<T>) {
                // Double: the M - 1 multipliers, then the N - C - 1 entries of the pivot row, each as one batch of independent
                // exact divisions with a single fallback test (t5_recip.cuh)
                T f[M];
                bool all_fast = true;
                #pragma unroll
                for (int i = 0; i < M; ++i)
                    if (i != R) f[i] = rc.div_try(w[i][C], all_fast);
                if (!all_fast) {
                    #pragma unroll
                    for (int i = 0; i < M; ++i)
                        if (i != R) f[i] = rc.div_fix(w[i][C], f[i]);
                }
                static_for<0, M>([&](auto i_) {
                    constexpr int i = decltype(i_)::value;
                    if constexpr (i != R) {
                        #pragma unroll
                        for (int j = C + 1; j < N; ++j) w[i][j] = w[i][j] - f[i] * w[R][j];
                        w[i][C] = T(0);
                    }
                });
                T q[N];
                all_fast = true;
                #pragma unroll
                for (int j = C + 1; j < N; ++j) q[j] = rc.div_try(w[R][j], all_fast);
                if (!all_fast) {
                    #pragma unroll
                    for (int j = C + 1; j < N; ++j) q[j] = rc.div_fix(w[R][j], q[j]);
                }
                #pragma unroll
                for (int j = C + 1; j < N; ++j) w[R][j] = q[j];
            } else {
                static_for<0, M>([&](auto i_) {
                    constexpr int i = decltype(i_)::value;
                    if constexpr (i != R) {
                        const T f = rc.div(w[i][C]);
                        #pragma unroll
                        for (int j = C + 1; j < N; ++j) w[i][j] = w[i][j] - f * w[R][j];
                        w[i][C] = T(0);
                    }
                });
                #pragma unroll
                for (int j = C + 1; j < N; ++j) w[R][j] = rc.div(w[R][j]);
            }
            w[R][C] = T(1);
            rank = R + 1;
        }
    }

    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* o, unsigned char* rank_out, const OpParams& prm) {
        T w[M][N];
        #pragma unroll
        for (int i = 0; i < M; ++i)
            #pragma unroll
            for (int j = 0; j < N; ++j) w[i][j] = a[i * N + j];
        int rank = 0;
        const int pivot_cols = prm.n;
        static_for<0, N>([&](auto c_) {
            constexpr int c = decltype(c_)::value;
            if (c < pivot_cols) {
                const int at = rank;              // the rank BEFORE this column: one instance at most
                static_for<0, (c < M - 1 ? c : M - 1) + 1>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    if (at == r) step<c, r>(w, rank);
                });
            }
        });
        #pragma unroll
        for (int i = 0; i < M; ++i)
            #pragma unroll
            for (int j = 0; j < N; ++j) o[i * N + j] = w[i][j];
        rank_out[0] = (unsigned char)rank;
    }
};

template <class T, int M, int N> struct Geo<RowEchelonOp<T, M, N>> {
    // Float records are half the bytes for the same arithmetic: larger tiles, or the tile scheduler's atomic
    // becomes the limit (4Mi 4x5 Float with 64-record tiles: 0.52 of the HBM peak, the Double twin 0.99)
    using type = std::conditional_t<(M * N * sizeof(T) <= 72), G<128, 1, 2, 6>,
                 std::conditional_t<(M >= 5), std::conditional_t<(M * N * sizeof(T) >= 288), G<64, 1, 1>, G<64, 1, 2>>,
                 std::conditional_t<(sizeof(T) == 4 && M * N <= 20), G<128, 2, 2, 3>,
                 std::conditional_t<(sizeof(T) == 4), G<128, 1, 2, 4>, G<64, 1, 2>>>>>;
};

}  // namespace t5
