// Register-level arithmetic of the tiny fixed shapes.
// Each Op is a pure function on register arrays; t5_tile.cuh supplies the data
// movement (AoS tile pipeline or SoA skeleton).  Every TU that includes this must be
// compiled with -fmad=false (multiply-round-add-round, bit-identical to the oracle).
#pragma once
#include "t5_tile.cuh"
#include "t5_recip.cuh"

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

// ---- outer product (prod 0 / tensor product of tiny tensors): C[p,q] = A[p] * B[q], ONE multiply per output
// (no `0 +`: the sign of a zero product is the multiply's; §8a a8) ------------------------
template <class T, int P, int Q>
struct OuterOp {
    using In0 = Rec<T, P>; using In1 = Rec<T, Q>;
    using Out0 = Rec<T, P * Q>; using Out1 = NoRec;
    __device__ __forceinline__ static void run(const T* a, const T* b, T* c, unsigned char*) {
        using R = Arith<T>;
        #pragma unroll
        for (int i = 0; i < P; ++i)
            #pragma unroll
            for (int j = 0; j < Q; ++j) c[i * Q + j] = R::mul(a[i], b[j]);
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

template <int K> using k_of = std::integral_constant<int, K>;


template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---- Gaussian elimination, first-non-zero pivot (§8a pivot rule) --------------------
// w is N x W in registers; every index is a compile-time constant after unrolling.  The
// pivot search only runs when the diagonal entry is an exact zero (rare): inside, a chain
// of predicated row swaps picks the first row below with a non-zero entry.
template <class T, int N, int W>
__device__ __forceinline__ bool eliminate(T (&w)[N][W], int& swaps, Recip<T> (&rc)[N]) {
    bool singular = false;
    swaps = 0;
    // static_for (template recursion), not `#pragma unroll`: for n >= 5 the unroller gives up on
    // the pivot chain and on the row loop (the inlined division is large), and one runtime row
    // index is enough to push the whole working matrix into local memory.
    static_for<0, N>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        if (k + 1 < N && w[k][k] == T(0)) {
            static_for<k + 1, N>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const bool sw = (w[k][k] == T(0)) && (w[r][k] != T(0));
                // columns < k of rows >= k are exact zeros already: swapping them is a no-op
                #pragma unroll
                for (int j = k; j < W; ++j) {
                    T x = w[k][j], y = w[r][j];
                    w[k][j] = sw ? y : x;
                    w[r][j] = sw ? x : y;
                }
                swaps += sw ? 1 : 0;
            });
        }
        singular = singular || (w[k][k] == T(0));
        rc[k] = Recip<T>(w[k][k]);
        if constexpr (kBatchDiv<T>) {
            // the multipliers of the step as one batch of independent exact divisions (one fallback test for all of them)
            T f[N];
            bool all_fast = true;
            #pragma unroll
            for (int i = k + 1; i < N; ++i) f[i] = rc[k].div_try(w[i][k], all_fast);
            if (!all_fast) {
                #pragma unroll
                for (int i = k + 1; i < N; ++i) f[i] = rc[k].div_fix(w[i][k], f[i]);
            }
            static_for<k + 1, N>([&](auto i_) {
                constexpr int i = decltype(i_)::value;
                #pragma unroll
                for (int j = k + 1; j < W; ++j) w[i][j] = w[i][j] - f[i] * w[k][j];
                w[i][k] = T(0);
            });
        } else {
            static_for<k + 1, N>([&](auto i_) {
                constexpr int i = decltype(i_)::value;
                const T f = rc[k].div(w[i][k]);
                #pragma unroll
                for (int j = k + 1; j < W; ++j) w[i][j] = w[i][j] - f * w[k][j];
                w[i][k] = T(0);
            });
        }
    });
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
        Recip<T> rc[N];
        const bool ok = eliminate<T, N, N>(w, swaps, rc);
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
    Recip<T> rc[N];
    const bool ok = eliminate<T, N, N + NRHS>(w, swaps, rc);
    T xs[N][NRHS];
    static_for<0, N>([&](auto kk_) {
        constexpr int k = N - 1 - decltype(kk_)::value;
        static_for<0, NRHS>([&](auto c_) {
            constexpr int c = decltype(c_)::value;
            T s = T(0);
            #pragma unroll
            for (int j = k + 1; j < N; ++j) s = s + w[k][j] * xs[j][c];
            xs[k][c] = rc[k].div(w[k][N + c] - s);
        });
    });
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

// ---- the rest of the SquareMatrix surface: tr, polyEval, charPoly ----------
// tr: fold_{i asc} (acc + A[i,i]) from 0.
template <class T, int N>
struct TraceOp {
    using In0 = Rec<T, N * N>; using In1 = NoRec;
    using Out0 = Rec<T, 1>; using Out1 = NoRec;
    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* o, unsigned char*) {
        using R = Arith<T>;
        T acc = R::zero();
        #pragma unroll
        for (int i = 0; i < N; ++i) acc = R::add(acc, a[i * N + i]);
        o[0] = acc;
    }
};

// polyEval A [c0..cd] by Horner: R = cd I; for j = d-1..0: R = (R .*. A) + cj I.
template <class T, int N>
struct PolyEvalOp {
    static constexpr bool uses_params = true;
    using In0 = Rec<T, N * N>; using In1 = NoRec;
    using Out0 = Rec<T, N * N>; using Out1 = NoRec;
    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* o, unsigned char*, const OpParams& prm) {
        using R = Arith<T>;
        T r[N][N];
        const T lead = prm.n > 0 ? param_as<T>(prm.v[prm.n - 1]) : R::zero();
        #pragma unroll
        for (int i = 0; i < N; ++i)
            #pragma unroll
            for (int j = 0; j < N; ++j) r[i][j] = (i == j) ? lead : R::zero();
        for (int c = prm.n - 2; c >= 0; --c) {
            const T cj = param_as<T>(prm.v[c]);
            T t[N][N];
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int j = 0; j < N; ++j) {
                    T acc = R::zero();
                    #pragma unroll
                    for (int k = 0; k < N; ++k) acc = R::add(acc, R::mul(r[i][k], a[k * N + j]));
                    t[i][j] = (i == j) ? R::add(acc, cj) : acc;
                }
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int j = 0; j < N; ++j) r[i][j] = t[i][j];
        }
        #pragma unroll
        for (int i = 0; i < N; ++i)
            #pragma unroll
            for (int j = 0; j < N; ++j) o[i * N + j] = r[i][j];
    }
};

// charPoly by Faddeev-LeVerrier: M = I; for k = 1..N: AM = A .*. M; c[N-k] = -(tr AM / k);
// M = AM + c[N-k] I.  Output c[0..N], c[N] = 1.
template <class T, int N>
struct CharPolyOp {
    using In0 = Rec<T, N * N>; using In1 = NoRec;
    using Out0 = Rec<T, N + 1>; using Out1 = NoRec;
    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* c, unsigned char*) {
        T m[N][N];
        #pragma unroll
        for (int i = 0; i < N; ++i)
            #pragma unroll
            for (int j = 0; j < N; ++j) m[i][j] = (i == j) ? T(1) : T(0);
        c[N] = T(1);
        #pragma unroll
        for (int k = 1; k <= N; ++k) {
            T am[N][N];
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int j = 0; j < N; ++j) {
                    T acc = T(0);
                    #pragma unroll
                    for (int l = 0; l < N; ++l) acc = acc + a[i * N + l] * m[l][j];
                    am[i][j] = acc;
                }
            T tr = T(0);
            #pragma unroll
            for (int i = 0; i < N; ++i) tr = tr + am[i][i];
            const T ck = -(tr / T(k));
            c[N - k] = ck;
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int j = 0; j < N; ++j) m[i][j] = (i == j) ? am[i][j] + ck : am[i][j];
        }
    }
};

// minPoly: first exact linear dependency among vec(I), vec(A), vec(A^2), ..; the
// oracle's Krylov elimination (oracle/t5_oracle.cpp min_poly_item) with every loop unrolled at
// compile time.  The only run-time indices are the pivot positions (first non-zero entry of a
// reduced candidate): they are read and cleared through select chains, so the basis, the
// combinations and the candidate all stay in registers (the any-n kernel keeps them in local
// memory: 4Mi 4x4 Double 5.7 ms).  Operation order per element is the oracle's.
template <class T, int LEN>
__device__ __forceinline__ T pick(const T (&v)[LEN], int p) {
    T r = v[0];
    #pragma unroll
    for (int t = 1; t < LEN; ++t) r = (t == p) ? v[t] : r;
    return r;
}

template <class T, int N>
struct MinPolyOp {
    static constexpr int L = N * N;
    using In0 = Rec<T, L>; using In1 = NoRec;
    using Out0 = Rec<T, N + 1>; using Out1 = Rec<unsigned char, 1>;
    __device__ __forceinline__ static void run(const T* a, const unsigned char*, T* c, unsigned char* degree) {
        T basis[N][L], coef[N][N + 1], P[L];
        int pivot[N] = {};
        int deg = -1;
        static_for<0, N>([&](auto m_) {
            constexpr int m = decltype(m_)::value;
            if (deg < 0) {   // every earlier candidate joined the basis: it has exactly m members
                if constexpr (m == 0) {
                    #pragma unroll
                    for (int t = 0; t < L; ++t) P[t] = (t / N == t % N) ? T(1) : T(0);
                } else {
                    T pn[L];
                    #pragma unroll
                    for (int i = 0; i < N; ++i)
                        #pragma unroll
                        for (int j = 0; j < N; ++j) {
                            T acc = T(0);
                            #pragma unroll
                            for (int k = 0; k < N; ++k) acc = acc + P[i * N + k] * a[k * N + j];
                            pn[i * N + j] = acc;
                        }
                    #pragma unroll
                    for (int t = 0; t < L; ++t) P[t] = pn[t];
                }
                T v[L], g[N + 1];
                #pragma unroll
                for (int t = 0; t < L; ++t) v[t] = P[t];
                #pragma unroll
                for (int j = 0; j <= N; ++j) g[j] = (j == m) ? T(1) : T(0);
                static_for<0, m>([&](auto i_) {
                    constexpr int i = decltype(i_)::value;
                    const int p = pivot[i];
                    const T f = pick(v, p) / pick(basis[i], p);
                    #pragma unroll
                    for (int t = 0; t < L; ++t) {
                        const T d = v[t] - f * basis[i][t];
                        v[t] = (t == p) ? T(0) : d;
                    }
                    #pragma unroll
                    for (int j = 0; j <= N; ++j) g[j] = g[j] - f * coef[i][j];
                });
                int p = L;
                #pragma unroll
                for (int t = L - 1; t >= 0; --t) p = (v[t] != T(0)) ? t : p;
                if (p == L) {
                    #pragma unroll
                    for (int j = 0; j <= N; ++j) c[j] = (j <= m) ? g[j] : T(0);
                    deg = m;
                } else {
                    #pragma unroll
                    for (int t = 0; t < L; ++t) basis[m][t] = v[t];
                    #pragma unroll
                    for (int j = 0; j <= N; ++j) coef[m][j] = g[j];
                    pivot[m] = p;
                }
            }
        });
        if (deg < 0) {   // I .. A^(N-1) independent: the characteristic polynomial
            CharPolyOp<T, N>::run(a, nullptr, c, nullptr);
            deg = N;
        }
        degree[0] = (unsigned char)deg;
    }
};

// inverse(A) = solve(A, I) for n <= 4: the whole [A | I] working matrix in registers.
template <class T, int N, class = void>
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

// n >= 5: [A | I] is N x 2N (128 doubles for 8x8) and the result another N x N, more than a thread
// has registers for.  The columns of the right-hand side never interact, so the elimination is split
// in time instead: (1) eliminate A alone, KEEPING each multiplier f(i,k) in the entry it zeroes (row
// swaps move whole rows, multipliers included) and the pivot row chosen in column k; (2) replay those
// row operations on a block of identity columns at a time: a row carries its right-hand side through every
// swap, so the columns start as the PERMUTED identity (integer bookkeeping) and then see
// x[i] -= f(i,k) x[k] for i > k, k ascending, with no swap left in the arithmetic; back-substitute, and
// hand the finished entries to the writer before the next block starts.  Every element sees the
// operations of the one-pass elimination in the same order (bit-identical to the oracle);
// the live set is N x N + 2N + the reciprocals (8x8 Double: ~100 doubles instead of 192+).
// (1) of the split elimination: LU of A with the pivot rule and multipliers of `eliminate`, the
// multiplier f(i,k) kept in w[i][k]; piv[k] = the row that was swapped into position k.
template <class T, int N>
__device__ __forceinline__ bool lu_factor(T (&w)[N][N], int (&piv)[N], Recip<T> (&rc)[N]) {
    bool singular = false;
    static_for<0, N>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        piv[k] = k;
        if (k + 1 < N && w[k][k] == T(0)) {
            static_for<k + 1, N>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const bool sw = (w[k][k] == T(0)) && (w[r][k] != T(0));
                #pragma unroll
                for (int j = 0; j < N; ++j) {      // whole rows: the multipliers of earlier steps travel with their row
                    T x = w[k][j], y = w[r][j];
                    w[k][j] = sw ? y : x;
                    w[r][j] = sw ? x : y;
                }
                piv[k] = sw ? r : piv[k];
            });
        }
        singular = singular || (w[k][k] == T(0));
        rc[k] = Recip<T>(w[k][k]);
        if constexpr (kBatchDiv<T>) {
            // the multipliers of the step as one batch of independent exact divisions (one fallback test for all of them)
            T f[N];
            bool all_fast = true;
            #pragma unroll
            for (int i = k + 1; i < N; ++i) f[i] = rc[k].div_try(w[i][k], all_fast);
            if (!all_fast) {
                #pragma unroll
                for (int i = k + 1; i < N; ++i) f[i] = rc[k].div_fix(w[i][k], f[i]);
            }
            static_for<k + 1, N>([&](auto i_) {
                constexpr int i = decltype(i_)::value;
                #pragma unroll
                for (int j = k + 1; j < N; ++j) w[i][j] = w[i][j] - f[i] * w[k][j];
                w[i][k] = f[i];
            });
        } else {
            static_for<k + 1, N>([&](auto i_) {
                constexpr int i = decltype(i_)::value;
                const T f = rc[k].div(w[i][k]);
                #pragma unroll
                for (int j = k + 1; j < N; ++j) w[i][j] = w[i][j] - f * w[k][j];
                w[i][k] = f;
            });
        }
    });
    return !singular;
}

// (2): the row operations replayed on CB right-hand-side columns that are ALREADY row-permuted
// (x = P b), then the back-substitution; emit(k, c, value) receives each finished entry.
template <class T, int N, int CB, class Emit>
__device__ __forceinline__ void lu_solve_block(const T (&w)[N][N], const Recip<T> (&rc)[N], T (&x)[N][CB], Emit&& emit) {
    static_for<0, N>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        static_for<k + 1, N>([&](auto i_) {
            constexpr int i = decltype(i_)::value;
            #pragma unroll
            for (int c = 0; c < CB; ++c) x[i][c] = x[i][c] - w[i][k] * x[k][c];
        });
    });
    static_for<0, N>([&](auto kk_) {
        constexpr int k = N - 1 - decltype(kk_)::value;
        if constexpr (kBatchDiv<T>) {
            T num[CB];
            bool all_fast = true;
            static_for<0, CB>([&](auto c_) {
                constexpr int c = decltype(c_)::value;
                T s = T(0);
                #pragma unroll
                for (int j = k + 1; j < N; ++j) s = s + w[k][j] * x[j][c];
                num[c] = x[k][c] - s;
                x[k][c] = rc[k].div_try(num[c], all_fast);
            });
            if (!all_fast) {
                #pragma unroll
                for (int c = 0; c < CB; ++c) x[k][c] = rc[k].div_fix(num[c], x[k][c]);
            }
            static_for<0, CB>([&](auto c_) { emit(k_of<k>{}, c_, x[k][decltype(c_)::value]); });
        } else {
            static_for<0, CB>([&](auto c_) {
                constexpr int c = decltype(c_)::value;
                T s = T(0);
                #pragma unroll
                for (int j = k + 1; j < N; ++j) s = s + w[k][j] * x[j][c];
                x[k][c] = rc[k].div(x[k][c] - s);
                emit(k_of<k>{}, c_, x[k][c]);
            });
        }
    });
}

// (3) for `inverse`: the factors leave the registers.  After (1) the thread writes L and U back over its own
// input record in shared memory (`scratch_in0`) and the registers are free for ALL N columns of the permuted
// identity at once; the replay then reads each factor entry once (one LDS per N multiply-subtracts) and runs N
// independent recurrences instead of two.  The earlier two-columns-at-a-time version kept the factors in
// registers (255 of them, two warps per scheduler) and stalled on the FP64 latency of the back-substitution
// chains, whose order the specification fixes: ncu showed 49 % of the stall samples as fixed-latency `wait` and
// the FP64 pipe 26 % busy.  Finished row k of the inverse overwrites row k of the factors, which nothing reads
// any more (L row k was consumed by the forward sweep, U row k by back-substitution step k), so in the AoS
// pipeline the result still leaves through the thread's input slot.
// (Tried and dropped for 8x8 Double in the AoS pipeline: streaming the result rows straight to global memory so that
// nothing aliases the factors, two passes of four columns and a 168-register cap for a third warp per scheduler -
// 0.340 -> 0.526 ms with the cap (256 B of spills), 0.457 ms without: a thread's 32-byte row pieces 512 B apart cost
// more than the in-place slot + coalesced tile store saves.  Float fits the cap without spilling: 0.202 -> 0.165 ms.)
template <class T, int N>
struct InverseOp<T, N, std::enable_if_t<(N >= 5)>> {
    static constexpr bool streams_out0 = true;
    static constexpr bool scratch_in0 = true;
    using In0 = Rec<T, N * N>; using In1 = NoRec;
    using Out0 = Rec<T, N * N>; using Out1 = Rec<unsigned char, 1>;
    template <class W, class S>
    __device__ __forceinline__ static void run_stream(const T* a, const unsigned char*, const W& out, unsigned char* status, const S& scr) {
        int piv[N];
        Recip<T> rc[N];
        bool ok;
        {
            T w[N][N];
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int j = 0; j < N; ++j) w[i][j] = a[i * N + j];
            ok = lu_factor<T, N>(w, piv, rc);
            static_for<0, N>([&](auto i_) {
                static_for<0, N>([&](auto j_) {
                    constexpr int i = decltype(i_)::value, j = decltype(j_)::value;
                    if (i != j) scr.template put<i * N + j>(w[i][j]);       // (the diagonal lives on in rc[])
                });
            });
        }
        // All N columns of the permuted identity at once: in the in-place AoS pipeline a finished row may only overwrite
        // factor rows nothing reads any more, which rules out a second pass over the factors.
        constexpr int CB = N;
        static_for<0, (N + CB - 1) / CB>([&](auto blk_) {
            constexpr int c0 = decltype(blk_)::value * CB;
            constexpr int cb = (N - c0 < CB) ? N - c0 : CB;
            // a row keeps its right-hand side through every swap, so column c of the identity starts as e_pos with the 1
            // wherever original row c ended up
            T x[N][cb];
            #pragma unroll
            for (int c = 0; c < cb; ++c) {
                int pos = c0 + c;
                #pragma unroll
                for (int k = 0; k < N; ++k) pos = (pos == k) ? piv[k] : ((pos == piv[k]) ? k : pos);
                #pragma unroll
                for (int i = 0; i < N; ++i) x[i][c] = (i == pos) ? T(1) : T(0);
            }
            static_for<0, N>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                static_for<k + 1, N>([&](auto i_) {
                    constexpr int i = decltype(i_)::value;
                    const T l = scr.template get<i * N + k>();
                    #pragma unroll
                    for (int c = 0; c < cb; ++c) x[i][c] = x[i][c] - l * x[k][c];
                });
            });
            static_for<0, N>([&](auto kk_) {
                constexpr int k = N - 1 - decltype(kk_)::value;
                T s[cb];
                #pragma unroll
                for (int c = 0; c < cb; ++c) s[c] = T(0);
                static_for<k + 1, N>([&](auto j_) {
                    constexpr int j = decltype(j_)::value;
                    const T u = scr.template get<k * N + j>();
                    #pragma unroll
                    for (int c = 0; c < cb; ++c) s[c] = s[c] + u * x[j][c];
                });
                // the cb quotients of this row share their divisor and are independent: one batch without a branch between
                // them (t5_recip.cuh), so their chains overlap instead of each waiting behind its own fallback test
                if constexpr (kBatchDiv<T>) {
                    T num[cb];
                    bool all_fast = true;
                    #pragma unroll
                    for (int c = 0; c < cb; ++c) { num[c] = x[k][c] - s[c]; x[k][c] = rc[k].div_try(num[c], all_fast); }
                    if (!all_fast) {
                        #pragma unroll
                        for (int c = 0; c < cb; ++c) x[k][c] = rc[k].div_fix(num[c], x[k][c]);
                    }
                } else {
                    #pragma unroll
                    for (int c = 0; c < cb; ++c) x[k][c] = rc[k].div(x[k][c] - s[c]);
                }
                static_for<0, cb>([&](auto c_) {
                    constexpr int c = decltype(c_)::value;
                    out.template put<k * N + c0 + c>(ok ? x[k][c] : T(0));
                });
            });
        });
        status[0] = ok ? 0 : 1;
    }
};

// solve A X = B with a MATRIX right-hand side for n >= 5: the same split.  B is not loaded into
// registers: each block of two columns is read from the operand's shared-memory slot when its turn
// comes (`lazy_in1`; the row permutation is a run-time index there, which shared memory does not mind),
// and X is written over B's slot column block by column block (over A's when they are the same size).
template <class T, int N, int NRHS>
struct SolveManyOp {
    static constexpr bool streams_out0 = true;
    static constexpr bool lazy_in1 = true;
    static constexpr int CB = 2;
    using In0 = Rec<T, N * N>; using In1 = Rec<T, N * NRHS>;
    using Out0 = Rec<T, N * NRHS>; using Out1 = Rec<unsigned char, 1>;
    template <class Rd, class W>
    __device__ __forceinline__ static void run_stream(const T* a, const Rd& b, const W& out, unsigned char* status) {
        T w[N][N];
        #pragma unroll
        for (int i = 0; i < N; ++i)
            #pragma unroll
            for (int j = 0; j < N; ++j) w[i][j] = a[i * N + j];
        int piv[N];
        Recip<T> rc[N];
        const bool ok = lu_factor<T, N>(w, piv, rc);
        int perm[N];                      // position i holds original row perm[i]
        #pragma unroll
        for (int i = 0; i < N; ++i) perm[i] = i;
        static_for<0, N>([&](auto k_) {
            constexpr int k = decltype(k_)::value;
            static_for<k + 1, N>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const bool sw = piv[k] == r;
                const int p = perm[k], q = perm[r];
                perm[k] = sw ? q : p;
                perm[r] = sw ? p : q;
            });
        });
        static_for<0, (NRHS + CB - 1) / CB>([&](auto blk_) {
            constexpr int c0 = decltype(blk_)::value * CB;
            constexpr int cb = (NRHS - c0 < CB) ? NRHS - c0 : CB;
            T x[N][cb];
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int c = 0; c < cb; ++c) x[i][c] = b.get(perm[i] * NRHS + c0 + c);
            lu_solve_block<T, N, cb>(w, rc, x, [&](auto k_, auto c_, T v) {
                out.template put<decltype(k_)::value * NRHS + c0 + decltype(c_)::value>(ok ? v : T(0));
            });
        });
        status[0] = ok ? 0 : 1;
    }
};

}  // namespace t5
