// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the reference's (tensor5/tensor) storage / index
// conventions and of the batched linear-algebra hot path that SURVEY.md §8a
// specifies.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg may load this library; the product (tensor_b200/) never
// does and fails loudly when its CUDA library is missing.
//
// PARITY STATUS
//   * transpose / linearize / unlinearize follow reference code that exists
//     (src/Data/Tensor/Vector.hs:150-168, :295-301, :306-307,
//      src/Data/Indexable.hs:101-102) and are pinned in tests/ against the
//     hand-derived known answers of SURVEY.md §3a and against an independent
//     restatement of the GADT backend's unRev/unConcat
//     (src/Data/Tensor.hs:229-255; the reference's own tests
//      tests/Tensor.hs:62-79 tie unRev to unConcat).
//   * matmul (.*.), dot, prod n, det, inverse, solve: the reference checkout
//     has NO implementation (CHANGELOG.md:24-29 is the only trace; SURVEY.md
//     §0 F1).  Their arithmetic is the specification of SURVEY.md §8a.
//     **PARITY UNPINNED** for these: the oracle is the specification.
//   * No Haskell toolchain exists in this image, so oracle/_ref cannot be built.
//
// Build:  g++ -O2 -ffp-contract=off -fno-fast-math  (see oracle/build.sh)
// -ffp-contract=off is REQUIRED: the specification is "multiply, round, add,
// round" (GHC primops on Double/Float never fuse).
//
// Element types: 0 = Double (f64), 1 = Float (f32), 2 = Int (i64, two's
// complement wrap-around like GHC's Int).

#include <cstdint>
#include <cstring>
#include <cstddef>
#include <vector>
#include <thread>
#include <algorithm>

namespace {

// ---------------------------------------------------------------------------
// Index conventions.  Vector.hs:150-158 `linearize ds is`: 1-based multi-index
// -> 0-based row-major offset; Vector.hs:160-168 `unlinearize ds i`: inverse.
// ---------------------------------------------------------------------------
static inline uint64_t product_tail(const uint64_t* ds, int r, int from) {
    uint64_t p = 1;
    for (int a = from; a < r; ++a) p *= ds[a];
    return p;
}

// go es js acc = acc + (head js - 1) * product (tail es)        (Vector.hs:158)
static uint64_t linearize(const uint64_t* ds, const uint64_t* is, int r) {
    uint64_t acc = 0;
    for (int a = 0; a < r; ++a) acc += (is[a] - 1) * product_tail(ds, r, a + 1);
    return acc;
}

// (q,r) = quotRem j (product t); emit q+1                       (Vector.hs:167-168)
static void unlinearize(const uint64_t* ds, int r, uint64_t i, uint64_t* out) {
    uint64_t j = i;
    for (int a = 0; a < r; ++a) {
        uint64_t p = product_tail(ds, r, a + 1);
        out[a] = j / p + 1;
        j = j % p;
    }
}

// transposeV ds = linearize ds . reverse . unlinearize (reverse ds)  (Vector.hs:306-307)
static uint64_t transposeV(const uint64_t* ds, int r, uint64_t i) {
    uint64_t rd[16], ix[16], rix[16];
    for (int a = 0; a < r; ++a) rd[a] = ds[r - 1 - a];
    unlinearize(rd, r, i, ix);
    for (int a = 0; a < r; ++a) rix[a] = ix[r - 1 - a];
    return linearize(ds, rix, r);
}

// ---------------------------------------------------------------------------
// Element arithmetic.  Int = wrap-around (computed in uint64_t).
// ---------------------------------------------------------------------------
template <class T> struct Arith {
    static inline T zero() { return T(0); }
    static inline T one() { return T(1); }
    static inline T mul(T a, T b) { return a * b; }
    static inline T add(T a, T b) { return a + b; }
    static inline T sub(T a, T b) { return a - b; }
};
template <> struct Arith<int64_t> {
    static inline int64_t zero() { return 0; }
    static inline int64_t one() { return 1; }
    static inline int64_t mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
    static inline int64_t add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
    static inline int64_t sub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
};

// C[i,k] = fold_{j ascending} (acc + A[i,j]*B[j,k]), acc0 = 0.
template <class T>
static void matmul_item(const T* A, const T* B, T* C, int m, int k, int n) {
    using R = Arith<T>;
    for (int i = 0; i < m; ++i)
        for (int c = 0; c < n; ++c) {
            T acc = R::zero();
            for (int j = 0; j < k; ++j) acc = R::add(acc, R::mul(A[i * k + j], B[j * n + c]));
            C[i * n + c] = acc;
        }
}

// contract the last nc indices of A with the first nc of B.
// nc = 0: one multiply per output, no add.
template <class T>
static void prod_item(const T* A, const T* B, T* C, int P, int K, int Q, int nc) {
    using R = Arith<T>;
    if (nc == 0) {
        for (int p = 0; p < P; ++p)
            for (int q = 0; q < Q; ++q) C[(size_t)p * Q + q] = R::mul(A[p], B[q]);
    } else {
        matmul_item(A, B, C, P, K, Q);
    }
}

// SURVEY §8a "Pivot rule": in column k the pivot is the FIRST row r >= k with
// W[r,k] != 0 (exact compare); swap rows k and r; for i > k:
// f = W[i,k] / W[k,k]; W[i,j] -= f*W[k,j] for j = k+1..n-1 then the augmented
// columns; W[i,k] = 0.  W is n x (n+nrhs) row-major.  Returns false when some
// column has no non-zero pivot.  *swaps counts row exchanges.
template <class T>
static bool forward_eliminate(T* W, int n, int nrhs, int* swaps) {
    const int w = n + nrhs;
    *swaps = 0;
    for (int k = 0; k < n; ++k) {
        int r = k;
        while (r < n && W[r * w + k] == T(0)) ++r;
        if (r == n) return false;
        if (r != k) {
            for (int j = 0; j < w; ++j) std::swap(W[k * w + j], W[r * w + j]);
            ++*swaps;
        }
        for (int i = k + 1; i < n; ++i) {
            T f = W[i * w + k] / W[k * w + k];
            for (int j = k + 1; j < w; ++j) W[i * w + j] = W[i * w + j] - f * W[k * w + j];
            W[i * w + k] = T(0);
        }
    }
    return true;
}

// a9: det = (-1)^swaps * prod_{k ascending} pivot_k, product folded from 1;
// singular -> exact +0.
template <class T>
static T det_item(const T* A, int n) {
    std::vector<T> Wv((size_t)n * n);            // any n (the reference's shapes are unbounded unary naturals, MultiIndex.hs:46-48)
    T* W = Wv.data();
    std::memcpy(W, A, sizeof(T) * n * n);
    int swaps;
    if (!forward_eliminate(W, n, 0, &swaps)) return T(0);
    T acc = T(1);
    for (int k = 0; k < n; ++k) acc = acc * W[k * n + k];
    return (swaps & 1) ? -acc : acc;
}

// a11: forward elimination on [A | B] then back-substitution, k = n-1..0,
// s = fold_{j = k+1..n-1 ascending} (s + U[k,j]*X[j,c]) from 0,
// X[k,c] = (Y[k,c] - s) / U[k,k].   status 1 (and zero-filled X) when singular.
template <class T>
static uint8_t solve_item(const T* A, const T* B, T* X, int n, int nrhs) {
    const int w = n + nrhs;
    std::vector<T> Wv((size_t)n * w);
    T* W = Wv.data();
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) W[i * w + j] = A[i * n + j];
        for (int c = 0; c < nrhs; ++c) W[i * w + n + c] = B[i * nrhs + c];
    }
    int swaps;
    if (!forward_eliminate(W, n, nrhs, &swaps)) {
        for (int i = 0; i < n * nrhs; ++i) X[i] = T(0);
        return 1;
    }
    for (int k = n - 1; k >= 0; --k)
        for (int c = 0; c < nrhs; ++c) {
            T s = T(0);
            for (int j = k + 1; j < n; ++j) s = s + W[k * w + j] * X[j * nrhs + c];
            X[k * nrhs + c] = (W[k * w + n + c] - s) / W[k * w + k];
        }
    return 0;
}

// a10: inverse(A) = solve(A, I) (Gauss-Jordan on [A | I] carried out as forward
// elimination + back-substitution, every arithmetic step performed even on the
// structural zeros of I so that the GPU kernel can share one code path).
template <class T>
static uint8_t inverse_item(const T* A, T* X, int n) {
    std::vector<T> Iv((size_t)n * n);
    T* I = Iv.data();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) I[i * n + j] = (i == j) ? T(1) : T(0);
    return solve_item(A, I, X, n, n);
}

// ---------------------------------------------------------------------------
// The rest of the old SquareMatrix surface (CHANGELOG.md:28-29 names `polyEval`
// and `minPoly`; `tr` / `charPoly` per SURVEY.md §8f-4).  No reference code exists
// for any of them: the definitions below ARE the specification
// — **PARITY UNPINNED**.
// ---------------------------------------------------------------------------

// tr: fold_{i ascending} (acc + A[i,i]), acc0 = 0.
template <class T>
static T trace_item(const T* A, int n) {
    using R = Arith<T>;
    T acc = R::zero();
    for (int i = 0; i < n; ++i) acc = R::add(acc, A[i * n + i]);
    return acc;
}

// polyEval A [c0, c1, .., cd] = c0 I + c1 A + .. + cd A^d by Horner's rule:
// R = cd I; for j = d-1 .. 0: R = (R .*. A) with cj added to the diagonal.
// The product is the .*. fold of a6.  No coefficients -> the zero matrix.
template <class T>
static void poly_eval_item(const T* A, const T* c, int ncoef, T* OUT, int n) {
    using R = Arith<T>;
    T Rm[16 * 16], Tm[16 * 16];
    for (int i = 0; i < n * n; ++i) Rm[i] = R::zero();
    if (ncoef > 0)
        for (int i = 0; i < n; ++i) Rm[i * n + i] = c[ncoef - 1];
    for (int j = ncoef - 2; j >= 0; --j) {
        matmul_item(Rm, A, Tm, n, n, n);
        for (int i = 0; i < n; ++i) Tm[i * n + i] = R::add(Tm[i * n + i], c[j]);
        std::memcpy(Rm, Tm, sizeof(T) * n * n);
    }
    std::memcpy(OUT, Rm, sizeof(T) * n * n);
}

// charPoly: coefficients c[0..n] (low order first, c[n] = 1) of det(x I - A) by
// Faddeev-LeVerrier: M = I; for k = 1..n: AM = A .*. M; c[n-k] = negate (tr AM / k);
// M = AM with c[n-k] added to the diagonal.  Fractional element types only.
template <class T>
static void char_poly_item(const T* A, T* c, int n) {
    T M[16 * 16], AM[16 * 16];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) M[i * n + j] = (i == j) ? T(1) : T(0);
    c[n] = T(1);
    for (int k = 1; k <= n; ++k) {
        matmul_item(A, M, AM, n, n, n);
        const T ck = -(trace_item(AM, n) / T(k));
        c[n - k] = ck;
        for (int i = 0; i < n; ++i) AM[i * n + i] = AM[i * n + i] + ck;
        std::memcpy(M, AM, sizeof(T) * n * n);
    }
}

// minPoly: the monic polynomial of least degree m with p(A) = 0, found as the first
// EXACT linear dependency among vec(I), vec(A), vec(A^2), .. (powers by P_m = P_{m-1} .*. A).
// Candidate m is reduced against the echelon basis built from the lower powers, in
// insertion order: f = v[p_i] / b_i[p_i]; v -= f b_i (every position, then v[p_i] := 0);
// the combination g (g = e_m initially) follows: g -= f g_i.  If v is exactly zero the
// answer is g[0..m] (g[m] = 1).  Otherwise v joins the basis with pivot = its first
// non-zero position (the first-non-zero rule of a9-a11).  If I .. A^(n-1) are independent
// the degree is n and the coefficients are charPoly's.  c has n+1 entries, zero above the
// degree; returns the degree.
template <class T>
static uint8_t min_poly_item(const T* A, T* c, int n) {
    const int L = n * n;
    T P[16 * 16], Pn[16 * 16];
    std::vector<T> basis((size_t)n * L), coef((size_t)n * (n + 1));
    int pivot[16];
    int nb = 0;
    for (int m = 0; m < n; ++m) {
        T v[16 * 16], g[17];
        if (m == 0) {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) P[i * n + j] = (i == j) ? T(1) : T(0);
        } else {
            matmul_item(P, A, Pn, n, n, n);
            std::memcpy(P, Pn, sizeof(T) * L);
        }
        std::memcpy(v, P, sizeof(T) * L);
        for (int j = 0; j <= n; ++j) g[j] = (j == m) ? T(1) : T(0);
        for (int i = 0; i < nb; ++i) {
            const T* b = &basis[(size_t)i * L];
            const T* gi = &coef[(size_t)i * (n + 1)];
            const T f = v[pivot[i]] / b[pivot[i]];
            for (int t = 0; t < L; ++t) v[t] = v[t] - f * b[t];
            v[pivot[i]] = T(0);
            for (int j = 0; j <= n; ++j) g[j] = g[j] - f * gi[j];
        }
        int p = 0;
        while (p < L && v[p] == T(0)) ++p;
        if (p == L) {
            for (int j = 0; j <= n; ++j) c[j] = (j <= m) ? g[j] : T(0);
            return (uint8_t)m;
        }
        std::memcpy(&basis[(size_t)nb * L], v, sizeof(T) * L);
        std::memcpy(&coef[(size_t)nb * (n + 1)], g, sizeof(T) * (n + 1));
        pivot[nb++] = p;
    }
    char_poly_item(A, c, n);
    return (uint8_t)n;
}

// rowEchelonForm (SURVEY §8f-4 "rowEchelonForm-style outputs"; the class method itself left the
// library with Data.Tensor.LinearAlgebra, CHANGELOG.md:24-29 - **PARITY UNPINNED**, this IS the
// specification).  Reduced row echelon form of an m x n matrix by Gauss-Jordan with the
// first-non-zero pivot rule and the quotient multipliers of a9-a11 (so that below the pivots the
// matrix evolves exactly as in det / inverse / solve: an exactly repeated row is an exact zero
// row here too, and rank == n iff inverse succeeds), pivot rows normalised afterwards:
//   rank = 0; for c = 0 .. pivot_cols-1 while rank < m:
//     p = first row >= rank with W[p,c] /= 0 (exact compare); none -> next column
//     swap rows p and rank
//     every other row i (ascending): f = W[i,c] / W[rank,c];
//                                    W[i,j] = W[i,j] - f*W[rank,j] (j > c);  W[i,c] = 0
//     d = W[rank,c];  W[rank,j] = W[rank,j] / d  (j > c);  W[rank,c] = 1
//     rank += 1
// Only the first pivot_cols columns are searched for pivots: pivot_cols = n is rowEchelonForm,
// pivot_cols = cols(A) on [A | B] is the old triangularSolve (A reduced, B carried along).
// Returns the rank (number of pivots).
template <class T>
static uint8_t row_echelon_item(const T* A, T* W, int m, int n, int pivot_cols) {
    std::memcpy(W, A, sizeof(T) * m * n);
    int rank = 0;
    for (int c = 0; c < pivot_cols && rank < m; ++c) {
        int p = rank;
        while (p < m && W[p * n + c] == T(0)) ++p;
        if (p == m) continue;
        if (p != rank)
            for (int j = 0; j < n; ++j) std::swap(W[p * n + j], W[rank * n + j]);
        const T d = W[rank * n + c];
        for (int i = 0; i < m; ++i) {
            if (i == rank) continue;
            const T f = W[i * n + c] / d;
            for (int j = c + 1; j < n; ++j) W[i * n + j] = W[i * n + j] - f * W[rank * n + j];
            W[i * n + c] = T(0);
        }
        for (int j = c + 1; j < n; ++j) W[rank * n + j] = W[rank * n + j] / d;
        W[rank * n + c] = T(1);
        ++rank;
    }
    return (uint8_t)rank;
}

// Run f(b0, b1) over contiguous chunks of [0, batch) on `threads` std::threads.
template <class F>
static void par_chunks(size_t batch, int threads, F f) {
    if (threads <= 1 || batch < 1024) { f(size_t(0), batch); return; }
    std::vector<std::thread> th;
    size_t per = (batch + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t b0 = std::min(batch, per * t), b1 = std::min(batch, per * (t + 1));
        if (b0 < b1) th.emplace_back([=] { f(b0, b1); });
    }
    for (auto& x : th) x.join();
}

template <class T>
static void matmul_batch(const void* A, const void* B, void* C, int m, int k, int n, size_t batch, int threads) {
    const T* a = (const T*)A; const T* b = (const T*)B; T* c = (T*)C;
    par_chunks(batch, threads, [=](size_t b0, size_t b1) {
        for (size_t i = b0; i < b1; ++i)
            matmul_item<T>(a + i * m * k, b + i * k * n, c + i * m * n, m, k, n);
    });
}
template <class T>
static void prod_batch(const void* A, const void* B, void* C, int P, int K, int Q, int nc, size_t batch, int threads) {
    const T* a = (const T*)A; const T* b = (const T*)B; T* c = (T*)C;
    par_chunks(batch, threads, [=](size_t b0, size_t b1) {
        for (size_t i = b0; i < b1; ++i)
            prod_item<T>(a + i * (size_t)P * K, b + i * (size_t)K * Q, c + i * (size_t)P * Q, P, K, Q, nc);
    });
}

static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" {

// ---- index conventions (Vector.hs:150-168, :306-307) ----------------------
uint64_t t5o_linearize(const uint64_t* ds, const uint64_t* is, int r) { return linearize(ds, is, r); }
void t5o_unlinearize(const uint64_t* ds, int r, uint64_t i, uint64_t* out) { unlinearize(ds, r, i, out); }
uint64_t t5o_transposeV(const uint64_t* ds, int r, uint64_t i) { return transposeV(ds, r, i); }

// ---- transpose = unRev . t0  (Indexable.hs:101-102, Vector.hs:295-301) ------
// unRev with sh = [] : zh = reverse rh; out[i] = content ! transposeV rh i.
// Pure permutation: elem_bytes in {4, 8}.
int t5o_transpose(int elem_bytes, int rank, const uint64_t* dims, size_t batch, const void* A, void* OUT, int threads) {
    if (rank < 0 || rank > 16) return 1;
    uint64_t d = 1;
    for (int a = 0; a < rank; ++a) d *= dims[a];
    std::vector<uint64_t> perm(d);
    for (uint64_t i = 0; i < d; ++i) perm[i] = transposeV(dims, rank, i);
    const char* in = (const char*)A; char* out = (char*)OUT;
    const uint64_t* pm = perm.data();
    par_chunks(batch, threads, [=](size_t b0, size_t b1) {
        for (size_t b = b0; b < b1; ++b)
            for (uint64_t i = 0; i < d; ++i)
                std::memcpy(out + (b * d + i) * elem_bytes, in + (b * d + pm[i]) * elem_bytes, elem_bytes);
    });
    return 0;
}

// ---- .*. (a6) -------------------------------------------------------------
int t5o_matmul(int dtype, int m, int k, int n, size_t batch, const void* A, const void* B, void* C, int threads) {
    switch (dtype) {
        case 0: matmul_batch<double>(A, B, C, m, k, n, batch, threads); return 0;
        case 1: matmul_batch<float>(A, B, C, m, k, n, batch, threads); return 0;
        case 2: matmul_batch<int64_t>(A, B, C, m, k, n, batch, threads); return 0;
    }
    return 2;
}

// ---- dot (a7): one scalar per pair ------------------------------------------
int t5o_dot(int dtype, int n, size_t batch, const void* X, const void* Y, void* OUT, int threads) {
    return t5o_matmul(dtype, 1, n, 1, batch, X, Y, OUT, threads);
}

// ---- batch-sum of dot (a7, H4): Float/Double partial dots are accumulated in
// double (the stated comparison target for the GPU tree reduction); Int wraps.
int t5o_dot_sum(int dtype, int n, size_t batch, const void* X, const void* Y, void* out_scalar) {
    if (dtype == 2) {
        const int64_t* x = (const int64_t*)X; const int64_t* y = (const int64_t*)Y;
        uint64_t s = 0;
        for (size_t b = 0; b < batch; ++b) {
            int64_t d; matmul_item<int64_t>(x + b * n, y + b * n, &d, 1, n, 1);
            s += (uint64_t)d;
        }
        *(int64_t*)out_scalar = (int64_t)s;
        return 0;
    }
    // Neumaier-compensated double accumulation of the per-pair dots.
    double s = 0.0, comp = 0.0;
    for (size_t b = 0; b < batch; ++b) {
        double d;
        if (dtype == 0) { matmul_item<double>((const double*)X + b * n, (const double*)Y + b * n, &d, 1, n, 1); }
        else if (dtype == 1) { float f; matmul_item<float>((const float*)X + b * n, (const float*)Y + b * n, &f, 1, n, 1); d = f; }
        else return 2;
        double t = s + d;
        if ((s >= 0 ? s : -s) >= (d >= 0 ? d : -d)) comp += (s - t) + d; else comp += (d - t) + s;
        s = t;
    }
    s += comp;
    if (dtype == 0) *(double*)out_scalar = s; else *(float*)out_scalar = (float)s;
    return 0;
}

// ---- prod n (a8) ------------------------------------------------------------
int t5o_prod(int dtype, int rankA, const uint64_t* dimsA, int rankB, const uint64_t* dimsB, int ncontract,
             size_t batch, const void* A, const void* B, void* C, int threads) {
    if (ncontract < 0 || ncontract > rankA || ncontract > rankB) return 1;
    uint64_t P = 1, K = 1, Q = 1;
    for (int a = 0; a < rankA - ncontract; ++a) P *= dimsA[a];
    for (int a = 0; a < ncontract; ++a) {
        if (dimsA[rankA - ncontract + a] != dimsB[a]) return 1;
        K *= dimsB[a];
    }
    for (int a = ncontract; a < rankB; ++a) Q *= dimsB[a];
    switch (dtype) {
        case 0: prod_batch<double>(A, B, C, (int)P, (int)K, (int)Q, ncontract, batch, threads); return 0;
        case 1: prod_batch<float>(A, B, C, (int)P, (int)K, (int)Q, ncontract, batch, threads); return 0;
        case 2: prod_batch<int64_t>(A, B, C, (int)P, (int)K, (int)Q, ncontract, batch, threads); return 0;
    }
    return 2;
}

// ---- det / inverse / solve (a9-a11): Fractional element types only ---------
int t5o_det(int dtype, int n, size_t batch, const void* A, void* OUT, int threads) {
    if (n < 1 || n > 4096) return 1;
    if (dtype == 0) {
        const double* a = (const double*)A; double* o = (double*)OUT;
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) o[b] = det_item<double>(a + b * n * n, n); });
        return 0;
    }
    if (dtype == 1) {
        const float* a = (const float*)A; float* o = (float*)OUT;
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) o[b] = det_item<float>(a + b * n * n, n); });
        return 0;
    }
    return 2;
}
int t5o_inverse(int dtype, int n, size_t batch, const void* A, void* OUT, uint8_t* status, int threads) {
    if (n < 1 || n > 4096) return 1;
    if (dtype == 0) {
        const double* a = (const double*)A; double* o = (double*)OUT;
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) status[b] = inverse_item<double>(a + b * n * n, o + b * n * n, n); });
        return 0;
    }
    if (dtype == 1) {
        const float* a = (const float*)A; float* o = (float*)OUT;
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) status[b] = inverse_item<float>(a + b * n * n, o + b * n * n, n); });
        return 0;
    }
    return 2;
}
int t5o_solve(int dtype, int n, int nrhs, size_t batch, const void* A, const void* B, void* X, uint8_t* status, int threads) {
    if (n < 1 || n > 4096 || nrhs < 1 || nrhs > 4096) return 1;
    if (dtype == 0) {
        const double* a = (const double*)A; const double* bb = (const double*)B; double* x = (double*)X;
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) status[b] = solve_item<double>(a + b * n * n, bb + b * n * nrhs, x + b * n * nrhs, n, nrhs); });
        return 0;
    }
    if (dtype == 1) {
        const float* a = (const float*)A; const float* bb = (const float*)B; float* x = (float*)X;
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) status[b] = solve_item<float>(a + b * n * n, bb + b * n * nrhs, x + b * n * nrhs, n, nrhs); });
        return 0;
    }
    return 2;
}

// ---- tr / polyEval / charPoly / minPoly (§8f-4; parity unpinned) ---------------
int t5o_trace(int dtype, int n, size_t batch, const void* A, void* OUT, int threads) {
    if (n < 1 || n > 16) return 1;
#define T5O_TR(T) { const T* a = (const T*)A; T* o = (T*)OUT; \
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) o[b] = trace_item<T>(a + b * n * n, n); }); return 0; }
    switch (dtype) { case 0: T5O_TR(double) case 1: T5O_TR(float) case 2: T5O_TR(int64_t) }
#undef T5O_TR
    return 2;
}
int t5o_poly_eval(int dtype, int n, const void* coeffs, int ncoef, size_t batch, const void* A, void* OUT, int threads) {
    if (n < 1 || n > 16 || ncoef < 0) return 1;
#define T5O_PE(T) { const T* a = (const T*)A; T* o = (T*)OUT; const T* c = (const T*)coeffs; \
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) poly_eval_item<T>(a + b * n * n, c, ncoef, o + b * n * n, n); }); return 0; }
    switch (dtype) { case 0: T5O_PE(double) case 1: T5O_PE(float) case 2: T5O_PE(int64_t) }
#undef T5O_PE
    return 2;
}
int t5o_char_poly(int dtype, int n, size_t batch, const void* A, void* OUT, int threads) {
    if (n < 1 || n > 16) return 1;
#define T5O_CP(T) { const T* a = (const T*)A; T* o = (T*)OUT; \
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) char_poly_item<T>(a + b * n * n, o + b * (n + 1), n); }); return 0; }
    switch (dtype) { case 0: T5O_CP(double) case 1: T5O_CP(float) }
#undef T5O_CP
    return 2;
}
int t5o_min_poly(int dtype, int n, size_t batch, const void* A, void* OUT, uint8_t* degree, int threads) {
    if (n < 1 || n > 16) return 1;
#define T5O_MP(T) { const T* a = (const T*)A; T* o = (T*)OUT; \
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) degree[b] = min_poly_item<T>(a + b * n * n, o + b * (n + 1), n); }); return 0; }
    switch (dtype) { case 0: T5O_MP(double) case 1: T5O_MP(float) }
#undef T5O_MP
    return 2;
}

int t5o_row_echelon(int dtype, int m, int n, int pivot_cols, size_t batch, const void* A, void* OUT, uint8_t* rank, int threads) {
    if (m < 1 || n < 1 || m > 16 || n > 32 || pivot_cols < 0 || pivot_cols > n) return 1;
#define T5O_RE(T) { const T* a = (const T*)A; T* o = (T*)OUT; \
        par_chunks(batch, threads, [=](size_t b0, size_t b1) { for (size_t b = b0; b < b1; ++b) rank[b] = row_echelon_item<T>(a + b * m * n, o + b * m * n, m, n, pivot_cols); }); return 0; }
    switch (dtype) { case 0: T5O_RE(double) case 1: T5O_RE(float) }
#undef T5O_RE
    return 2;
}

// ---- synthetic inputs -----------------------------------------
// x = splitmix64(seed ^ (op_id << 56) + global_element_index); Double =
// ((x >> 11) * 2^-53) * 2 - 1; Float = ((x >> 40) * 2^-24) * 2 - 1; Int = x.
// `first_elem` lets a shard generate its slice of the global stream.
int t5o_fill(int dtype, uint64_t seed, uint64_t op_id, uint64_t first_elem, size_t count, void* out) {
    const uint64_t base = seed ^ (op_id << 56);
    for (size_t i = 0; i < count; ++i) {
        uint64_t x = splitmix64(base + first_elem + i);
        if (dtype == 0) ((double*)out)[i] = ((double)(x >> 11) * 0x1.0p-53) * 2.0 - 1.0;
        else if (dtype == 1) ((float*)out)[i] = ((float)(x >> 40) * 0x1.0p-24f) * 2.0f - 1.0f;
        else if (dtype == 2) ((int64_t*)out)[i] = (int64_t)x;
        else return 2;
    }
    return 0;
}

// elimination configs — every `swap_every`-th matrix gets
// A[1,1] := 0 (forces a row swap), every `sing_every`-th gets row 2 := row 1
// (exactly singular).  first_item = global index of matrix 0 of this slice.
int t5o_plant_specials(int dtype, int n, uint64_t first_item, size_t batch, void* A, uint64_t swap_every, uint64_t sing_every) {
    if (dtype != 0 && dtype != 1) return 2;
    const size_t es = dtype == 0 ? 8 : 4;
    char* a = (char*)A;
    for (size_t b = 0; b < batch; ++b) {
        uint64_t g = first_item + b;
        char* m = a + b * (size_t)n * n * es;
        if (swap_every && g % swap_every == 0) std::memset(m, 0, es);
        if (sing_every && g % sing_every == 0 && n >= 2) std::memcpy(m + (size_t)n * es, m, (size_t)n * es);
    }
    return 0;
}

}  // extern "C"
