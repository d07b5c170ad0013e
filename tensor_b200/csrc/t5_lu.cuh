// (kernel + launch templates; instantiated per operation by t5_lu_det.cu / t5_lu_solve.cu / t5_lu_inverse.cu so that the
// three operations compile in parallel)
// Gaussian elimination for 9 <= n <= 128: ONE CTA PER MATRIX, the working matrix distributed over the CTA's
// registers 2D-cyclically (SquareMatrix det / inverse, LinearSystem solve of SURVEY.md §8a a9-a11 for the sizes the
// record-per-thread and row-per-lane kernels do not reach; the reference's shapes are unbounded unary naturals,
// src/Data/MultiIndex.hs:46-48).
//
// Layout.  The CTA is a TG x TG grid of threads (TG = 8: n <= 32, TG = 16: n <= 128).  Thread (ty, tx) owns the
// elements w[i][j] with i = ty (mod TG), j = tx (mod TG): an NB x NB register tile, NB = ceil(n / TG).  Cyclic,
// not blocked: as the elimination front k advances every thread keeps about the same share of the trailing
// matrix.  All register indices are compile-time (the run-time front only appears in predicates), so the
// matrix never touches local memory.
//
// One elimination step k (row k and the raw column k were published by step k - 1):
//   1. pivot = row[k]; exactly zero -> slow path: first row p > k with a non-zero entry in column k (atomicMin),
//      swap rows k and p WHOLE (stored multipliers included), republish; none -> singular
//   2. the multipliers f(i) = column[i] / pivot for i > k.  n <= 64 (TG <= 8): ONE division per row - thread t divides
//      for row t and the quotients travel through shared memory (one more group barrier).  TG = 16: every thread divides
//      for ITS rows through the pivot's hoisted reciprocal, as one batch (no second barrier: with only the TG owners of
//      column k dividing, 15/16 of a 256-thread CTA idled through the division latency every step - 3.5x slower at
//      n = 128).  solve / inverse also keep f(i) in the shared factor array, where the zero would go.
//   3. every thread: w[i][j] = w[i][j] - f(i) * row[j] for its i > k, j > k  (2 NB shared loads per NB^2 updates)
//   4. the owners of row k + 1 and of column k + 1 publish them (double-buffered by the parity of k)   -- barrier
// The tile index of the front, A = k / TG, is the OUTER compile-time loop (static_for) and the owner index k mod TG the
// inner run-time one, so every register index and loop bound in the step is a constant with no per-step dispatch.
// (History: run-time select chains over the tile rows - 57 % FSEL / ISETP / LOP3; then a `switch` on k / TG three
// times per step - indirect branches, 10 % of the issued instructions and the largest stall sites; ncu.)  Register
// columns the front has passed are dead - nothing reads them again - so they take part in the update unconditionally
// (garbage in, garbage out, column by column) instead of costing a predicate per element.
// Groups that share a warp (4 x 4 threads, two matrices per warp) run in lockstep: same trip counts (a group past the
// end of the batch recomputes the last item and stores nothing), the pivot slow path entered warp-wide, a singular
// group computing on and zero-filling at the end - so their barrier is the plain full-mask __syncwarp.
// Per element the operations and their order are the oracle's (oracle/t5_oracle.cpp forward_eliminate): updates
// in ascending k, multiply-round-subtract-round (this TU is compiled with -fmad=false -prec-div=true), pivot =
// FIRST non-zero, so det, the factors and everything derived from them are BIT-IDENTICAL to the oracle.
//
// det = (+-) fold_k (acc * pivot_k) from 1, singular -> +0.
// solve / inverse: the factors (U above, the kept multipliers below the diagonal) are dumped to shared memory
// once.  Because a swap moves whole rows - multipliers included - the interleaved swap / update sequence the
// oracle applies to the right-hand sides equals: permute the rows of B by the composed permutation, then
// y[i] -= f(i,k) * y[k] for k ascending with NO further swaps (the same values meet in the same order).  So the
// right-hand sides are loaded already permuted (inverse: the permuted identity, integer bookkeeping), 2D-cyclic
// like the matrix, and take one barrier per step (pivot row of Y published, f(i,k) read straight from the shared
// factors).  Back-substitution is a sequential recurrence per column whose order the specification fixes
// (s = fold_{j>k ascending}(s + u[k,j] x[j]) from 0, x[k] = (y[k] - s) / u[k,k]): Y goes to shared memory and
// thread c runs column c, reading u[k,j] as a broadcast.  Singular: status 1, item zero-filled.
#pragma once
#include <type_traits>

#include "t5_device.cuh"
#include "t5_internal.h"
#include "t5_recip.cuh"

namespace t5 {
namespace lu {

struct Args {
    const void* A;
    const void* B;          // n x nrhs per item (solve); unused for det / inverse
    void* X;                // determinant / solution / inverse
    unsigned char* status;  // unused for det
    unsigned long long count;
    int n, nrhs;
    int ldl, ldy;           // row pitches (elements) of the shared factors and of the shared right-hand sides
    int chunk;              // right-hand-side columns processed per pass (multiple of TG, <= TG * NB)
    unsigned group_smem;    // shared-memory bytes of one group (a CTA holds G groups)
};

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// CTA-uniform dispatch on a tile index: fn(std::integral_constant<int, A>) with A == a_rt
template <int NB, int A = 0, class F>
__device__ __forceinline__ void with_tile_index(int a_rt, F&& fn) {
    if constexpr (A < NB) {
        if (a_rt == A) fn(std::integral_constant<int, A>{});
        else with_tile_index<NB, A + 1>(a_rt, fn);
    }
}

// MODE 0: det, 1: solve (rhs from B), 2: inverse (rhs = identity, nrhs == n)
// G groups of TG x TG threads per CTA, one matrix per group.  Groups never interact: each has its own slice of shared
// memory and its own barrier - the CTA barrier when the CTA is one group, a masked __syncwarp when a group is part of a
// warp (TG = 4: two matrices per warp, every warp independent of the others - no CTA-wide barrier anywhere).
template <class T, int TG, int NB, int MODE, int G, bool ONEDIV, int NBR>
__global__ void __launch_bounds__(TG * TG * G) lu_kernel(Args p) {   // (a 64-register cap for 8 CTAs per SM spilled: det 16x16 0.665 -> 0.729 ms)
    constexpr int THREADS = TG * TG;                   // threads of ONE group
    constexpr int NMAX = TG * NB;
    static_assert(G == 1 || THREADS <= 32, "several groups per CTA only when a group fits in a warp");
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int gid = threadIdx.x / THREADS;
    unsigned char* smem_raw = smem_all + (size_t)gid * p.group_smem;
    // (the groups of a warp run in lockstep - same n, same trip counts, the pivot slow path entered warp-wide - so a group
    // barrier is the plain full-mask __syncwarp; a run-time member mask costs MATCH / REDUX / VOTE per barrier)
    auto group_sync = [&]() {
        if constexpr (THREADS >= 32) __syncthreads(); else __syncwarp();
    };
    T* prow = reinterpret_cast<T*>(smem_raw);          // [2][NMAX]   published pivot row (double-buffered)
    T* fcol = prow + 2 * NMAX;                          // [2][NMAX]   published raw column of the step (the numerators of the multipliers)
    T* xrow = fcol + 2 * NMAX;                          // [2][NMAX]   row exchange for swaps
    T* fmul = xrow + 2 * NMAX;                          // [NMAX]      (ONEDIV) multipliers f(i, k) of the current step, one division per thread
    constexpr bool RIDE = NBR > 0;                      // right-hand sides eliminated together with A, NBR tile columns of them
    constexpr int RW = TG * (NBR > 0 ? NBR : 1);
    T* prb = fmul + NMAX;                               // [2][RW]     (RIDE) right-hand-side entries of the published pivot row
    T* xrb = prb + 2 * RW;                              // [2][RW]     (RIDE) right-hand-side entries of the rows exchanged by a swap
    int* perm = reinterpret_cast<int*>(xrb + 2 * RW);    // [NMAX] row i of the permuted system = row perm[i] of the input
    int* ctl = perm + NMAX;                             // [4] pivot search result, swap count
    T* LU = reinterpret_cast<T*>(ctl + 4);              // [n][ldl]  (MODE != 0): U on and above the diagonal, multipliers below
    T* Y = LU + (MODE != 0 ? (size_t)p.n * p.ldl : 0);  // [n][ldy]  (MODE != 0)

    const int tid = threadIdx.x % THREADS, ty = tid / TG, tx = tid % TG;
    const int n = p.n;
    const T* gA = (const T*)p.A;
    // number of this thread's tile rows / columns that lie inside the matrix
    const int a_end = ty < n ? (n - ty + TG - 1) / TG : 0;

    // (groups of one warp stay in lockstep: a group past the end of the batch recomputes the last item and stores nothing)
    for (unsigned long long base = (unsigned long long)blockIdx.x * G; base < p.count; base += (unsigned long long)gridDim.x * G) {
        const bool live = base + gid < p.count;
        const unsigned long long item = live ? base + gid : p.count - 1;
        // ---- load: thread (ty, tx) takes A[ty + TG a][tx + TG b]; TG consecutive threads read TG consecutive elements
        T w[NB][NB];
        const T* a_item = gA + item * (unsigned long long)n * n;
        #pragma unroll
        for (int a = 0; a < NB; ++a)
            #pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int i = ty + TG * a, j = tx + TG * b;
                w[a][b] = (i < n && j < n) ? __ldg(a_item + (size_t)i * n + j) : T(0);
            }
        // RIDE (solve with at most TG * NBR right-hand sides; inverse for NB <= 4): the right-hand sides ride along as NBR
        // more tile columns of the working matrix - row i of B next to row i of A, column c in thread column c mod TG - and
        // take every swap and update together with their row, exactly as the oracle eliminates [A | B]; the separate
        // forward substitution (n more barriers) is gone (profiles/r02ar_time_lu.txt, r02as_time_lu.txt).
        T wb[RIDE ? NB : 1][RIDE ? NBR : 1];
        if constexpr (RIDE) {
            const T* b_item = MODE == 1 ? (const T*)p.B + item * (unsigned long long)n * p.nrhs : nullptr;
            #pragma unroll
            for (int a = 0; a < NB; ++a)
                #pragma unroll
                for (int b = 0; b < NBR; ++b) {
                    const int i = ty + TG * a, c = tx + TG * b;
                    if constexpr (MODE == 2) wb[a][b] = (i < n && i == c) ? T(1) : T(0);
                    else wb[a][b] = (i < n && c < p.nrhs) ? __ldg(b_item + (size_t)i * p.nrhs + c) : T(0);
                }
        }
        for (int i = tid; i < NMAX; i += THREADS) perm[i] = i;     // (a 4 x 4 group has fewer threads than rows)
        if (tid == 0) ctl[1] = 0;
        T acc = T(1);                                   // thread 0: running product of the pivots
        bool singular = false;
        group_sync();

        // ---- elimination of A: ONE barrier per step.  At the top of step k the pivot row and the raw column k (both as
        // left by step k - 1) are already published.  Every thread divides for its own NB rows (f = column / pivot with
        // the pivot's reciprocal hoisted, t5_recip.cuh: bit-identical to `/`) - redundantly across the TG threads of a
        // thread-row, but in parallel, instead of TG owner threads dividing while the other 15/16 of the CTA wait at a
        // second barrier - updates its tile, and the owners of row / column k + 1 publish them for the next step.
        // Row k1 and column k1 of the CURRENT registers -> buffer k1 & 1.  A1 = k1 / TG is a compile-time tile index.
        auto publish = [&](auto A1_, int k1) {
            constexpr int A1 = decltype(A1_)::value;
            if constexpr (A1 < NB) {
                const int r1 = k1 - A1 * TG;
                T* pr1 = prow + (k1 & 1) * NMAX;
                T* cr1 = fcol + (k1 & 1) * NMAX;
                if (ty == r1) {
                    #pragma unroll
                    for (int b = A1; b < NB; ++b) pr1[tx + TG * b] = w[A1][b];     // (columns before the front are dead)
                    if constexpr (RIDE) {
                        #pragma unroll
                        for (int b = 0; b < NBR; ++b) prb[(k1 & 1) * RW + tx + TG * b] = wb[A1][b];
                    }
                }
                if (tx == r1) {
                    #pragma unroll
                    for (int a = A1; a < NB; ++a) cr1[ty + TG * a] = w[a][A1];
                }
            }
        };
        publish(std::integral_constant<int, 0>{}, 0);
        group_sync();
        // The tile index of the front, A = k / TG, is the OUTER, compile-time loop; the owner index rk = k mod TG the inner
        // run-time one: every register index below is a constant without any per-step dispatch.  (The first version looped
        // over k and switched on k / TG three times per step - update, row publish, column publish: indirect branches whose
        // targets the front end had to fetch, 10 % of the issued instructions and the largest stall sites in ncu.)
        static_for<0, NB>([&](auto A_) {
            constexpr int A = decltype(A_)::value;
            const int rk_end = (THREADS >= 32 && singular) ? 0 : (n - A * TG < TG ? n - A * TG : TG);
            for (int rk = 0; rk < rk_end; ++rk) {
                const int k = A * TG + rk;                  // row / column k: owner thread-row (= thread-column) rk, tile index A
                const T* pr = prow + (k & 1) * NMAX;
                const T* cr = fcol + (k & 1) * NMAX;
                T piv = pr[k];
                const bool need = piv == T(0);
                // several groups per warp: the slow path is entered by the whole warp (a group that does not need it runs it
                // as a no-op with pidx = k), so every barrier below is a plain full-mask __syncwarp
                const bool enter = THREADS >= 32 ? need : __any_sync(0xffffffffu, need);
                if (enter) {
                    // slow path (rare: planted zeros, exactly singular items): first row p > k with a non-zero in column k
                    if (tid == 0) ctl[0] = n;
                    group_sync();
                    if (need)
                        for (int i = k + 1 + tid; i < n; i += THREADS)
                            if (cr[i] != T(0)) atomicMin(&ctl[0], i);
                    group_sync();
                    int pidx = need ? ctl[0] : k;
                    if (pidx >= n) {                                      // no pivot: singular (uniform within the group)
                        singular = true;
                        if constexpr (THREADS >= 32) break;               // the CTA is one group: stop here
                        pidx = k;                                         // part of a warp: keep in step with the other group, results are discarded
                    }
                    const int ap = pidx / TG, rp = pidx - ap * TG;
                    // exchange rows k and pidx: the live part in registers, the kept multipliers (columns < k) in the shared factors
                    if (ty == rk) {
                        #pragma unroll
                        for (int b = 0; b < NB; ++b) xrow[tx + TG * b] = w[A][b];
                        if constexpr (RIDE) {
                            #pragma unroll
                            for (int b = 0; b < NBR; ++b) xrb[tx + TG * b] = wb[A][b];
                        }
                    }
                    if (ty == rp)
                        with_tile_index<NB>(ap, [&](auto P_) {
                            constexpr int P = decltype(P_)::value;
                            #pragma unroll
                            for (int b = 0; b < NB; ++b) xrow[NMAX + tx + TG * b] = w[P][b];
                            if constexpr (RIDE) {
                                #pragma unroll
                                for (int b = 0; b < NBR; ++b) xrb[RW + tx + TG * b] = wb[P][b];
                            }
                        });
                    group_sync();
                    if (ty == rp && pidx != k)                            // first: when one thread owns both rows they are different tile rows
                        with_tile_index<NB>(ap, [&](auto P_) {
                            constexpr int P = decltype(P_)::value;
                            #pragma unroll
                            for (int b = 0; b < NB; ++b) w[P][b] = xrow[tx + TG * b];
                            if constexpr (RIDE) {
                                #pragma unroll
                                for (int b = 0; b < NBR; ++b) wb[P][b] = xrb[tx + TG * b];
                            }
                        });
                    if (ty == rk) {
                        #pragma unroll
                        for (int b = 0; b < NB; ++b) w[A][b] = xrow[NMAX + tx + TG * b];
                        if constexpr (RIDE) {
                            #pragma unroll
                            for (int b = 0; b < NBR; ++b) wb[A][b] = xrb[RW + tx + TG * b];
                        }
                    }
                    if (MODE != 0 && !RIDE && pidx != k)
                        for (int j = tid; j < k; j += THREADS) {
                            const T t = LU[(size_t)k * p.ldl + j];
                            LU[(size_t)k * p.ldl + j] = LU[(size_t)pidx * p.ldl + j];
                            LU[(size_t)pidx * p.ldl + j] = t;
                        }
                    if (tid == 0 && pidx != k) { const int t = perm[k]; perm[k] = perm[pidx]; perm[pidx] = t; ctl[1] += 1; }
                    publish(A_, k);                                       // (same buffer: every thread of the group is inside this branch)
                    group_sync();
                    piv = pr[k];
                }
                acc = acc * piv;                                          // (every thread: cheaper than a predicate; thread 0's copy is used)
                if constexpr (ONEDIV) {
                    // One division per ROW, not per (row, thread of the thread-row): thread t divides for row t (a group has at
                    // least as many threads as rows except 4 x 4 groups with NB > 4), the multipliers go through shared memory
                    // and the update below is loads + multiply-subtract only.
                    for (int i = k + 1 + tid; i < n; i += THREADS) {
                        const T f = cr[i] / piv;
                        fmul[i] = f;
                        if (MODE != 0 && !RIDE) LU[(size_t)i * p.ldl + k] = f;    // the multiplier goes where the zero would go (RIDE: nothing replays it)
                    }
                    group_sync();
                    T prj[NB];
                    #pragma unroll
                    for (int b = A; b < NB; ++b) prj[b] = pr[tx + TG * b];
                    T prbj[RIDE ? NBR : 1];
                    #pragma unroll
                    for (int b = 0; b < (RIDE ? NBR : 1); ++b) prbj[b] = RIDE ? prb[(k & 1) * RW + tx + TG * b] : T(0);
                    (void)prbj;
                    #pragma unroll
                    for (int a = A; a < NB; ++a) {
                        // rows of tile row A are past the front only for ty > rk; later tile rows are, whole; rows >= n do not exist
                        if ((a > A || ty > rk) && a < a_end) {
                            const T f = fmul[ty + TG * a];
                            #pragma unroll
                            for (int b = A; b < NB; ++b) w[a][b] = w[a][b] - f * prj[b];
                            if constexpr (RIDE) {
                                #pragma unroll
                                for (int b = 0; b < NBR; ++b) wb[a][b] = wb[a][b] - f * prbj[b];
                            }
                        }
                    }
                } else {
                    const Recip<T> rc(piv);
                    T prj[NB], f[NB];
                    #pragma unroll
                    for (int b = A; b < NB; ++b) prj[b] = pr[tx + TG * b];
                    // the multipliers of all this thread's tile rows first, as one batch of independent exact divisions with one
                    // fallback branch for the whole batch (rows that are not past the front divide stale values: unused, and kept
                    // out of the fallback test - stale zeros / denormals would send every step down the slow division)
                    // (worth it only where a thread has many rows and few warps cover for it: 16 x 16 threads, NB = 8:
                    // n = 128 det 7.8 -> 5.9 ms; smaller tiles lost 5-50 % to the divisions of rows that are not past the front)
                    constexpr bool BATCH = TG == 16 && NB == 8;
                    if constexpr (BATCH) {
                        bool all_fast = true;
                        #pragma unroll
                        for (int a = A; a < NB; ++a) {
                            bool ok = true;
                            f[a] = rc.div_try(cr[ty + TG * a], ok);
                            all_fast = all_fast && (ok || !((a > A || ty > rk) && a < a_end));
                        }
                        if (!all_fast) {
                            #pragma unroll
                            for (int a = A; a < NB; ++a) f[a] = rc.div_fix(cr[ty + TG * a], f[a]);
                        }
                    }
                    #pragma unroll
                    for (int a = A; a < NB; ++a) {
                        if ((a > A || ty > rk) && a < a_end) {
                            if constexpr (!BATCH) f[a] = rc.div(cr[ty + TG * a]);
                            if (MODE != 0 && !RIDE && tx == rk) LU[(size_t)(ty + TG * a) * p.ldl + k] = f[a];   // the multiplier goes where the zero would go
                            #pragma unroll
                            for (int b = A; b < NB; ++b) w[a][b] = w[a][b] - f[a] * prj[b];
                            if constexpr (RIDE) {
                                #pragma unroll
                                for (int b = 0; b < NBR; ++b) wb[a][b] = wb[a][b] - f[a] * prb[(k & 1) * RW + tx + TG * b];
                            }
                        }
                    }
                }
                if (k + 1 < n) {
                    if (rk + 1 < TG) publish(A_, k + 1);
                    else publish(std::integral_constant<int, A + 1>{}, k + 1);
                }
                group_sync();
            }
        });

        if (MODE == 0) {
            if (tid == 0 && live) {
                T d = singular ? T(0) : ((ctl[1] & 1) ? -acc : acc);
                ((T*)p.X)[item] = d;
            }
            group_sync();                            // perm / ctl are rewritten by the next item
            continue;
        }

        // ---- solve / inverse
        const int nrhs = p.nrhs;
        T* gX = (T*)p.X + item * (unsigned long long)n * nrhs;
        if (tid == 0 && live) p.status[item] = singular ? 1 : 0;
        if (THREADS >= 32 && singular) {                 // (a group that shares its warp runs on in step and zero-fills at the end)
            for (int e = tid; e < n * nrhs; e += THREADS) gX[e] = T(0);
            group_sync();
            continue;
        }
        // U (row i, columns j >= i) is final in the registers of its owners
        #pragma unroll
        for (int a = 0; a < NB; ++a)
            #pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int i = ty + TG * a, j = tx + TG * b;
                if (i < n && j < n && j >= i) LU[(size_t)i * p.ldl + j] = w[a][b];
            }
        group_sync();
        const T* gB = MODE == 1 ? (const T*)p.B + item * (unsigned long long)n * nrhs : nullptr;
        for (int c0 = 0; c0 < nrhs; c0 += p.chunk) {
            // right-hand sides of this pass, rows already permuted; same cyclic distribution (NB x NB tile, columns c0 + tx + TG b)
            T y[NB][NB];
            if constexpr (RIDE) {
                // the right-hand sides were eliminated together with A (one pass: nrhs <= TG * NBR = chunk)
                #pragma unroll
                for (int a = 0; a < NB; ++a)
                    #pragma unroll
                    for (int b = 0; b < NB; ++b) y[a][b] = b < NBR ? wb[a][b < NBR ? b : 0] : T(0);
            } else {
                #pragma unroll
                for (int a = 0; a < NB; ++a)
                    #pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const int i = ty + TG * a, c = c0 + tx + TG * b;
                        T v = T(0);
                        if (i < n && c < nrhs && TG * b < p.chunk) {
                            const int src = perm[i];
                            v = MODE == 2 ? (src == c ? T(1) : T(0)) : __ldg(gB + (size_t)src * nrhs + c);
                        }
                        y[a][b] = v;
                    }
                const int nbl = p.chunk / TG;                 // column tiles of this pass that hold right-hand sides (1 for a single vector)
                static_for<0, NB>([&](auto A_) {              // tile index of row k outside, owner index inside: no per-step dispatch
                    constexpr int A = decltype(A_)::value;
                    const int rk_end = n - A * TG < TG ? n - A * TG : TG;
                    for (int rk = 0; rk < rk_end; ++rk) {
                        const int k = A * TG + rk;
                        T* pr = prow + (k & 1) * NMAX;
                        if (ty == rk) {
                            #pragma unroll
                            for (int b = 0; b < NB; ++b)
                                if (b < nbl) pr[tx + TG * b] = y[A][b];
                        }
                        group_sync();
                        T prj[NB];
                        #pragma unroll
                        for (int b = 0; b < NB; ++b) prj[b] = b < nbl ? pr[tx + TG * b] : T(0);
                        #pragma unroll
                        for (int a = A; a < NB; ++a) {
                            if ((a > A || ty > rk) && a < a_end) {
                                const T f = LU[(size_t)(ty + TG * a) * p.ldl + k];
                                #pragma unroll
                                for (int b = 0; b < NB; ++b)
                                    if (b < nbl) y[a][b] = y[a][b] - f * prj[b];
                            }
                        }
                    }
                });
            }
            #pragma unroll
            for (int a = 0; a < NB; ++a)
                #pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int i = ty + TG * a, cc = tx + TG * b;
                    if (i < n && cc < p.chunk) Y[(size_t)i * p.ldy + cc] = y[a][b];
                }
            group_sync();
            // back-substitution: thread cc owns column c0 + cc; u[k][j] is a broadcast read
            for (int cc = tid; cc < p.chunk && c0 + cc < nrhs; cc += THREADS) {
                for (int k = n - 1; k >= 0; --k) {
                    T s = T(0);
                    const T* urow = LU + (size_t)k * p.ldl;
                    const T* ycol = Y + cc;
                    int j = k + 1;
                    // the sum is one dependent chain (its order is the specification's), but the operands are not: four
                    // products are formed ahead of the adds that consume them, so the chain runs at the add latency
                    // instead of shared-load + multiply + add per term
                    for (; j + 4 <= n; j += 4) {
                        const T p0 = urow[j] * ycol[(size_t)j * p.ldy], p1 = urow[j + 1] * ycol[(size_t)(j + 1) * p.ldy];
                        const T p2 = urow[j + 2] * ycol[(size_t)(j + 2) * p.ldy], p3 = urow[j + 3] * ycol[(size_t)(j + 3) * p.ldy];
                        s = s + p0; s = s + p1; s = s + p2; s = s + p3;
                    }
                    for (; j < n; ++j) s = s + urow[j] * ycol[(size_t)j * p.ldy];
                    Y[(size_t)k * p.ldy + cc] = (ycol[(size_t)k * p.ldy] - s) / urow[k];
                }
            }
            group_sync();
            const int width = (nrhs - c0) < p.chunk ? (nrhs - c0) : p.chunk;
            for (int e = tid; e < n * width && live; e += THREADS) {
                const int i = e / width, cc = e - i * width;
                gX[(size_t)i * nrhs + c0 + cc] = singular ? T(0) : Y[(size_t)i * p.ldy + cc];
            }
            group_sync();
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

template <class T, int TG, int NB, int MODE, int G = 1, int NBR = 0, bool ONEDIV = (TG <= 8)>
static int launch(const Shard& s, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    constexpr int NMAX = TG * NB;
    Args p;
    p.A = A; p.B = B; p.X = X; p.status = (unsigned char*)status; p.count = s.count; p.n = n; p.nrhs = nrhs;
    p.ldl = n | 1;                                           // odd pitch: rows a thread-row apart land in different banks
    size_t smem = sizeof(T) * (7 * NMAX + 4 * TG * (NBR > 0 ? NBR : 1)) + sizeof(int) * (NMAX + 4);
    p.chunk = TG; p.ldy = 1;
    if (MODE != 0) {
        // as many right-hand-side columns per pass as the tile holds and shared memory allows (the factors stay resident)
        const size_t budget = 200 * 1024 / G;
        int chunk = ((nrhs + TG - 1) / TG) * TG;
        if (chunk > NMAX) chunk = NMAX;
        while (chunk > TG && smem + sizeof(T) * ((size_t)n * p.ldl + (size_t)n * (chunk | 1)) > budget) chunk -= TG;
        if (NBR > 0 && chunk < nrhs) return launch<T, TG, NB, MODE, G, 0, ONEDIV>(s, n, nrhs, A, B, X, status);   // (does not fit one pass)
        p.chunk = chunk;
        p.ldy = chunk | 1;
        smem += sizeof(T) * ((size_t)n * p.ldl + (size_t)n * p.ldy);
    }
    smem = (smem + 15) / 16 * 16;
    p.group_smem = (unsigned)smem;
    smem *= G;
    if (smem > 227 * 1024) return T5I_UNSUPPORTED;
    auto kern = lu_kernel<T, TG, NB, MODE, G, ONEDIV, NBR>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return T5I_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TG * TG * G, smem) != cudaSuccess || per_sm < 1) return T5I_CUDA;
    const unsigned long long slots = (unsigned long long)per_sm * sm_count();
    const unsigned long long want = (s.count + G - 1) / G;
    const int grid = (int)(want < slots ? want : slots);
    kern<<<grid, TG * TG * G, smem, s.stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? T5I_OK : T5I_CUDA;
}

template <class T, int MODE>
static int dispatch(const Shard& s, int n, int nrhs, const void* A, const void* B, void* X, void* status) {
    if (s.layout != T5L_AOS) return T5I_NOT_SPECIALISED;
    if (n < 1 || n > 128) return T5I_NOT_SPECIALISED;     // (n <= 8 only gets here for column counts without a faster kernel)
    if (s.count == 0) return T5I_OK;
    // Geometry from tools/time_lu.py sweeps (profiles/r02i_lu_geometry.txt): an 8 x 8 thread grid as far as its tile fits in
    // registers (n <= 64: 64 doubles per thread) - two-warp CTAs whose barriers are cheap and of which many are resident,
    // each thread amortising the per-step overhead over NB^2 updates: n = 48 det 3.34 ms (16 x 16 threads) -> 2.03 ms,
    // n = 64 3.34 -> 2.49 ms; beyond that 16 x 16 threads.
    // n <= 16 (and the light operations up to 32): 4 x 4 threads per matrix, EIGHT matrices per 128-thread CTA, two per
    // warp - a group synchronises with a masked __syncwarp, so no warp ever waits for another: det 16x16 Double
    // 1.85 -> 0.95 ms (1.13 TB/s), inverse 5.29 -> 2.72 ms, 9x9 det 2.17 -> 1.08 ms (tools/time_lu.py, profiles/r02m_*).
    // Heavier tiles (inverse from n = 17, everything dense from n = 25) do better spread over 64 threads.
    // right-hand sides that ride along with the elimination (NBR tile columns): one column tile for solves with at most TG
    // right-hand sides at any n; up to n right-hand sides while the tile is small (NB <= 4); the whole identity of `inverse`
    // for 17 <= n <= 32 (8 x 8 threads: 32x32 Double 4.84 -> 4.05 ms, 24x24 4.60 -> 3.95; with 4 x 4 threads per matrix the
    // doubled register tile cost more than the forward pass it saves: 16x16 2.27 -> 2.46)
#define LU_LAUNCH(TG_, NB_, G_) do { \
        constexpr int FULL_ = (NB_ <= 4) ? NB_ : 0; \
        if (MODE == 2 && FULL_ > 0 && TG_ == 8) return launch<T, TG_, NB_, MODE, G_, (MODE == 2 && TG_ == 8 ? FULL_ : 0)>(s, n, nrhs, A, B, X, status); \
        if (MODE == 1 && nrhs <= TG_) return launch<T, TG_, NB_, MODE, G_, (MODE == 1 ? 1 : 0)>(s, n, nrhs, A, B, X, status); \
        if (MODE == 1 && FULL_ > 1 && nrhs <= TG_ * NB_) return launch<T, TG_, NB_, MODE, G_, (MODE == 1 ? FULL_ : 0)>(s, n, nrhs, A, B, X, status); \
        return launch<T, TG_, NB_, MODE, G_>(s, n, nrhs, A, B, X, status); } while (0)
    const bool light = MODE == 0 || (MODE == 1 && nrhs <= 4);
    if (n <= 8) LU_LAUNCH(4, 2, 8);
    if (n <= 16) LU_LAUNCH(4, 4, 8);
    if (n <= 24 && light) LU_LAUNCH(4, 6, 8);
    if (n <= 32 && MODE == 1 && nrhs <= 4) LU_LAUNCH(4, 8, 8);
    if (n <= 24) LU_LAUNCH(8, 3, 1);
    if (n <= 32) LU_LAUNCH(8, 4, 1);
    if (n <= 40) LU_LAUNCH(8, 5, 1);
    if (n <= 48) LU_LAUNCH(8, 6, 1);
    if (n <= 56) LU_LAUNCH(8, 7, 1);
    if (n <= 64) LU_LAUNCH(8, 8, 1);
    if (n <= 80) LU_LAUNCH(16, 5, 1);
    if (n <= 96) LU_LAUNCH(16, 6, 1);
    if (n <= 112) LU_LAUNCH(16, 7, 1);
    LU_LAUNCH(16, 8, 1);
#undef LU_LAUNCH
}

}  // namespace lu

}  // namespace t5
