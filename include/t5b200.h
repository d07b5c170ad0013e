/* t5b200.h — C ABI of the B200-native batched fixed-dimension linear-algebra
 * backend for the Haskell `tensor` library (tensor5/tensor).
 *
 * This is the drop-in boundary: the entry points below are exactly what a
 * `foreign import ccall safe` binding in a device-backed sibling of
 * Data.Tensor.Vector would bind (see INTEGRATION.md for the Haskell stubs).
 * The reference has no FFI today; each entry cites the
 * reference interface (file:line under the reference tree) whose job it takes
 * over on the device.
 *
 * Conventions
 *   - Every tensor of a batch is row-major, last index fastest — the reference's
 *     wire order (src/Data/Tensor/Vector.hs:150-158 linearize, :360-367
 *     fromList/toList).  A batch is `batch` such tensors.
 *   - Host data is always AoS (tensors concatenated).  A device buffer is either
 *     T5_AOS (same order) or T5_SOA (element-major: element e of item b of a
 *     shard lives at [e * shard_capacity + b]).
 *   - The batch axis is sharded over the context's devices by contiguous ranges
 *     (t5_partition).  No collective runs on any path except t5_dot_sum.
 *   - Every entry returns an int status (T5_OK == 0).  Nothing aborts, no C++
 *     exception crosses the boundary, and there is NO CPU fallback: an
 *     unsupported (dtype, shape, layout) is T5_ERR_UNSUPPORTED.
 *   - Launches are asynchronous on the context's per-device streams;
 *     t5_download, t5_sync, t5_dot_sum and t5_timer_end synchronise.
 *   - Inputs are never written.  Outputs must be distinct buffers.
 *   - Thread safety: a context serialises its entries with an internal mutex and
 *     sets the current device at the top of every entry, so `ccall safe` callers
 *     on several OS threads (GHC -threaded) and ForeignPtr finalizers running
 *     t5_free on arbitrary threads are legal.
 */
#ifndef T5B200_H
#define T5B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t5_ctx t5_ctx;
typedef struct t5_buf t5_buf;

/* Element types.  Double / Float / Int of the reference's element parameter `e`
 * (src/Data/Tensor/Vector.hs:121-124 `content :: V.Vector e`); T5_U8 is used
 * only for the per-item status bytes of inverse / solve (the `Maybe` result). */
enum { T5_F64 = 0, T5_F32 = 1, T5_I64 = 2, T5_U8 = 3 };
enum { T5_AOS = 0, T5_SOA = 1 };

enum {
    T5_OK = 0,
    T5_ERR_SHAPE = 1,       /* length / shape mismatch  <-> TensorException WrongListLength (src/Data/Tensor.hs:96-100) */
    T5_ERR_UNSUPPORTED = 2, /* (dtype, shape, layout) has no device kernel; never falls back to the CPU */
    T5_ERR_CUDA = 3,
    T5_ERR_NCCL = 4,
    T5_ERR_OOM = 5,
    T5_ERR_INVALID = 6      /* NULL handle, buffers of different contexts / batch counts, aliasing output */
};

const char* t5_strerror(int code);
/* Last CUDA / NCCL error text seen by this context (empty string when none). */
const char* t5_last_error(t5_ctx* ctx);
/* Library build tag, e.g. "t5b200 sm_100a r2". */
const char* t5_version(void);

/* ---- context ------------------------------------------------------------- */
/* devices == NULL && ndev == 0: every visible device.  One stream per device. */
int t5_init(const int* devices, int ndev, t5_ctx** out);
/* Synchronises, then releases every device resource of the context - the device memory of
 * buffers that are still alive included.  Such buffers become ORPHANS: every operation on
 * them returns T5_ERR_INVALID, and t5_free on an orphan is legal at any later time, from any
 * thread, and only releases the handle (the last one also releases the context's host
 * bookkeeping).  This is the order a Haskell program produces: `bracket t5_init t5_shutdown`
 * ends before the garbage collector runs the ForeignPtr finalizers (&t5_free) of the
 * batches it created (SURVEY.md §8b Ownership).  The ctx pointer itself must not be used
 * after t5_shutdown; a second t5_shutdown while orphans exist returns T5_ERR_INVALID. */
int t5_shutdown(t5_ctx* ctx);
/* Buffers allocated from this context and not yet freed. */
int t5_live_buffers(t5_ctx* ctx);
int t5_ndev(t5_ctx* ctx);
int t5_sync(t5_ctx* ctx);

/* Batch partitioner (pure function, no device needed): shard `g` of `ndev` owns
 * items [first, first + count) of a batch of `batch` items.  Contiguous split,
 * per-device share rounded up to a multiple of 256 items. */
int t5_partition(size_t batch, int ndev, int g, size_t* first, size_t* count);

/* ---- buffers --------------------------------------------------------------
 * A buffer holds `batch` items of `elems_per_item` elements of `dtype`
 * (the flat `content` vector of src/Data/Tensor/Vector.hs:121-124, unboxed),
 * sharded by t5_partition.  Haskell owns it through a ForeignPtr whose
 * finalizer is t5_free. */
int t5_alloc(t5_ctx* ctx, int dtype, size_t elems_per_item, size_t batch, int layout, t5_buf** out);
/* Stream-ordered and non-blocking: work that still reads or writes the buffer finishes
 * first.  Thread-safe, sets its own device, legal after t5_shutdown (see there). */
int t5_free(t5_buf* buf);
/* host is AoS; bytes must equal batch * elems_per_item * sizeof(dtype), else
 * T5_ERR_SHAPE (fromList's length check, src/Data/Tensor/Vector.hs:362-366). */
int t5_upload(t5_buf* buf, const void* host, size_t bytes);
int t5_download(t5_buf* buf, void* host, size_t bytes); /* toList, Vector.hs:367 */
/* AoS <-> SoA re-layout on the device (dst and src differ only in layout). */
int t5_relayout(t5_buf* dst, const t5_buf* src);
void* t5_pinned_alloc(size_t bytes);
void t5_pinned_free(void* p);
/* introspection used by the tests */
size_t t5_buf_batch(const t5_buf* buf);
size_t t5_buf_elems(const t5_buf* buf);
int t5_buf_dtype(const t5_buf* buf);
int t5_buf_layout(const t5_buf* buf);

/* ---- operations -----------------------------------------------------------
 * All operands of one call must belong to the same context and have the same
 * batch count and layout. */

/* MatrixProduct (.*.): C[i,c] = fold_{j ascending} (acc + A[i,j]*B[j,c]), acc0 = 0,
 * multiply then add, never fused.  A is m x k, B is k x n, C is m x n.  n == 1 is
 * matrix . vector, m == 1 vector . matrix.  Int64 wraps.  (Absent from the
 * reference checkout — CHANGELOG.md:24-25; specified in SURVEY.md §8a a6.)
 * Results are bit-identical to that fold for every element type and shape, with two
 * exceptions that run on the tensor cores and meet the stated tolerance instead
 * (|delta| <= tol * sum_j |A[i,j]||B[j,c]|); t5_matmul_is_exact answers for a given shape:
 *   Double, m >= 17, n >= 24, k >= 8, n and k even, and a 32- / 48- / 64- / 96- / 128-tiling of
 *   the m x n result that is at least a quarter full (24 x 24 and 32 x 32 items, four to a tile, are;
 *   72 x 72 is; 17 x 136 is not): FP64 DMMA, tol 1e-12;
 *   Float, k and n multiples of 4: items of at least 33 x 36 with k >= 32 whose 64- or 128-tiling of the m x n
 *   result is at least 2/7 full (every item up to 64 x 64 is: two to a tile; m = n = 96 is; 200 x 36 is not), and
 *   items of 17..32 rows and 20..32 columns with k >= 8 that hold at least 2048 elements in A, B and C together
 *   (32 x 32 is, 24 x 24 is not; four to a tile): tcgen05 3xTF32, tol 1e-5; items containing inf / NaN are
 *   computed with the exact fold.
 * The class depends on dtype and shape only, never on layout, batch size or device count.
 * Layout: the tile pipeline (records up to 8 x 8), outer products and contractions whose B fits
 * in shared memory read SoA directly.  Every other shape has AoS kernels only; given SoA
 * operands the library re-lays the shard out into stream-ordered AoS temporaries, runs the AoS
 * kernel and re-lays the result out (one extra read + write pass per operand).  SoA is the
 * tiny-record layout: for records of more than ~64 elements prefer AoS buffers. */
int t5_matmul(t5_ctx* ctx, int dtype, int m, int k, int n, const t5_buf* A, const t5_buf* B, t5_buf* C);
/* 1 when t5_matmul / t5_prod / t5_matmul_host on an (m x k).(k x n) view of this dtype is bit-identical to the
 * fold above, 0 when it runs on the tensor cores (tolerance class above); a negated T5_ERR_* (-T5_ERR_SHAPE,
 * -T5_ERR_UNSUPPORTED) for arguments t5_matmul would reject.  Pure: no context, no GPU. */
int t5_matmul_is_exact(int dtype, int m, int k, int n);

/* DotProduct (dot): OUT[b] = fold_{i ascending} (acc + X[b,i]*Y[b,i]).  §8a a7. */
int t5_dot(t5_ctx* ctx, int dtype, int n, const t5_buf* X, const t5_buf* Y, t5_buf* OUT);

/* sum_b dot(X_b, Y_b) -> *host_scalar (of dtype).  The only entry with a
 * cross-device exchange: per-device partials are summed with a 1-element
 * ncclAllReduce(ncclSum) when NCCL can be loaded, else (and always when
 * ndev == 1) in fixed device order on the host.  Float/Double partials are
 * accumulated in double.  Synchronises. */
int t5_dot_sum(t5_ctx* ctx, int dtype, int n, const t5_buf* X, const t5_buf* Y, void* host_scalar);
/* 1 when the last multi-device t5_dot_sum went through NCCL, 0 host-sum. */
int t5_dot_sum_used_nccl(t5_ctx* ctx);

/* TensorProduct / prod n: contract the last `ncontract` indices of A with the
 * first `ncontract` of B (dims must agree, else T5_ERR_SHAPE).  ncontract == 0 is
 * the outer product, one multiply per output.  §8a a8. */
int t5_prod(t5_ctx* ctx, int dtype, int rankA, const uint64_t* dimsA, int rankB, const uint64_t* dimsB,
            int ncontract, const t5_buf* A, const t5_buf* B, t5_buf* C);

/* transpose = unRev . t0: full index reversal, shape reversed
 * (src/Data/Indexable.hs:101-102; src/Data/Tensor/Vector.hs:295-301, :306-307).
 * Pure permutation, bit-exact for every dtype. */
int t5_transpose(t5_ctx* ctx, int dtype, int rank, const uint64_t* dims, const t5_buf* A, t5_buf* OUT);

/* SquareMatrix det / inverse and LinearSystem solve by Gaussian elimination
 * with the first-non-zero pivot rule of SURVEY.md §8a (a9-a11).  Double and
 * Float only (T5_I64 -> T5_ERR_UNSUPPORTED: needs Fractional).  STATUS is a
 * T5_U8 buffer with one byte per item: 0 = Just, 1 = Nothing (singular; the
 * item's output is zero-filled).
 * Sizes: 1 <= n <= 128, 1 <= nrhs (any count; beyond 128 columns they are processed in
 * passes); n > 128 is T5_ERR_UNSUPPORTED.  n <= 4: one record per thread in registers;
 * 5 <= n <= 8: record per thread with streamed results (nrhs in {1, 2, 3, 4, n}) or one row
 * per lane (other nrhs); just beyond that still a record per thread (det / solve with one
 * right-hand side / inverse: Double up to n = 12 / 13 / 11, Float up to 18 / 20 / 14);
 * otherwise up to n = 128: one CTA per matrix, the matrix 2D-cyclic in the CTA's registers.  Every size keeps the oracle's operation order: results are bit-identical
 * to oracle/t5_oracle.cpp, not merely within tolerance.  SoA: n <= 8 (nrhs in {1, 2, 3, 4, n})
 * directly, everything else through the AoS detour described at t5_matmul. */
int t5_det(t5_ctx* ctx, int dtype, int n, const t5_buf* A, t5_buf* OUT);
int t5_inverse(t5_ctx* ctx, int dtype, int n, const t5_buf* A, t5_buf* OUT, t5_buf* STATUS);
int t5_solve(t5_ctx* ctx, int dtype, int n, int nrhs, const t5_buf* A, const t5_buf* B, t5_buf* X, t5_buf* STATUS);

/* The rest of the old SquareMatrix class (CHANGELOG.md:28-29 names polyEval and minPoly;
 * tr / charPoly per SURVEY.md §8f-4).  No reference code exists for them: the arithmetic is
 * specified in oracle/t5_oracle.cpp (parity unpinned).
 *   t5_trace     OUT[b] = fold_{i asc} (acc + A[b,i,i]) from 0.  All element types.
 *   t5_poly_eval OUT_b = c0 I + c1 A_b + .. + cd A_b^d by Horner (R = cd I; R = R.A + cj I);
 *                `coeffs` is a HOST array of ncoef elements of dtype, low order first, shared
 *                by the whole batch; ncoef == 0 gives the zero matrix.  All element types.
 *   t5_char_poly OUT has n+1 elements per item: the coefficients of det(xI - A), low order
 *                first, OUT[n] = 1 (Faddeev-LeVerrier).  Double / Float.
 *   t5_min_poly  OUT as for t5_char_poly, zero above the degree; DEGREE (T5_U8) receives the
 *                degree: the first exact linear dependency among I, A, A^2, ..; degree n
 *                returns the characteristic polynomial.  Double / Float, n <= 6.
 * Sizes: t5_trace, t5_poly_eval, t5_char_poly n <= 8 (n <= 4 in registers, any ncoef); t5_min_poly n <= 6;
 * larger n is T5_ERR_UNSUPPORTED.  SoA: n <= 4 directly, else through the AoS detour (t5_matmul). */
int t5_trace(t5_ctx* ctx, int dtype, int n, const t5_buf* A, t5_buf* OUT);
int t5_poly_eval(t5_ctx* ctx, int dtype, int n, const void* coeffs, int ncoef, const t5_buf* A, t5_buf* OUT);
int t5_char_poly(t5_ctx* ctx, int dtype, int n, const t5_buf* A, t5_buf* OUT);
int t5_min_poly(t5_ctx* ctx, int dtype, int n, const t5_buf* A, t5_buf* OUT, t5_buf* DEGREE);

/* rowEchelonForm / triangularSolve of the old Data.Tensor.LinearAlgebra (removed with it,
 * CHANGELOG.md:24-29; SURVEY.md §8f-4 "rowEchelonForm-style outputs"; parity unpinned, the
 * arithmetic is specified in oracle/t5_oracle.cpp row_echelon_item).  OUT receives the REDUCED
 * row echelon form of each rows x cols item of A: Gauss-Jordan with the first-non-zero pivot
 * rule and the quotient multipliers of det/inverse/solve (rank == n iff t5_inverse succeeds),
 * the pivot column cleared in every other row, then the pivot row divided by its pivot.  Pivots are searched in the first pivot_cols columns only: pivot_cols = cols
 * is rowEchelonForm; on an augmented [A | B] item with pivot_cols = cols(A) it is
 * triangularSolve (A reduced, B carried along).  RANK (T5_U8) receives the number of pivots.
 * Double / Float, rows <= 8, cols <= 16 (else T5_ERR_UNSUPPORTED); n x n and n x (n+1) with n <= 8 and
 * n x 2n with n <= 4 run in registers (AoS and SoA), other shapes in a shared-memory kernel (AoS; SoA
 * through the AoS detour). */
int t5_row_echelon(t5_ctx* ctx, int dtype, int rows, int cols, int pivot_cols, const t5_buf* A, t5_buf* OUT, t5_buf* RANK);

/* Element-wise vector-space operations (Functor/Applicative instances,
 * src/Data/Tensor/Vector.hs:180-190, :250-251: liftA2 (+), liftA2 (-), scalar *).
 * op: 0 = a + b, 1 = a - b, 2 = a * b (Hadamard); t5_scale multiplies by *scalar. */
int t5_zip(t5_ctx* ctx, int dtype, int op, const t5_buf* A, const t5_buf* B, t5_buf* C);
int t5_scale(t5_ctx* ctx, int dtype, const void* scalar, const t5_buf* A, t5_buf* C);
/* `pure` of the Applicative instance (src/Data/Tensor/Vector.hs:185-188), generalised to a whole
 * record: every item of OUT becomes a copy of `record`, a HOST array of OUT's per-item element
 * count.  A record of equal elements is `pure a`; the identity matrix is the old SquareMatrix
 * `unit`; zeros are the VectorSpace `zero`.  All element types, AoS and SoA. */
int t5_replicate(t5_ctx* ctx, int dtype, const void* record, t5_buf* OUT);


/* ---- structural operations ---------------------------------------------------
 * The reshuffles of the IsTensor / MultiIndexable / Sliceable instances
 * (src/Data/Tensor/Vector.hs:260-354, :400-463) as device gathers.  Each one is a
 * pure index map, bit-exact for every dtype, AoS (table-driven gather) and SoA (a copy of
 * whole rows).  Indices and axes are 1-based like the reference's MultiIndex
 * (src/Data/Tensor/Vector.hs:150-168). */

/* append n x y (Vector.hs:322-335): concatenate along axis `axis`; dimsA and dimsB agree
 * everywhere else. */
int t5_append(t5_ctx* ctx, int dtype, int axis, int rank, const uint64_t* dimsA, const uint64_t* dimsB,
              const t5_buf* A, const t5_buf* B, t5_buf* OUT);
/* split n t (Vector.hs:336-354): the inverse of append; OUT1 gets the first `first_extent`
 * entries of axis `axis`, OUT2 the rest. */
int t5_split(t5_ctx* ctx, int dtype, int axis, int rank, const uint64_t* dims, uint64_t first_extent,
             const t5_buf* IN, t5_buf* OUT1, t5_buf* OUT2);
/* t `at` ix (Vector.hs:277-283): fix every index but the first (index has rank-1 entries):
 * the column extractor.  OUT has dims[0] elements per item. */
int t5_at(t5_ctx* ctx, int dtype, int rank, const uint64_t* dims, const uint64_t* index, const t5_buf* A, t5_buf* OUT);
/* [i] `ta` t: fix the first index: the row extractor.  Follows the tested GADT semantics
 * (src/Data/Tensor.hs:245-247); Vector.hs:284-288 is off by one. */
int t5_ta(t5_ctx* ctx, int dtype, int rank, const uint64_t* dims, uint64_t index, const t5_buf* A, t5_buf* OUT);
/* slice t sl (Vector.hs:400-429, :461-463): slicer[a] == 0 keeps axis a (Nothing),
 * slicer[a] == k > 0 fixes it to k (Just k). */
int t5_slice(t5_ctx* ctx, int dtype, int rank, const uint64_t* dims, const uint64_t* slicer, const t5_buf* A, t5_buf* OUT);
/* General axis permutation: output axis a is input axis perm[a] (0-based).  concat,
 * unConcat, rev and unRev of nested tensors (Vector.hs:260-301) are axis permutations of
 * the flat [outer][inner] layout; transpose is the full reversal. */
int t5_permute_axes(t5_ctx* ctx, int dtype, int rank, const uint64_t* dims, const int* perm, const t5_buf* A, t5_buf* OUT);
/* Pure host function (no device): the gather table of a structural operation, i.e. for
 * every output element offset the input element offset it copies.  op: 0 append (params =
 * axis, dimsB[rank]; entries >= prod(dims) refer to B), 1 / 2 split first / second half
 * (params = axis, first_extent), 3 at (params = index[rank-1]), 4 ta (params = i),
 * 5 slice (params = slicer[rank]), 6 permute_axes (params = perm[rank]).
 * Writes at most table_cap entries; *d_out receives the table length. */
int t5_structural_table(int op, int rank, const uint64_t* dims, const uint64_t* params, int nparams,
                        uint32_t* table_out, size_t table_cap, size_t* d_out);

/* ---- host-buffer (end-to-end) entries --------------------------------------
 * The same operations on HOST AoS arrays (`batch` items each): the call uploads chunks,
 * runs the kernel and downloads results, overlapping the three on separate streams, and
 * returns when the results are in host memory.  This is what a Haskell instance method
 * on an ordinary (host) tensor batch calls; semantics, element types and error codes are
 * those of the device-buffer entry of the same name.  Pinned host memory
 * (t5_pinned_alloc) gives full PCIe speed; pageable memory works but is staged by the
 * driver.  `status` arrays hold one byte per item. */
int t5_matmul_host(t5_ctx* ctx, int dtype, int m, int k, int n, size_t batch,
                   const void* A, const void* B, void* C);
int t5_dot_host(t5_ctx* ctx, int dtype, int n, size_t batch, const void* X, const void* Y, void* OUT);
int t5_prod_host(t5_ctx* ctx, int dtype, int rankA, const uint64_t* dimsA, int rankB, const uint64_t* dimsB,
                 int ncontract, size_t batch, const void* A, const void* B, void* C);
int t5_transpose_host(t5_ctx* ctx, int dtype, int rank, const uint64_t* dims, size_t batch, const void* A, void* OUT);
int t5_det_host(t5_ctx* ctx, int dtype, int n, size_t batch, const void* A, void* OUT);
int t5_inverse_host(t5_ctx* ctx, int dtype, int n, size_t batch, const void* A, void* OUT, uint8_t* status);
int t5_solve_host(t5_ctx* ctx, int dtype, int n, int nrhs, size_t batch, const void* A, const void* B, void* X,
                  uint8_t* status);

/* ---- measurement helpers ---------------------------------------------------
 * CUDA events on the context's own streams (the streams the kernels run on);
 * t5_timer_end synchronises and returns the MAX over devices in milliseconds. */
int t5_timer_begin(t5_ctx* ctx);
int t5_timer_end(t5_ctx* ctx, float* ms_max);
/* After t5_timer_end: the elapsed time of each device's own stream (ms[g], g < min(cap, ndev)) -
 * the per-device view of a sharded call (strong scaling: every device should take 1/ndev). */
int t5_timer_per_device(t5_ctx* ctx, float* ms, int cap);
/* Overwrite a > L2-sized scratch buffer on every device (L2 flush between timed
 * iterations of small workloads). */
int t5_flush_l2(t5_ctx* ctx);
/* Counter-based synthetic inputs generated on the device:
 * element i of the GLOBAL batch stream = f(splitmix64((seed ^ (op_id << 56)) + i)). */
int t5_fill_synthetic(t5_buf* buf, uint64_t seed, uint64_t op_id);
/* Every swap_every-th matrix gets A[1,1] := 0, every sing_every-th row 2 := row 1. */
int t5_plant_specials(t5_buf* buf, int n, uint64_t swap_every, uint64_t sing_every);
/* Number of kernels this context has launched since t5_init. */
uint64_t t5_launch_count(t5_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* T5B200_H */
