// Internal interface between the C-ABI host layer (t5_host.cu) and the kernel
// translation units.  Not installed; the public boundary is include/t5b200.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#define T5_MAX_DEVICES 16

namespace t5 {

enum { T5D_F64 = 0, T5D_F32 = 1, T5D_I64 = 2, T5D_U8 = 3 };
enum { T5L_AOS = 0, T5L_SOA = 1 };
// internal launcher results
enum { T5I_OK = 0, T5I_NOT_SPECIALISED = -1, T5I_CUDA = -2, T5I_UNSUPPORTED = -3 };

inline int dtype_size(int dtype) {
    switch (dtype) {
        case T5D_F64: case T5D_I64: return 8;
        case T5D_F32: return 4;
        case T5D_U8: return 1;
    }
    return 0;
}

// One device's slice of a batched call.
struct Shard {
    cudaStream_t stream;
    int layout;         // T5L_AOS / T5L_SOA
    size_t count;       // items in this shard
    size_t soa_stride;  // SoA element stride (shard capacity, items)
    size_t first;       // global index of the shard's first item
    unsigned long long* sched;  // this stream's tile-scheduler words {next tile, finished CTAs}
};

// ---- t5_small_ops.cu: register-resident thread-per-item kernels -----------------
int small_matmul(const Shard& s, int dtype, int m, int k, int n, const void* A, const void* B, void* C);
int small_outer(const Shard& s, int dtype, int p, int q, const void* A, const void* B, void* C);   // tiny outer products (tile pipeline)
int small_transpose(const Shard& s, int elem_bytes, int r, int c, const void* A, void* O);
int small_det(const Shard& s, int dtype, int n, const void* A, void* O);
// record-per-thread eliminations just beyond 8x8 (t5_small_ops_9.cu): Double up to 12 / 13 / 11 (det / solve / inverse), Float 18 / 20 / 14; AoS
int reg9_det(const Shard& s, int dtype, int n, const void* A, void* O);
int reg9_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status);
int reg9_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status);
int small_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status);
int small_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status);

int small_trace(const Shard& s, int dtype, int n, const void* A, void* O);
int small_poly_eval(const Shard& s, int dtype, int n, const void* coeffs_host, int ncoef, const void* A, void* O);
int small_char_poly(const Shard& s, int dtype, int n, const void* A, void* O);
int small_min_poly(const Shard& s, int dtype, int n, const void* A, void* O, void* degree);
// t5_solve_many.cu: matrix right-hand sides (2, 3, 4 or n columns) for 5 <= n <= 8, AoS and SoA
int many_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status);
// t5_echelon.cu: rowEchelonForm in registers (n x n, n x (n+1) for n = 2..8, n x 2n for n = 2..4), AoS and SoA
int small_row_echelon(const Shard& s, int dtype, int rows, int cols, int pivot_cols, const void* A, void* O, void* rank);

// ---- t5_mid_ops.cu: row-per-thread contractions for 0.25-4 KiB records (8x8, 16x16, ...) -----
int mid_matmul(const Shard& s, int dtype, int m, int k, int n, const void* A, const void* B, void* C);

// ---- t5_generic.cu: any-shape kernels (AoS only) ---------------------------------
// C[P x Q] = A[P x K] . B[K x Q] per item; ncontract == 0 -> single multiply.
int generic_stagedprod(const Shard& s, int dtype, int m, int k, int n, const void* A, const void* B, void* C);
int generic_longdot(const Shard& s, int dtype, int K, const void* X, const void* Y, void* O);
int generic_prod(const Shard& s, int dtype, int P, int K, int Q, int ncontract, const void* A, const void* B, void* C);
// out[b][i] = in[b][perm[i]], perm resident on the device (d entries)
int generic_transpose2d(const Shard& s, int elem_bytes, int R, int C, const void* A, void* O);
int generic_soa_gather(const Shard& s, int elem_bytes, size_t dA, size_t d_out, const uint32_t* table_dev, const void* A, const void* B, void* O);
int generic_permute(const Shard& s, int elem_bytes, size_t d, const uint32_t* perm_dev, const void* A, void* O);
// out[b][i] = in[b][table[i]] with d_in != d_out allowed (structural gathers); gather2: two sources (append)
int generic_gather(const Shard& s, int elem_bytes, size_t d_in, size_t d_out, const uint32_t* table_dev, const void* A, void* O);
int generic_gather2(const Shard& s, int elem_bytes, size_t dA, size_t dB, size_t d_out, const uint32_t* table_dev, const void* A, const void* B, void* O);
// gathers of staged records, up to two sources and two results (t2 / O2 may be null with d2 = 0); T5I_NOT_SPECIALISED when the
// records are too large to stage
bool staged_gather_fits(int elem_bytes, size_t d_in_total, size_t d_out_total, bool two_results);
int staged_gather(const Shard& s, int elem_bytes, size_t dA, size_t dB, size_t d1, size_t d2, const uint32_t* t1, const uint32_t* t2,
                  const void* A, const void* B, void* O1, void* O2);
int generic_scatter2(const Shard& s, int elem_bytes, size_t d_in, size_t d1, size_t d2, const uint32_t* table_dev, const void* A, void* O1, void* O2);
int generic_det(const Shard& s, int dtype, int n, const void* A, void* O);
int generic_solve(const Shard& s, int dtype, int n, int nrhs, bool identity_rhs, const void* A, const void* B, void* X, void* status);
// SquareMatrix extras for any n <= 8 (min_poly: n <= 6); coeffs_dev holds ncoef elements of dtype on the device
int generic_trace(const Shard& s, int dtype, int n, const void* A, void* O);
int generic_poly_eval(const Shard& s, int dtype, int n, const void* coeffs_dev, int ncoef, const void* A, void* O);
int generic_char_poly(const Shard& s, int dtype, int n, const void* A, void* O);
int generic_min_poly(const Shard& s, int dtype, int n, const void* A, void* O, void* degree);
int generic_row_echelon(const Shard& s, int dtype, int rows, int cols, int pivot_cols, const void* A, void* O, void* rank);
int generic_zip(const Shard& s, int dtype, int op, size_t elems, const void* A, const void* B, void* C);
int generic_replicate(const Shard& s, int elem_bytes, size_t d, const void* rec_dev, void* O);
int generic_scale(const Shard& s, int dtype, const void* scalar_host, size_t elems, const void* A, void* C);
// per-device partial of sum_b dot(x_b, y_b): *partial_dev is double (F64/F32) or int64
// (SoA: rows of row_stride items, the first row_valid are data; AoS: row_stride = 0)
int dot_sum_partial(const Shard& s, int dtype, size_t elems, size_t row_stride, size_t row_valid, const void* X, const void* Y, void* scratch_dev, void* partial_dev);
int fill_synthetic(const Shard& s, int dtype, size_t elems_per_item, uint64_t seed, uint64_t op_id, void* dst);
int plant_specials(const Shard& s, int dtype, int n, uint64_t swap_every, uint64_t sing_every, void* A);
// AoS [count][d] <-> SoA [d][stride]
int relayout(cudaStream_t st, int elem_bytes, size_t d, size_t count, size_t soa_stride, bool to_soa, const void* src, void* dst);

// ---- t5_dgemm.cu: FP64 tensor-core (DMMA) batched product: items from 17 x 24 up with k, n multiples of 8 -------
// returns T5I_NOT_SPECIALISED unless dmma_takes(m, k, n); within north_star's 1e-12 Double tolerance, not bit-exact
bool dmma_takes(int m, int k, int n, int* geo_out = nullptr);
int dmma_matmul(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C);

// ---- t5_tf32x3.cu: Float products on tcgen05 tensor cores (3xTF32 split, TMEM accumulators) for shapes
// on the 128 x 128 x 32 tile grid; within north_star's 1e-5 Float tolerance, not bit-exact
bool tf32x3_takes(int m, int k, int n, int* geo_out = nullptr);
int tf32x3_matmul(const Shard& s, int m, int k, int n, const void* A, const void* B, void* C);

// ---- t5_rowelim.cu: row-per-lane (8 lanes per matrix, warp shuffles) elimination for 5 <= n <= 8 -----
int rowlane_det(const Shard& s, int dtype, int n, const void* A, void* O);
int rowlane_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status);
int rowlane_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status);

// ---- t5_lu.cuh (t5_lu_det.cu / t5_lu_solve.cu / t5_lu_inverse.cu): CTA-per-matrix elimination for 9 <= n <= 128, the matrix 2D-cyclic in the CTA's registers (AoS);
// bit-identical to the oracle like every other elimination kernel
int lu_det(const Shard& s, int dtype, int n, const void* A, void* O);
// ---- t5_warp_lu.cuh (t5_warp_lu_det.cu / _solve.cu / _inverse.cu): a matrix row per lane, shuffles only (AoS): Float
// 9 <= n <= 16 and 30 <= n <= 32; T5I_NOT_SPECIALISED otherwise (Double, other sizes)
int warp_det(const Shard& s, int dtype, int n, const void* A, void* O);
int warp_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status);
int warp_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status);
int lu_inverse(const Shard& s, int dtype, int n, const void* A, void* O, void* status);
int lu_solve(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* status);

// ---- t5_xgemm.cu: exact (oracle-order, bit-identical) register-tiled product for shapes without a
// faster specialisation: large Float / Int64 matrices, Double shapes off the DMMA tile grid
int exact_blocked_matmul(const Shard& s, int dtype, int P, int K, int Q, const void* A, const void* B, void* C);

}  // namespace t5
