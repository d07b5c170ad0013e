// C++ host-mirror test: drives tensor_b200/cxx/tensor_b200.hpp (type-indexed Batch API over
// the C ABI) on a GPU and checks every result bit-for-bit against the CPU oracle
// (oracle/libt5oracle.so).  Reads like the reference's own property tests
// (tests/Tensor.hs): build values, apply the class methods, compare.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../tensor_b200/cxx/tensor_b200.hpp"

extern "C" {
int t5o_fill(int dtype, uint64_t seed, uint64_t op_id, uint64_t first_elem, size_t count, void* out);
int t5o_matmul(int dtype, int m, int k, int n, size_t batch, const void* A, const void* B, void* C, int threads);
int t5o_dot(int dtype, int n, size_t batch, const void* X, const void* Y, void* OUT, int threads);
int t5o_transpose(int elem_bytes, int rank, const uint64_t* dims, size_t batch, const void* A, void* OUT, int threads);
int t5o_prod(int dtype, int rankA, const uint64_t* dimsA, int rankB, const uint64_t* dimsB, int nc, size_t batch,
             const void* A, const void* B, void* C, int threads);
int t5o_det(int dtype, int n, size_t batch, const void* A, void* OUT, int threads);
int t5o_inverse(int dtype, int n, size_t batch, const void* A, void* OUT, uint8_t* status, int threads);
int t5o_solve(int dtype, int n, int nrhs, size_t batch, const void* A, const void* B, void* X, uint8_t* status, int threads);
int t5o_trace(int dtype, int n, size_t batch, const void* A, void* OUT, int threads);
int t5o_poly_eval(int dtype, int n, const void* coeffs, int ncoef, size_t batch, const void* A, void* OUT, int threads);
int t5o_char_poly(int dtype, int n, size_t batch, const void* A, void* OUT, int threads);
int t5o_min_poly(int dtype, int n, size_t batch, const void* A, void* OUT, uint8_t* degree, int threads);
int t5o_row_echelon(int dtype, int m, int n, int pivot_cols, size_t batch, const void* A, void* OUT, uint8_t* rank, int threads);
}

static int failures = 0;
template <class T>
static void expect_bits(const char* what, const std::vector<T>& got, const std::vector<T>& want) {
    bool ok = got.size() == want.size() && (got.empty() || !std::memcmp(got.data(), want.data(), got.size() * sizeof(T)));
    std::printf("%-40s %s\n", what, ok ? "ok" : "MISMATCH");
    if (!ok) ++failures;
}
template <class T> static std::vector<T> synth(int dt, uint64_t seed, uint64_t op, size_t n) {
    std::vector<T> v(n);
    t5o_fill(dt, seed, op, 0, n, v.data());
    return v;
}

int main() {
    using namespace t5;
    const uint64_t SA = 0x5EED000000000001ull, SB = 0x5EED000000000002ull;
    Context ctx({0});
    const size_t n = 4099;

    {   // (.*.) on 3x3 Double and Int64, matrix . vector on 4x4 Float
        auto a = synth<double>(0, SA, 1, n * 9), b = synth<double>(0, SB, 1, n * 9);
        auto A = Batch<double, 3, 3>::from_list(ctx, n, a), B = Batch<double, 3, 3>::from_list(ctx, n, b);
        std::vector<double> want(n * 9);
        t5o_matmul(0, 3, 3, 3, n, a.data(), b.data(), want.data(), 1);
        expect_bits("Matrix 3 3 Double .*.", (A * B).to_list(), want);

        auto ai = synth<int64_t>(2, SA, 1, n * 9), bi = synth<int64_t>(2, SB, 1, n * 9);
        std::vector<int64_t> wi(n * 9);
        t5o_matmul(2, 3, 3, 3, n, ai.data(), bi.data(), wi.data(), 1);
        expect_bits("Matrix 3 3 Int64 .*. (wrap-around)",
                    (Batch<int64_t, 3, 3>::from_list(ctx, n, ai) * Batch<int64_t, 3, 3>::from_list(ctx, n, bi)).to_list(), wi);

        auto m = synth<float>(1, SA, 2, n * 16), x = synth<float>(1, SB, 2, n * 4);
        std::vector<float> wy(n * 4), wd(n);
        t5o_matmul(1, 4, 4, 1, n, m.data(), x.data(), wy.data(), 1);
        auto M = Batch<float, 4, 4>::from_list(ctx, n, m);
        auto X = Batch<float, 4>::from_list(ctx, n, x);
        Batch<float, 4> Y = M * X;
        expect_bits("Matrix 4 4 Float .*. Vector 4", Y.to_list(), wy);
        t5o_dot(1, 4, n, x.data(), wy.data(), wd.data(), 1);
        expect_bits("dot (Vector 4 Float)", dot(X, Y).to_list(), wd);
    }
    {   // transpose: shape reversed at the type level, (A^T)^T == A
        auto a = synth<double>(0, SA, 3, n * 24);
        auto A = Batch<double, 2, 3, 4>::from_list(ctx, n, a);
        Batch<double, 4, 3, 2> T = transpose(A);
        std::vector<double> want(n * 24);
        const uint64_t d[3] = {2, 3, 4};
        t5o_transpose(8, 3, d, n, a.data(), want.data(), 1);
        expect_bits("transpose (Tensor [2,3,4])", T.to_list(), want);
        expect_bits("transpose . transpose == id", transpose(T).to_list(), a);
    }
    {   // tensor product
        auto a = synth<double>(0, SA, 4, n * 6), b = synth<double>(0, SB, 4, n * 4);
        Batch<double, 2, 3, 2, 2> T = tensor_product(Batch<double, 2, 3>::from_list(ctx, n, a), Batch<double, 2, 2>::from_list(ctx, n, b));
        std::vector<double> want(n * 24);
        const uint64_t da[2] = {2, 3}, db[2] = {2, 2};
        t5o_prod(0, 2, da, 2, db, 0, n, a.data(), b.data(), want.data(), 1);
        expect_bits("tensor product [2,3] (x) [2,2]", T.to_list(), want);
    }
    {   // det / inverse / solve with planted swaps and singular items
        auto a = synth<double>(0, SA, 5, n * 16), b = synth<double>(0, SB, 5, n * 4);
        for (size_t i = 0; i < n; i += 16) a[i * 16] = 0.0;
        for (size_t i = 0; i < n; i += 128) std::memcpy(&a[i * 16 + 4], &a[i * 16], 4 * sizeof(double));
        auto A = Batch<double, 4, 4>::from_list(ctx, n, a);
        std::vector<double> wd(n), wi(n * 16), wx(n * 4);
        std::vector<uint8_t> ws(n), ws2(n);
        t5o_det(0, 4, n, a.data(), wd.data(), 1);
        expect_bits("det (Matrix 4 4 Double)", det(A).to_list(), wd);
        t5o_inverse(0, 4, n, a.data(), wi.data(), ws.data(), 1);
        auto [inv, st] = inverse(A);
        expect_bits("inverse", inv.to_list(), wi);
        expect_bits("inverse status (Maybe)", st.to_list(), ws);
        t5o_solve(0, 4, 1, n, a.data(), b.data(), wx.data(), ws2.data(), 1);
        auto [x, st2] = solve(A, Batch<double, 4>::from_list(ctx, n, b));
        expect_bits("solve", x.to_list(), wx);
        expect_bits("solve status", st2.to_list(), ws2);
    }
    {   // the rest of SquareMatrix: tr, polyEval, charPoly, minPoly
        auto a = synth<double>(0, SA, 7, n * 9);
        for (size_t i = 0; i < n; i += 7) {              // exact structure: 2 I has minimal polynomial x - 2
            for (int e = 0; e < 9; ++e) a[i * 9 + e] = (e % 4 == 0) ? 2.0 : 0.0;
        }
        auto A = Batch<double, 3, 3>::from_list(ctx, n, a);
        std::vector<double> wt(n), wp(n * 9), wc(n * 4), wm(n * 4);
        std::vector<uint8_t> wdeg(n);
        const std::vector<double> coeffs = {0.5, -1.25, 2.0, 0.75};
        t5o_trace(0, 3, n, a.data(), wt.data(), 1);
        expect_bits("tr (Matrix 3 3 Double)", tr(A).to_list(), wt);
        t5o_poly_eval(0, 3, coeffs.data(), 4, n, a.data(), wp.data(), 1);
        expect_bits("polyEval", poly_eval(A, coeffs).to_list(), wp);
        t5o_char_poly(0, 3, n, a.data(), wc.data(), 1);
        Batch<double, 4> cp = char_poly(A);
        expect_bits("charPoly", cp.to_list(), wc);
        t5o_min_poly(0, 3, n, a.data(), wm.data(), wdeg.data(), 1);
        auto [mp, deg] = min_poly(A);
        expect_bits("minPoly", mp.to_list(), wm);
        expect_bits("minPoly degree", deg.to_list(), wdeg);
    }
    {   // pure / unit: A .*. unit == A .*. I bit for bit, pure a is constant
        auto a = synth<double>(0, SA, 10, n * 9);
        auto A = Batch<double, 3, 3>::from_list(ctx, n, a);
        auto U = unit<double, 3>(ctx, n);
        std::vector<double> eye(n * 9, 0.0), want(n * 9);
        for (size_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) eye[i * 9 + d * 4] = 1.0;
        expect_bits("unit", U.to_list(), eye);
        t5o_matmul(0, 3, 3, 3, n, a.data(), eye.data(), want.data(), 1);
        expect_bits("A .*. unit", (A * U).to_list(), want);
        expect_bits("pure 2.5", Batch<double, 2, 2>::pure(ctx, n, 2.5).to_list(), std::vector<double>(n * 4, 2.5));
    }
    {   // rowEchelonForm on a 3 x 4 batch (every 5th item has an exactly repeated row), all columns / first 3 eligible
        auto a = synth<double>(0, SA, 9, n * 12);
        for (size_t i = 0; i < n; i += 5)
            for (int j = 0; j < 4; ++j) a[i * 12 + 8 + j] = a[i * 12 + j];
        auto A = Batch<double, 3, 4>::from_list(ctx, n, a);
        std::vector<double> we(n * 12);
        std::vector<uint8_t> wr(n);
        t5o_row_echelon(0, 3, 4, 4, n, a.data(), we.data(), wr.data(), 1);
        auto [e, rk] = row_echelon_form(A);
        expect_bits("rowEchelonForm (Matrix 3 4 Double)", e.to_list(), we);
        expect_bits("rowEchelonForm rank", rk.to_list(), wr);
        t5o_row_echelon(0, 3, 4, 3, n, a.data(), we.data(), wr.data(), 1);
        auto [e3, rk3] = row_echelon_form(A, 3);
        expect_bits("triangularSolve ([A | b], 3 pivot columns)", e3.to_list(), we);
        expect_bits("triangularSolve rank", rk3.to_list(), wr);
    }
    {   // prod n, vector-space ops, structural ops, host-array entries
        auto a = synth<double>(0, SA, 8, n * 6), b = synth<double>(0, SB, 8, n * 12);
        auto A = Batch<double, 2, 3>::from_list(ctx, n, a);
        auto B = Batch<double, 3, 4>::from_list(ctx, n, b);
        Batch<double, 2, 4> P = prod<1>(A, B);                      // [2,3] x [3,4] contracted over 1 index
        std::vector<double> wp(n * 8);
        t5o_matmul(0, 2, 3, 4, n, a.data(), b.data(), wp.data(), 1);
        expect_bits("prod 1 == .*. ([2,3] x [3,4])", P.to_list(), wp);
        expect_bits("prod 1 == operator*", (A * B).to_list(), wp);
        Batch<double, 2, 3, 3, 4> O = prod<0>(A, B);
        std::vector<double> wo(n * 72);
        const uint64_t da[2] = {2, 3}, db[2] = {3, 4};
        t5o_prod(0, 2, da, 2, db, 0, n, a.data(), b.data(), wo.data(), 1);
        expect_bits("prod 0 == tensor product", O.to_list(), wo);

        auto c = synth<double>(0, SB, 9, n * 6);
        auto C = Batch<double, 2, 3>::from_list(ctx, n, c);
        std::vector<double> ws(n * 6), wd(n * 6), wh(n * 6), wk(n * 6);
        for (size_t i = 0; i < n * 6; ++i) { ws[i] = a[i] + c[i]; wd[i] = a[i] - c[i]; wh[i] = a[i] * c[i]; wk[i] = 3.0 * a[i]; }
        expect_bits("liftA2 (+)", (A + C).to_list(), ws);
        expect_bits("liftA2 (-)", (A - C).to_list(), wd);
        expect_bits("liftA2 (*)", hadamard(A, C).to_list(), wh);
        expect_bits("fmap (3 *)", (3.0 * A).to_list(), wk);

        // [A | col]: append along axis 2, then split it back (tests/Tensor.hs:25-29)
        auto col = synth<double>(0, SB, 10, n * 2);
        auto Col = Batch<double, 2, 1>::from_list(ctx, n, col);
        Batch<double, 2, 4> Aug = append<2>(A, Col);
        std::vector<double> wa(n * 8);
        for (size_t i = 0; i < n; ++i)
            for (int r = 0; r < 2; ++r) {
                for (int j = 0; j < 3; ++j) wa[i * 8 + r * 4 + j] = a[i * 6 + r * 3 + j];
                wa[i * 8 + r * 4 + 3] = col[i * 2 + r];
            }
        expect_bits("append 2 [2,3] [2,1]", Aug.to_list(), wa);
        auto [L, R] = split<2, 3>(Aug);
        expect_bits("split . append == id (left)", L.to_list(), a);
        expect_bits("split . append == id (right)", R.to_list(), col);
        Batch<double, 3> row2 = ta(2, A);
        Batch<double, 2> col3 = at(A, {3});
        std::vector<double> wr(n * 3), wc(n * 2);
        for (size_t i = 0; i < n; ++i) {
            for (int j = 0; j < 3; ++j) wr[i * 3 + j] = a[i * 6 + 3 + j];
            for (int r = 0; r < 2; ++r) wc[i * 2 + r] = a[i * 6 + r * 3 + 2];
        }
        expect_bits("[2] `ta` t (row 2)", row2.to_list(), wr);
        expect_bits("t `at` [3] (column 3)", col3.to_list(), wc);
        expect_bits("slice [Nothing, Just 3] == at [3]", slice<2>(A, {0, 3}).to_list(), wc);

        // the numerical contract is queryable without a device: tiny products are the exact fold, large Double / Float ones are not
        if (!product_is_exact<double, 3, 3, 3>() || !product_is_exact<int64_t, 128, 128, 128>() ||
            product_is_exact<double, 128, 128, 128>() || product_is_exact<float, 64, 64, 64>()) {
            std::printf("FAIL product_is_exact\n");
            ++failures;
        }
        std::vector<double> hc(n * 8), hinv(n * 16), hi_want(n * 16);
        matmul_host<double, 2, 3, 4>(ctx, n, a.data(), b.data(), hc.data());
        expect_bits("matmul_host (host arrays, one call)", hc, wp);
        auto m4 = synth<double>(0, SA, 11, n * 16);
        std::vector<uint8_t> hst(n), wst(n);
        inverse_host<double, 4>(ctx, n, m4.data(), hinv.data(), hst.data());
        t5o_inverse(0, 4, n, m4.data(), hi_want.data(), wst.data(), 1);
        expect_bits("inverse_host", hinv, hi_want);
        expect_bits("inverse_host status", hst, wst);
    }
    {   // ownership: a finalizer may run t5_free AFTER t5_shutdown.  The buffers become orphans:
        // operations on them are refused, a late free only releases the handle, nothing touches freed memory.
        t5_ctx* h = nullptr;
        const int dev0 = 0;
        bool ok = t5_init(&dev0, 1, &h) == T5_OK;
        t5_buf *a = nullptr, *b = nullptr;
        ok = ok && t5_alloc(h, T5_F64, 9, 1000, T5_AOS, &a) == T5_OK && t5_alloc(h, T5_F64, 9, 1000, T5_SOA, &b) == T5_OK;
        ok = ok && t5_live_buffers(h) == 2;
        ok = ok && t5_free(b) == T5_OK && t5_live_buffers(h) == 1;
        ok = ok && t5_shutdown(h) == T5_OK;                  // one buffer still alive
        std::vector<double> host(9000);
        ok = ok && t5_download(a, host.data(), host.size() * 8) == T5_ERR_INVALID;   // orphan: refused, no crash
        ok = ok && t5_shutdown(h) == T5_ERR_INVALID;         // struct still alive because of the orphan: detected
        ok = ok && t5_free(a) == T5_OK;                      // the last orphan takes the context with it
        std::printf("%-40s %s\n", "t5_free after t5_shutdown (orphans)", ok ? "ok" : "FAILED");
        if (!ok) ++failures;
        // the same through the C++ mirror: a Batch (shared_ptr deleter = t5_free) that outlives its Context
        std::unique_ptr<Batch<double, 3, 3>> late;
        {
            Context inner({0});
            late.reset(new Batch<double, 3, 3>(Batch<double, 3, 3>::pure(inner, 100, 1.5)));
        }
        late.reset();
        std::printf("%-40s ok\n", "Batch destroyed after its Context");
    }
    {   // fromList length check
        bool threw = false;
        try { Batch<double, 3, 3>::from_list(ctx, 2, std::vector<double>(17)); } catch (const WrongListLength&) { threw = true; }
        std::printf("%-40s %s\n", "fromList wrong length -> WrongListLength", threw ? "ok" : "MISSING");
        if (!threw) ++failures;
        bool unsup = false;
        try { det(Batch<int64_t, 3, 3>(ctx, 4)); } catch (const Unsupported&) { unsup = true; }
        std::printf("%-40s %s\n", "det on Int64 -> Unsupported", unsup ? "ok" : "MISSING");
        if (!unsup) ++failures;
    }
    std::printf(failures ? "FAILED (%d)\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}
