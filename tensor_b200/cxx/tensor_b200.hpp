// tensor_b200.hpp — C++17 host-side mirror of the reference's type-indexed class API over
// the C ABI (include/t5b200.h).  The reference is compiled Haskell whose shapes live in the
// type (`Tensor (is :: [PI]) e`, src/Data/Tensor/Vector.hs:121-124; MultiIndex.hs:46-61);
// with no GHC in this image the same static discipline is restated with C++ templates:
//
//   Batch<double, 3, 3> a = Batch<double, 3, 3>::from_list(ctx, n, values);   // fromList
//   auto c  = a * b;                    // (.*.)  : Batch<T,I,J> * Batch<T,J,K> -> Batch<T,I,K>
//   auto y  = a * x;                    //          Batch<T,I,J> * Batch<T,J>   -> Batch<T,I>
//   auto s  = dot(x, y);                // dot    : Batch<T,N> x Batch<T,N>     -> Batch<T>
//   auto t  = tensor_product(a, b);     // (⊗)    : Batch<T,Is...> ⊗ Batch<T,Js...> -> Batch<T,Is...,Js...>
//   auto at = transpose(a);             // transpose = unRev . t0 (Indexable.hs:101-102)
//   auto d  = det(a);  auto [inv, st] = inverse(a);  auto [x, st] = solve(a, b);
//   auto p  = prod<1>(a, b);            // prod n : contract the last n indices of a with the first n of b
//   auto z  = append<2>(a, col);        // append / split / at / ta / slice (Vector.hs:277-354, :400-463)
//   matmul_host<double, 3, 3, 3>(ctx, n, pa, pb, pc);   // ordinary host arrays in and out, one call
//
// A shape error is a COMPILE error, as in Haskell; a wrong list length throws
// WrongListLength (src/Data/Tensor.hs:96-100); an unsupported request throws Unsupported
// (there is no CPU fallback).  Header-only; link with libt5b200.so.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/t5b200.h"

namespace t5 {

struct TensorException : std::runtime_error { using std::runtime_error::runtime_error; };
struct WrongListLength : TensorException { WrongListLength() : TensorException("WrongListLength") {} };
struct Unsupported : TensorException { using TensorException::TensorException; };
struct DeviceError : std::runtime_error { using std::runtime_error::runtime_error; };

template <class T> struct dtype_of;
template <> struct dtype_of<double> { static constexpr int value = T5_F64; };
template <> struct dtype_of<float> { static constexpr int value = T5_F32; };
template <> struct dtype_of<int64_t> { static constexpr int value = T5_I64; };
template <> struct dtype_of<uint8_t> { static constexpr int value = T5_U8; };

class Context {
  public:
    explicit Context(const std::vector<int>& devices = {}) {
        int rc = t5_init(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &h_);
        if (rc != T5_OK) throw DeviceError(std::string("t5_init: ") + t5_strerror(rc));
    }
    ~Context() { if (h_) t5_shutdown(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    t5_ctx* handle() const { return h_; }
    int ndev() const { return t5_ndev(h_); }
    void sync() const { check(t5_sync(h_), "t5_sync"); }
    void check(int rc, const char* what) const {
        if (rc == T5_OK) return;
        if (rc == T5_ERR_SHAPE) throw WrongListLength();
        if (rc == T5_ERR_UNSUPPORTED) throw Unsupported(std::string(what) + ": " + t5_strerror(rc));
        throw DeviceError(std::string(what) + ": " + t5_strerror(rc) + " [" + t5_last_error(h_) + "]");
    }
  private:
    t5_ctx* h_ = nullptr;
};

template <size_t... Is> struct Shape {
    static constexpr size_t rank = sizeof...(Is);
    static constexpr size_t size = (size_t(1) * ... * Is);
    static std::vector<uint64_t> dims() { return {uint64_t(Is)...}; }
};

// `count` tensors of shape Is... and element type T on the context's devices.
template <class T, size_t... Is>
class Batch {
  public:
    using shape = Shape<Is...>;
    Batch(const Context& ctx, size_t count) : ctx_(&ctx), count_(count) {
        t5_buf* b = nullptr;
        ctx.check(t5_alloc(ctx.handle(), dtype_of<T>::value, shape::size, count, T5_AOS, &b), "t5_alloc");
        buf_ = std::shared_ptr<t5_buf>(b, [](t5_buf* p) { t5_free(p); });   // the ForeignPtr finalizer
    }
    // fromList (src/Data/Tensor/Vector.hs:362-366): row-major, length checked
    static Batch from_list(const Context& ctx, size_t count, const std::vector<T>& xs) {
        if (xs.size() != count * shape::size) throw WrongListLength();
        Batch b(ctx, count);
        ctx.check(t5_upload(b.buf(), xs.data(), xs.size() * sizeof(T)), "t5_upload");
        return b;
    }
    // every item a copy of one record; a record of equal elements is `pure a` (Vector.hs:185-188)
    static Batch replicate(const Context& ctx, size_t count, const std::vector<T>& record) {
        if (record.size() != shape::size) throw WrongListLength();
        Batch b(ctx, count);
        ctx.check(t5_replicate(ctx.handle(), dtype_of<T>::value, record.data(), b.buf()), "t5_replicate");
        return b;
    }
    static Batch pure(const Context& ctx, size_t count, T a) { return replicate(ctx, count, std::vector<T>(shape::size, a)); }
    // toList (Vector.hs:367)
    std::vector<T> to_list() const {
        std::vector<T> out(count_ * shape::size);
        ctx_->check(t5_download(buf(), out.data(), out.size() * sizeof(T)), "t5_download");
        return out;
    }
    size_t count() const { return count_; }
    t5_buf* buf() const { return buf_.get(); }
    const Context& ctx() const { return *ctx_; }
  private:
    const Context* ctx_;
    size_t count_;
    std::shared_ptr<t5_buf> buf_;
};

// Is (.*.) of an I x J by a J x K matrix of this element type the bit-exact sequential fold (true), or does the library run it
// on the tensor cores within BASELINE.json's tolerance (include/t5b200.h, t5_matmul)?  Pure: shape and element type only.
template <class T, size_t I, size_t J, size_t K>
inline bool product_is_exact() { return t5_matmul_is_exact(dtype_of<T>::value, (int)I, (int)J, (int)K) == 1; }

template <class T, size_t I, size_t J, size_t K>
Batch<T, I, K> operator*(const Batch<T, I, J>& a, const Batch<T, J, K>& b) {     // (.*.) matrix . matrix
    Batch<T, I, K> c(a.ctx(), a.count());
    a.ctx().check(t5_matmul(a.ctx().handle(), dtype_of<T>::value, (int)I, (int)J, (int)K, a.buf(), b.buf(), c.buf()), "t5_matmul");
    return c;
}
template <class T, size_t I, size_t J>
Batch<T, I> operator*(const Batch<T, I, J>& a, const Batch<T, J>& x) {           // (.*.) matrix . vector
    Batch<T, I> y(a.ctx(), a.count());
    a.ctx().check(t5_matmul(a.ctx().handle(), dtype_of<T>::value, (int)I, (int)J, 1, a.buf(), x.buf(), y.buf()), "t5_matmul");
    return y;
}
template <class T, size_t N>
Batch<T> dot(const Batch<T, N>& x, const Batch<T, N>& y) {
    Batch<T> s(x.ctx(), x.count());
    x.ctx().check(t5_dot(x.ctx().handle(), dtype_of<T>::value, (int)N, x.buf(), y.buf(), s.buf()), "t5_dot");
    return s;
}
template <class T, size_t N>
T dot_sum(const Batch<T, N>& x, const Batch<T, N>& y) {
    T out{};
    x.ctx().check(t5_dot_sum(x.ctx().handle(), dtype_of<T>::value, (int)N, x.buf(), y.buf(), &out), "t5_dot_sum");
    return out;
}
template <class T, size_t... Is, size_t... Js>
Batch<T, Is..., Js...> tensor_product(const Batch<T, Is...>& a, const Batch<T, Js...>& b) {   // (⊗) = prod 0
    Batch<T, Is..., Js...> c(a.ctx(), a.count());
    auto da = Shape<Is...>::dims(); auto db = Shape<Js...>::dims();
    a.ctx().check(t5_prod(a.ctx().handle(), dtype_of<T>::value, (int)da.size(), da.data(), (int)db.size(), db.data(), 0,
                          a.buf(), b.buf(), c.buf()), "t5_prod");
    return c;
}

namespace detail {
template <class T, class Seq, size_t... Is> struct reversed;
template <class T, size_t... Rs> struct reversed<T, std::index_sequence<Rs...>> { using type = Batch<T, Rs...>; };
template <class T, size_t... Rs, size_t I, size_t... Is>
struct reversed<T, std::index_sequence<Rs...>, I, Is...> : reversed<T, std::index_sequence<I, Rs...>, Is...> {};
}  // namespace detail

template <class T, size_t... Is>
typename detail::reversed<T, std::index_sequence<>, Is...>::type transpose(const Batch<T, Is...>& a) {
    typename detail::reversed<T, std::index_sequence<>, Is...>::type t(a.ctx(), a.count());
    auto d = Shape<Is...>::dims();
    a.ctx().check(t5_transpose(a.ctx().handle(), dtype_of<T>::value, (int)d.size(), d.data(), a.buf(), t.buf()), "t5_transpose");
    return t;
}
template <class T, size_t N>
Batch<T> det(const Batch<T, N, N>& a) {
    Batch<T> d(a.ctx(), a.count());
    a.ctx().check(t5_det(a.ctx().handle(), dtype_of<T>::value, (int)N, a.buf(), d.buf()), "t5_det");
    return d;
}
// inverse :: ... -> Maybe: status byte 0 = Just, 1 = Nothing (item zero-filled)
template <class T, size_t N>
std::pair<Batch<T, N, N>, Batch<uint8_t>> inverse(const Batch<T, N, N>& a) {
    Batch<T, N, N> x(a.ctx(), a.count());
    Batch<uint8_t> st(a.ctx(), a.count());
    a.ctx().check(t5_inverse(a.ctx().handle(), dtype_of<T>::value, (int)N, a.buf(), x.buf(), st.buf()), "t5_inverse");
    return {x, st};
}
template <class T, size_t N>
std::pair<Batch<T, N>, Batch<uint8_t>> solve(const Batch<T, N, N>& a, const Batch<T, N>& b) {
    Batch<T, N> x(a.ctx(), a.count());
    Batch<uint8_t> st(a.ctx(), a.count());
    a.ctx().check(t5_solve(a.ctx().handle(), dtype_of<T>::value, (int)N, 1, a.buf(), b.buf(), x.buf(), st.buf()), "t5_solve");
    return {x, st};
}
// the rest of SquareMatrix (CHANGELOG.md:28-29)
template <class T, size_t N>
Batch<T> tr(const Batch<T, N, N>& a) {
    Batch<T> t(a.ctx(), a.count());
    a.ctx().check(t5_trace(a.ctx().handle(), dtype_of<T>::value, (int)N, a.buf(), t.buf()), "t5_trace");
    return t;
}
template <class T, size_t N>
Batch<T, N, N> poly_eval(const Batch<T, N, N>& a, const std::vector<T>& coeffs) {  // c0 I + c1 a + ..
    Batch<T, N, N> r(a.ctx(), a.count());
    a.ctx().check(t5_poly_eval(a.ctx().handle(), dtype_of<T>::value, (int)N, coeffs.data(), (int)coeffs.size(), a.buf(), r.buf()), "t5_poly_eval");
    return r;
}
template <class T, size_t N>
Batch<T, N + 1> char_poly(const Batch<T, N, N>& a) {                               // low order first, leading 1
    Batch<T, N + 1> c(a.ctx(), a.count());
    a.ctx().check(t5_char_poly(a.ctx().handle(), dtype_of<T>::value, (int)N, a.buf(), c.buf()), "t5_char_poly");
    return c;
}
template <class T, size_t N>
std::pair<Batch<T, N + 1>, Batch<uint8_t>> min_poly(const Batch<T, N, N>& a) {      // (coefficients, degree)
    Batch<T, N + 1> c(a.ctx(), a.count());
    Batch<uint8_t> deg(a.ctx(), a.count());
    a.ctx().check(t5_min_poly(a.ctx().handle(), dtype_of<T>::value, (int)N, a.buf(), c.buf(), deg.buf()), "t5_min_poly");
    return {c, deg};
}
// unit: a batch of identity matrices (the old SquareMatrix method)
template <class T, size_t N>
Batch<T, N, N> unit(const Context& ctx, size_t count) {
    std::vector<T> id(N * N, T(0));
    for (size_t i = 0; i < N; ++i) id[i * N + i] = T(1);
    return Batch<T, N, N>::replicate(ctx, count, id);
}
// rowEchelonForm (reduced; pivots looked for in the first pivot_cols columns, default all): (form, rank)
template <class T, size_t R, size_t C>
std::pair<Batch<T, R, C>, Batch<uint8_t>> row_echelon_form(const Batch<T, R, C>& a, int pivot_cols = (int)C) {
    Batch<T, R, C> e(a.ctx(), a.count());
    Batch<uint8_t> rank(a.ctx(), a.count());
    a.ctx().check(t5_row_echelon(a.ctx().handle(), dtype_of<T>::value, (int)R, (int)C, pivot_cols, a.buf(), e.buf(), rank.buf()), "t5_row_echelon");
    return {e, rank};
}
template <class T, size_t... Is>
Batch<T, Is...> operator+(const Batch<T, Is...>& a, const Batch<T, Is...>& b) {   // liftA2 (+)
    Batch<T, Is...> c(a.ctx(), a.count());
    a.ctx().check(t5_zip(a.ctx().handle(), dtype_of<T>::value, 0, a.buf(), b.buf(), c.buf()), "t5_zip");
    return c;
}

template <class T, size_t... Is>
Batch<T, Is...> operator-(const Batch<T, Is...>& a, const Batch<T, Is...>& b) {   // liftA2 (-)
    Batch<T, Is...> c(a.ctx(), a.count());
    a.ctx().check(t5_zip(a.ctx().handle(), dtype_of<T>::value, 1, a.buf(), b.buf(), c.buf()), "t5_zip");
    return c;
}
template <class T, size_t... Is>
Batch<T, Is...> hadamard(const Batch<T, Is...>& a, const Batch<T, Is...>& b) {    // liftA2 (*)
    Batch<T, Is...> c(a.ctx(), a.count());
    a.ctx().check(t5_zip(a.ctx().handle(), dtype_of<T>::value, 2, a.buf(), b.buf(), c.buf()), "t5_zip");
    return c;
}
template <class T, size_t... Is>
Batch<T, Is...> operator*(T k, const Batch<T, Is...>& a) {                        // fmap (k *)
    Batch<T, Is...> c(a.ctx(), a.count());
    a.ctx().check(t5_scale(a.ctx().handle(), dtype_of<T>::value, &k, a.buf(), c.buf()), "t5_scale");
    return c;
}

// ---- type-level shape arithmetic for prod n / append / split / at / ta ----------------------
namespace detail {
template <class T, class S> struct batch_of;
template <class T, size_t... Is> struct batch_of<T, Shape<Is...>> { using type = Batch<T, Is...>; };
template <class A, class B> struct cat;
template <size_t... Is, size_t... Js> struct cat<Shape<Is...>, Shape<Js...>> { using type = Shape<Is..., Js...>; };
template <size_t N, class S> struct drop { using type = S; };
template <size_t N, size_t I, size_t... Is> struct drop<N, Shape<I, Is...>> {
    using type = std::conditional_t<N == 0, Shape<I, Is...>, typename drop<(N ? N - 1 : 0), Shape<Is...>>::type>;
};
template <size_t N, class S> struct take { using type = Shape<>; };
template <size_t N, size_t I, size_t... Is> struct take<N, Shape<I, Is...>> {
    using type = std::conditional_t<N == 0, Shape<>, typename cat<Shape<I>, typename take<(N ? N - 1 : 0), Shape<Is...>>::type>::type>;
};
template <size_t N, class S> struct nth;
template <size_t N, size_t I, size_t... Is> struct nth<N, Shape<I, Is...>> {
    static constexpr size_t value = N == 0 ? I : nth<(N ? N - 1 : 0), Shape<Is..., 0>>::value;   // padded: never read past the end
};
template <size_t N> struct nth<N, Shape<>> { static constexpr size_t value = 0; };
// replace entry AX (0-based) of a shape
template <size_t AX, size_t V, class S> struct set_axis {
    using type = typename cat<typename cat<typename take<AX, S>::type, Shape<V>>::type, typename drop<AX + 1, S>::type>::type;
};
template <class A, class B> constexpr bool same_shape = std::is_same<A, B>::value;
}  // namespace detail

// prod n: contract the last NC indices of a with the first NC of b; a mismatch is a compile error
template <size_t NC, class T, size_t... Is, size_t... Js>
auto prod(const Batch<T, Is...>& a, const Batch<T, Js...>& b) {
    using SA = Shape<Is...>; using SB = Shape<Js...>;
    static_assert(NC <= SA::rank && NC <= SB::rank, "prod n: n exceeds a rank");
    static_assert(detail::same_shape<typename detail::drop<SA::rank - NC, SA>::type, typename detail::take<NC, SB>::type>,
                  "prod n: the contracted dimensions differ");
    using SC = typename detail::cat<typename detail::take<SA::rank - NC, SA>::type, typename detail::drop<NC, SB>::type>::type;
    typename detail::batch_of<T, SC>::type c(a.ctx(), a.count());
    auto da = SA::dims(); auto db = SB::dims();
    a.ctx().check(t5_prod(a.ctx().handle(), dtype_of<T>::value, (int)da.size(), da.data(), (int)db.size(), db.data(), (int)NC,
                          a.buf(), b.buf(), c.buf()), "t5_prod");
    return c;
}

// append n x y / split n t (Vector.hs:322-354); AXIS is 1-based like the reference's SOne, SSucc ..
template <size_t AXIS, class T, size_t... Is, size_t... Js>
auto append(const Batch<T, Is...>& x, const Batch<T, Js...>& y) {
    using SX = Shape<Is...>; using SY = Shape<Js...>;
    static_assert(SX::rank == SY::rank && AXIS >= 1 && AXIS <= SX::rank, "append: ranks differ or bad axis");
    constexpr size_t EX = detail::nth<AXIS - 1, SX>::value, EY = detail::nth<AXIS - 1, SY>::value;
    static_assert(detail::same_shape<typename detail::set_axis<AXIS - 1, 0, SX>::type, typename detail::set_axis<AXIS - 1, 0, SY>::type>,
                  "append: the other dimensions differ");
    using SZ = typename detail::set_axis<AXIS - 1, EX + EY, SX>::type;
    typename detail::batch_of<T, SZ>::type z(x.ctx(), x.count());
    auto dx = SX::dims(); auto dy = SY::dims();
    x.ctx().check(t5_append(x.ctx().handle(), dtype_of<T>::value, (int)AXIS, (int)dx.size(), dx.data(), dy.data(), x.buf(), y.buf(), z.buf()), "t5_append");
    return z;
}
template <size_t AXIS, size_t FIRST, class T, size_t... Is>
auto split(const Batch<T, Is...>& t) {
    using S = Shape<Is...>;
    static_assert(AXIS >= 1 && AXIS <= S::rank, "split: bad axis");
    constexpr size_t E = detail::nth<AXIS - 1, S>::value;
    static_assert(FIRST >= 1 && FIRST < E, "split: both parts must be non-empty");
    typename detail::batch_of<T, typename detail::set_axis<AXIS - 1, FIRST, S>::type>::type a(t.ctx(), t.count());
    typename detail::batch_of<T, typename detail::set_axis<AXIS - 1, E - FIRST, S>::type>::type b(t.ctx(), t.count());
    auto d = S::dims();
    t.ctx().check(t5_split(t.ctx().handle(), dtype_of<T>::value, (int)AXIS, (int)d.size(), d.data(), FIRST, t.buf(), a.buf(), b.buf()), "t5_split");
    return std::make_pair(a, b);
}
// [i] `ta` t: fix the first index (1-based); t `at` ix: fix every index but the first
template <class T, size_t I, size_t... Is>
Batch<T, Is...> ta(size_t i, const Batch<T, I, Is...>& t) {
    Batch<T, Is...> o(t.ctx(), t.count());
    auto d = Shape<I, Is...>::dims();
    t.ctx().check(t5_ta(t.ctx().handle(), dtype_of<T>::value, (int)d.size(), d.data(), i, t.buf(), o.buf()), "t5_ta");
    return o;
}
template <class T, size_t I, size_t... Is>
Batch<T, I> at(const Batch<T, I, Is...>& t, const std::vector<uint64_t>& ix) {
    if (ix.size() != sizeof...(Is)) throw WrongListLength();
    Batch<T, I> o(t.ctx(), t.count());
    auto d = Shape<I, Is...>::dims();
    t.ctx().check(t5_at(t.ctx().handle(), dtype_of<T>::value, (int)d.size(), d.data(), ix.data(), t.buf(), o.buf()), "t5_at");
    return o;
}
// slice t sl: slicer entry 0 keeps an axis (Nothing), k > 0 fixes it (Just k); the result shape is given explicitly
// and checked by the library (WrongListLength on a mismatch)
template <size_t... Os, class T, size_t... Is>
Batch<T, Os...> slice(const Batch<T, Is...>& t, const std::vector<uint64_t>& slicer) {
    if (slicer.size() != sizeof...(Is)) throw WrongListLength();
    Batch<T, Os...> o(t.ctx(), t.count());
    auto d = Shape<Is...>::dims();
    t.ctx().check(t5_slice(t.ctx().handle(), dtype_of<T>::value, (int)d.size(), d.data(), slicer.data(), t.buf(), o.buf()), "t5_slice");
    return o;
}

// ---- ordinary host arrays in and out: one call per operation (the end-to-end entries) -------------
template <class T, size_t I, size_t J, size_t K>
void matmul_host(const Context& ctx, size_t count, const T* a, const T* b, T* c) {
    ctx.check(t5_matmul_host(ctx.handle(), dtype_of<T>::value, (int)I, (int)J, (int)K, count, a, b, c), "t5_matmul_host");
}
template <class T, size_t N>
void det_host(const Context& ctx, size_t count, const T* a, T* out) {
    ctx.check(t5_det_host(ctx.handle(), dtype_of<T>::value, (int)N, count, a, out), "t5_det_host");
}
template <class T, size_t N>
void inverse_host(const Context& ctx, size_t count, const T* a, T* out, uint8_t* status) {
    ctx.check(t5_inverse_host(ctx.handle(), dtype_of<T>::value, (int)N, count, a, out, status), "t5_inverse_host");
}
template <class T, size_t N>
void solve_host(const Context& ctx, size_t count, const T* a, const T* b, T* x, uint8_t* status) {
    ctx.check(t5_solve_host(ctx.handle(), dtype_of<T>::value, (int)N, 1, count, a, b, x, status), "t5_solve_host");
}
template <class T, size_t... Is>
void transpose_host(const Context& ctx, size_t count, const T* a, T* out) {
    auto d = Shape<Is...>::dims();
    ctx.check(t5_transpose_host(ctx.handle(), dtype_of<T>::value, (int)d.size(), d.data(), count, a, out), "t5_transpose_host");
}

}  // namespace t5
