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
//
// A shape error is a COMPILE error, as in Haskell; a wrong list length throws
// WrongListLength (src/Data/Tensor.hs:96-100); an unsupported request throws Unsupported
// (there is no CPU fallback).  Header-only; link with libt5b200.so.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
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
template <class T, size_t... Is>
Batch<T, Is...> operator+(const Batch<T, Is...>& a, const Batch<T, Is...>& b) {   // liftA2 (+)
    Batch<T, Is...> c(a.ctx(), a.count());
    a.ctx().check(t5_zip(a.ctx().handle(), dtype_of<T>::value, 0, a.buf(), b.buf(), c.buf()), "t5_zip");
    return c;
}

}  // namespace t5
