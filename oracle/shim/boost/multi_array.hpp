// TEST INFRASTRUCTURE ONLY: stand-in for boost/multi_array.hpp: dense N-dimensional array, row-major, with the members the
// reference uses (extents[a][b][c], resize keeping the overlap, operator[] chains, shape, num_elements, data).
#pragma once
#include <vector>
#include <cstddef>
#include <algorithm>
namespace boost {
namespace detail_ma {
template <std::size_t N> struct extent_gen {
  std::size_t e[N ? N : 1];
  extent_gen<N + 1> operator[](std::size_t n) const { extent_gen<N + 1> r; for (std::size_t i = 0; i < N; ++i) r.e[i] = e[i]; r.e[N] = n; return r; }
};
template <class T, std::size_t D> struct view {      // D remaining dimensions
  T* p; const std::size_t* shape; const std::size_t* stride;
  view<T, D - 1> operator[](std::size_t i) const { view<T, D - 1> v = { p + i * stride[0], shape + 1, stride + 1 }; return v; }
  std::size_t size() const { return shape[0]; }
};
template <class T> struct view<T, 1> {
  T* p; const std::size_t* shape; const std::size_t* stride;
  T& operator[](std::size_t i) const { return p[i]; }
  std::size_t size() const { return shape[0]; }
  T* begin() const { return p; }
  T* end() const { return p + shape[0]; }
};
}
static const detail_ma::extent_gen<0> extents = detail_ma::extent_gen<0>();
template <class T, std::size_t N>
class multi_array {
public:
  typedef T element;
  typedef std::size_t size_type;
  typedef std::size_t index;
  multi_array() { for (std::size_t i = 0; i < N; ++i) shape_[i] = 0; strides(); }
  explicit multi_array(const detail_ma::extent_gen<N>& x) { for (std::size_t i = 0; i < N; ++i) shape_[i] = x.e[i]; strides(); d_.assign(total(), T()); }
  multi_array(const multi_array& o) : d_(o.d_) { std::copy(o.shape_, o.shape_ + N, shape_); strides(); }
  multi_array& operator=(const multi_array& o) { d_ = o.d_; std::copy(o.shape_, o.shape_ + N, shape_); strides(); return *this; }
  void resize(const detail_ma::extent_gen<N>& x) {
    multi_array n(x);
    std::size_t lim[N], idx[N];
    bool any = true;
    for (std::size_t i = 0; i < N; ++i) { lim[i] = std::min(shape_[i], n.shape_[i]); idx[i] = 0; if (!lim[i]) any = false; }
    while (any) {   // copy the overlapping block
      std::size_t a = 0, b = 0;
      for (std::size_t i = 0; i < N; ++i) { a += idx[i] * stride_[i]; b += idx[i] * n.stride_[i]; }
      n.d_[b] = d_[a];
      std::size_t k = N;
      while (k-- > 0) { if (++idx[k] < lim[k]) break; idx[k] = 0; if (k == 0) any = false; }
    }
    *this = n;
  }
  const size_type* shape() const { return shape_; }
  size_type num_elements() const { return d_.size(); }
  size_type size() const { return shape_[0]; }
  static size_type num_dimensions() { return N; }
  T* data() { return d_.empty() ? 0 : &d_[0]; }
  const T* data() const { return d_.empty() ? 0 : &d_[0]; }
  T* origin() { return data(); }
  const T* origin() const { return data(); }
  typename std::conditional<N == 1, T&, detail_ma::view<T, N - 1> >::type operator[](std::size_t i) { return at(i, std::integral_constant<bool, N == 1>()); }
  typename std::conditional<N == 1, const T&, detail_ma::view<const T, N - 1> >::type operator[](std::size_t i) const { return at(i, std::integral_constant<bool, N == 1>()); }
private:
  T& at(std::size_t i, std::true_type) { return d_[i]; }
  const T& at(std::size_t i, std::true_type) const { return d_[i]; }
  detail_ma::view<T, N - 1> at(std::size_t i, std::false_type) { detail_ma::view<T, N - 1> v = { data() + i * stride_[0], shape_ + 1, stride_ + 1 }; return v; }
  detail_ma::view<const T, N - 1> at(std::size_t i, std::false_type) const { detail_ma::view<const T, N - 1> v = { data() + i * stride_[0], shape_ + 1, stride_ + 1 }; return v; }
  std::size_t total() const { std::size_t t = 1; for (std::size_t i = 0; i < N; ++i) t *= shape_[i]; return t; }
  void strides() { std::size_t s = 1; for (std::size_t i = N; i-- > 0;) { stride_[i] = s; s *= shape_[i]; } }
  std::vector<T> d_;
  std::size_t shape_[N], stride_[N];
};
}
