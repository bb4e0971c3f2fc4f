// TEST INFRASTRUCTURE ONLY: stand-in for boost/dynamic_bitset.hpp (64-bit blocks) with the members the reference uses.
// Semantics follow Boost: bit 0 is the least significant; operator>>= moves bits towards lower indices.
#pragma once
#include <vector>
#include <limits>
#include <cstddef>
#include <memory>
#include <algorithm>
#include <iosfwd>
namespace boost {
template <class Block = unsigned long, class Alloc = std::allocator<Block> >
class dynamic_bitset {
  typedef unsigned long long W;
public:
  typedef std::size_t size_type;
  static const size_type npos = static_cast<size_type>(-1);
  class reference {
    W* w_; W m_;
  public:
    reference(W* w, W m) : w_(w), m_(m) {}
    operator bool() const { return (*w_ & m_) != 0; }
    bool operator~() const { return (*w_ & m_) == 0; }
    reference& operator=(bool v) { if (v) *w_ |= m_; else *w_ &= ~m_; return *this; }
    reference& operator=(const reference& o) { return *this = (bool)o; }
    reference& operator|=(bool v) { if (v) *w_ |= m_; return *this; }
    reference& operator&=(bool v) { if (!v) *w_ &= ~m_; return *this; }
    reference& operator^=(bool v) { if (v) *w_ ^= m_; return *this; }
    reference& flip() { *w_ ^= m_; return *this; }
  };
  dynamic_bitset() : n_(0) {}
  explicit dynamic_bitset(size_type n, unsigned long v = 0) : n_(n), w_((n + 63) / 64, 0) { if (!w_.empty()) { w_[0] = v; trim(); } }
  size_type size() const { return n_; }
  bool empty() const { return n_ == 0; }
  size_type num_blocks() const { return w_.size(); }
  void resize(size_type n, bool v = false) {
    const size_type old = n_;
    w_.resize((n + 63) / 64, v ? ~W(0) : W(0));
    if (v && n > old && old % 64) w_[old / 64] |= ~W(0) << (old % 64);
    n_ = n; trim();
  }
  void clear() { n_ = 0; w_.clear(); }
  void swap(dynamic_bitset& o) { std::swap(n_, o.n_); w_.swap(o.w_); }
  void push_back(bool v) { resize(n_ + 1); if (v) set(n_ - 1); }
  bool test(size_type i) const { return (w_[i >> 6] >> (i & 63)) & 1; }
  bool operator[](size_type i) const { return test(i); }
  reference operator[](size_type i) { return reference(&w_[i >> 6], W(1) << (i & 63)); }
  dynamic_bitset& set(size_type i, bool v = true) { if (v) w_[i >> 6] |= W(1) << (i & 63); else w_[i >> 6] &= ~(W(1) << (i & 63)); return *this; }
  dynamic_bitset& set() { std::fill(w_.begin(), w_.end(), ~W(0)); trim(); return *this; }
  dynamic_bitset& reset(size_type i) { return set(i, false); }
  dynamic_bitset& reset() { std::fill(w_.begin(), w_.end(), W(0)); return *this; }
  dynamic_bitset& flip(size_type i) { w_[i >> 6] ^= W(1) << (i & 63); return *this; }
  dynamic_bitset& flip() { for (size_type k = 0; k < w_.size(); ++k) w_[k] = ~w_[k]; trim(); return *this; }
  size_type count() const { size_type c = 0; for (size_type k = 0; k < w_.size(); ++k) c += __builtin_popcountll(w_[k]); return c; }
  bool any() const { for (size_type k = 0; k < w_.size(); ++k) if (w_[k]) return true; return false; }
  bool none() const { return !any(); }
  size_type find_first() const { return find_from(0); }
  size_type find_next(size_type p) const { return p + 1 >= n_ ? npos : find_from(p + 1); }
  bool operator==(const dynamic_bitset& o) const { return n_ == o.n_ && w_ == o.w_; }
  bool operator!=(const dynamic_bitset& o) const { return !(*this == o); }
  bool operator<(const dynamic_bitset& o) const {   // Boost compares as unsigned integers, most significant block first
    for (size_type k = w_.size(); k-- > 0;) if (w_[k] != o.w_[k]) return w_[k] < o.w_[k];
    return false;
  }
  dynamic_bitset& operator&=(const dynamic_bitset& o) { for (size_type k = 0; k < w_.size(); ++k) w_[k] &= o.w_[k]; return *this; }
  dynamic_bitset& operator|=(const dynamic_bitset& o) { for (size_type k = 0; k < w_.size(); ++k) w_[k] |= o.w_[k]; return *this; }
  dynamic_bitset& operator^=(const dynamic_bitset& o) { for (size_type k = 0; k < w_.size(); ++k) w_[k] ^= o.w_[k]; return *this; }
  dynamic_bitset& operator-=(const dynamic_bitset& o) { for (size_type k = 0; k < w_.size(); ++k) w_[k] &= ~o.w_[k]; return *this; }
  dynamic_bitset operator~() const { dynamic_bitset r(*this); r.flip(); return r; }
  dynamic_bitset& operator>>=(size_type s) {   // bit i <- bit i + s
    if (s >= n_) return reset();
    const size_type ws = s >> 6, bs = s & 63, nw = w_.size();
    for (size_type k = 0; k < nw; ++k) {
      W lo = k + ws < nw ? w_[k + ws] : 0, hi = k + ws + 1 < nw ? w_[k + ws + 1] : 0;
      w_[k] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
    }
    return *this;
  }
  dynamic_bitset& operator<<=(size_type s) {   // bit i + s <- bit i
    if (s >= n_) return reset();
    const size_type ws = s >> 6, bs = s & 63, nw = w_.size();
    for (size_type k = nw; k-- > 0;) {
      W hi = k >= ws ? w_[k - ws] : 0, lo = k >= ws + 1 ? w_[k - ws - 1] : 0;
      w_[k] = bs ? (hi << bs) | (lo >> (64 - bs)) : hi;
    }
    trim();
    return *this;
  }
  dynamic_bitset operator>>(size_type s) const { dynamic_bitset r(*this); r >>= s; return r; }
  dynamic_bitset operator<<(size_type s) const { dynamic_bitset r(*this); r <<= s; return r; }
  bool is_subset_of(const dynamic_bitset& o) const { for (size_type k = 0; k < w_.size(); ++k) if (w_[k] & ~o.w_[k]) return false; return true; }
  bool intersects(const dynamic_bitset& o) const { for (size_type k = 0; k < w_.size() && k < o.w_.size(); ++k) if (w_[k] & o.w_[k]) return true; return false; }
private:
  void trim() { if (n_ % 64 && !w_.empty()) w_.back() &= (W(1) << (n_ % 64)) - 1; }
  size_type find_from(size_type p) const {
    if (p >= n_) return npos;
    size_type k = p >> 6;
    W x = w_[k] & (~W(0) << (p & 63));
    while (true) {
      if (x) return (k << 6) + __builtin_ctzll(x);
      if (++k >= w_.size()) return npos;
      x = w_[k];
    }
  }
  size_type n_;
  std::vector<W> w_;
};
template <class B, class A> dynamic_bitset<B,A> operator&(dynamic_bitset<B,A> a, const dynamic_bitset<B,A>& b) { a &= b; return a; }
template <class B, class A> dynamic_bitset<B,A> operator|(dynamic_bitset<B,A> a, const dynamic_bitset<B,A>& b) { a |= b; return a; }
template <class B, class A> dynamic_bitset<B,A> operator^(dynamic_bitset<B,A> a, const dynamic_bitset<B,A>& b) { a ^= b; return a; }
template <class B, class A> dynamic_bitset<B,A> operator-(dynamic_bitset<B,A> a, const dynamic_bitset<B,A>& b) { a -= b; return a; }
template <class C, class T, class B, class A>
std::basic_ostream<C, T>& operator<<(std::basic_ostream<C, T>& os, const dynamic_bitset<B,A>& b) { for (std::size_t i = b.size(); i-- > 0;) os << (b.test(i) ? '1' : '0'); return os; }
}
