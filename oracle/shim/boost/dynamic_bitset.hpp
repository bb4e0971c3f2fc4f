// Minimal stand-in for boost/dynamic_bitset.hpp over std::vector<bool> (oracle build only).
#pragma once
#include <vector>
#include <limits>
#include <cstddef>
#include <memory>
namespace boost {
template <class Block = unsigned long, class Alloc = std::allocator<Block> >
class dynamic_bitset {
public:
  typedef std::size_t size_type;
  static const size_type npos = static_cast<size_type>(-1);
  dynamic_bitset() {}
  explicit dynamic_bitset(size_type n, unsigned long v = 0) : b_(n, false) {
    for (size_type i = 0; i < n && i < 8 * sizeof(v); ++i) b_[i] = (v >> i) & 1ul;
  }
  size_type size() const { return b_.size(); }
  void resize(size_type n, bool v = false) { b_.resize(n, v); }
  void clear() { b_.clear(); }
  bool test(size_type i) const { return b_[i]; }
  bool operator[](size_type i) const { return b_[i]; }
  std::vector<bool>::reference operator[](size_type i) { return b_[i]; }
  dynamic_bitset& set(size_type i, bool v = true) { b_[i] = v; return *this; }
  dynamic_bitset& set() { b_.assign(b_.size(), true); return *this; }
  dynamic_bitset& reset(size_type i) { b_[i] = false; return *this; }
  dynamic_bitset& reset() { b_.assign(b_.size(), false); return *this; }
  dynamic_bitset& flip(size_type i) { b_[i] = !b_[i]; return *this; }
  dynamic_bitset& flip() { b_.flip(); return *this; }
  size_type count() const { size_type c = 0; for (size_type i = 0; i < b_.size(); ++i) c += b_[i]; return c; }
  bool any() const { return count() != 0; }
  bool none() const { return !any(); }
  size_type find_first() const { for (size_type i = 0; i < b_.size(); ++i) if (b_[i]) return i; return npos; }
  size_type find_next(size_type p) const { for (size_type i = p + 1; i < b_.size(); ++i) if (b_[i]) return i; return npos; }
  void push_back(bool v) { b_.push_back(v); }
  bool operator==(const dynamic_bitset& o) const { return b_ == o.b_; }
  bool operator!=(const dynamic_bitset& o) const { return b_ != o.b_; }
  bool operator<(const dynamic_bitset& o) const { return b_ < o.b_; }
  dynamic_bitset& operator&=(const dynamic_bitset& o) { for (size_type i = 0; i < b_.size(); ++i) b_[i] = b_[i] && o.b_[i]; return *this; }
  dynamic_bitset& operator|=(const dynamic_bitset& o) { for (size_type i = 0; i < b_.size(); ++i) b_[i] = b_[i] || o.b_[i]; return *this; }
  dynamic_bitset& operator^=(const dynamic_bitset& o) { for (size_type i = 0; i < b_.size(); ++i) b_[i] = b_[i] != o.b_[i]; return *this; }
  dynamic_bitset operator~() const { dynamic_bitset r(*this); r.flip(); return r; }
private:
  std::vector<bool> b_;
};
template <class B, class A> dynamic_bitset<B,A> operator&(dynamic_bitset<B,A> a, const dynamic_bitset<B,A>& b) { a &= b; return a; }
template <class B, class A> dynamic_bitset<B,A> operator|(dynamic_bitset<B,A> a, const dynamic_bitset<B,A>& b) { a |= b; return a; }
template <class B, class A> dynamic_bitset<B,A> operator^(dynamic_bitset<B,A> a, const dynamic_bitset<B,A>& b) { a ^= b; return a; }
}
