// Minimal stand-in for boost/pool/pool_alloc.hpp: plain new/delete allocators (oracle build only).
#pragma once
#include <memory>
#include <cstddef>
namespace boost {
template <class T> struct fast_pool_allocator : public std::allocator<T> {
  template <class U> struct rebind { typedef fast_pool_allocator<U> other; };
  fast_pool_allocator() {}
  template <class U> fast_pool_allocator(const fast_pool_allocator<U>&) {}
  static T* allocate() { return static_cast<T*>(::operator new(sizeof(T))); }
  static T* allocate(std::size_t n) { return static_cast<T*>(::operator new(n * sizeof(T))); }
  static void deallocate(T* p) { ::operator delete(p); }
  static void deallocate(T* p, std::size_t) { ::operator delete(p); }
};
template <class T> struct pool_allocator : public fast_pool_allocator<T> {
  template <class U> struct rebind { typedef pool_allocator<U> other; };
};
}
