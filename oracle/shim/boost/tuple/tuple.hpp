// TEST INFRASTRUCTURE ONLY: stand-in for boost/tuple/tuple.hpp over std::tuple (get<N>() member and free function,
// make_tuple, tie assignable from std::pair).
#pragma once
#include <tuple>
#include <utility>
namespace boost {
template <class... Ts>
struct tuple : std::tuple<Ts...> {
  typedef std::tuple<Ts...> base;
  tuple() {}
  tuple(const Ts&... a) : base(a...) {}
  template <class... Us> tuple(const std::tuple<Us...>& o) : base(o) {}
  template <class U, class V> tuple& operator=(const std::pair<U, V>& p) { std::get<0>(*this) = p.first; std::get<1>(*this) = p.second; return *this; }
  template <std::size_t I> typename std::tuple_element<I, base>::type& get() { return std::get<I>(static_cast<base&>(*this)); }
  template <std::size_t I> const typename std::tuple_element<I, base>::type& get() const { return std::get<I>(static_cast<const base&>(*this)); }
};
template <std::size_t I, class... Ts> typename std::tuple_element<I, std::tuple<Ts...> >::type& get(tuple<Ts...>& t) { return t.template get<I>(); }
template <std::size_t I, class... Ts> const typename std::tuple_element<I, std::tuple<Ts...> >::type& get(const tuple<Ts...>& t) { return t.template get<I>(); }
template <class... Ts> tuple<Ts...> make_tuple(const Ts&... a) { return tuple<Ts...>(a...); }
template <class... Ts> tuple<Ts&...> tie(Ts&... a) { return tuple<Ts&...>(a...); }
namespace tuples { using boost::tuple; using boost::get; using boost::make_tuple; using boost::tie; }
}
