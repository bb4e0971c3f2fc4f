// TEST INFRASTRUCTURE ONLY: the visitor vocabulary of BGL's depth-first searches as the reference spells it
// (make_dfs_visitor(make_pair(record_predecessors(p, on_tree_edge()), record_distances(d, on_tree_edge())))).
#pragma once
#include "boost/graph/adjacency_list.hpp"
namespace boost {
struct on_tree_edge {};
template <class P> struct shim_pred_recorder { P p; };
template <class P> struct shim_dist_recorder { P p; };
template <class P> shim_pred_recorder<P> record_predecessors(P p, on_tree_edge) { shim_pred_recorder<P> r = { p }; return r; }
template <class P> shim_dist_recorder<P> record_distances(P p, on_tree_edge) { shim_dist_recorder<P> r = { p }; return r; }
template <class A, class B> struct shim_dfs_visitor { A a; B b; };
template <class A, class B> shim_dfs_visitor<A, B> make_dfs_visitor(const std::pair<A, B>& p) { shim_dfs_visitor<A, B> v = { p.first, p.second }; return v; }
template <class P> inline void shim_tree_edge(shim_pred_recorder<P>& r, std::size_t u, std::size_t v) { r.p[v] = u; }
template <class P> inline void shim_tree_edge(shim_dist_recorder<P>& r, std::size_t u, std::size_t v) { r.p[v] = r.p[u] + 1; }
}
