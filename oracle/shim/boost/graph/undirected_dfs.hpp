// TEST INFRASTRUCTURE ONLY: undirected_dfs(g, root_vertex(v).visitor(vis).edge_color_map(m)) on the shim graph.
#pragma once
#include "boost/graph/depth_first_search.hpp"
namespace boost {
template <class Vis> struct shim_dfs_params {
  std::size_t root; Vis vis;
  template <class M> shim_dfs_params edge_color_map(const M&) const { return *this; }
  template <class M> shim_dfs_params vertex_color_map(const M&) const { return *this; }
};
struct shim_root_param {
  std::size_t root;
  template <class Vis> shim_dfs_params<Vis> visitor(const Vis& v) const { shim_dfs_params<Vis> p = { root, v }; return p; }
};
inline shim_root_param root_vertex(std::size_t v) { shim_root_param r = { v }; return r; }
template <class G, class Vis>
void shim_undirected_visit(G& g, std::size_t u, std::vector<char>& vcol, std::vector<char>& ecol, Vis& vis)
{
  vcol[u] = 1;
  for (std::size_t k = 0; k < g.out_[u].size(); ++k) {
    const std::size_t id = g.out_[u][k];
    if (ecol[id]) continue;
    ecol[id] = 1;
    const std::size_t v = g.edges_[id].s == u ? g.edges_[id].t : g.edges_[id].s;
    if (!vcol[v]) { shim_tree_edge(vis.a, u, v); shim_tree_edge(vis.b, u, v); shim_undirected_visit(g, v, vcol, ecol, vis); }
  }
  vcol[u] = 2;
}
template <class G, class Vis>
void undirected_dfs(G& g, shim_dfs_params<Vis> p)
{
  std::vector<char> vcol(g.out_.size(), 0), ecol(g.edges_.size(), 0);
  if (p.root < g.out_.size()) shim_undirected_visit(g, p.root, vcol, ecol, p.vis);
  for (std::size_t u = 0; u < g.out_.size(); ++u) if (!vcol[u]) shim_undirected_visit(g, u, vcol, ecol, p.vis);
}
}
