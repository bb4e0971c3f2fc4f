// TEST INFRASTRUCTURE ONLY: topological_sort on the shim graph = BGL's: a depth-first search over the vertices in ascending
// order, out-edges in insertion order, every vertex written to the output when it FINISHES (reverse topological order).
#pragma once
#include <stdexcept>
#include "boost/graph/adjacency_list.hpp"
namespace boost {
struct not_a_dag : public std::invalid_argument { not_a_dag() : std::invalid_argument("The graph must be a DAG.") {} };
template <class G, class Out>
void shim_topo_visit(const G& g, std::size_t u, std::vector<char>& col, Out& out)
{
  col[u] = 1;
  for (std::size_t k = 0; k < g.out_[u].size(); ++k) {
    const std::size_t v = g.edges_[g.out_[u][k]].t;
    if (col[v] == 1) throw not_a_dag();
    if (!col[v]) shim_topo_visit(g, v, col, out);
  }
  col[u] = 2;
  *out++ = u;
}
template <class G, class Out>
void topological_sort(const G& g, Out out)
{
  std::vector<char> col(g.out_.size(), 0);
  for (std::size_t u = 0; u < g.out_.size(); ++u) if (!col[u]) shim_topo_visit(g, u, col, out);
}
}
