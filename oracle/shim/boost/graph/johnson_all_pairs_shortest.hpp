// TEST INFRASTRUCTURE ONLY: all-pairs shortest paths on the shim graph (Floyd-Warshall; BGL's "infinity" = max int).
#pragma once
#include "boost/graph/adjacency_list.hpp"
namespace boost {
template <class G, class Matrix>
bool johnson_all_pairs_shortest_paths(G& g, Matrix& D)
{
  const std::size_t n = g.out_.size();
  const int inf = (std::numeric_limits<int>::max)();
  for (std::size_t i = 0; i < n; ++i) for (std::size_t j = 0; j < n; ++j) D[i][j] = i == j ? 0 : inf;
  for (std::size_t k = 0; k < g.edges_.size(); ++k) {
    const shim_edge& e = g.edges_[k];
    if (g.weight_[k] < D[e.s][e.t]) D[e.s][e.t] = g.weight_[k];
    if (!G::is_directed && g.weight_[k] < D[e.t][e.s]) D[e.t][e.s] = g.weight_[k];
  }
  for (std::size_t k = 0; k < n; ++k) for (std::size_t i = 0; i < n; ++i) { if (D[i][k] == inf) continue; for (std::size_t j = 0; j < n; ++j) { if (D[k][j] == inf) continue; const long long s = (long long)D[i][k] + D[k][j]; if (s < D[i][j]) D[i][j] = (int)s; } }
  for (std::size_t i = 0; i < n; ++i) if (D[i][i] < 0) return false;
  return true;
}
}
