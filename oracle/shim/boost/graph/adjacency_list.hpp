// TEST INFRASTRUCTURE ONLY: stand-in for the slice of the Boost Graph Library the reference uses
// (LM/ProgressiveAligner.cpp:2976-3046 findMidpoint, LM/Backbone.cpp:742-786): adjacency_list<vecS, vecS, ...> built from an
// edge range, an int edge-weight map, johnson_all_pairs_shortest_paths, undirected_dfs with predecessor/distance recorders,
// topological_sort.  Traversal orders follow BGL's (vertices ascending, out-edges in insertion order), which is what makes
// topological_sort's output reproducible.
#pragma once
#include <vector>
#include <utility>
#include <limits>
#include <cstddef>
#include <iterator>
namespace boost {
struct vecS {}; struct listS {}; struct setS {};
struct undirectedS {}; struct directedS {}; struct bidirectionalS {};
struct no_property {};
enum default_color_type { white_color, gray_color, green_color, red_color, black_color };
struct edge_weight_t {}; struct edge_color_t {}; struct vertex_color_t {}; struct vertex_index_t {};
static const edge_weight_t edge_weight = edge_weight_t();
static const edge_color_t edge_color = edge_color_t();
static const vertex_color_t vertex_color = vertex_color_t();
template <class Tag, class T, class Next = no_property> struct property {};

struct shim_edge { std::size_t s, t, id; };
inline bool operator==(const shim_edge& a, const shim_edge& b) { return a.id == b.id; }
inline bool operator!=(const shim_edge& a, const shim_edge& b) { return a.id != b.id; }

template <class OutS, class VertS, class Dir, class VP = no_property, class EP = no_property, class GP = no_property, class EL = listS>
class adjacency_list {
public:
  typedef std::size_t vertex_descriptor;
  typedef shim_edge edge_descriptor;
  typedef std::size_t vertices_size_type;
  typedef std::size_t edges_size_type;
  typedef std::vector<shim_edge>::const_iterator edge_iterator;
  static const bool is_directed = !std::is_same<Dir, undirectedS>::value;
  adjacency_list() {}
  explicit adjacency_list(vertices_size_type n) : out_(n) {}
  template <class It> adjacency_list(It first, It last, vertices_size_type n) : out_(n) {
    for (; first != last; ++first) add((std::size_t)first->first, (std::size_t)first->second);
  }
  void add(std::size_t s, std::size_t t) {
    if (std::max(s, t) >= out_.size()) out_.resize(std::max(s, t) + 1);   // BGL grows a vecS vertex set on demand
    shim_edge e = { s, t, edges_.size() };
    edges_.push_back(e);
    weight_.push_back(0);
    out_[s].push_back(e.id);
    if (!is_directed) out_[t].push_back(e.id);
  }
  std::vector<std::vector<std::size_t> > out_;   // per vertex: edge ids in insertion order
  std::vector<shim_edge> edges_;
  std::vector<int> weight_;
};
template <class G> struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor;
  typedef typename G::edge_descriptor edge_descriptor;
  typedef typename G::vertices_size_type vertices_size_type;
  typedef typename G::edges_size_type edges_size_type;
  typedef typename G::edge_iterator edge_iterator;
};
template <class G> struct shim_weight_map { G* g; int& operator[](const shim_edge& e) const { return g->weight_[e.id]; } };
struct shim_dummy_map {};
template <class G, class Tag> struct property_map { typedef shim_dummy_map type; typedef shim_dummy_map const_type; };
template <class G> struct property_map<G, edge_weight_t> { typedef shim_weight_map<G> type; typedef shim_weight_map<G> const_type; };
template <class O, class V, class D, class VP, class EP, class GP, class EL>
shim_weight_map<adjacency_list<O, V, D, VP, EP, GP, EL> > get(edge_weight_t, adjacency_list<O, V, D, VP, EP, GP, EL>& g) { shim_weight_map<adjacency_list<O, V, D, VP, EP, GP, EL> > m = { &g }; return m; }
template <class G> shim_dummy_map get(edge_color_t, G&) { return shim_dummy_map(); }
template <class G> shim_dummy_map get(vertex_color_t, G&) { return shim_dummy_map(); }
template <class O, class V, class D, class VP, class EP, class GP, class EL>
std::pair<std::vector<shim_edge>::const_iterator, std::vector<shim_edge>::const_iterator> edges(const adjacency_list<O, V, D, VP, EP, GP, EL>& g) { return std::make_pair(g.edges_.begin(), g.edges_.end()); }
template <class O, class V, class D, class VP, class EP, class GP, class EL>
std::size_t num_vertices(const adjacency_list<O, V, D, VP, EP, GP, EL>& g) { return g.out_.size(); }
template <class O, class V, class D, class VP, class EP, class GP, class EL>
std::size_t vertex(std::size_t n, const adjacency_list<O, V, D, VP, EP, GP, EL>&) { return n; }
template <class O, class V, class D, class VP, class EP, class GP, class EL>
std::pair<shim_edge, bool> add_edge(std::size_t s, std::size_t t, adjacency_list<O, V, D, VP, EP, GP, EL>& g) { g.add(s, t); return std::make_pair(g.edges_.back(), true); }
}
