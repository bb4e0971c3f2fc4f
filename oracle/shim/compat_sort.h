// TEST INFRASTRUCTURE ONLY (force-included when compiling LM/Aligner.cpp for the full-pipeline oracle).
// LabeledMemComparator and MatchLeftEndComparator declare copy constructors taking a NON-const reference (LM/Aligner.h:52),
// which today's libstdc++ std::sort cannot copy internally.  The reference calls `sort(first, last, comparator_lvalue)`; these
// more specialised overloads hand std::sort a copyable wrapper that forwards to the caller's comparator object: the same
// algorithm makes the same comparisons, so the resulting order is the one std::sort would produce.
#pragma once
#include <algorithm>
namespace mems { class LabeledMemComparator; class MatchLeftEndComparator; }
namespace oracle_compat {
template <class C> struct by_ref {
  C* c;
  template <class A, class B> bool operator()(const A& a, const B& b) const { return (*c)(a, b); }
};
}
namespace std {
template <class It> inline void sort(It f, It l, mems::LabeledMemComparator& c) { oracle_compat::by_ref<mems::LabeledMemComparator> w = { &c }; std::sort(f, l, w); }
template <class It> inline void sort(It f, It l, mems::MatchLeftEndComparator& c) { oracle_compat::by_ref<mems::MatchLeftEndComparator> w = { &c }; std::sort(f, l, w); }
}
