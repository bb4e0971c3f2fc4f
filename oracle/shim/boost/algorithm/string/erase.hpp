// TEST INFRASTRUCTURE ONLY: stand-in for boost/algorithm/string/erase.hpp
#pragma once
#include <string>
namespace boost { namespace algorithm {
inline void erase_all(std::string& s, const std::string& what) { if (what.empty()) return; std::string::size_type p = 0; while ((p = s.find(what, p)) != std::string::npos) s.erase(p, what.size()); }
inline std::string erase_all_copy(std::string s, const std::string& what) { erase_all(s, what); return s; }
} using algorithm::erase_all; using algorithm::erase_all_copy; }
