// TEST INFRASTRUCTURE ONLY: stand-in for boost/algorithm/string.hpp (replace_all + the erase / case headers).
#pragma once
#include <string>
#include "boost/algorithm/string/erase.hpp"
#include "boost/algorithm/string/case_conv.hpp"
namespace boost { namespace algorithm {
inline void replace_all(std::string& s, const std::string& from, const std::string& to) {
  if (from.empty()) return;
  std::string::size_type p = 0;
  while ((p = s.find(from, p)) != std::string::npos) { s.replace(p, from.size(), to); p += to.size(); }
}
} using algorithm::replace_all; }
