// Minimal stand-in for boost/algorithm/string.hpp (replace_all, erase_all, to_lower/upper) (oracle build only).
#pragma once
#include <string>
#include <cctype>
namespace boost { namespace algorithm {
inline void replace_all(std::string& s, const std::string& from, const std::string& to) {
  if (from.empty()) return;
  std::string::size_type p = 0;
  while ((p = s.find(from, p)) != std::string::npos) { s.replace(p, from.size(), to); p += to.size(); }
}
inline void erase_all(std::string& s, const std::string& what) { replace_all(s, what, ""); }
inline void to_lower(std::string& s) { for (std::string::size_type i = 0; i < s.size(); ++i) s[i] = (char)std::tolower((unsigned char)s[i]); }
inline void to_upper(std::string& s) { for (std::string::size_type i = 0; i < s.size(); ++i) s[i] = (char)std::toupper((unsigned char)s[i]); }
}
using algorithm::replace_all; using algorithm::erase_all; using algorithm::to_lower; using algorithm::to_upper;
}
