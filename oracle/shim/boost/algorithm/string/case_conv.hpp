// TEST INFRASTRUCTURE ONLY: stand-in for boost/algorithm/string/case_conv.hpp
#pragma once
#include <string>
#include <cctype>
namespace boost { namespace algorithm {
inline void to_upper(std::string& s) { for (std::string::size_type i = 0; i < s.size(); ++i) s[i] = (char)std::toupper((unsigned char)s[i]); }
inline void to_lower(std::string& s) { for (std::string::size_type i = 0; i < s.size(); ++i) s[i] = (char)std::tolower((unsigned char)s[i]); }
inline std::string to_upper_copy(std::string s) { to_upper(s); return s; }
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
} using algorithm::to_upper; using algorithm::to_lower; using algorithm::to_upper_copy; using algorithm::to_lower_copy; }
