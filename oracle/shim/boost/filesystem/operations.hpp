// Minimal stand-in for boost/filesystem (path, remove, exists) (oracle build only).
#pragma once
#include <string>
#include <cstdio>
#include <sys/stat.h>
namespace boost { namespace filesystem {
class path {
public:
  path() {}
  path(const std::string& s) : s_(s) {}
  path(const char* s) : s_(s) {}
  path& operator=(const std::string& s) { s_ = s; return *this; }
  const std::string& string() const { return s_; }
  const char* c_str() const { return s_.c_str(); }
  path branch_path() const { std::string::size_type k = s_.find_last_of('/'); return k == std::string::npos ? path() : path(s_.substr(0, k)); }
  path parent_path() const { return branch_path(); }
  path leaf() const { std::string::size_type k = s_.find_last_of('/'); return k == std::string::npos ? path(s_) : path(s_.substr(k + 1)); }
  path filename() const { return leaf(); }
  path& operator/=(const path& o) { if (!s_.empty() && s_[s_.size() - 1] != '/') s_ += '/'; s_ += o.s_; return *this; }
  path& operator/=(const std::string& o) { return (*this) /= path(o); }
  void clear() { s_.clear(); }
  bool empty() const { return s_.empty(); }
private:
  std::string s_;
};
inline path operator/(path a, const path& b) { a /= b; return a; }
inline bool remove(const path& p) { return std::remove(p.string().c_str()) == 0; }
inline bool exists(const path& p) { struct stat st; return stat(p.string().c_str(), &st) == 0; }
}}
