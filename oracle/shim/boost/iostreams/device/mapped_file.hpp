// Minimal stand-in for boost/iostreams/device/mapped_file.hpp via POSIX mmap (oracle build only).
#pragma once
#include <string>
#include <cstddef>
#include <stdexcept>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
namespace boost { namespace iostreams {
class mapped_file_source {
public:
  typedef std::size_t size_type;
  mapped_file_source() : p_(0), base_(0), n_(0), maplen_(0) {}
  ~mapped_file_source() {}
  void open(const std::string& path, size_type length = static_cast<size_type>(-1), long long offset = 0) {
    close();
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("mapped_file_source: cannot open " + path);
    struct stat st; fstat(fd, &st);
    size_type fsz = st.st_size;
    if (length == static_cast<size_type>(-1) || offset + length > fsz) length = fsz - offset;
    long pg = sysconf(_SC_PAGESIZE);
    long long aoff = (offset / pg) * pg;
    maplen_ = length + (offset - aoff);
    if (maplen_ == 0) { ::close(fd); p_ = 0; n_ = 0; base_ = 0; return; }
    void* m = mmap(0, maplen_, PROT_READ, MAP_PRIVATE, fd, aoff);
    ::close(fd);
    if (m == MAP_FAILED) throw std::runtime_error("mapped_file_source: mmap failed");
    base_ = static_cast<char*>(m); p_ = base_ + (offset - aoff); n_ = length;
  }
  bool is_open() const { return base_ != 0; }
  void close() { if (base_) munmap(base_, maplen_); base_ = 0; p_ = 0; n_ = 0; maplen_ = 0; }
  const char* data() const { return p_; }
  size_type size() const { return n_; }
  const char* begin() const { return p_; }
  const char* end() const { return p_ + n_; }
private:
  const char* p_; char* base_; size_type n_, maplen_;
};
}}
