// Minimal stand-in for boost/type_traits/remove_pointer.hpp (oracle build only; test infrastructure).
#pragma once
#include <type_traits>
namespace boost { template <class T> struct remove_pointer { typedef typename std::remove_pointer<T>::type type; }; }
