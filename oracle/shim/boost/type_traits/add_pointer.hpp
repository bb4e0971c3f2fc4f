// Minimal stand-in for boost/type_traits/add_pointer.hpp (oracle build only; test infrastructure).
#pragma once
#include <type_traits>
namespace boost { template <class T> struct add_pointer { typedef typename std::add_pointer<T>::type type; }; }
