// Stand-in for boost::core::demangle (error-message formatting of the Examples
// framework only).  TEST INFRASTRUCTURE, see Eigen/Core in this directory.
#pragma once
#include <cstdlib>
#include <cxxabi.h>
#include <string>
namespace boost::core {
inline std::string demangle(const char* name) {
  int status = 0;
  char* p = abi::__cxa_demangle(name, nullptr, nullptr, &status);
  std::string out = (status == 0 && p != nullptr) ? p : name;
  std::free(p);
  return out;
}
}  // namespace boost::core
