// Stand-in for the one boost::algorithm function the Examples framework calls
// (replace_all, type-name prettifying).  TEST INFRASTRUCTURE, see Eigen/Core.
#pragma once
#include <string>
#include <string_view>
namespace boost::algorithm {
inline void replace_all(std::string& s, std::string_view from, std::string_view to) {
  if (from.empty()) return;
  for (std::size_t pos = s.find(from); pos != std::string::npos; pos = s.find(from, pos + to.size())) {
    s.replace(pos, from.size(), to);
  }
}
}  // namespace boost::algorithm
