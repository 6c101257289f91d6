// Stand-in for boost::container::static_vector (only named by a header the
// Examples framework includes; never exercised by the seeding path).
// TEST INFRASTRUCTURE, see Eigen/Core in this directory.
#pragma once
#include <cassert>
#include <cstddef>
#include <vector>
namespace boost::container {
template <typename T, std::size_t N, typename Options = void>
class static_vector : public std::vector<T> {
 public:
  using std::vector<T>::vector;
};
}  // namespace boost::container
