// Stand-in for boost::container::small_vector (TEST INFRASTRUCTURE, see Eigen/Core
// in this directory): same observable behaviour as a std::vector; the inline
// capacity is only an allocation optimisation in Boost.
#pragma once
#include <cstddef>
#include <vector>
namespace boost::container {
template <typename T, std::size_t N, typename Allocator = void, typename Options = void>
class small_vector : public std::vector<T> {
 public:
  using std::vector<T>::vector;
};
}  // namespace boost::container
