#pragma once
#include <cstddef>
namespace tbb {
template <typename T>
class blocked_range {
 public:
  blocked_range(T b, T e, std::size_t = 1) : m_b(b), m_e(e) {}
  T begin() const { return m_b; }
  T end() const { return m_e; }
 private:
  T m_b, m_e;
};
template <typename R, typename F>
void parallel_for(const R& r, const F& f) { f(r); }
}  // namespace tbb
