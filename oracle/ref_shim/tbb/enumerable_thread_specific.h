#pragma once
namespace tbb {
template <typename T>
class enumerable_thread_specific {
 public:
  T& local() { return m_value; }
  T* begin() { return &m_value; }
  T* end() { return &m_value + 1; }
 private:
  T m_value{};
};
}  // namespace tbb
