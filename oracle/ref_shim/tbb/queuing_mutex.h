#pragma once
#include <mutex>
namespace tbb {
class queuing_mutex {
 public:
  class scoped_lock {
   public:
    scoped_lock() = default;
    explicit scoped_lock(queuing_mutex& m) : m_lock(m.m_mutex) {}
   private:
    std::unique_lock<std::mutex> m_lock;
  };
 private:
  std::mutex m_mutex;
};
}  // namespace tbb
