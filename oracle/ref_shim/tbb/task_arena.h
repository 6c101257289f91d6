// Declaration-level stand-ins for the TBB types named by the Examples framework
// headers (Sequencer.hpp / tbbWrap.hpp).  The oracle driver never runs a Sequencer;
// everything here executes inline on the calling thread.  TEST INFRASTRUCTURE.
#pragma once
#include <utility>
namespace tbb {
class task_arena {
 public:
  static constexpr int automatic = -1;
  explicit task_arena(int = automatic, unsigned = 1) {}
  template <typename F>
  void execute(F&& f) { std::forward<F>(f)(); }
};
}  // namespace tbb
