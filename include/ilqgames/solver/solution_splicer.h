// SolutionSplicer (reference: include/ilqgames/solver/solution_splicer.h:57-87,
// src/solution_splicer.cpp:57-131): the receding-horizon caller's running plan.  Host-side
// bookkeeping on OperatingPoint / Strategy (no arithmetic): a newly solved horizon is written over
// the tail of the stored plan from the time step at which it starts, keeping up to five already
// executed steps in front of it for a downstream path follower.
#ifndef ILQGAMES_B200_SOLVER_SOLUTION_SPLICER_H
#define ILQGAMES_B200_SOLVER_SOLUTION_SPLICER_H

#include <ilqgames/b200/solvers.h>

#include <algorithm>
#include <vector>

namespace ilqgames {

class SolutionSplicer {
 public:
  ~SolutionSplicer() {}
  explicit SolutionSplicer(const SolverLog& log)
      : strategies_(log.FinalStrategies()), operating_point_(log.FinalOperatingPoint()) {}

  // Splice in the final iterate of `log`, a horizon that starts at or after the stored plan's t0.
  void Splice(const SolverLog& log) {
    const OperatingPoint& fresh = log.FinalOperatingPoint();
    const std::vector<Strategy>& fresh_strategies = log.FinalStrategies();
    const size_t horizon = time::kNumTimeSteps;
    CHECK_GE(fresh.t0, operating_point_.t0);
    CHECK_GE(operating_point_.xs.size(), horizon);
    CHECK_EQ(fresh.xs.size(), horizon);

    // time step of the stored plan at which the new horizon begins (the 1e-4 guards the truncation)
    const size_t begin = static_cast<size_t>(1e-4 + (fresh.t0 - operating_point_.t0) / time::kTimeStep);
    constexpr size_t kKeep = 5;  // executed steps kept in front of the new horizon
    const size_t drop = begin < kKeep ? 0 : begin - kKeep;  // stored steps that fall off the front
    const size_t kept = begin - drop;
    const size_t total = kept + horizon;

    // the kept steps move to the front, the new horizon follows them
    auto rebuild = [&](auto& stored, const auto& incoming) {
      using Vec = typename std::remove_reference<decltype(stored)>::type;
      Vec out;
      out.reserve(total);
      for (size_t k = 0; k < kept; k++) out.push_back(stored[drop + k]);
      for (size_t k = 0; k < horizon; k++) out.push_back(incoming[k]);
      stored.swap(out);
    };
    rebuild(operating_point_.xs, fresh.xs);
    rebuild(operating_point_.us, fresh.us);
    for (size_t ii = 0; ii < strategies_.size(); ii++) {
      rebuild(strategies_[ii].Ps, fresh_strategies[ii].Ps);
      rebuild(strategies_[ii].alphas, fresh_strategies[ii].alphas);
    }
    operating_point_.t0 += drop * time::kTimeStep;
  }

  // Is `t` covered by the stored plan?
  bool ContainsTime(Time t) const {
    return operating_point_.t0 <= t && operating_point_.t0 + operating_point_.xs.size() * time::kTimeStep >= t;
  }

  const std::vector<Strategy>& CurrentStrategies() const { return strategies_; }
  const OperatingPoint& CurrentOperatingPoint() const { return operating_point_; }

 private:
  std::vector<Strategy> strategies_;
  OperatingPoint operating_point_;
};

}  // namespace ilqgames

#endif
