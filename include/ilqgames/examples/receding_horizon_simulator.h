// RecedingHorizonSimulator (reference: include/ilqgames/examples/receding_horizon_simulator.h:59-63,
// src/receding_horizon_simulator.cpp:65-137): a stand-in for online operation.  The game is solved
// once; afterwards, repeatedly, the state is carried along the running plan for a quarter second
// of "driving", the problem is re-based there (Problem::SetUpNextRecedingHorizon) and re-solved
// under a time budget, the state is carried on for however long that solve took, and the new
// horizon is spliced into the running plan if the solve converged.  One log per solver call.
//
// Host-side sequencing only: the solves (ILQSolver / AugmentedLagrangianSolver), the re-basing
// (ilqg_setup_next_receding_horizon) and the integration along the plan (ilqg_integrate_plan) all
// run on the device.  The optional `now` argument (seconds, monotonic) replaces the wall clock
// (and lifts the solver's own wall-clock budget), which makes a run reproducible; the default is
// std::chrono::system_clock and a budget of planner_runtime per solve, like the reference.
#ifndef ILQGAMES_B200_EXAMPLES_RECEDING_HORIZON_SIMULATOR_H
#define ILQGAMES_B200_EXAMPLES_RECEDING_HORIZON_SIMULATOR_H

#include <ilqgames/b200/solvers.h>
#include <ilqgames/solver/solution_splicer.h>

#include <chrono>
#include <functional>
#include <memory>
#include <vector>

namespace ilqgames {

inline std::vector<std::shared_ptr<const SolverLog>> RecedingHorizonSimulator(
    Time final_time, Time planner_runtime, GameSolver* solver, const std::function<Time()>& now = {}) {
  CHECK_NOTNULL(solver);
  const auto wall = [] {
    return std::chrono::duration<Time>(std::chrono::system_clock::now().time_since_epoch()).count();
  };
  const std::function<Time()> clock = now ? now : std::function<Time()>(wall);
  Problem& problem = solver->GetProblem();

  // one timed solver call; the log joins the list
  std::vector<std::shared_ptr<const SolverLog>> logs;
  auto timed_solve = [&](Time budget) {
    const Time started = clock();
    bool success = false;
    logs.push_back(solver->Solve(&success, budget));
    return std::make_pair(clock() - started, success);
  };

  const auto first = timed_solve(constants::kInfinity);
  CHECK(first.second) << "the initial solve failed";  // :79
  VLOG(1) << "Solved initial problem in " << first.first << " seconds, with " << logs.back()->NumIterates()
          << " iterations.";

  SolutionSplicer splicer(*logs.front());
  VectorXf x(problem.InitialState());
  Time t = splicer.CurrentOperatingPoint().t0;
  // carry x over [from, to] along the running plan
  auto drive = [&](Time from, Time to) {
    x = problem.Dynamics()->Integrate(from, to, x, splicer.CurrentOperatingPoint(), splicer.CurrentStrategies());
  };

  constexpr Time kDriveTime = 0.25;  // kExtraTime, :93
  for (;;) {
    t += kDriveTime;
    // stop at the end of the run, or when the plan cannot cover the next planning interval (:96-98)
    if (t >= final_time || !splicer.ContainsTime(t + planner_runtime + time::kTimeStep)) break;
    drive(t - kDriveTime, t);

    // the running plan becomes the problem's warm start, re-based to where the solve will end
    problem.OverwriteSolution(splicer.CurrentOperatingPoint(), splicer.CurrentStrategies());
    problem.SetUpNextRecedingHorizon(x, t, planner_runtime);
    // with a scripted clock the solver's own wall-clock cut-off is lifted as well: the run is meant
    // to be reproducible, and a loaded host must not decide how many iterations a solve gets
    const Time elapsed = timed_solve(now ? constants::kInfinity : planner_runtime).first;
    CHECK_LE(elapsed, planner_runtime);  // :118
    VLOG(1) << "t = " << t << ": Solved warm-started problem in " << elapsed << " seconds.";

    t += elapsed;
    if (t >= final_time || !splicer.ContainsTime(t)) break;
    drive(t - elapsed, t);  // the world moved while the solver ran
    if (logs.back()->WasConverged()) splicer.Splice(*logs.back());
  }
  return logs;
}

}  // namespace ilqgames

#endif
