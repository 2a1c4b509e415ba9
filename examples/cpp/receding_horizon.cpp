// The counterpart of the reference's exec/receding_horizon_example (RecedingHorizonSimulator,
// src/receding_horizon_simulator.cpp:65-137) without the GUI: plan, drive a quarter second along
// the plan, re-plan from where the cars are for a horizon that starts when the solver will be done,
// splice.  The proximity constraints are left out: the reference cannot re-base a constrained
// problem either (include/ilqg.h, ilqg_setup_next_receding_horizon).
//
//   examples/cpp/receding_horizon [final_time=3.0] [planner_runtime=0.25]
#include <ilqgames/examples/receding_horizon_simulator.h>
#include <ilqgames/solver/ilq_solver.h>

#include "intersection_problem.h"

#include <cstdio>
#include <cstdlib>

using namespace ilqgames;

namespace {
class UnconstrainedIntersection : public ilqgames_b200_examples::IntersectionProblem {
  bool WithProximityConstraints() const override { return false; }
};
}  // namespace

int main(int argc, char** argv) {
  const Time final_time = argc > 1 ? std::atof(argv[1]) : 3.0;
  const Time planner_runtime = argc > 2 ? std::atof(argv[2]) : 0.25;

  SolverParams params;
  params.max_backtracking_steps = 100;
  params.linesearch = true;
  params.expected_decrease_fraction = 0.001;
  params.initial_alpha_scaling = 0.1;
  params.convergence_tolerance = 1.0;
  params.max_solver_iters = 50;

  auto problem = std::make_shared<UnconstrainedIntersection>();
  problem->Initialize();
  ILQSolver solver(problem, params);
  const std::vector<std::shared_ptr<const SolverLog>> logs = RecedingHorizonSimulator(final_time, planner_runtime, &solver);

  std::printf("%zu solver calls up to t = %.2f s\n", logs.size(), final_time);
  for (size_t k = 0; k < logs.size(); k++) {
    const OperatingPoint& op = logs[k]->FinalOperatingPoint();
    std::printf("  call %zu: horizon starts at t0 = %.2f s, %zu iterates, converged = %d, car 1 at (%.2f, %.2f)\n", k, op.t0,
                logs[k]->NumIterates(), (int)logs[k]->WasConverged(), op.xs[0](0), op.xs[0](1));
  }
  return 0;
}
