// The counterpart of the reference's exec/three_player_intersection/main.cpp (:103-143) without the
// GUI: build the game, solve it with the augmented-Lagrangian solver, save the logs in the
// reference's format -- then what the reference cannot do: the same solve for a batch of perturbed
// initial states in one call (ILQSolver::SolveBatch, one game per CUDA warp group).
//
//   make -C examples/cpp            (links libilqg_b200.so; ILQG_LIB=... for another ilqg.h library)
//   examples/cpp/three_player_intersection [batch=256] [experiment_name]
#include <ilqgames/solver/augmented_lagrangian_solver.h>
#include <ilqgames/solver/ilq_solver.h>

#include "intersection_problem.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace ilqgames;

int main(int argc, char** argv) {
  const int batch = argc > 1 ? std::atoi(argv[1]) : 256;
  const std::string experiment = argc > 2 ? argv[2] : "three_player_intersection";

  // exec/three_player_intersection/main.cpp:109-120
  SolverParams params;
  params.max_backtracking_steps = 100;
  params.linesearch = true;
  params.expected_decrease_fraction = 0.001;
  params.initial_alpha_scaling = 0.1;
  params.convergence_tolerance = 1.0;
  params.unconstrained_solver_max_iters = 10;
  params.geometric_mu_scaling = 1.1;
  params.geometric_mu_downscaling = 0.5;
  params.geometric_lambda_downscaling = 0.5;

  auto problem = std::make_shared<ilqgames_b200_examples::IntersectionProblem>();
  problem->Initialize();

  // :123-129 -- one game, constraints through the augmented Lagrangian
  AugmentedLagrangianSolver solver(problem, params);
  auto start = std::chrono::steady_clock::now();
  bool success = false;
  const std::shared_ptr<const SolverLog> log = solver.Solve(&success);
  double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
  std::printf("AugmentedLagrangianSolver: %zu iterates in %.3f s, success = %d, total costs =", log->NumIterates(), seconds,
              (int)success);
  for (float c : log->TotalCosts()) std::printf(" %.3f", c);
  std::printf("\n");
  if (log->Save(true, experiment)) std::printf("saved the final iterate under $ILQGAMES_LOG_DIR/%s\n", experiment.c_str());

  // a batch of games: the same intersection from perturbed initial states, one inner solve each
  std::mt19937 rng(4096);
  std::uniform_real_distribution<float> metres(-2.0f, 2.0f), scale(0.8f, 1.2f);
  std::vector<VectorXf> x0s(batch, problem->InitialState());
  for (VectorXf& x0 : x0s)
    for (int first : {0, 6, 12}) {
      x0(first) += metres(rng);
      x0(first + 1) += metres(rng);
      x0(first == 12 ? first + 3 : first + 4) *= scale(rng);
    }
  SolverParams inner = params;
  inner.max_solver_iters = params.unconstrained_solver_max_iters;
  ILQSolver batched(problem, inner);
  start = std::chrono::steady_clock::now();
  const std::vector<BatchSolution> solutions = batched.SolveBatch(x0s);
  seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
  long iterations = 0;
  int succeeded = 0;
  for (const BatchSolution& s : solutions) {
    iterations += s.iterations;
    succeeded += s.success;
  }
  std::printf("ILQSolver::SolveBatch: %d games, %ld iLQ iterations in %.3f s (%.0f instance-iterations/s), %d succeeded\n",
              batch, iterations, seconds, iterations / seconds, succeeded);
  return 0;
}
