// host_api_test.cpp -- the C++ host classes (include/ilqgames/**) exercised the way the
// reference's own tests and executables use them; links against any implementation of ilqg.h
// (the CPU oracle in `-m "not gpu"` runs, libilqg_b200.so on the GPU box).
//
//   host_api_test <out.bin>      writes named float arrays that tests/test_host_api.py compares
//                                with the same computation driven through ctypes
// Self-checking parts exit non-zero on failure.
#include <ilqgames/solver/augmented_lagrangian_solver.h>
#include <ilqgames/solver/ilq_solver.h>
#include <ilqgames/solver/lq_feedback_solver.h>
#include <ilqgames/examples/receding_horizon_simulator.h>
#include <ilqgames/solver/solution_splicer.h>

#include "../../examples/cpp/intersection_problem.h"

#ifdef DROPIN_REFERENCE_EXAMPLE
// the reference's own problem definition, compiled unchanged (tests/test_host_api.py builds this
// variant only where /root/reference exists)
#include <ilqgames/examples/air_3d_example.h>
#include <ilqgames/examples/dubins_origin_example.h>
#include <ilqgames/examples/modified_air_3d_example.h>
#include <ilqgames/examples/modified_three_player_intersection_example.h>
#include <ilqgames/examples/skeleton_example.h>
#include <ilqgames/examples/three_player_intersection_reachability_example.h>
#include <ilqgames/examples/one_player_reachability_example.h>
#include <ilqgames/examples/roundabout_merging_example.h>
#include <ilqgames/examples/three_player_collision_avoidance_reachability_example.h>
#include <ilqgames/examples/three_player_intersection_example.h>
#include <ilqgames/examples/three_player_overtaking_example.h>
#include <ilqgames/examples/two_player_collision_avoidance_reachability_example.h>
#include <ilqgames/examples/two_player_collision_example.h>
#include <ilqgames/examples/two_player_reachability_example.h>
#endif

#include <cstdio>
#include <cstdlib>

using namespace ilqgames;

namespace {

FILE* g_out = nullptr;

void Dump(const char* name, const float* data, size_t count) {
  std::fprintf(g_out, "%s %zu\n", name, count);
  std::fwrite(data, sizeof(float), count, g_out);
}
void Dump(const char* name, const std::vector<float>& v) { Dump(name, v.data(), v.size()); }

std::vector<float> Flatten(const OperatingPoint& op) {
  std::vector<float> out;
  for (size_t k = 0; k < op.xs.size(); k++) {
    for (long a = 0; a < op.xs[k].size(); a++) out.push_back(op.xs[k](a));
    for (const VectorXf& u : op.us[k])
      for (long a = 0; a < u.size(); a++) out.push_back(u(a));
  }
  return out;
}
std::vector<float> Flatten(const std::vector<Strategy>& strategies) {
  std::vector<float> out;
  for (size_t k = 0; k < strategies[0].Ps.size(); k++)
    for (const Strategy& s : strategies) {
      for (long r = 0; r < s.Ps[k].rows(); r++)
        for (long c = 0; c < s.Ps[k].cols(); c++) out.push_back(s.Ps[k](r, c));
      for (long r = 0; r < s.alphas[k].size(); r++) out.push_back(s.alphas[k](r));
    }
  return out;
}

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "%s:%d EXPECT failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

// ---- the LQ game of the reference's test/test_lq_solver.cpp:143-177, 227-248 ------------------
class TwoPlayerPointMass1D : public MultiPlayerDynamicalSystem {
 public:
  TwoPlayerPointMass1D() : MultiPlayerDynamicalSystem(2) {}
  Dimension UDim(PlayerIndex) const override { return 1; }
  PlayerIndex NumPlayers() const override { return 2; }
  std::vector<Dimension> PositionDimensions() const override { return {0}; }
  LinearDynamicsApproximation Linearize() const {
    LinearDynamicsApproximation lin(*this);       // A = I, B = 0
    lin.A(0, 1) += 1.0f * time::kTimeStep;        // A = I + dt [[0 1] [0 0]]
    lin.Bs[0](0, 0) = 0.05f * time::kTimeStep;
    lin.Bs[0](1, 0) = 1.0f * time::kTimeStep;
    lin.Bs[1](0, 0) = 0.032f * time::kTimeStep;
    lin.Bs[1](1, 0) = 0.11f * time::kTimeStep;
    return lin;
  }
};

// LQFeedbackSolverTest.MatchesLyapunovIterations (test/test_lq_solver.cpp:292-345): the
// feedback gains at k = 0 equal the fixed point of the coupled Lyapunov iteration.
void TestLQMatchesLyapunovIterations() {
  auto dynamics = std::make_shared<TwoPlayerPointMass1D>();
  const size_t T = time::kNumTimeSteps;
  LQFeedbackSolver solver(dynamics, T);
  const LinearDynamicsApproximation lin = dynamics->Linearize();
  std::vector<LinearDynamicsApproximation> linearization(T, lin);
  std::vector<QuadraticCostApproximation> quad_k(2, QuadraticCostApproximation(2));
  quad_k[0].state.hess = MatrixXf::Identity(2, 2);
  quad_k[1].state.hess = 0.1f * MatrixXf::Identity(2, 2);
  const float R11 = 1.0f, R12 = 0.1f, R21 = 0.1f, R22 = 1.0f;
  auto one = [](float v) { MatrixXf m(1, 1); m(0, 0) = v; return m; };
  quad_k[0].control.emplace(0, SingleCostApproximation(one(R11), VectorXf::Zero(1)));
  quad_k[0].control.emplace(1, SingleCostApproximation(one(R12), VectorXf::Zero(1)));
  quad_k[1].control.emplace(0, SingleCostApproximation(one(R21), VectorXf::Zero(1)));
  quad_k[1].control.emplace(1, SingleCostApproximation(one(R22), VectorXf::Zero(1)));
  std::vector<std::vector<QuadraticCostApproximation>> quadraticization(T, quad_k);

  std::vector<VectorXf> delta_xs;
  const std::vector<Strategy> strategies = solver.Solve(linearization, quadraticization, VectorXf::Zero(2), &delta_xs);
  EXPECT(strategies.size() == 2 && delta_xs.size() == T);

  // coupled Lyapunov iteration from the test: P <- S^-1 Y, Z_i <- F' Z_i F + Q_i + sum_j P_j' R_ij P_j
  const MatrixXf &A = lin.A, &B1 = lin.Bs[0], &B2 = lin.Bs[1];
  MatrixXf Z1 = quad_k[0].state.hess, Z2 = quad_k[1].state.hess, P1(1, 2), P2(1, 2);
  for (int it = 0; it < 100; it++) {
    const MatrixXf B1Z = B1.transpose() * Z1, B2Z = B2.transpose() * Z2;
    const float s11 = R11 + (B1Z * B1)(0, 0), s12 = (B1Z * B2)(0, 0), s21 = (B2Z * B1)(0, 0), s22 = R22 + (B2Z * B2)(0, 0);
    const MatrixXf y1 = B1Z * A, y2 = B2Z * A;
    const float det = s11 * s22 - s12 * s21;
    for (int c = 0; c < 2; c++) {
      P1(0, c) = (s22 * y1(0, c) - s12 * y2(0, c)) / det;
      P2(0, c) = (-s21 * y1(0, c) + s11 * y2(0, c)) / det;
    }
    const MatrixXf F = A - B1 * P1 - B2 * P2;
    Z1 = F.transpose() * Z1 * F + quad_k[0].state.hess + R11 * (P1.transpose() * P1) + R12 * (P2.transpose() * P2);
    Z2 = F.transpose() * Z2 * F + quad_k[1].state.hess + R21 * (P1.transpose() * P1) + R22 * (P2.transpose() * P2);
  }
  const float tol = 1e-4f;  // the reference's tolerance (:313-316)
  EXPECT((strategies[0].Ps[0] - P1).cwiseAbsMax() < tol);
  EXPECT((strategies[1].Ps[0] - P2).cwiseAbsMax() < tol);
  // zero nominal and zero gradients: alphas and the delta-x forward pass vanish
  EXPECT(strategies[0].alphas[0].norm() == 0.0f && delta_xs[T - 1].norm() == 0.0f);
  Dump("lq_P1_k0", strategies[0].Ps[0].data(), 2);
  Dump("lq_P2_k0", strategies[1].Ps[0].data(), 2);
  std::printf("LQFeedbackSolver matches Lyapunov iterations: P1 = [%.6f %.6f], P2 = [%.6f %.6f]\n",
              strategies[0].Ps[0](0, 0), strategies[0].Ps[0](0, 1), strategies[1].Ps[0](0, 0), strategies[1].Ps[0](0, 1));
}

// LQOpenLoopSolverTest.NashEquilibrium (test/test_lq_solver.cpp:347-379): on the same game with
// nominal 0.5 and x0 = (1, 1), no player lowers its own open-loop cost by moving one of its
// alphas by -+0.1 (NumericalCheckLocalNashEquilibrium(..., open_loop = true),
// src/check_local_nash_equilibrium.cpp:60-135; costs as ComputeStrategyCosts with open_loop:
// T - 1 steps, state cost at the next state, src/compute_strategy_costs.cpp:61-105).
void TestLQOpenLoopIsNash() {
  auto dynamics = std::make_shared<TwoPlayerPointMass1D>();
  const size_t T = time::kNumTimeSteps;
  LQOpenLoopSolver solver(dynamics, T);
  const LinearDynamicsApproximation lin = dynamics->Linearize();
  std::vector<LinearDynamicsApproximation> linearization(T, lin);
  const float nominal = 0.5f, rel = 0.1f;
  const float Qw[2] = {1.0f, rel}, Rw[2][2] = {{1.0f, rel}, {rel, 1.0f}};
  std::vector<QuadraticCostApproximation> quad_k(2, QuadraticCostApproximation(2));
  auto one = [](float v) { MatrixXf m(1, 1); m(0, 0) = v; return m; };
  auto onev = [](float v) { VectorXf m(1); m(0) = v; return m; };
  for (int i = 0; i < 2; i++) {
    quad_k[i].state.hess = Qw[i] * MatrixXf::Identity(2, 2);
    quad_k[i].state.grad = VectorXf::Constant(2, -Qw[i] * nominal);  // weight * (x - nominal) at x = 0
    for (int j = 0; j < 2; j++)
      quad_k[i].control.emplace((PlayerIndex)j, SingleCostApproximation(one(Rw[i][j]), onev(-Rw[i][j] * nominal)));
  }
  std::vector<std::vector<QuadraticCostApproximation>> quadraticization(T, quad_k);
  VectorXf x0 = VectorXf::Constant(2, 1.0f);
  std::vector<VectorXf> delta_xs;
  const std::vector<Strategy> strategies = solver.Solve(linearization, quadraticization, x0, &delta_xs);
  EXPECT(strategies.size() == 2 && delta_xs.size() == T);
  for (size_t k = 0; k < T; k++) EXPECT(strategies[0].Ps[k].cwiseAbsMax() == 0.0f && strategies[1].Ps[k].cwiseAbsMax() == 0.0f);

  auto cost = [&](int player, int pk, int pi, double delta) {
    double x[2] = {1.0, 1.0}, total = 0.0;
    for (size_t k = 0; k + 1 < T; k++) {
      double u[2];
      for (int i = 0; i < 2; i++) u[i] = -(double)strategies[i].alphas[k](0) - (((int)k == pk && i == pi) ? delta : 0.0);
      double nx[2];
      for (int a = 0; a < 2; a++)
        nx[a] = lin.A(a, 0) * x[0] + lin.A(a, 1) * x[1] + lin.Bs[0](a, 0) * u[0] + lin.Bs[1](a, 0) * u[1];
      x[0] = nx[0]; x[1] = nx[1];
      for (int a = 0; a < 2; a++) total += 0.5 * Qw[player] * (x[a] - nominal) * (x[a] - nominal);
      for (int j = 0; j < 2; j++) total += 0.5 * Rw[player][j] * (u[j] - nominal) * (u[j] - nominal);
    }
    return total;
  };
  for (int i = 0; i < 2; i++) {
    const double base = cost(i, -1, -1, 0.0);
    for (int k = 0; k + 1 < (int)T; k++)
      for (double delta : {-0.1, 0.1}) EXPECT(cost(i, k, i, delta) >= base - 1e-6);
  }
  // delta_xs is the optimal trajectory from x0
  EXPECT(std::fabs(delta_xs[0](0) - 1.0f) < 1e-6f && std::fabs(delta_xs[0](1) - 1.0f) < 1e-6f);
  std::vector<float> al;
  for (size_t k = 0; k < T; k++) { al.push_back(strategies[0].alphas[k](0)); al.push_back(strategies[1].alphas[k](0)); }
  Dump("lq_open_loop_alphas", al);
  std::printf("LQOpenLoopSolver is an open-loop Nash equilibrium: alpha[0] = [%.6f %.6f]\n", al[0], al[1]);
}

template <typename ProblemType>
std::shared_ptr<Problem> MakeProblem() {
  auto problem = std::make_shared<ProblemType>();
  problem->Initialize();
  return problem;
}

void TestProblemDescriptor(const std::shared_ptr<Problem>& problem, const char* tag, int players = 3, int xdim = 16,
                           int costs = 18, int polylines = 3) {
  ilqg_problem_desc desc;
  EXPECT(b200::DescribeProblem(*problem, &desc));
  // a negative expectation = "whatever the source says" (descriptors taken from the reference's own source)
  EXPECT(players < 0 || (desc.num_players == players && desc.xdim == xdim && desc.num_costs == costs && desc.num_polylines == polylines));
  std::vector<float> raw(sizeof(desc) / sizeof(float));
  std::memcpy(raw.data(), &desc, sizeof(desc));
  Dump((std::string("desc_") + tag).c_str(), raw);
  Dump((std::string("x0_") + tag).c_str(), problem->InitialState().data(), (size_t)problem->InitialState().size());
}

// exec/three_player_intersection/main.cpp:109-120
SolverParams IntersectionParams() {
  SolverParams params;
  params.max_backtracking_steps = 100;
  params.max_solver_iters = 100;
  params.unconstrained_solver_max_iters = 10;
  params.linesearch = true;
  params.expected_decrease_fraction = 0.001;
  params.initial_alpha_scaling = 0.1;
  params.convergence_tolerance = 1.0;
  params.geometric_mu_scaling = 1.1;
  params.geometric_mu_downscaling = 0.5;
  params.geometric_lambda_downscaling = 0.5;
  return params;
}

void TestILQSolver(const std::shared_ptr<Problem>& problem) {
  SolverParams params = IntersectionParams();
  params.max_solver_iters = 4;
  ILQSolver solver(problem, params);
  bool success = false;
  const std::shared_ptr<SolverLog> log = solver.Solve(&success);
  EXPECT(log->NumIterates() >= 1 && log->NumIterates() <= 5);
  EXPECT(success == (log->NumIterates() == 5 || log->WasConverged()));
  EXPECT(log->FinalOperatingPoint().xs.size() == time::kNumTimeSteps);
  // every iterate starts at the problem's initial state (src/ilq_solver.cpp:88-89)
  for (size_t it = 0; it < log->NumIterates(); it++)
    EXPECT((log->State(it, 0) - problem->InitialState()).norm() == 0.0f);
  const float n_iter = (float)log->NumIterates();
  Dump("ilq_num_iterates", &n_iter, 1);
  Dump("ilq_final_op", Flatten(log->FinalOperatingPoint()));
  Dump("ilq_final_strategies", Flatten(log->FinalStrategies()));
  Dump("ilq_total_costs", log->TotalCosts());
  // SolverLog::Save: the reference's text format (src/solver_log.cpp:113-171) under $ILQGAMES_LOG_DIR
  EXPECT(log->Save(false, "host_api_test"));
  EXPECT(log->Save(true, "host_api_test_last"));
  std::printf("ILQSolver::Solve: %zu iterates, success = %d, total costs = [%.4f %.4f %.4f]\n", log->NumIterates(),
              (int)success, log->TotalCosts()[0], log->TotalCosts()[1], log->TotalCosts()[2]);

  // SolveBatch: game 0 = the problem's own x0 must reproduce Solve(); the problem is untouched
  std::vector<VectorXf> x0s(3, problem->InitialState());
  x0s[1](0) += 1.0f;
  x0s[2](7) -= 2.0f;
  const std::vector<BatchSolution> batch = solver.SolveBatch(x0s);
  EXPECT(batch.size() == 3);
  const std::vector<float> a = Flatten(batch[0].operating_point), b = Flatten(log->FinalOperatingPoint());
  EXPECT(a == b);
  EXPECT(batch[1].operating_point.xs[0](0) == x0s[1](0));
  Dump("batch_op_1", Flatten(batch[1].operating_point));
  Dump("batch_op_2", Flatten(batch[2].operating_point));
  std::printf("ILQSolver::SolveBatch: 3 games, iterations = [%d %d %d]\n", batch[0].iterations, batch[1].iterations,
              batch[2].iterations);
}

// Problem::SetUpNextRecedingHorizon (src/problem.cpp:127-186) as a receding-horizon caller uses it
// (src/receding_horizon_simulator.cpp:100-107): solve, write the solution back, re-base the plan
// on a measured state, solve again from the shifted warm start.
class IntersectionWithoutConstraints : public ilqgames_b200_examples::IntersectionProblem {
  bool WithProximityConstraints() const override { return false; }
};
class IntersectionWithLaneBoundaries : public ilqgames_b200_examples::IntersectionProblem {
  bool WithLaneBoundaries() const override { return true; }
};

void TestRecedingHorizon() {
  const std::shared_ptr<Problem> problem = MakeProblem<IntersectionWithoutConstraints>();
  EXPECT(!problem->IsConstrained());
  SolverParams params = IntersectionParams();
  params.max_solver_iters = 3;
  ILQSolver solver(problem, params);
  bool success = false;
  std::shared_ptr<SolverLog> log = solver.Solve(&success);
  EXPECT(success);
  problem->OverwriteSolution(log->FinalOperatingPoint(), log->FinalStrategies());
  const OperatingPoint plan = log->FinalOperatingPoint();
  const Time t = 0.33, planner_runtime = 0.25;
  VectorXf x = plan.xs[3];
  x(0) += 0.05f;  // the measured state is a little off the plan
  x(7) -= 0.03f;
  // MultiPlayerIntegrableSystem::Integrate(t0, t, ...) along the plan
  // (src/multi_player_integrable_system.cpp:54-83): the measured state a little later
  const VectorXf x_later = problem->Dynamics()->Integrate(t, 0.81, x, plan, log->FinalStrategies());
  EXPECT(x_later.size() == x.size());
  EXPECT((x_later - plan.xs[8]).norm() < 0.5f && (x_later - x).norm() > 0.1f);  // it moved, and stayed near the plan
  Dump("ip_out", x_later.data(), (size_t)x_later.size());
  problem->SetUpNextRecedingHorizon(x, t, planner_runtime);
  const OperatingPoint& op = problem->CurrentOperatingPoint();
  // :95 / :107 -- the new problem starts within one time step of t + planner_runtime
  EXPECT(std::fabs(t + planner_runtime - op.t0) <= time::kTimeStep);
  EXPECT(op.xs.size() == time::kNumTimeSteps && problem->CurrentStrategies()[0].Ps.size() == time::kNumTimeSteps);
  // the first plan state is one of the old plan's states (the nearest one), the ego part of the new
  // initial state is taken from it (ConcatenatedDynamicalSystem::Stitch)
  size_t first = 0;
  for (size_t k = 0; k < plan.xs.size(); k++)
    if ((plan.xs[k] - op.xs[0]).norm() == 0.0f) first = k;
  EXPECT(first >= 4 && first <= 8);
  for (int a = 0; a < 6; a++) EXPECT(problem->InitialState()(a) == op.xs[0](a));
  // shifted rows are copies; the tail is the zero-control extension
  EXPECT((plan.xs[first + 10] - op.xs[10]).norm() == 0.0f);
  const size_t T = time::kNumTimeSteps;
  for (size_t k = T - first; k < T; k++)
    for (const Strategy& st : problem->CurrentStrategies()) EXPECT(st.Ps[k].cwiseAbsMax() == 0.0f && st.alphas[k].norm() == 0.0f);
  const float t0f = (float)op.t0;
  Dump("rh_t0", &t0f, 1);
  Dump("rh_x0", problem->InitialState().data(), (size_t)problem->InitialState().size());
  Dump("rh_op", Flatten(op));
  Dump("rh_strategies", Flatten(problem->CurrentStrategies()));
  // the next solve starts from the shifted warm start at the new initial state
  log = solver.Solve(&success);
  EXPECT((log->State(0, 0) - problem->InitialState()).norm() == 0.0f);
  Dump("rh_next_op", Flatten(log->FinalOperatingPoint()));
  std::printf("Problem::SetUpNextRecedingHorizon: first time step of the new problem = %zu, t0 = %.4f\n", first, op.t0);
}

// RecedingHorizonSimulator (src/receding_horizon_simulator.cpp:65-137) with a scripted clock (every
// reading is 20 ms after the last, so each solve "takes" 20 ms): one log per solver call, each new
// problem starts about planner_runtime after the time it was set up at, from where the state
// really is; the same script gives the same run.
void TestRecedingHorizonSimulator() {
  const Time final_time = 1.5, planner_runtime = 0.5, tick = 0.02;
  std::vector<std::vector<float>> runs;
  for (int run = 0; run < 2; run++) {
    const std::shared_ptr<Problem> problem = MakeProblem<IntersectionWithoutConstraints>();
    SolverParams params = IntersectionParams();
    params.max_solver_iters = 30;
    ILQSolver solver(problem, params);
    Time fake_now = 0.0;
    const std::vector<std::shared_ptr<const SolverLog>> logs =
        RecedingHorizonSimulator(final_time, planner_runtime, &solver, [&] { return fake_now += tick; });
    // t advances by 0.25 s of driving + 20 ms of solving per round: rounds at t = 0.25, 0.52, ..., 1.33
    EXPECT(logs.size() == 6);
    std::vector<float> summary;
    for (size_t k = 0; k < logs.size(); k++) {
      const OperatingPoint& op = logs[k]->FinalOperatingPoint();
      EXPECT(op.xs.size() == time::kNumTimeSteps);
      if (k > 0) {
        const Time set_up_at = 0.25 * k + tick * (k - 1);
        EXPECT(std::fabs(set_up_at + planner_runtime - op.t0) <= time::kTimeStep);  // src/problem.cpp:123
        EXPECT(op.t0 > logs[k - 1]->FinalOperatingPoint().t0);
      }
      summary.push_back((float)op.t0);
      summary.push_back((float)logs[k]->NumIterates());
      summary.push_back(logs[k]->WasConverged() ? 1.0f : 0.0f);
      for (long a = 0; a < op.xs[0].size(); a++) summary.push_back(op.xs[0](a));
      for (long a = 0; a < op.xs[50].size(); a++) summary.push_back(op.xs[50](a));
    }
    // the cars drive on: player 1's position at the start of each new problem keeps moving
    const OperatingPoint& first = logs.front()->FinalOperatingPoint();
    const OperatingPoint& last = logs.back()->FinalOperatingPoint();
    EXPECT((last.xs[0].head(2) - first.xs[0].head(2)).norm() > 1.0f);
    runs.push_back(summary);
    if (run == 0) {
      Dump("sim_summary", summary);
      std::printf("RecedingHorizonSimulator: %zu solver calls, last problem starts at t0 = %.3f\n", logs.size(),
                  last.t0);
    }
  }
  EXPECT(runs[0] == runs[1]);
}

// SolutionSplicer (src/solution_splicer.cpp:57-131) on the synthetic logs of
// oracle/ref_driver.cpp:ilqg_ref_splice (same fill formula), for a new horizon that starts before
// and after the five-step keep window; tests/test_host_api.py compares the dumps with what the
// reference's own SolutionSplicer produced (tests/golden/ref_splice.npz).
class SpliceDynamics : public MultiPlayerDynamicalSystem {
 public:
  SpliceDynamics() : MultiPlayerDynamicalSystem(3) {}
  Dimension UDim(PlayerIndex i) const override { return i == 0 ? 1 : 2; }
  PlayerIndex NumPlayers() const override { return 2; }
  std::vector<Dimension> PositionDimensions() const override { return {0, 1}; }
};

void FillLog(SolverLog* log, const std::shared_ptr<const MultiPlayerIntegrableSystem>& dyn, Time t0, float tag) {
  const size_t T = time::kNumTimeSteps;
  OperatingPoint op(T, t0, dyn);
  std::vector<Strategy> st;
  for (PlayerIndex i = 0; i < 2; i++) st.emplace_back(T, dyn->XDim(), dyn->UDim(i));
  for (size_t k = 0; k < T; k++) {
    for (int d = 0; d < 3; d++) op.xs[k](d) = tag * 1000.f + (float)k + 0.01f * d;
    int c = 0;
    for (PlayerIndex i = 0; i < 2; i++)
      for (int d = 0; d < dyn->UDim(i); d++, c++) {
        op.us[k][i](d) = -(tag * 1000.f + (float)k + 0.01f * c);
        st[i].alphas[k](d) = tag * 100.f + 0.5f * (float)k + 0.01f * c;
        st[i].Ps[k](d, 0) = tag * 10.f + 0.25f * (float)k + (float)i;
      }
  }
  log->AddSolverIterate(op, st, std::vector<float>(2, 0.f), 0.0, true);
}

void TestSolutionSplicer() {
  const std::shared_ptr<const MultiPlayerIntegrableSystem> dyn(new SpliceDynamics);
  int which = 0;
  for (const Time new_t0 : {0.3, 1.2, 0.0}) {
    SolverLog stored, fresh;
    FillLog(&stored, dyn, 0.0, 1.f);
    FillLog(&fresh, dyn, new_t0, 2.f);
    SolutionSplicer splicer(stored);
    EXPECT(splicer.ContainsTime(5.0) && !splicer.ContainsTime(10.5));
    splicer.Splice(fresh);
    const OperatingPoint& op = splicer.CurrentOperatingPoint();
    std::vector<float> xs, us, al, p00;
    for (size_t k = 0; k < op.xs.size(); k++) {
      for (int d = 0; d < 3; d++) xs.push_back(op.xs[k](d));
      for (PlayerIndex i = 0; i < 2; i++) {
        for (int d = 0; d < dyn->UDim(i); d++) {
          us.push_back(op.us[k][i](d));
          al.push_back(splicer.CurrentStrategies()[i].alphas[k](d));
        }
        p00.push_back(splicer.CurrentStrategies()[i].Ps[k](0, 0));
      }
    }
    const std::string tag = "splice" + std::to_string(which++) + "_";
    const float t0f = (float)op.t0;
    Dump((tag + "t0").c_str(), &t0f, 1);
    Dump((tag + "xs").c_str(), xs);
    Dump((tag + "us").c_str(), us);
    Dump((tag + "alphas").c_str(), al);
    Dump((tag + "P00").c_str(), p00);
  }
  std::printf("SolutionSplicer: spliced 3 synthetic horizons\n");
}

void TestAugmentedLagrangianSolver(const std::shared_ptr<Problem>& problem) {
  SolverParams params = IntersectionParams();
  params.max_solver_iters = 30;  // NumIterates cap of the outer loop
  AugmentedLagrangianSolver solver(problem, params);
  bool success = false;
  const std::shared_ptr<SolverLog> log = solver.Solve(&success);
  EXPECT(log->NumIterates() >= 2);   // the constrained game goes around the outer loop
  EXPECT(solver.NumIterates() >= (int)log->NumIterates());
  const float stats[3] = {(float)log->NumIterates(), (float)solver.NumIterates(), success ? 1.0f : 0.0f};
  Dump("al_stats", stats, 3);
  Dump("al_final_op", Flatten(log->FinalOperatingPoint()));
  std::printf("AugmentedLagrangianSolver::Solve: %zu inner solves, %d iterates, success = %d\n", log->NumIterates(),
              solver.NumIterates(), (int)success);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <out.bin>\n", argv[0]);
    return 2;
  }
  g_out = std::fopen(argv[1], "wb");
  if (!g_out) return 2;
  TestLQMatchesLyapunovIterations();
  TestLQOpenLoopIsNash();
  const std::shared_ptr<Problem> problem = MakeProblem<ilqgames_b200_examples::IntersectionProblem>();
  TestProblemDescriptor(problem, "own");
  // the same game with its lane boundaries as Polyline2SignedDistanceConstraints (24 records)
  TestProblemDescriptor(MakeProblem<IntersectionWithLaneBoundaries>(), "lanes", 3, 16, 24, 3);
#ifdef DROPIN_REFERENCE_EXAMPLE
  TestProblemDescriptor(MakeProblem<ThreePlayerIntersectionExample>(), "reference");
  // src/roundabout_merging_example.cpp (+ roundabout_lane_center.cpp, initialize_along_route.cpp) and
  // src/air_3d_example.cpp (+ draw_shapes.cpp), all compiled unchanged
  TestProblemDescriptor(MakeProblem<RoundaboutMergingExample>(), "roundabout", 4, 24, 44, 4);
  TestProblemDescriptor(MakeProblem<Air3DExample>(), "air3d", 2, 3, 8, 1);
  // src/three_player_overtaking_example.cpp: a fourth example, made of the same record kinds (n = 18)
  TestProblemDescriptor(MakeProblem<ThreePlayerOvertakingExample>(), "overtaking", 3, 18, 22, 2);
  // src/two_player_collision_example.cpp: FinalTimeCost-wrapped goal costs -> records with a time gate
  TestProblemDescriptor(MakeProblem<TwoPlayerCollisionExample>(), "collision", 2, 12, 22, 6);
  // src/two_player_collision_avoidance_reachability_example.cpp: SinglePlayerCar5D + SignedDistanceCost
  TestProblemDescriptor(MakeProblem<TwoPlayerCollisionAvoidanceReachabilityExample>(), "reachability2", 2, 10, 4, 0);
  // src/three_player_collision_avoidance_reachability_example.cpp: ExtremeValueCost -> grouped records
  TestProblemDescriptor(MakeProblem<ThreePlayerCollisionAvoidanceReachabilityExample>(), "reachability3", 3, 15, 21, 0);
  // src/one_player_reachability_example.cpp: a single player on a SinglePlayerDubinsCar
  TestProblemDescriptor(MakeProblem<OnePlayerReachabilityExample>(), "reachability1", 1, 3, 4, 1);
  // src/dubins_origin_example.cpp: two Dubins cars, QuadraticDifferenceCost
  TestProblemDescriptor(MakeProblem<DubinsOriginExample>(), "dubins_origin", 2, 6, 5, 0);
  // src/two_player_reachability_example.cpp: TwoPlayerUnicycle4D, a coupled system like Air3D
  TestProblemDescriptor(MakeProblem<TwoPlayerReachabilityExample>(), "reachability_2p", 2, 4, 4, 1);
  // src/modified_air_3d_example.cpp: two SinglePlayerPointMass2D
  TestProblemDescriptor(MakeProblem<ModifiedAir3DExample>(), "modified_air3d", 2, 8, 4, 0);
  // three more whose descriptors tests/golden/make_ref_golden.py takes from here instead of from a
  // hand-written builder in problems.py (they are made of record kinds that exist already)
  TestProblemDescriptor(MakeProblem<ModifiedThreePlayerIntersectionExample>(), "modified_intersection", -1);
  TestProblemDescriptor(MakeProblem<SkeletonExample>(), "skeleton", -1);
  TestProblemDescriptor(MakeProblem<ThreePlayerIntersectionReachabilityExample>(), "intersection_reachability", -1);
#endif
  TestILQSolver(problem);
  TestAugmentedLagrangianSolver(problem);
  TestRecedingHorizon();
  TestRecedingHorizonSimulator();
  TestSolutionSplicer();
  std::fclose(g_out);
  std::printf("host_api_test: all checks passed\n");
  return 0;
}
