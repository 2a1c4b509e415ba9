// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not part of the product.
//
// A thin extern "C" driver around the UNMODIFIED reference sources (compiled in place from
// /root/reference/src by oracle/Makefile, target `ref`, against oracle/ref_shim's stand-ins for
// Eigen3/glog/gflags) so that tests/golden/make_ref_golden.py can run the reference's own
// ILQSolver::Solve / AugmentedLagrangianSolver::Solve / LQFeedbackSolver::Solve on chosen initial
// states and commit what they produce as fixtures.  No algorithm lives here: this file only
// constructs the reference's example Problems, calls the reference's solvers, and copies their
// logs into flat arrays.
//
// Layouts written (all fp32, row-major over the listed indices; matrices row-major [row][col]):
//   xs     [iterate][T][n]        us     [iterate][T][M]   (players' controls concatenated)
//   Ps     [T][M][n]              alphas [T][M]            (final strategies)
//   costs  [N]                                             (final iterate)
#include <ilqgames/constraint/constraint.h>
#include <ilqgames/constraint/polyline2_signed_distance_constraint.h>
#include <ilqgames/geometry/polyline2.h>
#include <ilqgames/cost/player_cost.h>
#include <ilqgames/examples/air_3d_example.h>
#include <ilqgames/examples/dubins_origin_example.h>
#include <ilqgames/examples/modified_air_3d_example.h>
#include <ilqgames/examples/modified_three_player_intersection_example.h>
#include <ilqgames/examples/one_player_reachability_example.h>
#include <ilqgames/examples/skeleton_example.h>
#include <ilqgames/examples/three_player_intersection_reachability_example.h>
#include <ilqgames/examples/roundabout_lane_center.h>
#include <ilqgames/examples/roundabout_merging_example.h>
#include <ilqgames/geometry/draw_shapes.h>
#include <ilqgames/examples/three_player_collision_avoidance_reachability_example.h>
#include <ilqgames/examples/three_player_intersection_example.h>
#include <ilqgames/examples/three_player_overtaking_example.h>
#include <ilqgames/examples/two_player_collision_avoidance_reachability_example.h>
#include <ilqgames/examples/two_player_collision_example.h>
#include <ilqgames/examples/two_player_reachability_example.h>
#include <ilqgames/solver/augmented_lagrangian_solver.h>
#include <ilqgames/solver/ilq_solver.h>
#include <ilqgames/solver/lq_feedback_solver.h>
#include <ilqgames/solver/lq_open_loop_solver.h>
#include <ilqgames/solver/problem.h>
#include <ilqgames/solver/solution_splicer.h>
#include <ilqgames/solver/solver_params.h>
#include <ilqgames/utils/solver_log.h>
#include <ilqgames/utils/relative_time_tracker.h>
#include <ilqgames/utils/types.h>

#include <cstdint>
#include <limits>
#include <memory>
#include <vector>

using namespace ilqgames;

extern "C" {

// Mirror of include/ilqgames/solver/solver_params.h (the fields the in-scope solvers read).
struct ilqg_ref_params {
  float convergence_tolerance;
  int32_t max_solver_iters;
  int32_t linesearch;
  float initial_alpha_scaling;
  float geometric_alpha_scaling;
  int32_t max_backtracking_steps;
  float expected_decrease_fraction;
  int32_t open_loop;
  int32_t unconstrained_solver_max_iters;
  float geometric_mu_scaling;
  float geometric_mu_downscaling;
  float geometric_lambda_downscaling;
  float constraint_error_tolerance;
};

enum { ILQG_REF_INTERSECTION = 0, ILQG_REF_ROUNDABOUT = 1, ILQG_REF_AIR3D = 2, ILQG_REF_OVERTAKING = 3, ILQG_REF_COLLISION = 4, ILQG_REF_REACHABILITY2 = 5, ILQG_REF_REACHABILITY3 = 6, ILQG_REF_REACHABILITY1 = 7, ILQG_REF_DUBINS_ORIGIN = 8, ILQG_REF_REACHABILITY_2P = 9, ILQG_REF_MODIFIED_AIR3D = 10,
       ILQG_REF_MODIFIED_INTERSECTION = 11, ILQG_REF_SKELETON = 12, ILQG_REF_INTERSECTION_REACHABILITY = 13 };
enum { ILQG_REF_ILQ = 0, ILQG_REF_AL = 1 };

}  // extern "C"

namespace {

std::shared_ptr<Problem> MakeProblem(int which) {
  std::shared_ptr<Problem> p;
  if (which == ILQG_REF_INTERSECTION) p = std::make_shared<ThreePlayerIntersectionExample>();
  else if (which == ILQG_REF_ROUNDABOUT) p = std::make_shared<RoundaboutMergingExample>();
  else if (which == ILQG_REF_AIR3D) p = std::make_shared<Air3DExample>();
  else if (which == ILQG_REF_OVERTAKING) p = std::make_shared<ThreePlayerOvertakingExample>();
  else if (which == ILQG_REF_COLLISION) p = std::make_shared<TwoPlayerCollisionExample>();
  else if (which == ILQG_REF_REACHABILITY2) p = std::make_shared<TwoPlayerCollisionAvoidanceReachabilityExample>();
  else if (which == ILQG_REF_REACHABILITY3) p = std::make_shared<ThreePlayerCollisionAvoidanceReachabilityExample>();
  else if (which == ILQG_REF_REACHABILITY1) p = std::make_shared<OnePlayerReachabilityExample>();
  else if (which == ILQG_REF_DUBINS_ORIGIN) p = std::make_shared<DubinsOriginExample>();
  else if (which == ILQG_REF_REACHABILITY_2P) p = std::make_shared<TwoPlayerReachabilityExample>();
  else if (which == ILQG_REF_MODIFIED_AIR3D) p = std::make_shared<ModifiedAir3DExample>();
  else if (which == ILQG_REF_MODIFIED_INTERSECTION) p = std::make_shared<ModifiedThreePlayerIntersectionExample>();
  else if (which == ILQG_REF_SKELETON) p = std::make_shared<SkeletonExample>();
  else if (which == ILQG_REF_INTERSECTION_REACHABILITY) p = std::make_shared<ThreePlayerIntersectionReachabilityExample>();
  else return nullptr;
  p->Initialize();
  return p;
}

SolverParams ToParams(const ilqg_ref_params& q) {
  SolverParams p;
  p.convergence_tolerance = q.convergence_tolerance;
  p.max_solver_iters = (size_t)q.max_solver_iters;
  p.linesearch = q.linesearch != 0;
  p.initial_alpha_scaling = q.initial_alpha_scaling;
  p.geometric_alpha_scaling = q.geometric_alpha_scaling;
  p.max_backtracking_steps = (size_t)q.max_backtracking_steps;
  p.expected_decrease_fraction = q.expected_decrease_fraction;
  p.open_loop = q.open_loop != 0;
  p.unconstrained_solver_max_iters = (size_t)q.unconstrained_solver_max_iters;
  p.geometric_mu_scaling = q.geometric_mu_scaling;
  p.geometric_mu_downscaling = q.geometric_mu_downscaling;
  p.geometric_lambda_downscaling = q.geometric_lambda_downscaling;
  p.constraint_error_tolerance = q.constraint_error_tolerance;
  // Keep multipliers after an AL solve so they can be read back (they are reset by hand below).
  p.reset_lambdas = false;
  p.reset_mu = false;
  return p;
}

void ResetMultipliers(Problem& problem) {
  for (auto& pc : problem.PlayerCosts()) {
    for (const auto& c : pc.StateConstraints()) c->ScaleLambdas(constants::kDefaultLambda);
    for (const auto& pair : pc.ControlConstraints()) pair.second->ScaleLambdas(constants::kDefaultLambda);
  }
  Constraint::GlobalMu() = constants::kDefaultMu;
}

void CopyIterate(const SolverLog& log, size_t it, const MultiPlayerIntegrableSystem& dyn, float* xs,
                 float* us) {
  const size_t T = time::kNumTimeSteps;
  const int n = dyn.XDim(), M = dyn.TotalUDim(), N = dyn.NumPlayers();
  for (size_t k = 0; k < T; k++) {
    if (xs) for (int d = 0; d < n; d++) xs[(it * T + k) * n + d] = log.State(it, k, d);
    int off = 0;
    for (int i = 0; i < N; i++) {
      const int m = dyn.UDim(i);
      if (us) for (int d = 0; d < m; d++) us[(it * T + k) * M + off + d] = log.Control(it, k, (PlayerIndex)i, d);
      off += m;
    }
  }
}

// SolverLog's per-iterate strategy accessors are declared but not linkable in the reference
// (`inline` definitions in src/solver_log.cpp:173-197), so only the final strategies are copied.
void CopyStrategies(const std::vector<Strategy>& st, const MultiPlayerIntegrableSystem& dyn, float* Ps,
                    float* alphas) {
  const size_t T = time::kNumTimeSteps;
  const int n = dyn.XDim(), M = dyn.TotalUDim(), N = dyn.NumPlayers();
  for (size_t k = 0; k < T; k++) {
    int off = 0;
    for (int i = 0; i < N; i++) {
      const int m = dyn.UDim(i);
      for (int d = 0; d < m; d++) {
        if (alphas) alphas[k * M + off + d] = st[i].alphas[k](d);
        if (Ps) for (int c = 0; c < n; c++) Ps[(k * M + off + d) * n + c] = st[i].Ps[k](d, c);
      }
      off += m;
    }
  }
}

}  // namespace

extern "C" {

int ilqg_ref_dims(int which, int* n, int* M, int* N, int* T, int* num_constraints) {
  auto p = MakeProblem(which);
  if (!p) return -1;
  *n = p->Dynamics()->XDim();
  *M = p->Dynamics()->TotalUDim();
  *N = p->Dynamics()->NumPlayers();
  *T = (int)time::kNumTimeSteps;
  int c = 0;
  for (auto& pc : p->PlayerCosts()) c += (int)pc.StateConstraints().size() + (int)pc.ControlConstraints().size();
  *num_constraints = c;
  return 0;
}

int ilqg_ref_x0(int which, float* x0) {
  auto p = MakeProblem(which);
  if (!p) return -1;
  for (int d = 0; d < p->Dynamics()->XDim(); d++) x0[d] = p->InitialState()(d);
  return 0;
}

// RoundaboutLaneCenter (src/roundabout_lane_center.cpp) / DrawCircle (src/draw_shapes.cpp): the
// polylines the examples build their lane and target costs on.  Returns the point count.
int ilqg_ref_roundabout_lane(float entrance_angle, float exit_angle, float distance, float* pts, int max_pts) {
  const PointList2 p = RoundaboutLaneCenter(entrance_angle, exit_angle, distance);
  for (size_t i = 0; i < p.size() && (int)i < max_pts; i++) { pts[2 * i] = p[i].x(); pts[2 * i + 1] = p[i].y(); }
  return (int)p.size();
}

int ilqg_ref_draw_circle(float cx, float cy, float radius, int num_segments, float* pts, int max_pts) {
  const Polyline2 c = DrawCircle(Point2(cx, cy), radius, (size_t)num_segments);
  const auto& segs = c.Segments();
  int k = 0;
  for (size_t i = 0; i < segs.size() && k < max_pts; i++, k++) { pts[2 * k] = segs[i].FirstPoint().x(); pts[2 * k + 1] = segs[i].FirstPoint().y(); }
  if (k < max_pts && !segs.empty()) { pts[2 * k] = segs.back().SecondPoint().x(); pts[2 * k + 1] = segs.back().SecondPoint().y(); k++; }
  return k;
}

// One game from x0 (zero initial operating point and strategies, as Problem::Initialize leaves
// them).  solver = ILQG_REF_ILQ: ILQSolver::Solve(success, inf) with the given multipliers state
// (lambda = 0, mu as given); ILQG_REF_AL: AugmentedLagrangianSolver::Solve(success, inf).
// Logs up to max_log iterates (the log's iterate 0 is the initial rollout).  lambdas_out is
// [constraint][T] in the order players -> state constraints -> control constraints.
int ilqg_ref_solve(int which, int solver, const float* x0, const ilqg_ref_params* q, float mu0,
                   int max_log, float* xs, float* us, float* Ps, float* alphas, float* costs,
                   int* num_iterates, int* success, float* lambdas_out, float* mu_out) {
  auto problem = MakeProblem(which);
  if (!problem) return -1;
  const auto& dyn = *problem->Dynamics();
  VectorXf x(dyn.XDim());
  for (int d = 0; d < dyn.XDim(); d++) x(d) = x0[d];
  problem->ResetInitialState(x);
  ResetMultipliers(*problem);
  Constraint::GlobalMu() = mu0;

  const SolverParams params = ToParams(*q);
  bool ok = false;
  std::shared_ptr<SolverLog> log;
  const Time inf = std::numeric_limits<Time>::infinity();
  if (solver == ILQG_REF_ILQ) {
    ILQSolver s(problem, params);
    log = s.Solve(&ok, inf);
  } else {
    // Not infinity: the outer loop's guard is `elapsed < max_runtime - bound` with elapsed starting
    // at max_runtime / max_solver_iters (src/augmented_lagrangian_solver.cpp:83-110), which is
    // false for inf < inf, so an infinite budget would skip the multiplier loop altogether.
    AugmentedLagrangianSolver s(problem, params);
    log = s.Solve(&ok, 1e9);
  }
  const int iterates = (int)log->NumIterates();
  *num_iterates = iterates;
  *success = ok ? 1 : 0;
  for (int it = 0; it < iterates && it < max_log; it++)
    CopyIterate(*log, (size_t)it, dyn, xs, us);
  CopyStrategies(log->FinalStrategies(), dyn, Ps, alphas);
  if (costs) {
    // SolverLog keeps total costs only for the last iterate through its public interface.
    const auto c = log->TotalCosts();
    for (size_t i = 0; i < c.size(); i++) costs[i] = c[i];
  }
  if (lambdas_out) {
    size_t c = 0;
    const size_t T = time::kNumTimeSteps;
    for (auto& pc : problem->PlayerCosts()) {
      for (const auto& con : pc.StateConstraints()) {
        for (size_t k = 0; k < T; k++) lambdas_out[c * T + k] = con->Lambda(time::kTimeStep * (Time)k + 1e-6);
        c++;
      }
      for (const auto& pair : pc.ControlConstraints()) {
        for (size_t k = 0; k < T; k++) lambdas_out[c * T + k] = pair.second->Lambda(time::kTimeStep * (Time)k + 1e-6);
        c++;
      }
    }
  }
  if (mu_out) *mu_out = Constraint::GlobalMu();
  ResetMultipliers(*problem);
  return 0;
}

// ILQSolver::Solve from x0, Problem::OverwriteSolution with the final iterate, then
// Problem::SetUpNextRecedingHorizon(x_meas, t, planner_runtime) (src/problem.cpp:127-186).
// Returns what the Problem holds afterwards: InitialState (x0_out [n]), CurrentOperatingPoint
// (xs [T][n], us [T][M], t0), CurrentStrategies (Ps [T][M][n], alphas [T][M]).
int ilqg_ref_receding_horizon(int which, const float* x0, const ilqg_ref_params* q, const float* x_meas,
                              double t, double planner_runtime, float* x0_out, float* xs, float* us,
                              float* Ps, float* alphas, double* t0_out) {
  auto problem = MakeProblem(which);
  if (!problem) return -1;
  const auto& dyn = *problem->Dynamics();
  const int n = dyn.XDim(), M = dyn.TotalUDim(), N = dyn.NumPlayers();
  VectorXf x(n), xm(n);
  for (int d = 0; d < n; d++) { x(d) = x0[d]; xm(d) = x_meas[d]; }
  problem->ResetInitialState(x);
  ResetMultipliers(*problem);
  const SolverParams params = ToParams(*q);
  ILQSolver s(problem, params);
  bool ok = false;
  auto log = s.Solve(&ok, std::numeric_limits<Time>::infinity());
  problem->OverwriteSolution(log->FinalOperatingPoint(), log->FinalStrategies());
  problem->SetUpNextRecedingHorizon(xm, t, planner_runtime);
  const OperatingPoint& op = problem->CurrentOperatingPoint();
  const size_t T = time::kNumTimeSteps;
  for (int d = 0; d < n; d++) x0_out[d] = problem->InitialState()(d);
  for (size_t k = 0; k < T; k++) {
    for (int d = 0; d < n; d++) xs[k * n + d] = op.xs[k](d);
    int off = 0;
    for (int i = 0; i < N; i++) {
      for (int d = 0; d < dyn.UDim(i); d++) us[k * M + off + d] = op.us[k][i](d);
      off += dyn.UDim(i);
    }
  }
  CopyStrategies(problem->CurrentStrategies(), dyn, Ps, alphas);
  *t0_out = op.t0;
  RelativeTimeTracker::ResetInitialTime(0.0);
  return ok ? 0 : 1;
}

namespace {
// A plan of `steps` time steps (possibly longer than the horizon, as a SolutionSplicer holds) from
// flat arrays: xs [steps][n], us [steps][M], Ps [steps][M][n], alphas [steps][M].
void LoadPlan(const MultiPlayerIntegrableSystem& dyn, int steps, const float* xs, const float* us, const float* Ps,
              const float* alphas, double plan_t0, OperatingPoint* op, std::vector<Strategy>* strategies) {
  const int n = dyn.XDim(), M = dyn.TotalUDim(), N = dyn.NumPlayers();
  *op = OperatingPoint((size_t)steps, (PlayerIndex)N, plan_t0);
  strategies->clear();
  for (int i = 0; i < N; i++) strategies->emplace_back((size_t)steps, dyn.XDim(), dyn.UDim(i));
  for (int k = 0; k < steps; k++) {
    op->xs[k] = VectorXf(n);
    for (int d = 0; d < n; d++) op->xs[k](d) = xs[(size_t)k * n + d];
    int off = 0;
    for (int i = 0; i < N; i++) {
      const int m = dyn.UDim(i);
      op->us[k][i] = VectorXf(m);
      for (int r = 0; r < m; r++) {
        op->us[k][i](r) = us[(size_t)k * M + off + r];
        (*strategies)[i].alphas[k](r) = alphas[(size_t)k * M + off + r];
        for (int d = 0; d < n; d++) (*strategies)[i].Ps[k](r, d) = Ps[((size_t)k * M + off + r) * n + d];
      }
      off += m;
    }
  }
}
}  // namespace

// MultiPlayerIntegrableSystem::Integrate(t0, t, x0, operating_point, strategies)
// (src/multi_player_integrable_system.cpp:54-83) of problem `which`'s dynamics under the given plan.
int ilqg_ref_integrate(int which, int steps, const float* xs, const float* us, const float* Ps,
                       const float* alphas, double plan_t0, const float* x0, double t0, double t, float* x_out) {
  auto problem = MakeProblem(which);
  if (!problem) return -1;
  const auto& dyn = *problem->Dynamics();
  OperatingPoint op(1, 1, 0.0);
  std::vector<Strategy> strategies;
  LoadPlan(dyn, steps, xs, us, Ps, alphas, plan_t0, &op, &strategies);
  VectorXf x(dyn.XDim());
  for (int d = 0; d < dyn.XDim(); d++) x(d) = x0[d];
  const VectorXf out = dyn.Integrate(t0, t, x, op, strategies);
  for (int d = 0; d < dyn.XDim(); d++) x_out[d] = out(d);
  return 0;
}

// Problem::OverwriteSolution with a plan of `steps` >= kNumTimeSteps time steps (a spliced plan),
// then Problem::SetUpNextRecedingHorizon(x_meas, t, planner_runtime): the sequence of
// src/receding_horizon_simulator.cpp:105-109.  Outputs as ilqg_ref_receding_horizon.
int ilqg_ref_receding_from_plan(int which, int steps, const float* plan_xs, const float* plan_us,
                                const float* plan_Ps, const float* plan_alphas, double plan_t0,
                                const float* x_meas, double t, double planner_runtime, float* x0_out,
                                float* xs, float* us, float* Ps, float* alphas, double* t0_out) {
  auto problem = MakeProblem(which);
  if (!problem) return -1;
  const auto& dyn = *problem->Dynamics();
  const int n = dyn.XDim(), M = dyn.TotalUDim(), N = dyn.NumPlayers();
  OperatingPoint plan(1, 1, 0.0);
  std::vector<Strategy> strategies;
  LoadPlan(dyn, steps, plan_xs, plan_us, plan_Ps, plan_alphas, plan_t0, &plan, &strategies);
  problem->OverwriteSolution(plan, strategies);
  VectorXf xm(n);
  for (int d = 0; d < n; d++) xm(d) = x_meas[d];
  problem->SetUpNextRecedingHorizon(xm, t, planner_runtime);
  const OperatingPoint& op = problem->CurrentOperatingPoint();
  const size_t T = time::kNumTimeSteps;
  if (op.xs.size() != T) return 2;
  for (int d = 0; d < n; d++) x0_out[d] = problem->InitialState()(d);
  for (size_t k = 0; k < T; k++) {
    for (int d = 0; d < n; d++) xs[k * n + d] = op.xs[k](d);
    int off = 0;
    for (int i = 0; i < N; i++) {
      for (int d = 0; d < dyn.UDim(i); d++) us[k * M + off + d] = op.us[k][i](d);
      off += dyn.UDim(i);
    }
  }
  CopyStrategies(problem->CurrentStrategies(), dyn, Ps, alphas);
  *t0_out = op.t0;
  RelativeTimeTracker::ResetInitialTime(0.0);
  return 0;
}

// SolutionSplicer (src/solution_splicer.cpp:57-131) on synthetic logs: a stored plan of T steps
// starting at t0 = 0 whose entries encode (tag 1, time step, dimension), spliced with a new
// horizon starting at new_t0 (tag 2).  Two-player system, n = 3, m = (1, 2).  Outputs the spliced
// plan: xs [steps][3], us [steps][3], alphas [steps][3], P(0, 0) per player [steps][2], t0.
namespace {
struct SpliceDynamics : public MultiPlayerDynamicalSystem {
  SpliceDynamics() : MultiPlayerDynamicalSystem(3) {}
  Dimension UDim(PlayerIndex i) const { return i == 0 ? 1 : 2; }
  PlayerIndex NumPlayers() const { return 2; }
  VectorXf Evaluate(Time, const VectorXf& x, const std::vector<VectorXf>&) const { return x; }
  LinearDynamicsApproximation Linearize(Time, const VectorXf&, const std::vector<VectorXf>&) const {
    return LinearDynamicsApproximation(*this);
  }
  float DistanceBetween(const VectorXf& a, const VectorXf& b) const { return (a - b).squaredNorm(); }
  std::vector<Dimension> PositionDimensions() const { return {0, 1}; }
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
}  // namespace

int ilqg_ref_splice(double new_t0, int max_steps, float* xs, float* us, float* alphas, float* P00, double* t0_out) {
  const std::shared_ptr<const MultiPlayerIntegrableSystem> dyn(new SpliceDynamics);
  SolverLog stored, fresh;
  FillLog(&stored, dyn, 0.0, 1.f);
  FillLog(&fresh, dyn, new_t0, 2.f);
  SolutionSplicer splicer(stored);
  splicer.Splice(fresh);
  const OperatingPoint& op = splicer.CurrentOperatingPoint();
  const int steps = (int)op.xs.size();
  for (int k = 0; k < steps && k < max_steps; k++) {
    for (int d = 0; d < 3; d++) xs[k * 3 + d] = op.xs[k](d);
    int c = 0;
    for (PlayerIndex i = 0; i < 2; i++) {
      for (int d = 0; d < dyn->UDim(i); d++, c++) {
        us[k * 3 + c] = op.us[k][i](d);
        alphas[k * 3 + c] = splicer.CurrentStrategies()[i].alphas[k](d);
      }
      P00[k * 2 + i] = splicer.CurrentStrategies()[i].Ps[k](0, 0);
    }
  }
  *t0_out = op.t0;
  return steps;
}

// ILQSolver with max_solver_iters = 1 from x0: returns the linearization and quadraticization
// members after the solve -- i.e. A, B of the INITIAL rollout and Q, l, R_ii, r_ii of the
// accepted operating point (src/ilq_solver.cpp:437-490).  A [T][n][n], B [T][n][M],
// Q [T][N][n][n], l [T][N][n], R [T][sum m_i^2] (row-major blocks, player order), r [T][M].
int ilqg_ref_lin_quad(int which, const float* x0, const ilqg_ref_params* q, float mu0, float* A,
                      float* B, float* Q, float* l, float* R, float* r) {
  auto problem = MakeProblem(which);
  if (!problem) return -1;
  const auto& dyn = *problem->Dynamics();
  const int n = dyn.XDim(), M = dyn.TotalUDim(), N = dyn.NumPlayers();
  VectorXf x(n);
  for (int d = 0; d < n; d++) x(d) = x0[d];
  problem->ResetInitialState(x);
  ResetMultipliers(*problem);
  Constraint::GlobalMu() = mu0;
  SolverParams params = ToParams(*q);
  params.max_solver_iters = 1;
  ILQSolver s(problem, params);
  bool ok = false;
  s.Solve(&ok, std::numeric_limits<Time>::infinity());
  const auto& lin = *s.Linearization();
  const auto& quad = *s.Quadraticization();
  const size_t T = time::kNumTimeSteps;
  int sumsq = 0;
  for (int i = 0; i < N; i++) sumsq += dyn.UDim(i) * dyn.UDim(i);
  for (size_t k = 0; k < T; k++) {
    for (int a = 0; a < n; a++)
      for (int b = 0; b < n; b++) A[(k * n + a) * n + b] = lin[k].A(a, b);
    int off = 0, roff = 0;
    for (int i = 0; i < N; i++) {
      const int m = dyn.UDim(i);
      for (int a = 0; a < n; a++)
        for (int d = 0; d < m; d++) B[(k * n + a) * M + off + d] = lin[k].Bs[i](a, d);
      for (int a = 0; a < n; a++) {
        l[(k * N + i) * n + a] = quad[k][i].state.grad(a);
        for (int b = 0; b < n; b++) Q[((k * N + i) * n + a) * n + b] = quad[k][i].state.hess(a, b);
      }
      const auto& cq = quad[k][i].control.at(i);
      for (int a = 0; a < m; a++) {
        r[k * M + off + a] = cq.grad(a);
        for (int b = 0; b < m; b++) R[k * sumsq + roff + a * m + b] = cq.hess(a, b);
      }
      off += m;
      roff += m * m;
    }
  }
  ResetMultipliers(*problem);
  return ok ? 0 : 1;
}

// The reference's own Polyline2SignedDistanceConstraint (src/polyline2_signed_distance_constraint.cpp) at
// `count` points: no example adds one (the intersection example constructs six and leaves them commented
// out), so the class itself is the only thing there is to pin the record kind against.
// `lambda_step` != 0: IncrementLambda(0, lambda_step) first (lambda = max(0, mu lambda_step), constraint.h:96-100).
// out [count][7] = g, lambda, d/dx, d/dy, d2/dx2, d2/dxdy, d2/dy2 of the augmented Lagrangian terms.
int ilqg_ref_polyline_constraint(const float* pts, int npts, float threshold, int keep_left, float mu,
                                 float lambda_step, const float* xy, int count, float* out) {
  PointList2 points;
  for (int k = 0; k < npts; k++) points.push_back(Point2(pts[2 * k], pts[2 * k + 1]));
  const Polyline2 polyline(points);
  Polyline2SignedDistanceConstraint constraint(polyline, {0, 1}, threshold, keep_left != 0);
  const float saved_mu = Constraint::GlobalMu();
  Constraint::GlobalMu() = mu;
  if (lambda_step != 0.0f) constraint.IncrementLambda(0.0, lambda_step);
  for (int k = 0; k < count; k++) {
    VectorXf input(2);
    input(0) = xy[2 * k];
    input(1) = xy[2 * k + 1];
    MatrixXf hess = MatrixXf::Zero(2, 2);
    VectorXf grad = VectorXf::Zero(2);
    constraint.Quadraticize(0.0, input, &hess, &grad);
    float* o = out + 7 * k;
    o[0] = constraint.Evaluate(input);
    o[1] = constraint.Lambda(0.0);
    o[2] = grad(0);
    o[3] = grad(1);
    o[4] = hess(0, 0);
    o[5] = hess(0, 1);
    o[6] = hess(1, 1);
  }
  Constraint::GlobalMu() = saved_mu;
  return 0;
}

}  // extern "C"
