// solvers.h -- Problem, ILQSolver, LQFeedbackSolver and AugmentedLagrangianSolver of the
// ilqgames API, re-authored as thin hosts over the C ABI (include/ilqg.h).  Class names,
// constructor arguments, Solve() signatures, ownership and error behaviour follow the reference
// headers cited at each class; the numerical bodies are the CUDA kernels behind ilqg_*.
// Link against ilqgames_b200/lib/libilqg_b200.so (or, in tests, any other implementation of
// ilqg.h).
#ifndef ILQGAMES_B200_SOLVERS_H
#define ILQGAMES_B200_SOLVERS_H

#include <ilqg.h>
#include <ilqgames/b200/core.h>
#include <ilqgames/b200/costs.h>
#include <ilqgames/b200/dynamics.h>

#include <cstring>

namespace ilqgames {

// ---- include/ilqgames/solver/problem.h:61-176 --------------------------------------------------
class Problem {
 public:
  virtual ~Problem() {}

  virtual void Initialize() {
    ConstructDynamics();
    ConstructPlayerCosts();
    ConstructInitialState();
    ConstructInitialOperatingPoint();
    ConstructInitialStrategies();
    initialized_ = true;
  }

  void ResetInitialTime(Time t0) { CHECK(initialized_); operating_point_->t0 = t0; }
  void ResetInitialState(const VectorXf& x0) { CHECK(initialized_); x0_ = x0; }

  // src/problem.cpp:188-194
  virtual void OverwriteSolution(const OperatingPoint& operating_point, const std::vector<Strategy>& strategies) {
    CHECK(initialized_);
    *operating_point_ = operating_point;
    *strategies_ = strategies;
  }

  // src/problem.cpp:64-186 (SyncToExistingProblem + SetUpNextRecedingHorizon): integrate the measured
  // state x0 (taken at t0) along the current plan for ~planner_runtime, start the new problem at
  // the nearest plan state, shift operating point and strategies, extend them with zero controls.
  // Runs on the device (ilqg_setup_next_receding_horizon); defined below b200::Handle.
  virtual void SetUpNextRecedingHorizon(const VectorXf& x0, Time t0, Time planner_runtime = 0.1);

  bool IsConstrained() const {
    for (const auto& pc : player_costs_)
      if (pc.IsConstrained()) return true;
    return false;
  }

  virtual Time InitialTime() const { return operating_point_->t0; }
  const VectorXf& InitialState() const { return x0_; }
  std::vector<PlayerCost>& PlayerCosts() { return player_costs_; }
  const std::vector<PlayerCost>& PlayerCosts() const { return player_costs_; }
  const std::shared_ptr<const MultiPlayerIntegrableSystem>& Dynamics() const { return dynamics_; }
  virtual const OperatingPoint& CurrentOperatingPoint() const { return *operating_point_; }
  virtual const std::vector<Strategy>& CurrentStrategies() const { return *strategies_; }

 protected:
  Problem() : initialized_(false) {}

  virtual void ConstructDynamics() = 0;
  virtual void ConstructPlayerCosts() = 0;
  virtual void ConstructInitialState() = 0;
  virtual void ConstructInitialOperatingPoint() {
    operating_point_.reset(new OperatingPoint(time::kNumTimeSteps, 0.0, dynamics_));
  }
  virtual void ConstructInitialStrategies() {
    strategies_.reset(new std::vector<Strategy>());
    for (PlayerIndex ii = 0; ii < dynamics_->NumPlayers(); ii++)
      strategies_->emplace_back(time::kNumTimeSteps, dynamics_->XDim(), dynamics_->UDim(ii));
  }

  std::shared_ptr<const MultiPlayerIntegrableSystem> dynamics_;
  std::vector<PlayerCost> player_costs_;
  VectorXf x0_;
  std::unique_ptr<OperatingPoint> operating_point_;
  std::unique_ptr<std::vector<Strategy>> strategies_;
  bool initialized_;
};

// include/ilqgames/solver/top_down_renderable_problem.h:52-64
class TopDownRenderableProblem : public Problem {
 public:
  virtual ~TopDownRenderableProblem() {}
  virtual std::vector<float> Xs(const VectorXf& x) const = 0;
  virtual std::vector<float> Ys(const VectorXf& x) const = 0;
  virtual std::vector<float> Thetas(const VectorXf& x) const = 0;

 protected:
  TopDownRenderableProblem() : Problem() {}
};

// ---------------------------------------------------------------------------------------------
// glue between the class API and the POD ABI
// ---------------------------------------------------------------------------------------------
namespace b200 {

// Problem -> ilqg_problem_desc.  Records per player in the reference's accumulation order
// (src/player_cost.cpp:194-215): state costs, control costs, state constraints, control
// constraints.  Returns false if some object has no device implementation.
inline bool DescribeProblem(const Problem& problem, ilqg_problem_desc* desc) {
  std::memset(desc, 0, sizeof(*desc));
  const auto& dyn = problem.Dynamics();
  const std::vector<PlayerCost>& pcs = problem.PlayerCosts();
  if (!dyn || dyn->NumPlayers() > ILQG_MAX_PLAYERS || pcs.size() != dyn->NumPlayers()) return false;
  desc->num_time_steps = (int32_t)time::kNumTimeSteps;
  desc->time_step = time::kTimeStep;
  desc->initial_time = problem.InitialTime();
  desc->num_players = dyn->NumPlayers();
  desc->xdim = dyn->XDim();
  for (PlayerIndex ii = 0; ii < dyn->NumPlayers(); ii++) desc->udim[ii] = dyn->UDim(ii);
  if (!dyn->Describe(desc)) return false;
  DescribeContext ctx{desc};
  auto push = [&](int player, int arg, bool is_equality, auto&& describe) {
    if (desc->num_costs >= ILQG_MAX_COSTS) return false;
    ilqg_cost_desc& rec = desc->costs[desc->num_costs];
    std::memset(&rec, 0, sizeof(rec));
    rec.polyline = -1;
    if (!describe(&rec)) return false;
    rec.player = player;
    rec.arg = arg;
    rec.is_equality = is_equality;
    desc->num_costs++;
    return true;
  };
  // a Cost may contribute several records (ExtremeValueCost: one per member)
  auto push_cost = [&](int player, int arg, const Cost& cost) {
    std::vector<ilqg_cost_desc> records;
    if (!cost.DescribeAll(&records, &ctx)) return false;
    for (const ilqg_cost_desc& described : records)
      if (!push(player, arg, false, [&](ilqg_cost_desc* r) { *r = described; return true; })) return false;
    return true;
  };
  for (size_t ii = 0; ii < pcs.size(); ii++) {
    const PlayerCost& pc = pcs[ii];
    desc->state_regularization[ii] = pc.StateRegularization();
    desc->control_regularization[ii] = pc.ControlRegularization();
    desc->cost_structure[ii] = pc.CostStructure();
    for (const auto& c : pc.StateCosts())
      if (!push_cost((int)ii, -1, *c)) return false;
    for (const auto& pr : pc.ControlCosts())
      if (!push_cost((int)ii, pr.first, *pr.second)) return false;
    for (const auto& c : pc.StateConstraints())
      if (!push((int)ii, -1, c->IsEquality(), [&](ilqg_cost_desc* r) { return c->Describe(r, &ctx); })) return false;
    for (const auto& pr : pc.ControlConstraints())
      if (!push((int)ii, pr.first, pr.second->IsEquality(), [&](ilqg_cost_desc* r) { return pr.second->Describe(r, &ctx); }))
        return false;
  }
  return true;
}

inline ilqg_solver_params ToAbi(const SolverParams& p) {
  ilqg_solver_params out;
  std::memset(&out, 0, sizeof(out));
  out.convergence_tolerance = p.convergence_tolerance;
  out.max_solver_iters = (int32_t)p.max_solver_iters;
  out.linesearch = p.linesearch;
  out.initial_alpha_scaling = p.initial_alpha_scaling;
  out.geometric_alpha_scaling = p.geometric_alpha_scaling;
  out.max_backtracking_steps = (int32_t)p.max_backtracking_steps;
  out.expected_decrease_fraction = p.expected_decrease_fraction;
  out.open_loop = p.open_loop;
  out.unconstrained_solver_max_iters = (int32_t)p.unconstrained_solver_max_iters;
  out.geometric_mu_scaling = p.geometric_mu_scaling;
  out.geometric_mu_downscaling = p.geometric_mu_downscaling;
  out.geometric_lambda_downscaling = p.geometric_lambda_downscaling;
  out.constraint_error_tolerance = p.constraint_error_tolerance;
  out.adaptive_regularization = 1;
  out.disable_convergence_exit = 0;
  return out;
}

#define ILQG_CALL(expr)                                                   \
  do {                                                                    \
    const int ilqg_rc_ = (expr);                                          \
    CHECK(ilqg_rc_ == ILQG_OK) << #expr << ": " << ilqg_strerror(ilqg_rc_); \
  } while (0)

// One batch of games on one device.
class Handle {
 public:
  Handle(const ilqg_problem_desc& desc, const ilqg_solver_params& params, int batch, int device = 0) : h_(nullptr) {
    ILQG_CALL(ilqg_create(&desc, &params, batch, device, &h_));
    ILQG_CALL(ilqg_get_layout(h_, &lo_));
  }
  // the batch sharded over several GPUs of the node (ilqg_create_multi)
  Handle(const ilqg_problem_desc& desc, const ilqg_solver_params& params, int batch, const std::vector<int>& devices)
      : h_(nullptr) {
    ILQG_CALL(ilqg_create_multi(&desc, &params, batch, devices.data(), (int)devices.size(), &h_));
    ILQG_CALL(ilqg_get_layout(h_, &lo_));
  }
  ~Handle() { if (h_) ilqg_destroy(h_); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  ilqg_handle get() const { return h_; }
  const ilqg_layout& layout() const { return lo_; }
  int B() const { return lo_.batch; }
  int T() const { return lo_.num_time_steps; }
  int n() const { return lo_.xdim; }
  int M() const { return lo_.total_udim; }
  int N() const { return lo_.num_players; }

  template <typename T>
  std::vector<T> Download(int what, size_t per_instance) const {
    std::vector<T> out((size_t)B() * per_instance);
    ILQG_CALL(ilqg_download(h_, what, out.data(), out.size() * sizeof(T)));
    return out;
  }

  // game b's current iterate as the reference's carriers
  OperatingPoint OperatingPointOf(int b, const std::vector<float>& xs, const std::vector<float>& us, Time t0) const {
    OperatingPoint op((size_t)T(), (PlayerIndex)N(), t0);
    for (int k = 0; k < T(); k++) {
      op.xs[k] = VectorXf(n());
      for (int a = 0; a < n(); a++) op.xs[k](a) = xs[((size_t)b * T() + k) * n() + a];
      for (int i = 0; i < N(); i++) {
        op.us[k][i] = VectorXf(lo_.udim[i]);
        for (int a = 0; a < lo_.udim[i]; a++) op.us[k][i](a) = us[((size_t)b * T() + k) * M() + lo_.u_offset[i] + a];
      }
    }
    return op;
  }
  std::vector<Strategy> StrategiesOf(int b, const std::vector<float>& Ps, const std::vector<float>& alphas) const {
    std::vector<Strategy> out;
    for (int i = 0; i < N(); i++) {
      out.emplace_back((size_t)T(), n(), lo_.udim[i]);
      for (int k = 0; k < T(); k++)
        for (int r = 0; r < lo_.udim[i]; r++) {
          const size_t row = ((size_t)b * T() + k) * M() + lo_.u_offset[i] + r;
          out[i].alphas[k](r) = alphas[row];
          for (int a = 0; a < n(); a++) out[i].Ps[k](r, a) = Ps[row * n() + a];
        }
    }
    return out;
  }
  // the reverse packing, game 0 .. B-1 all from the same carriers (Problem's warm start)
  void UploadWarmStart(const OperatingPoint& op, const std::vector<Strategy>& strategies) {
    std::vector<float> xs((size_t)B() * T() * n()), us((size_t)B() * T() * M()), Ps((size_t)B() * T() * M() * n()),
        al((size_t)B() * T() * M());
    for (int b = 0; b < B(); b++)
      for (int k = 0; k < T(); k++) {
        for (int a = 0; a < n(); a++) xs[((size_t)b * T() + k) * n() + a] = op.xs[k](a);
        for (int i = 0; i < N(); i++)
          for (int r = 0; r < lo_.udim[i]; r++) {
            const size_t row = ((size_t)b * T() + k) * M() + lo_.u_offset[i] + r;
            us[row] = op.us[k][i](r);
            al[row] = strategies[i].alphas[k](r);
            for (int a = 0; a < n(); a++) Ps[row * n() + a] = strategies[i].Ps[k](r, a);
          }
      }
    ILQG_CALL(ilqg_upload_warmstart(h_, xs.data(), us.data(), Ps.data(), al.data()));
  }

 private:
  ilqg_handle h_;
  ilqg_layout lo_;
};

}  // namespace b200

inline void Problem::SetUpNextRecedingHorizon(const VectorXf& x0, Time t0, Time planner_runtime) {
  CHECK(initialized_);
  // the stored plan may be longer than the horizon: a receding-horizon caller overwrites it with a
  // SolutionSplicer's plan first (src/receding_horizon_simulator.cpp:105-109, src/problem.cpp:160-169)
  const size_t horizon = time::kNumTimeSteps, stored = operating_point_->xs.size();
  CHECK_GE(stored, horizon);
  CHECK_LE(planner_runtime + t0, operating_point_->t0 + time::kTimeHorizon);  // :70
  ilqg_problem_desc desc;
  CHECK(b200::DescribeProblem(*this, &desc)) << "a cost, constraint or dynamics class has no device record";
  desc.num_time_steps = (int32_t)stored;
  b200::Handle h(desc, b200::ToAbi(SolverParams()), 1);
  h.UploadWarmStart(*operating_point_, *strategies_);
  std::vector<float> x((size_t)x0.size());
  for (long a = 0; a < x0.size(); a++) x[(size_t)a] = x0(a);
  double new_t0 = 0.0;
  ILQG_CALL(ilqg_setup_next_receding_horizon(h.get(), x.data(), t0, planner_runtime, &new_t0));
  const size_t T = h.T(), n = h.n(), M = h.M();
  *operating_point_ = h.OperatingPointOf(0, h.Download<float>(ILQG_WARM_XS, T * n),
                                         h.Download<float>(ILQG_WARM_US, T * M), new_t0);
  *strategies_ = h.StrategiesOf(0, h.Download<float>(ILQG_WARM_PS, T * M * n),
                                h.Download<float>(ILQG_WARM_ALPHAS, T * M));
  // the new problem is the first `horizon` steps of the re-based plan (:160-169)
  operating_point_->xs.resize(horizon);
  operating_point_->us.resize(horizon);
  for (Strategy& strategy : *strategies_) {
    strategy.Ps.resize(horizon);
    strategy.alphas.resize(horizon);
  }
  const std::vector<float> nx = h.Download<float>(ILQG_X0, n);
  for (size_t a = 0; a < n; a++) x0_((long)a) = nx[a];
}

inline VectorXf MultiPlayerIntegrableSystem::Integrate(Time t0, Time t, const VectorXf& x0,
                                                       const OperatingPoint& operating_point,
                                                       const std::vector<Strategy>& strategies) const {
  CHECK_GE(t, t0);
  CHECK_GE(t0, operating_point.t0);
  CHECK_EQ(strategies.size(), NumPlayers());
  CHECK_EQ((long)x0.size(), (long)XDim());
  // a dynamics-only descriptor: the plan's own length and start time, no cost records
  ilqg_problem_desc desc;
  std::memset(&desc, 0, sizeof(desc));
  desc.num_time_steps = (int32_t)operating_point.xs.size();
  desc.time_step = time::kTimeStep;
  desc.initial_time = operating_point.t0;
  desc.num_players = NumPlayers();
  desc.xdim = XDim();
  for (PlayerIndex ii = 0; ii < NumPlayers(); ii++) desc.udim[ii] = UDim(ii);
  CHECK(Describe(&desc)) << "this dynamics class has no device record";
  b200::Handle h(desc, b200::ToAbi(SolverParams()), 1);
  h.UploadWarmStart(operating_point, strategies);
  std::vector<float> x((size_t)x0.size()), out((size_t)x0.size());
  for (long a = 0; a < x0.size(); a++) x[(size_t)a] = x0(a);
  ILQG_CALL(ilqg_integrate_plan(h.get(), x.data(), t0, t, out.data()));
  VectorXf result(x0.size());
  for (long a = 0; a < x0.size(); a++) result(a) = out[(size_t)a];
  return result;
}


// ---- include/ilqgames/solver/game_solver.h:58-95 -----------------------------------------------
class GameSolver {
 public:
  virtual ~GameSolver() {}
  virtual std::shared_ptr<SolverLog> Solve(bool* success = nullptr, Time max_runtime = constants::kInfinity) = 0;
  Problem& GetProblem() { return *problem_; }

 protected:
  GameSolver(const std::shared_ptr<Problem>& problem, const SolverParams& params) : problem_(problem), params_(params) {
    CHECK_NOTNULL(problem_.get());
    CHECK_NOTNULL(problem_->Dynamics().get());
  }
  virtual std::shared_ptr<SolverLog> CreateNewLog() const { return std::make_shared<SolverLog>(); }
  const std::shared_ptr<Problem> problem_;
  const SolverParams params_;
};

// One game's final iterate of a batched solve (SolveBatch).
struct BatchSolution {
  OperatingPoint operating_point;
  std::vector<Strategy> strategies;
  std::vector<float> total_costs;
  int status;       // ILQG_STATUS_*
  int iterations;   // completed iterations of the while loop
  bool success;     // what Solve() would have stored in *success
};

// ---- include/ilqgames/solver/ilq_solver.h:66-196 -----------------------------------------------
class ILQSolver : public GameSolver {
 public:
  ~ILQSolver() {}
  ILQSolver(const std::shared_ptr<Problem>& problem, const SolverParams& params = SolverParams())
      : GameSolver(problem, params) {
    // the problem is described once, like the reference sizes its scratch here (:76-94)
    CHECK(b200::DescribeProblem(*problem_, &desc_)) << "a cost, constraint or dynamics class has no device record";
    abi_params_ = b200::ToAbi(params_);
    single_.reset(new b200::Handle(desc_, abi_params_, 1));
  }

  // src/ilq_solver.cpp:76-172.  The device keeps last_merit_function_value_ between calls on
  // the same solver, as the member does in the reference.
  std::shared_ptr<SolverLog> Solve(bool* success = nullptr, Time max_runtime = constants::kInfinity) override {
    const auto start = Clock::now();
    auto elapsed = [&]() { return std::chrono::duration<Time>(Clock::now() - start).count(); };
    std::shared_ptr<SolverLog> log = CreateNewLog();
    b200::Handle& h = *single_;
    const VectorXf& x0 = problem_->InitialState();
    CHECK_EQ(x0.size(), (long)h.n());
    ILQG_CALL(ilqg_upload_x0(h.get(), x0.data(), sizeof(float) * h.n()));
    h.UploadWarmStart(problem_->CurrentOperatingPoint(), problem_->CurrentStrategies());
    ILQG_CALL(ilqg_solve_begin(h.get()));                       // :86-107
    AppendIterate(h, log.get(), elapsed(), false);              // :111
    int running = 0;
    ILQG_CALL(ilqg_count_running(h.get(), &running));
    while (running > 0 && elapsed() < max_runtime) {            // :123-124
      ILQG_CALL(ilqg_iterate(h.get(), 1, nullptr));             // :127-158
      const int status = h.Download<int32_t>(ILQG_STATUS, 1)[0];
      if (status == ILQG_STATUS_LINESEARCH_FAILED) {            // :146-153
        if (success) *success = false;
        return log;
      }
      AppendIterate(h, log.get(), elapsed(), status == ILQG_STATUS_CONVERGED);  // :164
      running = status == ILQG_STATUS_RUNNING;
    }
    if (success) *success = true;
    return log;
  }

  // Additive: the same Solve() for many initial states at once, one game per x0, all starting
  // from the problem's current operating point and strategies.  The problem is not modified.
  std::vector<BatchSolution> SolveBatch(const std::vector<VectorXf>& x0s, int device = 0) {
    return SolveBatch(x0s, std::vector<int>{device});
  }

  // ... and the batch sharded over several GPUs of the node: devices[k] solves a contiguous slice of
  // the games (no exchange between devices: the games are independent), results come back in order.
  std::vector<BatchSolution> SolveBatch(const std::vector<VectorXf>& x0s, const std::vector<int>& devices) {
    const int B = (int)x0s.size();
    CHECK_GT(B, 0);
    CHECK(!devices.empty());
    if (!batch_ || batch_->B() != B || batch_devices_ != devices) {
      if (devices.size() == 1) batch_.reset(new b200::Handle(desc_, abi_params_, B, devices[0]));
      else batch_.reset(new b200::Handle(desc_, abi_params_, B, devices));
      batch_devices_ = devices;
    }
    b200::Handle& h = *batch_;
    std::vector<float> packed((size_t)B * h.n());
    for (int b = 0; b < B; b++) {
      CHECK_EQ(x0s[b].size(), (long)h.n());
      for (int a = 0; a < h.n(); a++) packed[(size_t)b * h.n() + a] = x0s[b](a);
    }
    ILQG_CALL(ilqg_upload_x0(h.get(), packed.data(), packed.size() * sizeof(float)));
    h.UploadWarmStart(problem_->CurrentOperatingPoint(), problem_->CurrentStrategies());
    ILQG_CALL(ilqg_reset(h.get(), ILQG_RESET_SOLVER));  // every game gets a fresh ILQSolver
    ILQG_CALL(ilqg_solve_begin(h.get()));
    int running = B;
    while (running > 0) {
      ILQG_CALL(ilqg_iterate(h.get(), 4, nullptr));
      ILQG_CALL(ilqg_count_running(h.get(), &running));
    }
    return Collect(h);
  }

 protected:
  static std::vector<BatchSolution> Collect(const b200::Handle& h) {
    const size_t T = h.T(), n = h.n(), M = h.M(), N = h.N();
    const auto xs = h.Download<float>(ILQG_XS, T * n), us = h.Download<float>(ILQG_US, T * M);
    const auto Ps = h.Download<float>(ILQG_PS, T * M * n), al = h.Download<float>(ILQG_ALPHAS, T * M);
    const auto costs = h.Download<float>(ILQG_TOTAL_COSTS, N);
    const auto status = h.Download<int32_t>(ILQG_STATUS, 1), iters = h.Download<int32_t>(ILQG_ITERS, 1);
    std::vector<BatchSolution> out;
    for (int b = 0; b < h.B(); b++)
      out.push_back(BatchSolution{h.OperatingPointOf(b, xs, us, 0.0), h.StrategiesOf(b, Ps, al),
                                  std::vector<float>(costs.begin() + b * N, costs.begin() + (b + 1) * N), status[b],
                                  iters[b], status[b] != ILQG_STATUS_LINESEARCH_FAILED});
    return out;
  }

  void AppendIterate(const b200::Handle& h, SolverLog* log, Time elapsed, bool converged) const {
    const size_t T = h.T(), n = h.n(), M = h.M(), N = h.N();
    const auto xs = h.Download<float>(ILQG_XS, T * n), us = h.Download<float>(ILQG_US, T * M);
    const auto Ps = h.Download<float>(ILQG_PS, T * M * n), al = h.Download<float>(ILQG_ALPHAS, T * M);
    log->AddSolverIterate(h.OperatingPointOf(0, xs, us, problem_->InitialTime()), h.StrategiesOf(0, Ps, al),
                          h.Download<float>(ILQG_TOTAL_COSTS, N), elapsed, converged);
  }

  friend class AugmentedLagrangianSolver;
  ilqg_problem_desc desc_;
  ilqg_solver_params abi_params_;
  std::unique_ptr<b200::Handle> single_, batch_;
  std::vector<int> batch_devices_;
};

// ---- include/ilqgames/solver/lq_solver.h:57-81, lq_feedback_solver.h:70-124,
// ---- lq_open_loop_solver.h:72-135 --------------------------------------------------------------
class LQSolver {
 public:
  virtual ~LQSolver() {}
  virtual std::vector<Strategy> Solve(
      const std::vector<LinearDynamicsApproximation>& linearization,
      const std::vector<std::vector<QuadraticCostApproximation>>& quadraticization, const VectorXf& x0,
      std::vector<VectorXf>* delta_xs = nullptr, std::vector<std::vector<VectorXf>>* costates = nullptr) = 0;

 protected:
  LQSolver(const std::shared_ptr<const MultiPlayerIntegrableSystem>& dynamics, size_t num_time_steps)
      : dynamics_(dynamics), num_time_steps_(num_time_steps) { CHECK_NOTNULL(dynamics.get()); }

  // Both solvers: flatten lin / quad into LQ records, one ilqg_lq_backward on a batch-1 handle.
  // quadraticization[k][i].control holds R_ij, r_ij for the players j whose control player i
  // penalises.  costates are not produced (the reference computes them and never reads them,
  // SURVEY Q4); passing a non-null pointer is an error.
  std::vector<Strategy> SolveOnDevice(bool open_loop, bool adaptive_regularization,
                                      const std::vector<LinearDynamicsApproximation>& linearization,
                                      const std::vector<std::vector<QuadraticCostApproximation>>& quadraticization,
                                      const VectorXf& x0, std::vector<VectorXf>* delta_xs,
                                      std::vector<std::vector<VectorXf>>* costates) {
    CHECK(costates == nullptr) << "costates are not produced on this path";
    CHECK_EQ(linearization.size(), num_time_steps_);
    CHECK_EQ(quadraticization.size(), num_time_steps_);
    const int T = (int)num_time_steps_, N = dynamics_->NumPlayers(), n = dynamics_->XDim();
    if (!handle_) BuildHandle(quadraticization.front(), open_loop, adaptive_regularization);
    b200::Handle& h = *handle_;
    const ilqg_layout& lo = h.layout();
    const int M = lo.total_udim;
    std::vector<float> A((size_t)T * n * n), Bs((size_t)T * n * M), Q((size_t)T * N * n * n), l((size_t)T * N * n),
        R((size_t)T * lo.R_floats), r((size_t)T * lo.r_floats);
    for (int k = 0; k < T; k++) {
      const LinearDynamicsApproximation& lin = linearization[k];
      for (int a = 0; a < n; a++) {
        for (int c = 0; c < n; c++) A[((size_t)k * n + a) * n + c] = lin.A(a, c);
        for (int i = 0; i < N; i++)
          for (int c = 0; c < lo.udim[i]; c++) Bs[((size_t)k * n + a) * M + lo.u_offset[i] + c] = lin.Bs[i](a, c);
      }
      for (int i = 0; i < N; i++) {
        const QuadraticCostApproximation& q = quadraticization[k][i];
        for (int a = 0; a < n; a++) {
          l[((size_t)k * N + i) * n + a] = q.state.grad(a);
          for (int c = 0; c < n; c++) Q[(((size_t)k * N + i) * n + a) * n + c] = q.state.hess(a, c);
        }
      }
      for (int p = 0; p < lo.num_pairs; p++) {
        const int i = lo.pair_player[p], j = lo.pair_arg[p], mj = lo.udim[j];
        const auto it = quadraticization[k][i].control.find((PlayerIndex)j);
        if (it == quadraticization[k][i].control.end()) continue;  // block stays zero
        for (int a = 0; a < mj; a++) {
          r[(size_t)k * lo.r_floats + lo.pair_r_offset[p] + a] = it->second.grad(a);
          for (int c = 0; c < mj; c++) R[(size_t)k * lo.R_floats + lo.pair_R_offset[p] + a * mj + c] = it->second.hess(a, c);
        }
      }
    }
    ILQG_CALL(ilqg_upload_lq(h.get(), A.data(), Bs.data(), Q.data(), l.data(), R.data(), r.data()));
    std::vector<float> x0f((size_t)n);
    for (int a = 0; a < n; a++) x0f[(size_t)a] = x0(a);
    ILQG_CALL(ilqg_upload(h.get(), ILQG_LQ_X0, x0f.data(), sizeof(float) * x0f.size()));
    ILQG_CALL(ilqg_lq_backward(h.get()));
    const auto Ps = h.Download<float>(ILQG_LQ_PS, (size_t)T * M * n), al = h.Download<float>(ILQG_LQ_ALPHAS, (size_t)T * M);
    std::vector<Strategy> strategies = h.StrategiesOf(0, Ps, al);
    if (delta_xs) {
      // lq_feedback_solver.cpp:217-241 (x* <- A x* - sum_i B_i alpha_i) resp.
      // lq_open_loop_solver.cpp:158-186 (the optimal state trajectory), computed on the device
      const auto dx = h.Download<float>(ILQG_DELTA_XS, (size_t)T * n);
      delta_xs->assign(T, VectorXf::Zero(n));
      for (int k = 0; k < T; k++)
        for (int a = 0; a < n; a++) (*delta_xs)[k](a) = dx[(size_t)k * n + a];
    }
    return strategies;
  }

  const std::shared_ptr<const MultiPlayerIntegrableSystem> dynamics_;
  const size_t num_time_steps_;

 private:
  void BuildHandle(const std::vector<QuadraticCostApproximation>& quad0, bool open_loop, bool adaptive_regularization) {
    ilqg_problem_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    const int N = dynamics_->NumPlayers();
    CHECK_LE(N, ILQG_MAX_PLAYERS);
    desc.num_time_steps = (int32_t)num_time_steps_;
    desc.time_step = time::kTimeStep;
    desc.num_players = N;
    desc.xdim = dynamics_->XDim();
    for (int i = 0; i < N; i++) desc.udim[i] = dynamics_->UDim((PlayerIndex)i);
    // a zero-weight control record declares that the (i, j) block exists
    for (int i = 0; i < N; i++)
      for (const auto& entry : quad0[i].control) {
        if ((int)entry.first == i) continue;
        ilqg_cost_desc& rec = desc.costs[desc.num_costs++];
        std::memset(&rec, 0, sizeof(rec));
        rec.kind = ILQG_COST_QUADRATIC;
        rec.player = i;
        rec.arg = entry.first;
        rec.dim[0] = -1;
        rec.polyline = -1;
      }
    ilqg_solver_params p = b200::ToAbi(SolverParams());
    p.adaptive_regularization = adaptive_regularization;
    p.open_loop = open_loop;
    handle_.reset(new b200::Handle(desc, p, 1));
  }

  std::unique_ptr<b200::Handle> handle_;
};

class LQFeedbackSolver : public LQSolver {
 public:
  ~LQFeedbackSolver() {}
  LQFeedbackSolver(const std::shared_ptr<const MultiPlayerIntegrableSystem>& dynamics, size_t num_time_steps,
                   bool adaptive_regularization = true)
      : LQSolver(dynamics, num_time_steps), adaptive_regularization_(adaptive_regularization) {}

  // src/lq_feedback_solver.cpp:71-244
  std::vector<Strategy> Solve(const std::vector<LinearDynamicsApproximation>& linearization,
                              const std::vector<std::vector<QuadraticCostApproximation>>& quadraticization,
                              const VectorXf& x0, std::vector<VectorXf>* delta_xs = nullptr,
                              std::vector<std::vector<VectorXf>>* costates = nullptr) override {
    return SolveOnDevice(false, adaptive_regularization_, linearization, quadraticization, x0, delta_xs, costates);
  }

 private:
  const bool adaptive_regularization_;
};

class LQOpenLoopSolver : public LQSolver {
 public:
  ~LQOpenLoopSolver() {}
  LQOpenLoopSolver(const std::shared_ptr<const MultiPlayerIntegrableSystem>& dynamics, size_t num_time_steps)
      : LQSolver(dynamics, num_time_steps) {}

  // src/lq_open_loop_solver.cpp:73-195: P stays zero, alphas are the (sign-flipped) open-loop controls
  std::vector<Strategy> Solve(const std::vector<LinearDynamicsApproximation>& linearization,
                              const std::vector<std::vector<QuadraticCostApproximation>>& quadraticization,
                              const VectorXf& x0, std::vector<VectorXf>* delta_xs = nullptr,
                              std::vector<std::vector<VectorXf>>* costates = nullptr) override {
    return SolveOnDevice(true, false, linearization, quadraticization, x0, delta_xs, costates);
  }
};

// ---- include/ilqgames/solver/augmented_lagrangian_solver.h:60-110 ------------------------------
class AugmentedLagrangianSolver : public GameSolver {
 public:
  ~AugmentedLagrangianSolver() {}
  AugmentedLagrangianSolver(const std::shared_ptr<Problem>& problem, const SolverParams& params)
      : GameSolver(problem, params) {
    SolverParams unconstrained_solver_params(params);
    unconstrained_solver_params.max_solver_iters = params.unconstrained_solver_max_iters;  // :82-83
    unconstrained_solver_.reset(new ILQSolver(problem, unconstrained_solver_params));
  }

  // src/augmented_lagrangian_solver.cpp:72-210.  The per-game state machine runs on the device
  // (ilqg_al_begin / ilqg_al_advance); the log receives the final iterate of every inner solve.
  std::shared_ptr<SolverLog> Solve(bool* success = nullptr, Time max_runtime = constants::kInfinity) override {
    const auto start = Clock::now();
    auto elapsed = [&]() { return std::chrono::duration<Time>(Clock::now() - start).count(); };
    std::shared_ptr<SolverLog> log = CreateNewLog();
    ILQSolver& inner = *unconstrained_solver_;
    b200::Handle& h = *inner.single_;
    const VectorXf& x0 = problem_->InitialState();
    ILQG_CALL(ilqg_upload_x0(h.get(), x0.data(), sizeof(float) * h.n()));
    h.UploadWarmStart(problem_->CurrentOperatingPoint(), problem_->CurrentStrategies());
    ILQG_CALL(ilqg_al_begin(h.get(), (int)params_.max_solver_iters, params_.constraint_error_tolerance));
    int active = 1;
    bool first_solve = true;  // the reference always runs the first unconstrained solve, whatever max_runtime (:92)
    while (active > 0 && (first_solve || elapsed() < max_runtime)) {
      first_solve = false;
      ILQG_CALL(ilqg_solve_begin(h.get()));
      int running = 1;
      while (running > 0) {
        ILQG_CALL(ilqg_iterate(h.get(), 4, nullptr));
        ILQG_CALL(ilqg_count_running(h.get(), &running));
      }
      inner.AppendIterate(h, log.get(), elapsed(), false);
      ILQG_CALL(ilqg_al_advance(h.get(), &active));
    }
    if (success) *success = h.Download<int32_t>(ILQG_AL_SUCCESS, 1)[0] != 0 && active == 0;
    num_iterates_ = h.Download<int32_t>(ILQG_AL_ITERATES, 1)[0];
    // :192-207.  In the reference `initial_op` / `initial_strategies` are REFERENCES to the Problem's own
    // members (:77-78), so its reset_problem restore copies the members onto themselves: whatever the loop
    // last wrote with OverwriteSolution (:159-162) stays.  The device's warm start is exactly that state.
    if (h.layout().num_constraints > 0) {
      const size_t T = h.T(), n = h.n(), M = h.M();
      problem_->OverwriteSolution(
          h.OperatingPointOf(0, h.Download<float>(ILQG_WARM_XS, T * n), h.Download<float>(ILQG_WARM_US, T * M),
                             problem_->CurrentOperatingPoint().t0),
          h.StrategiesOf(0, h.Download<float>(ILQG_WARM_PS, T * M * n), h.Download<float>(ILQG_WARM_ALPHAS, T * M)));
    }
    const int mask = (params_.reset_lambdas ? ILQG_RESET_LAMBDAS : 0) | (params_.reset_mu ? ILQG_RESET_MU : 0);
    if (mask) ILQG_CALL(ilqg_reset(h.get(), mask));
    return log;
  }

  // log->NumIterates() of the reference's accumulated log for the last Solve()
  int NumIterates() const { return num_iterates_; }

 private:
  std::unique_ptr<ILQSolver> unconstrained_solver_;
  int num_iterates_ = 0;
};

}  // namespace ilqgames

#endif
