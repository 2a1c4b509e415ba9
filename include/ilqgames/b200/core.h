// core.h -- host-side value types of the ilqgames API, re-authored on top of the C ABI
// (include/ilqg.h).  Names, members and meaning follow the reference headers cited at each type
// so that code written against HJReachability/ilqgames compiles against these; the bodies are
// new: the numerical work happens behind ilqg_* on the GPU, these types only carry data.
#ifndef ILQGAMES_B200_CORE_H
#define ILQGAMES_B200_CORE_H

#include <ilqgames/b200/eigen_shim.h>
#include <ilqgames/b200/log_shim.h>

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <limits>
#include <memory>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

namespace ilqgames {

// ---- include/ilqgames/utils/types.h:56-143 --------------------------------------------------
using Eigen::MatrixXf;
using Eigen::VectorXf;
using PlayerIndex = unsigned short;
using Dimension = int;
using Point2 = Eigen::Vector2f;
using PointList2 = std::vector<Point2>;
using Time = double;
using Clock = std::chrono::system_clock;

template <typename T> using PlayerPtrMap = std::unordered_map<PlayerIndex, std::shared_ptr<T>>;
template <typename T> using PlayerPtrMultiMap = std::unordered_multimap<PlayerIndex, std::shared_ptr<T>>;
template <typename T> using PlayerMap = std::unordered_map<PlayerIndex, T>;
template <typename T> using PlayerMultiMap = std::unordered_multimap<PlayerIndex, T>;
template <typename T> using PtrVector = std::vector<std::shared_ptr<T>>;

namespace constants {
static constexpr float kGravity = 9.81;
static constexpr float kSmallNumber = 1e-4;
static constexpr float kInfinity = std::numeric_limits<float>::infinity();
static constexpr float kInvalidValue = std::numeric_limits<float>::quiet_NaN();
static constexpr float kDefaultLambda = 0.0;
static constexpr float kDefaultMu = 10.0;
}  // namespace constants

namespace time {
static constexpr Time kTimeStep = 0.1;
static constexpr Time kTimeHorizon = 10.0;
static constexpr size_t kNumTimeSteps =
    static_cast<size_t>((kTimeHorizon + constants::kSmallNumber) / kTimeStep);  // = 100
}  // namespace time

template <typename T>
inline constexpr T sgn(T x) {
  if constexpr (std::is_signed<T>::value) return (T(0) < x) - (x < T(0));
  return T(0) < x;
}
template <typename T>
inline T signed_sqrt(T x) { return sgn(x) * std::sqrt(std::abs(x)); }

// ---- include/ilqgames/utils/operating_point.h:55-85 ------------------------------------------
struct OperatingPoint {
  std::vector<VectorXf> xs;
  std::vector<std::vector<VectorXf>> us;
  Time t0;

  OperatingPoint(size_t num_time_steps, PlayerIndex num_players, Time initial_time)
      : xs(num_time_steps), us(num_time_steps, std::vector<VectorXf>(num_players)), t0(initial_time) {}

  template <typename MultiPlayerSystemType>
  OperatingPoint(size_t num_time_steps, Time initial_time,
                 const std::shared_ptr<const MultiPlayerSystemType>& dynamics)
      : OperatingPoint(num_time_steps, dynamics->NumPlayers(), initial_time) {
    CHECK_NOTNULL(dynamics.get());
    for (size_t kk = 0; kk < num_time_steps; kk++) {
      xs[kk] = VectorXf::Zero(dynamics->XDim());
      for (PlayerIndex ii = 0; ii < dynamics->NumPlayers(); ii++)
        us[kk][ii] = VectorXf::Zero(dynamics->UDim(ii));
    }
  }

  void swap(OperatingPoint& other) {
    xs.swap(other.xs);
    us.swap(other.us);
    std::swap(t0, other.t0);
  }
};

// ---- include/ilqgames/utils/strategy.h:59-85 -------------------------------------------------
struct Strategy {
  std::vector<MatrixXf> Ps;
  std::vector<VectorXf> alphas;

  Strategy(size_t horizon, Dimension xdim, Dimension udim) : Ps(horizon), alphas(horizon) {
    for (size_t ii = 0; ii < horizon; ii++) {
      Ps[ii] = MatrixXf::Zero(udim, xdim);
      alphas[ii] = VectorXf::Zero(udim);
    }
  }

  // u = u_ref - P dx - alpha
  VectorXf operator()(size_t time_index, const VectorXf& delta_x, const VectorXf& u_ref) const {
    return u_ref - Ps[time_index] * delta_x - alphas[time_index];
  }

  size_t NumVariables() const {
    CHECK_EQ(Ps.size(), alphas.size());
    return Ps.size() * (Ps.front().size() + alphas.front().size());
  }
};

// ---- include/ilqgames/utils/linear_dynamics_approximation.h:53-72 ----------------------------
struct LinearDynamicsApproximation {
  MatrixXf A;
  std::vector<MatrixXf> Bs;

  LinearDynamicsApproximation() {}
  template <typename MultiPlayerSystemType>
  explicit LinearDynamicsApproximation(const MultiPlayerSystemType& system)
      : A(MatrixXf::Identity(system.XDim(), system.XDim())), Bs(system.NumPlayers()) {
    for (size_t ii = 0; ii < system.NumPlayers(); ii++)
      Bs[ii] = MatrixXf::Zero(system.XDim(), system.UDim(ii));
  }
};

// ---- include/ilqgames/utils/quadratic_cost_approximation.h:55-86 -----------------------------
struct SingleCostApproximation {
  MatrixXf hess;
  VectorXf grad;

  SingleCostApproximation(const MatrixXf& hessian, const VectorXf& gradient) : hess(hessian), grad(gradient) {
    CHECK_EQ(hess.rows(), hess.cols());
    CHECK_EQ(hess.rows(), grad.size());
  }
  SingleCostApproximation(Dimension dim, float regularization = 0.0)
      : hess(regularization * MatrixXf::Identity(dim, dim)), grad(VectorXf::Zero(dim)) {}
};

struct QuadraticCostApproximation {
  SingleCostApproximation state;
  PlayerMap<SingleCostApproximation> control;
  explicit QuadraticCostApproximation(Dimension xdim, float regularization = 0.0) : state(xdim, regularization) {}
};

// ---- include/ilqgames/utils/solver_log.h:60-175, src/solver_log.cpp:113-171 -------------------
class SolverLog {
 public:
  SolverLog() {}
  SolverLog(const SolverLog&) = delete;
  SolverLog& operator=(const SolverLog&) = delete;

  void AddSolverIterate(const OperatingPoint& operating_point, const std::vector<Strategy>& strategies,
                        const std::vector<float>& total_costs, Time cumulative_runtime, bool was_converged) {
    operating_points_.push_back(operating_point);
    strategies_.push_back(strategies);
    total_player_costs_.push_back(total_costs);
    cumulative_runtimes_.push_back(cumulative_runtime);
    was_converged_.push_back(was_converged);
  }
  void AddLog(const SolverLog& log) {
    for (size_t ii = 0; ii < log.NumIterates(); ii++)
      AddSolverIterate(log.operating_points_[ii], log.strategies_[ii], log.total_player_costs_[ii],
                       log.cumulative_runtimes_[ii], log.was_converged_[ii]);
  }
  bool WasConverged() const { return was_converged_.back(); }
  bool WasConverged(size_t idx) const { return was_converged_[idx]; }
  Time InitialTime() const { return NumIterates() > 0 ? operating_points_[0].t0 : 0.0; }
  PlayerIndex NumPlayers() const { return (PlayerIndex)strategies_[0].size(); }
  size_t NumIterates() const { return operating_points_.size(); }
  std::vector<float> TotalCosts() const { return total_player_costs_.back(); }
  const std::vector<Strategy>& InitialStrategies() const { return strategies_.front(); }
  const OperatingPoint& InitialOperatingPoint() const { return operating_points_.front(); }
  const std::vector<Strategy>& FinalStrategies() const { return strategies_.back(); }
  const OperatingPoint& FinalOperatingPoint() const { return operating_points_.back(); }
  VectorXf State(size_t iterate, size_t time_index) const { return operating_points_[iterate].xs[time_index]; }
  float State(size_t iterate, size_t time_index, Dimension dim) const { return operating_points_[iterate].xs[time_index](dim); }
  VectorXf Control(size_t iterate, size_t time_index, PlayerIndex player) const {
    return operating_points_[iterate].us[time_index][player];
  }
  MatrixXf P(size_t iterate, size_t time_index, PlayerIndex player) const { return strategies_[iterate][player].Ps[time_index]; }
  VectorXf alpha(size_t iterate, size_t time_index, PlayerIndex player) const { return strategies_[iterate][player].alphas[time_index]; }
  size_t TimeToIndex(Time t) const {
    return static_cast<size_t>(std::max<Time>(constants::kSmallNumber, t - InitialTime()) / time::kTimeStep);
  }
  Time IndexToTime(size_t idx) const { return InitialTime() + time::kTimeStep * static_cast<Time>(idx); }

  // The reference's on-disk format (src/solver_log.cpp:113-171): one directory per iterate under
  // <log dir>/<experiment_name>/ holding t0.txt, xs.txt (one state per line), u<player>.txt (one
  // control per line), costs.txt (one total cost per line) and cumulative_runtimes.txt.  Rows are
  // written like Eigen prints a transposed vector: default stream precision, single-space
  // separated, every entry right-aligned to the widest one.  The log directory is the
  // reference's compile-time ILQGAMES_LOG_DIR; here the environment variable of that name, or
  // "logs".
  bool Save(bool only_last_trajectory = false, const std::string& experiment_name = "experiment") const {
    const char* env = std::getenv("ILQGAMES_LOG_DIR");
    const std::string root = env ? env : "logs";
    const std::string dir_name = root + "/" + experiment_name;
    if (!MakeDirectory(root) || !MakeDirectory(dir_name)) return false;
    size_t start = 0;
    if (only_last_trajectory) start = operating_points_.size() - 1;
    for (size_t ii = start; ii < operating_points_.size(); ii++) {
      const OperatingPoint& op = operating_points_[ii];
      const std::string sub = dir_name + "/" + std::to_string(ii);
      if (!MakeDirectory(sub)) return false;
      std::ofstream file(sub + "/t0.txt");
      file << op.t0 << std::endl;
      file.close();
      file.open(sub + "/xs.txt");
      for (const VectorXf& x : op.xs) file << Row(x) << std::endl;
      file.close();
      file.open(sub + "/costs.txt");
      for (float c : total_player_costs_[ii]) file << c << std::endl;
      file.close();
      file.open(sub + "/cumulative_runtimes.txt");
      file << cumulative_runtimes_[ii] << std::endl;
      file.close();
      for (PlayerIndex jj = 0; jj < NumPlayers(); jj++) {
        file.open(sub + "/u" + std::to_string(jj) + ".txt");
        for (size_t kk = 0; kk < op.us.size(); kk++) file << Row(op.us[kk][jj]) << std::endl;
        file.close();
      }
      if (!file) return false;
    }
    return true;
  }

 private:
  static bool MakeDirectory(const std::string& name) {
    struct stat st;
    if (stat(name.c_str(), &st) == 0) return S_ISDIR(st.st_mode);
    return mkdir(name.c_str(), 0777) == 0;
  }
  static std::string Row(const VectorXf& v) {
    std::vector<std::string> cells;
    size_t width = 0;
    for (long a = 0; a < v.size(); a++) {
      std::ostringstream os;
      os << v(a);
      cells.push_back(os.str());
      width = std::max(width, cells.back().size());
    }
    std::ostringstream row;
    for (size_t a = 0; a < cells.size(); a++) row << (a ? " " : "") << std::setw((int)width) << cells[a];
    return row.str();
  }

  std::vector<OperatingPoint> operating_points_;
  std::vector<std::vector<Strategy>> strategies_;
  std::vector<std::vector<float>> total_player_costs_;
  std::vector<Time> cumulative_runtimes_;
  std::vector<bool> was_converged_;
};

// ---- include/ilqgames/solver/solver_params.h:50-84 (same fields, same defaults) --------------
struct SolverParams {
  float convergence_tolerance = 1e-1;
  size_t max_solver_iters = 1000;
  bool linesearch = true;
  float initial_alpha_scaling = 0.5;
  float geometric_alpha_scaling = 0.5;
  size_t max_backtracking_steps = 10;
  float expected_decrease_fraction = 0.1;
  float state_regularization = 0.0;    // dead in the reference too
  float control_regularization = 0.0;
  bool open_loop = false;
  size_t unconstrained_solver_max_iters = 10;
  float geometric_mu_scaling = 1.1;
  float geometric_mu_downscaling = 0.5;
  float geometric_lambda_downscaling = 0.5;
  float constraint_error_tolerance = 1e-1;
  bool reset_problem = true;
  bool reset_lambdas = true;
  bool reset_mu = true;
};

}  // namespace ilqgames

#endif
