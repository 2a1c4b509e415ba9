// ===========================================================================
// ilqg_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference algorithm (HJReachability/ilqgames) for the
// one hot path this repository accelerates.  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load the
// library built from this file; the product (ilqgames_b200/) never does.
//
// Parity pinning (details in DESIGN.md section 6):
//  1. against outputs of the reference itself.  The reference's build system cannot run here
//     (cmake + Eigen3 + glog + gflags are absent), but `make -C oracle ref` compiles the
//     reference's OWN src/*.cpp, unmodified, against the stand-in headers under oracle/ref_shim
//     (the reference's own 51-test gtest suite passes on that build: `make ref-test`).
//     tests/golden/make_ref_golden.py runs it and commits tests/golden/ref_*.npz; this
//     restatement reproduces those fixtures BIT FOR BIT (tests/test_ref_pins.py): ILQSolver
//     iterates, final strategies, AugmentedLagrangianSolver results with multipliers, the open-loop
//     solver, Problem::SetUpNextRecedingHorizon, MultiPlayerIntegrableSystem::Integrate(t0, t, ..),
//     on the three headline example problems and on the eleven other examples of the reference
//     that are not built on the flat systems (every dynamics and cost class those use is restated
//     here, also the ones the CUDA library does not implement yet).  What that does not pin is the
//     rounding of Eigen's own dense kernels (stand-in: plain loops).
//  2. against the reference's own tests, transcribed with file:line citations in
//     tests/test_oracle_pins.py:
//   * geometry golden values     test/test_polyline2.cpp:52-125,
//                                test/test_line_segment2.cpp:57-102
//   * LQ solve known answer      test/test_lq_solver.cpp:292-317 (Lyapunov, 1e-4), Nash checks
//                                :319-379, open loop vs feedback :381-436
//   * analytic derivatives vs finite differences
//                                test/test_quadraticization.cpp:138-201,
//                                test/test_linearization.cpp:142-196
//   * player cost known answers  test/test_player_cost.cpp:84-122
//
// Every function cites the reference file:line it follows.  Scalars are `real`
// (float by default = the reference's MatrixXf/VectorXf; build with
// -DILQG_ORACLE_DOUBLE for an fp64 accuracy yardstick).  Expressions keep the
// reference's literal types (double literals, float variables) so the C++
// usual arithmetic conversions reproduce its mixed precision.
// ===========================================================================
#include "../include/ilqg.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#ifdef ILQG_ORACLE_DOUBLE
typedef double real;
#else
typedef float real;
#endif

namespace {

// include/ilqgames/utils/types.h:113-126
constexpr float kSmallNumber = 1e-4;
constexpr real kInfinity = std::numeric_limits<real>::infinity();
constexpr float kDefaultMu = 10.0;

// include/ilqgames/utils/types.h:152-165 (sgn)
template <typename T>
inline T sgn(T x) {
  return (T(0) < x) - (x < T(0));
}

// ----------------------------- geometry ------------------------------------
struct Point2 {
  real x, y;
};

// include/ilqgames/geometry/line_segment2.h:52-62
struct Segment {
  Point2 p1, p2;
  real length;
  Point2 unit;
};

Segment MakeSegment(Point2 a, Point2 b) {
  Segment s;
  s.p1 = a;
  s.p2 = b;
  const real dx = a.x - b.x, dy = a.y - b.y;
  s.length = std::sqrt(dx * dx + dy * dy);
  s.unit.x = (b.x - a.x) / s.length;
  s.unit.y = (b.y - a.y) / s.length;
  return s;
}

// src/line_segment2.cpp:48-54
bool SegmentSide(const Segment& s, Point2 q) {
  const real rx = q.x - s.p1.x, ry = q.y - s.p1.y;
  const real cross = rx * s.unit.y - s.unit.x * ry;
  return cross > 0.0;
}

// src/line_segment2.cpp:56-100
Point2 SegmentClosestPoint(const Segment& s, Point2 q, bool* is_endpoint,
                           real* signed_sq) {
  const real rx = q.x - s.p1.x, ry = q.y - s.p1.y;
  const real dot = rx * s.unit.x + ry * s.unit.y;
  const real cross = rx * s.unit.y - s.unit.x * ry;
  const real cross_sign = sgn(cross);
  if (dot < 0.0) {
    *is_endpoint = true;
    *signed_sq = cross_sign * (rx * rx + ry * ry);
    return s.p1;
  } else if (dot > s.length) {
    *is_endpoint = true;
    const real ex = q.x - s.p2.x, ey = q.y - s.p2.y;
    *signed_sq = cross_sign * (ex * ex + ey * ey);
    return s.p2;
  }
  *is_endpoint = false;
  *signed_sq = cross_sign * cross * cross;
  Point2 p;
  p.x = s.p1.x + dot * s.unit.x;
  p.y = s.p1.y + dot * s.unit.y;
  return p;
}

struct Polyline {
  std::vector<Segment> segs;  // src/polyline2.cpp:49-60
};

inline bool SamePointExact(Point2 a, Point2 b) { return a.x == b.x && a.y == b.y; }

// src/polyline2.cpp:105-174
Point2 PolylineClosestPoint(const Polyline& pl, Point2 q, bool* is_vertex,
                            int* segment_idx_out, real* signed_sq_out,
                            bool* is_endpoint) {
  real closest_sq = kInfinity;
  Point2 closest = {0, 0};
  int segment_idx = 0;
  bool vertex = false;
  const int ns = (int)pl.segs.size();
  for (int c = 0; c < ns; c++) {
    const Segment& s = pl.segs[c];
    bool seg_endpoint;
    real cur_sq;
    const Point2 cur = SegmentClosestPoint(s, q, &seg_endpoint, &cur_sq);
    if (std::abs(cur_sq) < std::abs(closest_sq)) {
      if (seg_endpoint && (c > 0 || SamePointExact(cur, s.p2)) &&
          (c < ns - 1 || SamePointExact(cur, s.p1))) {
        const Segment shortcut =
            SamePointExact(cur, s.p1)
                ? MakeSegment(pl.segs[c - 1].p1, s.p2)
                : MakeSegment(s.p1, pl.segs[c + 1].p2);
        cur_sq *= SegmentSide(shortcut, q) ? sgn(cur_sq) : -sgn(cur_sq);
      }
      closest_sq = cur_sq;
      closest = cur;
      vertex = seg_endpoint;
      segment_idx = c;
    }
  }
  if (is_vertex) *is_vertex = vertex;
  if (segment_idx_out) *segment_idx_out = segment_idx;
  if (signed_sq_out) *signed_sq_out = closest_sq;
  if (is_endpoint) {
    auto same = [](Point2 a, Point2 b) {
      const real dx = a.x - b.x, dy = a.y - b.y;
      return dx * dx + dy * dy < kSmallNumber;
    };
    *is_endpoint = same(closest, pl.segs.front().p1) || same(closest, pl.segs.back().p2);
  }
  return closest;
}

// ------------------------------ problem ------------------------------------
struct Problem {
  ilqg_problem_desc d;
  ilqg_solver_params p;
  int T, N, n, M;
  int uoff[ILQG_MAX_PLAYERS + 1];
  std::vector<Polyline> polylines;
  // control-cost pairs (i,j) sorted by (i,j)
  int num_pairs;
  int pair_i[ILQG_MAX_PAIRS], pair_j[ILQG_MAX_PAIRS];
  int pair_Roff[ILQG_MAX_PAIRS], pair_roff[ILQG_MAX_PAIRS];
  int pair_of[ILQG_MAX_PLAYERS][ILQG_MAX_PLAYERS];
  int R_floats, r_floats;
  int num_constraints;
  int constraint_slot[ILQG_MAX_COSTS];  // cost record -> lambda slot or -1
  std::vector<int> lambda_index;         // kk -> TimeIndex
  int first_step[ILQG_MAX_COSTS];        // FinalTimeCost gate: record c counts at kk >= first_step[c]
};

// FinalTimeCost::Evaluate / Quadraticize (include/ilqgames/cost/final_time_cost.h:64-77): active
// iff t >= initial_time_ + threshold_time_, with t = RelativeTime(kk) = kk * kTimeStep
// (relative_time_tracker.h:63-65; src/ilq_solver.cpp:236,475) and initial_time_ the tracker's
// static, which Problem::SyncToExistingProblem moves (src/problem.cpp:120).
void UpdateCostGates(Problem* pr, double tracker_initial_time) {
  for (int c = 0; c < pr->d.num_costs; c++) {
    const double threshold = pr->d.costs[c].active_from;
    int first = 0;
    if (threshold != 0.0) {
      first = pr->T;
      for (int kk = 0; kk < pr->T; kk++)
        if (static_cast<double>(kk) * pr->d.time_step >= tracker_initial_time + threshold) { first = kk; break; }
    }
    pr->first_step[c] = first;
  }
}

inline bool IsConstraintKind(int kind) {
  return kind == ILQG_CONSTRAINT_PROXIMITY || kind == ILQG_CONSTRAINT_SINGLE_DIMENSION ||
         kind == ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE;
}

int BuildProblem(const ilqg_problem_desc* desc, const ilqg_solver_params* params,
                 Problem* pr) {
  pr->d = *desc;
  pr->p = *params;
  const ilqg_problem_desc& d = pr->d;
  if (d.num_time_steps < 2 || d.num_time_steps > ILQG_MAX_TIME_STEPS)
    return ILQG_ERR_INVALID_ARGUMENT;
  if (d.num_players < 1 || d.num_players > ILQG_MAX_PLAYERS) return ILQG_ERR_INVALID_ARGUMENT;
  if (d.xdim < 1 || d.xdim > ILQG_MAX_XDIM) return ILQG_ERR_INVALID_ARGUMENT;
  if (d.num_costs < 0 || d.num_costs > ILQG_MAX_COSTS) return ILQG_ERR_INVALID_ARGUMENT;
  pr->T = d.num_time_steps;
  pr->N = d.num_players;
  pr->n = d.xdim;
  pr->uoff[0] = 0;
  for (int i = 0; i < pr->N; i++) {
    if (d.udim[i] < 1) return ILQG_ERR_INVALID_ARGUMENT;
    pr->uoff[i + 1] = pr->uoff[i] + d.udim[i];
  }
  pr->M = pr->uoff[pr->N];
  if (pr->M > ILQG_MAX_UDIM) return ILQG_ERR_INVALID_ARGUMENT;

  // polylines: src/polyline2.cpp:49-60
  pr->polylines.clear();
  for (int p = 0; p < d.num_polylines; p++) {
    Polyline pl;
    for (int q = d.polyline_start[p] + 1; q < d.polyline_start[p + 1]; q++) {
      Point2 a = {(real)d.polyline_points[q - 1][0], (real)d.polyline_points[q - 1][1]};
      Point2 b = {(real)d.polyline_points[q][0], (real)d.polyline_points[q][1]};
      pl.segs.push_back(MakeSegment(a, b));
    }
    if (pl.segs.empty()) return ILQG_ERR_INVALID_ARGUMENT;
    pr->polylines.push_back(pl);
  }

  // Control pairs: player i always owns an (i,i) block -- LQFeedbackSolver
  // CHECKs it exists (src/lq_feedback_solver.cpp:139-141) -- plus any (i,j) a
  // control cost / constraint names (src/player_cost.cpp:59-86).
  bool has[ILQG_MAX_PLAYERS][ILQG_MAX_PLAYERS] = {};
  for (int i = 0; i < pr->N; i++) has[i][i] = true;
  for (int c = 0; c < d.num_costs; c++) {
    const ilqg_cost_desc& cd = d.costs[c];
    if (cd.player < 0 || cd.player >= pr->N || cd.arg >= pr->N) return ILQG_ERR_INVALID_ARGUMENT;
    if (cd.arg >= 0) has[cd.player][cd.arg] = true;
  }
  pr->num_pairs = 0;
  pr->R_floats = pr->r_floats = 0;
  for (int i = 0; i < pr->N; i++)
    for (int j = 0; j < pr->N; j++) {
      pr->pair_of[i][j] = -1;
      if (!has[i][j]) continue;
      const int p = pr->num_pairs++;
      pr->pair_of[i][j] = p;
      pr->pair_i[p] = i;
      pr->pair_j[p] = j;
      pr->pair_Roff[p] = pr->R_floats;
      pr->pair_roff[p] = pr->r_floats;
      pr->R_floats += d.udim[j] * d.udim[j];
      pr->r_floats += d.udim[j];
    }

  pr->num_constraints = 0;
  for (int c = 0; c < d.num_costs; c++) {
    pr->constraint_slot[c] = IsConstraintKind(d.costs[c].kind) ? pr->num_constraints++ : -1;
    if (pr->constraint_slot[c] >= 0 && d.costs[c].active_from != 0.0) return ILQG_ERR_INVALID_ARGUMENT;
    // ExtremeValueCost members: plain, ungated costs
    if (d.costs[c].group < 0) return ILQG_ERR_INVALID_ARGUMENT;
    if (d.costs[c].group > 0 && (pr->constraint_slot[c] >= 0 || d.costs[c].active_from != 0.0))
      return ILQG_ERR_INVALID_ARGUMENT;
  }
  UpdateCostGates(pr, d.initial_time);

  // RelativeTimeTracker::RelativeTime / TimeIndex in double,
  // include/ilqgames/utils/relative_time_tracker.h:63-72.
  pr->lambda_index.resize(pr->T);
  for (int kk = 0; kk < pr->T; kk++) {
    const double t = static_cast<double>(kk) * d.time_step;
    long idx = static_cast<long>(static_cast<size_t>((t - d.initial_time) / d.time_step));
    idx = std::max(0L, std::min<long>(idx, pr->T - 1));
    pr->lambda_index[kk] = (int)idx;
  }
  return ILQG_OK;
}

// ---------------------------- per-instance state ---------------------------
struct Instance {
  std::vector<real> x0;
  std::vector<real> lq_x0;  // ILQG_LQ_X0: the x0 argument of a stand-alone LQSolver::Solve
  // Problem::operating_point_ / strategies_ (problem.h:171-172): the warm start
  // every Solve() begins from; only OverwriteSolution changes it.
  std::vector<real> prob_xs, prob_us, prob_Ps, prob_alphas;
  std::vector<real> xs, us;        // current operating point [T][n], [T][M]
  std::vector<real> Ps, alphas;    // current strategies [T][M][n], [T][M]
  std::vector<real> lqPs, lqAlphas;  // raw LQ solution
  std::vector<real> A, B;          // [T][n][n], [T][n][M]
  std::vector<real> Q, l, R, r;    // [T][N][n][n], [T][N][n], [T][R_floats], [T][r_floats]
  std::vector<real> dxs;           // [T][n]
  std::vector<real> lambdas;       // [ncon][T]
  real mu;
  real last_merit, expected_decrease, step;
  std::vector<real> total_costs;
  // PlayerCost::time_of_extreme_cost_ (player_cost.h:151) seen from two places: TotalCosts
  // writes te_new; quadraticization reads te_quad.  te_quad <- te_new right before each LQ
  // solve, which reproduces the reference's ordering: the quadraticization an LQ solve
  // consumes was computed inside the previous MeritFunction call, i.e. BEFORE the TotalCosts
  // call that followed it (src/ilq_solver.cpp:146,158).
  std::vector<int> te_quad, te_new;
  int status, iters, backtracks;
  real max_constraint_error;
  // AugmentedLagrangianSolver::Solve per instance (ilqg_al_begin / ilqg_al_advance):
  // 0 = not in an AL solve, 1 = active, 2 = finished (solve_begin leaves it alone)
  int al_state, al_iterates, al_success;
};

// ------------------------------- dynamics ----------------------------------
// xdot for all subsystems: src/concatenated_dynamical_system.cpp:69-84 and the
// SinglePlayer*::Evaluate / Air3D::Evaluate bodies.
void EvaluateDynamics(const Problem& pr, const real* x, const real* u, real* xdot) {
  const ilqg_problem_desc& d = pr.d;
  for (int s = 0; s < d.num_subsystems; s++) {
    const ilqg_subsystem_desc& sd = d.subsystems[s];
    const real* xs = x + sd.x_offset;
    real* xd = xdot + sd.x_offset;
    const real* us = u + pr.uoff[sd.first_player];
    switch (sd.kind) {
      case ILQG_DYN_CAR6D: {  // single_player_car_6d.h:102-113
        const real L = sd.params[0];
        xd[0] = xs[4] * std::cos(xs[2]);
        xd[1] = xs[4] * std::sin(xs[2]);
        xd[2] = (xs[4] / L) * std::tan(xs[3]);
        xd[3] = us[0];
        xd[4] = xs[5];
        xd[5] = us[1];
        break;
      }
      case ILQG_DYN_CAR5D: {  // single_player_car_5d.h:102-113
        const real L = sd.params[0];
        xd[0] = xs[4] * std::cos(xs[2]);
        xd[1] = xs[4] * std::sin(xs[2]);
        xd[2] = (xs[4] / L) * std::tan(xs[3]);
        xd[3] = us[0];
        xd[4] = us[1];
        break;
      }
      case ILQG_DYN_DUBINS: {  // single_player_dubins_car.h:93-101
        const real v = sd.params[0];
        xd[0] = v * std::cos(xs[2]);
        xd[1] = v * std::sin(xs[2]);
        xd[2] = us[0];
        break;
      }
      case ILQG_DYN_UNICYCLE4D: {  // single_player_unicycle_4d.h:90-99
        xd[0] = xs[3] * std::cos(xs[2]);
        xd[1] = xs[3] * std::sin(xs[2]);
        xd[2] = us[0];
        xd[3] = us[1];
        break;
      }
      case ILQG_DYN_POINT_MASS_2D: {  // single_player_point_mass_2d.h:91-101
        xd[0] = xs[2];
        xd[1] = xs[3];
        xd[2] = us[0];
        xd[3] = us[1];
        break;
      }
      case ILQG_DYN_TWO_PLAYER_UNICYCLE4D: {  // two_player_unicycle_4d.h:104-117
        const real* u2 = u + pr.uoff[sd.first_player + 1];
        xd[0] = xs[3] * std::cos(xs[2]) + u2[0];
        xd[1] = xs[3] * std::sin(xs[2]) + u2[1];
        xd[2] = us[0];
        xd[3] = us[1];
        break;
      }
      case ILQG_DYN_AIR3D: {  // air_3d.h:114-127
        const real ve = sd.params[0], vp = sd.params[1];
        const real u1 = us[0];
        const real u2 = u[pr.uoff[sd.first_player + 1]];
        xd[0] = -ve + vp * std::cos(xs[2]) + u1 * xs[1];
        xd[1] = vp * std::sin(xs[2]) - u1 * xs[0];
        xd[2] = u2 - u1;
        break;
      }
      default:
        break;
    }
  }
}

// MultiPlayerDynamicalSystem::Integrate, src/multi_player_dynamical_system.cpp:52-77.
// RK4 with 2 substeps; the double `dt` narrows to the vector scalar type when it
// multiplies a VectorXf (Eigen scalar promotion), as do the 0.5/2.0/6.0 literals.
// `substeps` is how often the reference's loop `for (t = t0; t < t0 + interval - 0.5 * dt; t += dt)`
// runs: 2 for any interval that is not vanishingly small next to t0 (Rk4Substeps below).
void Integrate(const Problem& pr, const real* x0, const real* u, real* xout, double time_interval,
               int substeps = 2) {
  const int n = pr.n;
  const double dt_d = time_interval / static_cast<double>(2);
  const real dt = (real)dt_d;
  real x[ILQG_MAX_XDIM], k1[ILQG_MAX_XDIM], k2[ILQG_MAX_XDIM], k3[ILQG_MAX_XDIM],
      k4[ILQG_MAX_XDIM], tmp[ILQG_MAX_XDIM];
  for (int a = 0; a < n; a++) x[a] = x0[a];
  for (int sub = 0; sub < substeps; sub++) {
    EvaluateDynamics(pr, x, u, k1);
    for (int a = 0; a < n; a++) { k1[a] = dt * k1[a]; tmp[a] = x[a] + (real)0.5 * k1[a]; }
    EvaluateDynamics(pr, tmp, u, k2);
    for (int a = 0; a < n; a++) { k2[a] = dt * k2[a]; tmp[a] = x[a] + (real)0.5 * k2[a]; }
    EvaluateDynamics(pr, tmp, u, k3);
    for (int a = 0; a < n; a++) { k3[a] = dt * k3[a]; tmp[a] = x[a] + k3[a]; }
    EvaluateDynamics(pr, tmp, u, k4);
    for (int a = 0; a < n; a++) {
      k4[a] = dt * k4[a];
      x[a] += (k1[a] + (real)2.0 * (k2[a] + k3[a]) + k4[a]) / (real)6.0;
    }
  }
  for (int a = 0; a < n; a++) xout[a] = x[a];
}

void Integrate(const Problem& pr, const real* x0, const real* u, real* xout) {
  Integrate(pr, x0, u, xout, pr.d.time_step);
}

// The trip count of the RK4 loop of MultiPlayerDynamicalSystem::Integrate
// (src/multi_player_dynamical_system.cpp:61-65), evaluated in double like the reference does.
int Rk4Substeps(double t0, double time_interval) {
  const double dt = time_interval / static_cast<double>(2);
  int count = 0;
  for (double t = t0; t < t0 + time_interval - 0.5 * dt && count < 4; t += dt) count++;
  return count;
}

// ConcatenatedDynamicalSystem::Linearize, src/concatenated_dynamical_system.cpp:86-107
// + SinglePlayerCar6D::Linearize (single_player_car_6d.h:115-138),
// SinglePlayerUnicycle4D::Linearize (single_player_unicycle_4d.h:101-116),
// Air3D::Linearize (air_3d.h:129-149).  A = I (+ dt * df/dx), B = dt * df/du
// (forward Euler).  `kTimeStep` is a double: products with it are
// evaluated in double and narrowed on assignment.
void Linearize(const Problem& pr, const real* x, const real* u, real* A, real* B) {
  const int n = pr.n, M = pr.M;
  const double kTimeStep = pr.d.time_step;
  for (int a = 0; a < n * n; a++) A[a] = 0;
  for (int a = 0; a < n; a++) A[a * n + a] = 1;
  for (int a = 0; a < n * M; a++) B[a] = 0;
  for (int s = 0; s < pr.d.num_subsystems; s++) {
    const ilqg_subsystem_desc& sd = pr.d.subsystems[s];
    const int o = sd.x_offset;
    const real* xs = x + o;
    const int uo = pr.uoff[sd.first_player];
#define AA(r, c) A[(o + (r)) * n + (o + (c))]
#define BB(r, c) B[(o + (r)) * M + (uo + (c))]
    switch (sd.kind) {
      case ILQG_DYN_CAR6D: {
        const real L = sd.params[0];
        const real ctheta = std::cos(xs[2]) * kTimeStep;
        const real stheta = std::sin(xs[2]) * kTimeStep;
        const real cphi = std::cos(xs[3]);
        const real tphi = std::tan(xs[3]);
        AA(0, 2) += -xs[4] * stheta;
        AA(0, 4) += ctheta;
        AA(1, 2) += xs[4] * ctheta;
        AA(1, 4) += stheta;
        AA(2, 3) += xs[4] * kTimeStep / (L * cphi * cphi);
        AA(2, 4) += tphi * kTimeStep / L;
        AA(4, 5) += kTimeStep;
        BB(3, 0) = kTimeStep;
        BB(5, 1) = kTimeStep;
        break;
      }
      case ILQG_DYN_CAR5D: {  // single_player_car_5d.h:115-138
        const real L = sd.params[0];
        const real ctheta = std::cos(xs[2]) * kTimeStep;
        const real stheta = std::sin(xs[2]) * kTimeStep;
        const real cphi = std::cos(xs[3]);
        const real tphi = std::tan(xs[3]);
        AA(0, 2) += -xs[4] * stheta;
        AA(0, 4) += ctheta;
        AA(1, 2) += xs[4] * ctheta;
        AA(1, 4) += stheta;
        AA(2, 3) += xs[4] * kTimeStep / (L * cphi * cphi);
        AA(2, 4) += tphi * kTimeStep / L;
        BB(3, 0) = kTimeStep;
        BB(4, 1) = kTimeStep;
        break;
      }
      case ILQG_DYN_DUBINS: {  // single_player_dubins_car.h:103-116
        const real v = sd.params[0];
        const real ctheta = std::cos(xs[2]) * kTimeStep;
        const real stheta = std::sin(xs[2]) * kTimeStep;
        AA(0, 2) += -v * stheta;
        AA(1, 2) += v * ctheta;
        BB(2, 0) = kTimeStep;
        break;
      }
      case ILQG_DYN_UNICYCLE4D: {
        const real ctheta = std::cos(xs[2]) * kTimeStep;
        const real stheta = std::sin(xs[2]) * kTimeStep;
        AA(0, 2) += -xs[3] * stheta;
        AA(0, 3) += ctheta;
        AA(1, 2) += xs[3] * ctheta;
        AA(1, 3) += stheta;
        BB(2, 0) = kTimeStep;
        BB(3, 1) = kTimeStep;
        break;
      }
      case ILQG_DYN_POINT_MASS_2D: {  // single_player_point_mass_2d.h:103-111
        AA(0, 2) += kTimeStep;
        AA(1, 3) += kTimeStep;
        BB(2, 0) = kTimeStep;
        BB(3, 1) = kTimeStep;
        break;
      }
      case ILQG_DYN_TWO_PLAYER_UNICYCLE4D: {  // two_player_unicycle_4d.h:119-137
        const int uo2 = pr.uoff[sd.first_player + 1];
        const real ctheta = std::cos(xs[2]) * kTimeStep;
        const real stheta = std::sin(xs[2]) * kTimeStep;
        AA(0, 2) += -xs[3] * stheta;
        AA(0, 3) += ctheta;
        AA(1, 2) += xs[3] * ctheta;
        AA(1, 3) += stheta;
        BB(2, 0) = kTimeStep;
        BB(3, 1) = kTimeStep;
        B[(o + 0) * M + (uo2 + 0)] = kTimeStep;
        B[(o + 1) * M + (uo2 + 1)] = kTimeStep;
        break;
      }
      case ILQG_DYN_AIR3D: {
        const real vp = sd.params[1];
        const real u1 = u[uo];
        const int uo2 = pr.uoff[sd.first_player + 1];
        const real ctheta = std::cos(xs[2]) * kTimeStep;
        const real stheta = std::sin(xs[2]) * kTimeStep;
        AA(0, 1) += u1 * kTimeStep;
        AA(0, 2) -= vp * stheta;
        AA(1, 0) -= u1 * kTimeStep;
        AA(1, 2) += vp * ctheta;
        BB(0, 0) = xs[1] * kTimeStep;
        BB(1, 0) = -xs[0] * kTimeStep;
        BB(2, 0) = -kTimeStep;
        B[(o + 2) * M + uo2] = kTimeStep;
        break;
      }
      default:
        break;
    }
#undef AA
#undef BB
  }
}

// ------------------------------ costs --------------------------------------
// Constraint::Mu, include/ilqgames/constraint/constraint.h:112-117
inline real ConstraintMu(const ilqg_cost_desc& cd, real mu, real lambda, real g) {
  if (!cd.is_equality && g <= kSmallNumber && std::abs(lambda) <= kSmallNumber) return 0.0;
  return mu;
}

// Constraint::ModifyDerivatives, src/constraint.cpp:63-89
void ModifyDerivatives(const ilqg_cost_desc& cd, real lambda, real mu_global, real g,
                       real* dx, real* ddx, real* dy, real* ddy, real* dxdy) {
  const real mu = ConstraintMu(cd, mu_global, lambda, g);
  const real new_dx = lambda * *dx + mu * g * *dx;
  const real new_ddx = lambda * *ddx + mu * (*dx * *dx + g * *ddx);
  if (dy) {
    const real new_dy = lambda * *dy + mu * g * *dy;
    const real new_ddy = lambda * *ddy + mu * (*dy * *dy + g * *ddy);
    const real new_dxdy = lambda * *dxdy + mu * (*dy * *dx + g * *dxdy);
    *dy = new_dy;
    *ddy = new_ddy;
    *dxdy = new_dxdy;
  }
  *dx = new_dx;
  *ddx = new_ddx;
}

// Cost::Evaluate for every record kind (value of the cost, or g(x) for constraints).
real EvaluateRecord(const Problem& pr, const ilqg_cost_desc& cd, const real* in, int dim) {
  const real weight_ = cd.weight;
  switch (cd.kind) {
    case ILQG_COST_QUADRATIC: {  // src/quadratic_cost.cpp:51-63
      const real nominal_ = cd.value;
      if (cd.dim[0] >= 0) {
        const real delta = in[cd.dim[0]] - nominal_;
        return 0.5 * weight_ * delta * delta;
      }
      real sq = 0;
      for (int a = 0; a < dim; a++) sq += (in[a] - nominal_) * (in[a] - nominal_);
      return 0.5 * weight_ * sq;
    }
    case ILQG_COST_QUADRATIC_POLYLINE2: {  // src/quadratic_polyline2_cost.cpp:52-69
      real ssd;
      bool is_endpoint;
      PolylineClosestPoint(pr.polylines[cd.polyline], {in[cd.dim[0]], in[cd.dim[1]]}, nullptr,
                           nullptr, &ssd, &is_endpoint);
      if (is_endpoint) ssd = 0.0;
      return 0.5 * weight_ * std::abs(ssd);
    }
    case ILQG_COST_PROXIMITY: {  // src/proximity_cost.cpp:52-62
      const real threshold_ = cd.value;
      const real threshold_sq_ = threshold_ * threshold_;
      const real dx = in[cd.dim[0]] - in[cd.dim[2]];
      const real dy = in[cd.dim[1]] - in[cd.dim[3]];
      const real delta_sq = dx * dx + dy * dy;
      if (delta_sq >= threshold_sq_) return 0.0;
      const real gap = threshold_ - std::sqrt(delta_sq);
      return 0.5 * weight_ * gap * gap;
    }
    case ILQG_COST_SIGNED_DISTANCE: {  // src/signed_distance_cost.cpp:50-62
      const real dx = in[cd.dim[0]] - in[cd.dim[2]];
      const real dy = in[cd.dim[1]] - in[cd.dim[3]];
      const real cost = (real)cd.value - std::hypot(dx, dy);
      return cd.flag ? cost : -cost;
    }
    case ILQG_COST_QUADRATIC_DIFFERENCE: {  // src/quadratic_difference_cost.cpp:50-60
      real total = 0.0;
      for (int ii = 0; ii < cd.flag; ii++) {
        const real diff = in[cd.dim[ii]] - in[cd.dim[2 + ii]];
        total += diff * diff;
      }
      return 0.5 * weight_ * total;
    }
    case ILQG_COST_SEMIQUADRATIC: {  // src/semiquadratic_cost.cpp:51-59
      const real diff = in[cd.dim[0]] - cd.value;
      const bool oriented_right_ = cd.flag != 0;
      if ((diff > 0.0 && oriented_right_) || (diff < 0.0 && !oriented_right_))
        return 0.5 * weight_ * diff * diff;
      return 0.0;
    }
    case ILQG_COST_SEMIQUADRATIC_POLYLINE2: {  // src/semiquadratic_polyline2_cost.cpp:52-73
      const real threshold_ = cd.value;
      const real signed_squared_threshold_ = sgn(threshold_) * threshold_ * threshold_;
      const bool oriented_right_ = cd.flag != 0;
      real ssd;
      bool is_endpoint;
      PolylineClosestPoint(pr.polylines[cd.polyline], {in[cd.dim[0]], in[cd.dim[1]]}, nullptr,
                           nullptr, &ssd, &is_endpoint);
      if (is_endpoint) return 0.0;
      const bool active = (ssd > signed_squared_threshold_ && oriented_right_) ||
                          (ssd < signed_squared_threshold_ && !oriented_right_);
      if (!active) return 0.0;
      const real signed_distance = sgn(ssd) * std::sqrt(std::abs(ssd));
      const real diff = signed_distance - threshold_;
      return 0.5 * weight_ * diff * diff;
    }
    case ILQG_COST_POLYLINE2_SIGNED_DISTANCE: {  // src/polyline2_signed_distance_cost.cpp:52-65
      real ssd;
      PolylineClosestPoint(pr.polylines[cd.polyline], {in[cd.dim[0]], in[cd.dim[1]]}, nullptr,
                           nullptr, &ssd, nullptr);
      if (!cd.flag) ssd *= -1.0;
      return sgn(ssd) * std::sqrt(std::abs(ssd)) - cd.value;
    }
    case ILQG_CONSTRAINT_PROXIMITY: {  // src/proximity_constraint.cpp:56-62
      const real dx = in[cd.dim[0]] - in[cd.dim[2]];
      const real dy = in[cd.dim[1]] - in[cd.dim[3]];
      const real value = std::hypot(dx, dy) - cd.value;
      return cd.flag ? value : -value;
    }
    case ILQG_CONSTRAINT_SINGLE_DIMENSION: {  // single_dimension_constraint.h:68-70
      return cd.flag ? in[cd.dim[0]] - cd.value : cd.value - in[cd.dim[0]];
    }
    case ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE: {  // src/polyline2_signed_distance_constraint.cpp:58-70
      real ssd;
      PolylineClosestPoint(pr.polylines[cd.polyline], {in[cd.dim[0]], in[cd.dim[1]]}, nullptr,
                           nullptr, &ssd, nullptr);
      const real value = sgn(ssd) * std::sqrt(std::abs(ssd)) - cd.value;
      return cd.flag ? value : -value;
    }
  }
  return 0;
}

// Cost::Quadraticize for every record kind: accumulates into hess (dim x dim
// row-major) and grad.  `lambda`, `mu` only matter for constraints.
void QuadraticizeRecord(const Problem& pr, const ilqg_cost_desc& cd, const real* in, int dim,
                        real lambda, real mu, real* hess, real* grad) {
  const real weight_ = cd.weight;
#define H(r, c) hess[(r) * dim + (c)]
  switch (cd.kind) {
    case ILQG_COST_QUADRATIC: {  // src/quadratic_cost.cpp:65-94
      const real nominal_ = cd.value;
      if (cd.dim[0] >= 0) {
        const int d0 = cd.dim[0];
        const real delta = in[d0] - nominal_;
        const real dx = weight_ * delta;
        const real ddx = weight_;
        grad[d0] += dx;
        H(d0, d0) += ddx;
      } else {
        for (int a = 0; a < dim; a++) {
          grad[a] += weight_ * (in[a] - nominal_);
          H(a, a) = H(a, a) + weight_;
        }
      }
      break;
    }
    case ILQG_COST_QUADRATIC_POLYLINE2: {  // src/quadratic_polyline2_cost.cpp:71-126
      const int xidx_ = cd.dim[0], yidx_ = cd.dim[1];
      const Polyline& pl = pr.polylines[cd.polyline];
      const Point2 cur = {in[xidx_], in[yidx_]};
      bool is_vertex, is_endpoint;
      int seg;
      const Point2 closest = PolylineClosestPoint(pl, cur, &is_vertex, &seg, nullptr, &is_endpoint);
      if (is_endpoint) return;
      real ddx = weight_;
      real ddy = weight_;
      real dxdy = 0.0;
      real dx = weight_ * (cur.x - closest.x);
      real dy = weight_ * (cur.y - closest.y);
      if (!is_vertex) {
        const Segment& s = pl.segs[seg];
        const real relx = cur.x - s.p1.x, rely = cur.y - s.p1.y;
        ddx = weight_ * s.unit.y * s.unit.y;
        ddy = weight_ * s.unit.x * s.unit.x;
        dxdy = -weight_ * s.unit.x * s.unit.y;
        const real w_cross = weight_ * (relx * s.unit.y - rely * s.unit.x);
        dx = w_cross * s.unit.y;
        dy = -w_cross * s.unit.x;
      }
      grad[xidx_] += dx;
      grad[yidx_] += dy;
      H(xidx_, xidx_) += ddx;
      H(yidx_, yidx_) += ddy;
      H(xidx_, yidx_) += dxdy;
      H(yidx_, xidx_) += dxdy;
      break;
    }
    case ILQG_COST_PROXIMITY: {  // src/proximity_cost.cpp:63-122
      const int xidx1_ = cd.dim[0], yidx1_ = cd.dim[1], xidx2_ = cd.dim[2], yidx2_ = cd.dim[3];
      const real threshold_ = cd.value;
      const real threshold_sq_ = threshold_ * threshold_;
      const real dx = in[xidx1_] - in[xidx2_];
      const real dy = in[yidx1_] - in[yidx2_];
      const real delta_sq = dx * dx + dy * dy;
      if (delta_sq >= threshold_sq_) return;
      const real delta = std::sqrt(delta_sq);
      const real gap = threshold_ - delta;
      const real weight_delta = weight_ / delta;
      const real dx_delta = dx / delta;
      const real dy_delta = dy / delta;
      const real ddx1 = -weight_delta * gap * dx;
      const real ddy1 = -weight_delta * gap * dy;
      const real hess_x1x1 = weight_delta * (dx_delta * (gap * dx_delta + dx) - gap);
      const real hess_y1y1 = weight_delta * (dy_delta * (gap * dy_delta + dy) - gap);
      const real hess_x1y1 = weight_delta * (dx_delta * (gap * dy_delta + dy));
      grad[xidx1_] += ddx1;
      grad[xidx2_] -= ddx1;
      grad[yidx1_] += ddy1;
      grad[yidx2_] -= ddy1;
      H(xidx1_, xidx1_) += hess_x1x1;
      H(xidx1_, xidx2_) -= hess_x1x1;
      H(xidx2_, xidx1_) -= hess_x1x1;
      H(xidx2_, xidx2_) += hess_x1x1;
      H(yidx1_, yidx1_) += hess_y1y1;
      H(yidx1_, yidx2_) -= hess_y1y1;
      H(yidx2_, yidx1_) -= hess_y1y1;
      H(yidx2_, yidx2_) += hess_y1y1;
      H(xidx1_, yidx1_) += hess_x1y1;
      H(yidx1_, xidx1_) += hess_x1y1;
      H(xidx1_, yidx2_) -= hess_x1y1;
      H(yidx2_, xidx1_) -= hess_x1y1;
      H(xidx2_, yidx1_) -= hess_x1y1;
      H(yidx1_, xidx2_) -= hess_x1y1;
      H(xidx2_, yidx2_) += hess_x1y1;
      H(yidx2_, xidx2_) += hess_x1y1;
      break;
    }
    case ILQG_COST_SIGNED_DISTANCE: {  // src/signed_distance_cost.cpp:64-112
      const int xdim1_ = cd.dim[0], ydim1_ = cd.dim[1], xdim2_ = cd.dim[2], ydim2_ = cd.dim[3];
      const real s = cd.flag ? 1.0 : -1.0;
      const real delta_x = in[xdim1_] - in[xdim2_];
      const real delta_y = in[ydim1_] - in[ydim2_];
      const real norm = std::hypot(delta_x, delta_y);
      const real norm_3 = norm * norm * norm;
      const real dx1 = -s * delta_x / norm;
      const real dy1 = -s * delta_y / norm;
      const real ddx1 = -s * delta_y * delta_y / norm_3;
      const real ddy1 = -s * delta_x * delta_x / norm_3;
      const real dx1dy1 = s * delta_x * delta_y / norm_3;
      grad[xdim1_] += dx1;
      grad[ydim1_] += dy1;
      grad[xdim2_] -= dx1;
      grad[ydim2_] -= dy1;
      H(xdim1_, xdim1_) += ddx1;
      H(ydim1_, ydim1_) += ddy1;
      H(xdim1_, ydim1_) += dx1dy1;
      H(ydim1_, xdim1_) += dx1dy1;
      H(xdim2_, xdim2_) += ddx1;
      H(ydim2_, ydim2_) += ddy1;
      H(xdim2_, ydim2_) += dx1dy1;
      H(ydim2_, xdim2_) += dx1dy1;
      H(xdim1_, xdim2_) -= ddx1;
      H(xdim1_, ydim2_) -= dx1dy1;
      H(ydim1_, xdim2_) -= dx1dy1;
      H(ydim1_, ydim2_) -= ddy1;
      H(xdim2_, xdim1_) -= ddx1;
      H(xdim2_, ydim1_) -= dx1dy1;
      H(ydim2_, xdim1_) -= dx1dy1;
      H(ydim2_, ydim1_) -= ddy1;
      break;
    }
    case ILQG_COST_QUADRATIC_DIFFERENCE: {  // src/quadratic_difference_cost.cpp:62-91
      for (int ii = 0; ii < cd.flag; ii++) {
        const int a = cd.dim[ii], b = cd.dim[2 + ii];
        const real dx = weight_ * (in[a] - in[b]);
        const real dy = -dx;
        H(a, a) += weight_;
        H(b, b) += weight_;
        H(a, b) += -weight_;
        H(b, a) += -weight_;
        grad[a] += dx;
        grad[b] += dy;
      }
      break;
    }
    case ILQG_COST_SEMIQUADRATIC: {  // src/semiquadratic_cost.cpp:63-85
      const int d0 = cd.dim[0];
      const bool oriented_right_ = cd.flag != 0;
      const real diff = in[d0] - cd.value;
      if ((diff < 0.0 && oriented_right_) || (diff > 0.0 && !oriented_right_)) return;
      const real dx = weight_ * diff;
      const real ddx = weight_;
      grad[d0] += dx;
      H(d0, d0) += ddx;
      break;
    }
    case ILQG_COST_SEMIQUADRATIC_POLYLINE2: {  // src/semiquadratic_polyline2_cost.cpp:75-142
      const int xidx_ = cd.dim[0], yidx_ = cd.dim[1];
      const real threshold_ = cd.value;
      const real signed_squared_threshold_ = sgn(threshold_) * threshold_ * threshold_;
      const bool oriented_right_ = cd.flag != 0;
      const Polyline& pl = pr.polylines[cd.polyline];
      const Point2 cur = {in[xidx_], in[yidx_]};
      real ssd;
      bool is_vertex, is_endpoint;
      int seg;
      const Point2 closest = PolylineClosestPoint(pl, cur, &is_vertex, &seg, &ssd, &is_endpoint);
      const bool active = (ssd > signed_squared_threshold_ && oriented_right_) ||
                          (ssd < signed_squared_threshold_ && !oriented_right_);
      if (!active) return;
      if (is_endpoint) return;
      real ddx = weight_;
      real ddy = weight_;
      real dxdy = 0.0;
      real scaling = std::sqrt(std::abs(ssd));
      scaling = (scaling - std::abs(threshold_)) / scaling;
      real dx = weight_ * scaling * (cur.x - closest.x);
      real dy = weight_ * scaling * (cur.y - closest.y);
      if (!is_vertex) {
        const Segment& s = pl.segs[seg];
        const real relx = cur.x - s.p1.x, rely = cur.y - s.p1.y;
        ddx = weight_ * s.unit.y * s.unit.y;
        ddy = weight_ * s.unit.x * s.unit.x;
        const real cross_term = -weight_ * s.unit.x * s.unit.y;
        dxdy = cross_term;
        const real w_cross = weight_ * (relx * s.unit.y - rely * s.unit.x - threshold_);
        dx = w_cross * s.unit.y;
        dy = -w_cross * s.unit.x;
      }
      grad[xidx_] += dx;
      grad[yidx_] += dy;
      H(xidx_, xidx_) += ddx;
      H(yidx_, yidx_) += ddy;
      H(xidx_, yidx_) += dxdy;
      H(yidx_, xidx_) += dxdy;
      break;
    }
    case ILQG_COST_POLYLINE2_SIGNED_DISTANCE: {  // src/polyline2_signed_distance_cost.cpp:67-121
      const int xidx_ = cd.dim[0], yidx_ = cd.dim[1];
      const Polyline& pl = pr.polylines[cd.polyline];
      const Point2 cur = {in[xidx_], in[yidx_]};
      bool is_vertex;
      real ssd;
      int seg;
      const Point2 closest = PolylineClosestPoint(pl, cur, &is_vertex, &seg, &ssd, nullptr);
      if (!cd.flag) ssd *= -1.0;
      const real sign = sgn(ssd);
      const real distance = std::sqrt(std::abs(ssd));
      const real delta_x = cur.x - closest.x;
      const real delta_y = cur.y - closest.y;
      real dx = sign * delta_x / distance;
      real dy = sign * delta_y / distance;
      const real denom = ssd * distance;
      real ddx = delta_y * delta_y / denom;
      real ddy = delta_x * delta_x / denom;
      real dxdy = -delta_x * delta_y / denom;
      if (!is_vertex) {
        const Segment& s = pl.segs[seg];
        dx = s.unit.y;
        dy = -s.unit.x;
        ddx = 0.0;
        ddy = 0.0;
        dxdy = 0.0;
      }
      grad[xidx_] += dx;
      grad[yidx_] += dy;
      H(xidx_, xidx_) += ddx;
      H(yidx_, yidx_) += ddy;
      H(xidx_, yidx_) += dxdy;
      H(yidx_, xidx_) += dxdy;
      break;
    }
    case ILQG_CONSTRAINT_PROXIMITY: {  // src/proximity_constraint.cpp:64-116
      const int xidx1_ = cd.dim[0], yidx1_ = cd.dim[1], xidx2_ = cd.dim[2], yidx2_ = cd.dim[3];
      const real threshold_ = cd.value;
      const real dx = in[xidx1_] - in[xidx2_];
      const real dy = in[yidx1_] - in[yidx2_];
      const real prox = std::hypot(dx, dy);
      const real sign = (cd.flag) ? 1.0 : -1.0;
      const real g = sign * (prox - threshold_);
      const real rel_dx = dx / prox;
      const real rel_dy = dy / prox;
      real grad_x1 = sign * rel_dx;
      real grad_y1 = sign * rel_dy;
      real hess_x1x1 = sign * (1.0 - rel_dx * rel_dx) / prox;
      real hess_y1y1 = sign * (1.0 - rel_dy * rel_dy) / prox;
      real hess_x1y1 = -sign * rel_dx * rel_dy / prox;
      ModifyDerivatives(cd, lambda, mu, g, &grad_x1, &hess_x1x1, &grad_y1, &hess_y1y1, &hess_x1y1);
      grad[xidx1_] += grad_x1;
      grad[xidx2_] -= grad_x1;
      grad[yidx1_] += grad_y1;
      grad[yidx2_] -= grad_y1;
      H(xidx1_, xidx1_) += hess_x1x1;
      H(xidx1_, xidx2_) -= hess_x1x1;
      H(xidx2_, xidx1_) -= hess_x1x1;
      H(xidx2_, xidx2_) += hess_x1x1;
      H(yidx1_, yidx1_) += hess_y1y1;
      H(yidx1_, yidx2_) -= hess_y1y1;
      H(yidx2_, yidx1_) -= hess_y1y1;
      H(yidx2_, yidx2_) += hess_y1y1;
      H(xidx1_, yidx1_) += hess_x1y1;
      H(xidx1_, yidx2_) -= hess_x1y1;
      H(xidx2_, yidx1_) -= hess_x1y1;
      H(xidx2_, yidx2_) += hess_x1y1;
      H(yidx1_, xidx1_) += hess_x1y1;
      H(yidx1_, xidx2_) -= hess_x1y1;
      H(yidx2_, xidx1_) -= hess_x1y1;
      H(yidx2_, xidx2_) += hess_x1y1;
      break;
    }
    case ILQG_CONSTRAINT_SINGLE_DIMENSION: {  // single_dimension_constraint.h:74-96
      const int dim_ = cd.dim[0];
      const real sign = (cd.flag) ? 1.0 : -1.0;
      const real x = in[dim_];
      const real g = sign * (x - cd.value);
      real dx = sign;
      real ddx = 0.0;
      ModifyDerivatives(cd, lambda, mu, g, &dx, &ddx, nullptr, nullptr, nullptr);
      grad[dim_] += dx;
      H(dim_, dim_) += ddx;
      break;
    }
    case ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE: {  // src/polyline2_signed_distance_constraint.cpp:72-145
      const int xidx_ = cd.dim[0], yidx_ = cd.dim[1];
      const Polyline& pl = pr.polylines[cd.polyline];
      const real threshold_ = cd.value;
      const bool keep_left_ = cd.flag != 0;
      bool is_vertex;
      int seg;
      real signed_distance_sq;
      const Point2 closest_point =
          PolylineClosestPoint(pl, {in[xidx_], in[yidx_]}, &is_vertex, &seg, &signed_distance_sq, nullptr);
      const Segment& closest_segment = pl.segs[seg];
      const real s = sgn(signed_distance_sq);
      const real x = in[xidx_];
      const real y = in[yidx_];
      real px = closest_segment.p1.x;
      real py = closest_segment.p1.y;
      real rx = x - px;
      real ry = y - py;
      real d_sq = rx * rx + ry * ry;
      real d = std::sqrt(d_sq);
      const real ux = closest_segment.unit.x;
      const real uy = closest_segment.unit.y;
      const real sign = (keep_left_) ? 1.0 : -1.0;
      const real signed_root = sgn(signed_distance_sq) * std::sqrt(std::abs(signed_distance_sq));
      const real g = (keep_left_) ? signed_root - threshold_ : threshold_ - signed_root;
      real dx = sign * uy;
      real ddx = 0.0;
      real dy = -sign * ux;
      real ddy = 0.0;
      real dxdy = 0.0;
      if (is_vertex) {
        px = closest_point.x;
        py = closest_point.y;
        rx = x - px;
        ry = y - py;
        d_sq = (rx * rx + ry * ry);
        d = std::sqrt(d_sq);
        dx = sign * s * rx / d;
        ddx = sign * s * (d_sq - px * px - x * x + 2 * px * x) / (d_sq * d);
        dxdy = -sign * s * rx * ry / (d_sq * d);
        dy = sign * s * ry / d;
        ddy = sign * s * (d_sq - py * py - y * y + 2 * py * y) / (d_sq * d);
      }
      ModifyDerivatives(cd, lambda, mu, g, &dx, &ddx, &dy, &ddy, &dxdy);
      grad[xidx_] += dx;
      grad[yidx_] += dy;
      H(xidx_, xidx_) += ddx;
      H(xidx_, yidx_) += dxdy;
      H(yidx_, xidx_) += dxdy;
      H(yidx_, yidx_) += ddy;
      break;
    }
  }
#undef H
}

// ExtremeValueCost::ExtremeCost, src/extreme_value_cost.cpp:64-84: of the group of records that
// starts at `first` (consecutive records of one player and argument with the same group id) the
// member with the largest (group_is_min: smallest) value, the first one on ties.  *next = one past
// the group.  A NaN member never wins; if all are NaN the reference dereferences an unset pointer
// -- the first member stands in here.
int ExtremeMember(const Problem& pr, int first, const real* x, const real* u, int* next, real* value_out) {
  const ilqg_problem_desc& d = pr.d;
  const ilqg_cost_desc& head = d.costs[first];
  const bool is_min = head.group_is_min != 0;
  real extreme = is_min ? kInfinity : -kInfinity;
  int chosen = first, c = first;
  for (; c < d.num_costs; c++) {
    const ilqg_cost_desc& cd = d.costs[c];
    if (cd.group != head.group || cd.player != head.player || cd.arg != head.arg) break;
    const real value = cd.arg < 0 ? EvaluateRecord(pr, cd, x, pr.n)
                                  : EvaluateRecord(pr, cd, u + pr.uoff[cd.arg], d.udim[cd.arg]);
    if ((is_min && value < extreme) || (!is_min && value > extreme)) {
      extreme = value;
      chosen = c;
    }
  }
  *next = c;
  if (value_out) *value_out = extreme;
  return chosen;
}

// PlayerCost::Quadraticize / QuadraticizeControlCosts, src/player_cost.cpp:194-225,
// as dispatched by ILQSolver::ComputeCostQuadraticization (src/ilq_solver.cpp:471-490).
// Writes Q_i, l_i (all players) and R_p, r_p (all pairs) for time step kk.
void QuadraticizeStep(const Problem& pr, const Instance& in, int kk, const real* x, const real* u,
                      real* Q, real* l, real* R, real* r) {
  const int n = pr.n, N = pr.N;
  const ilqg_problem_desc& d = pr.d;
  for (int i = 0; i < N; i++) {
    real* Qi = Q + (size_t)i * n * n;
    real* li = l + (size_t)i * n;
    // QuadraticCostApproximation(xdim, state_regularization_): reg * I, zero grad.
    for (int a = 0; a < n * n; a++) Qi[a] = 0;
    for (int a = 0; a < n; a++) Qi[a * n + a] = d.state_regularization[i];
    for (int a = 0; a < n; a++) li[a] = 0;
  }
  for (int p = 0; p < pr.num_pairs; p++) {
    const int i = pr.pair_i[p], j = pr.pair_j[p], mj = d.udim[j];
    real* Rp = R + pr.pair_Roff[p];
    real* rp = r + pr.pair_roff[p];
    for (int a = 0; a < mj * mj; a++) Rp[a] = 0;
    for (int a = 0; a < mj; a++) rp[a] = 0;
    // A control block only exists in the reference if some control cost or
    // constraint names it (player_cost.cpp:70-80); it then starts at reg * I.
    // The (i,i) block always exists in every in-scope example.
    for (int a = 0; a < mj; a++) Rp[a * mj + a] = d.control_regularization[i];
  }
  for (int c0 = 0, next = 0; c0 < d.num_costs; c0 = next) {
    next = c0 + 1;
    // an ExtremeValueCost quadraticizes its extreme member only (src/extreme_value_cost.cpp:56-62)
    const int c = d.costs[c0].group > 0 ? ExtremeMember(pr, c0, x, u, &next, nullptr) : c0;
    const ilqg_cost_desc& cd = d.costs[c];
    if (kk < pr.first_step[c]) continue;  // FinalTimeCost::Quadraticize returns before its cost's
    const int i = cd.player;
    const bool full = d.cost_structure[i] == ILQG_COST_SUM || in.te_quad[i] == kk;
    const int slot = pr.constraint_slot[c];
    const bool is_con = slot >= 0;
    // QuadraticizeControlCosts keeps only control COSTS (no constraints).
    if (!full && (cd.arg < 0 || is_con)) continue;
    const real lambda = is_con ? in.lambdas[(size_t)slot * pr.T + pr.lambda_index[kk]] : (real)0;
    if (cd.arg < 0) {
      QuadraticizeRecord(pr, cd, x, n, lambda, in.mu, Q + (size_t)i * n * n, l + (size_t)i * n);
    } else {
      const int p = pr.pair_of[i][cd.arg];
      QuadraticizeRecord(pr, cd, u + pr.uoff[cd.arg], d.udim[cd.arg], lambda, in.mu,
                         R + pr.pair_Roff[p], r + pr.pair_roff[p]);
    }
  }
}

// PlayerCost::Evaluate(t, x, us), src/player_cost.cpp:128-144: state + control
// COSTS only (no constraints).
real EvaluatePlayerCost(const Problem& pr, int i, int kk, const real* x, const real* u) {
  real total = 0.0;
  const ilqg_problem_desc& d = pr.d;
  for (int c0 = 0, next = 0; c0 < d.num_costs; c0 = next) {
    next = c0 + 1;
    if (d.costs[c0].player != i) continue;
    // ExtremeValueCost::Evaluate: the extreme member's value (src/extreme_value_cost.cpp:50-54)
    const int c = d.costs[c0].group > 0 ? ExtremeMember(pr, c0, x, u, &next, nullptr) : c0;
    const ilqg_cost_desc& cd = d.costs[c];
    if (cd.player != i || pr.constraint_slot[c] >= 0) continue;
    if (kk < pr.first_step[c]) continue;  // FinalTimeCost::Evaluate is 0 before its threshold
    if (cd.arg < 0)
      total += EvaluateRecord(pr, cd, x, pr.n);
    else
      total += EvaluateRecord(pr, cd, u + pr.uoff[cd.arg], d.udim[cd.arg]);
  }
  return total;
}

// ILQSolver::TotalCosts, src/ilq_solver.cpp:220-257
void TotalCosts(const Problem& pr, Instance& in) {
  const int T = pr.T, N = pr.N, n = pr.n, M = pr.M;
  for (int i = 0; i < N; i++) {
    const int cs = pr.d.cost_structure[i];
    in.total_costs[i] = cs == ILQG_COST_SUM ? (real)0.0 : cs == ILQG_COST_MAX ? -kInfinity : kInfinity;
  }
  for (int kk = 0; kk < T; kk++)
    for (int i = 0; i < N; i++) {
      const real cur = EvaluatePlayerCost(pr, i, kk, &in.xs[(size_t)kk * n], &in.us[(size_t)kk * M]);
      const int cs = pr.d.cost_structure[i];
      if (cs == ILQG_COST_SUM)
        in.total_costs[i] += cur;
      else if (cs == ILQG_COST_MAX && cur > in.total_costs[i]) {
        in.total_costs[i] = cur;
        in.te_new[i] = kk;
      } else if (cs == ILQG_COST_MIN) {
        if (cur < in.total_costs[i]) {
          in.total_costs[i] = cur;
          in.te_new[i] = kk;
        }
      }
    }
}

// ILQSolver::CurrentOperatingPoint, src/ilq_solver.cpp:174-206 with
// Strategy::operator(), include/ilqgames/utils/strategy.h:73-76.
void Rollout(const Problem& pr, const real* last_xs, const real* last_us, const real* Ps,
             const real* alphas, real* xs, real* us) {
  const int T = pr.T, n = pr.n, M = pr.M;
  real x[ILQG_MAX_XDIM], dx[ILQG_MAX_XDIM], xn[ILQG_MAX_XDIM];
  for (int a = 0; a < n; a++) x[a] = last_xs[a];
  for (int kk = 0; kk < T; kk++) {
    for (int a = 0; a < n; a++) dx[a] = x[a] - last_xs[(size_t)kk * n + a];
    for (int a = 0; a < n; a++) xs[(size_t)kk * n + a] = x[a];
    for (int c = 0; c < M; c++) {
      real acc = 0;
      const real* Prow = Ps + ((size_t)kk * M + c) * n;
      for (int a = 0; a < n; a++) acc += Prow[a] * dx[a];
      us[(size_t)kk * M + c] = last_us[(size_t)kk * M + c] - acc - alphas[(size_t)kk * M + c];
    }
    if (kk < T - 1) {
      Integrate(pr, x, &us[(size_t)kk * M], xn);
      for (int a = 0; a < n; a++) x[a] = xn[a];
    }
  }
}

// ILQSolver::ComputeLinearization, src/ilq_solver.cpp:437-455
void LinearizeAll(const Problem& pr, Instance& in) {
  const int T = pr.T, n = pr.n, M = pr.M;
  for (int kk = 0; kk < T; kk++)
    Linearize(pr, &in.xs[(size_t)kk * n], &in.us[(size_t)kk * M], &in.A[(size_t)kk * n * n],
              &in.B[(size_t)kk * n * M]);
}

// ILQSolver::ComputeCostQuadraticization, src/ilq_solver.cpp:471-490
void QuadraticizeAll(const Problem& pr, Instance& in) {
  const int T = pr.T, n = pr.n, M = pr.M, N = pr.N;
  for (int kk = 0; kk < T; kk++)
    QuadraticizeStep(pr, in, kk, &in.xs[(size_t)kk * n], &in.us[(size_t)kk * M],
                     &in.Q[(size_t)kk * N * n * n], &in.l[(size_t)kk * N * n],
                     &in.R[(size_t)kk * pr.R_floats], &in.r[(size_t)kk * pr.r_floats]);
}

// ------------------------- Householder QR solve ----------------------------
// Restates Eigen's HouseholderQR (householder_qr_inplace_unblocked +
// MatrixBase::makeHouseholder + solve = Q^T b then back-substitution) as called
// at src/lq_feedback_solver.cpp:180.  Eigen is an un-vendored, unpinned
// dependency (cmake/Dependencies.cmake:5); this follows its published
// algorithm, not its vectorised summation order.
// Solves S X = Y in place: S is m x m (destroyed), Y is m x c (becomes X).
void HouseholderQrSolve(real* S, int m, real* Y, int c) {
  std::vector<real> tau(m);
  for (int k = 0; k < m; k++) {
    real tail_sq = 0;
    for (int i = k + 1; i < m; i++) tail_sq += S[i * m + k] * S[i * m + k];
    const real c0 = S[k * m + k];
    real beta, t;
    const real tol = std::numeric_limits<real>::min();
    if (tail_sq <= tol) {
      t = 0;
      beta = c0;
      for (int i = k + 1; i < m; i++) S[i * m + k] = 0;
    } else {
      beta = std::sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < m; i++) S[i * m + k] /= (c0 - beta);
      t = (beta - c0) / beta;
    }
    tau[k] = t;
    S[k * m + k] = beta;
    // apply H = I - tau v v^T (v = [1; essential]) to the remaining columns
    for (int j = k + 1; j < m; j++) {
      real w = S[k * m + j];
      for (int i = k + 1; i < m; i++) w += S[i * m + k] * S[i * m + j];
      w *= t;
      S[k * m + j] -= w;
      for (int i = k + 1; i < m; i++) S[i * m + j] -= S[i * m + k] * w;
    }
    // ... and to the right-hand sides (Q^T Y)
    for (int j = 0; j < c; j++) {
      real w = Y[k * c + j];
      for (int i = k + 1; i < m; i++) w += S[i * m + k] * Y[i * c + j];
      w *= t;
      Y[k * c + j] -= w;
      for (int i = k + 1; i < m; i++) Y[i * c + j] -= S[i * m + k] * w;
    }
  }
  // back substitution with the upper triangle R
  for (int j = 0; j < c; j++)
    for (int i = m - 1; i >= 0; i--) {
      real acc = Y[i * c + j];
      for (int q = i + 1; q < m; q++) acc -= S[i * m + q] * Y[q * c + j];
      Y[i * c + j] = acc / S[i * m + i];
    }
}

// LQFeedbackSolver::Solve, src/lq_feedback_solver.cpp:71-244.
// x0arg is the `x0` argument (ILQSolver passes x0 - xs[0], ilq_solver.cpp:140-143).
void LQFeedbackSolve(const Problem& pr, Instance& in, const real* x0arg) {
  const int T = pr.T, n = pr.n, M = pr.M, N = pr.N;
  const ilqg_problem_desc& d = pr.d;
  std::vector<real> Zs((size_t)T * N * n * n), zetas((size_t)T * N * n);
  std::vector<real> S(M * M), Y(M * (n + 1)), F(n * n), beta(n), BiZi(ILQG_MAX_UDIM * n),
      tmp(n * n), tmpv(n);
  std::fill(in.lqPs.begin(), in.lqPs.end(), (real)0);  // Strategy ctor zero-fills (strategy.h:64-70)
  std::fill(in.lqAlphas.begin(), in.lqAlphas.end(), (real)0);

  auto Z = [&](int k, int i) { return &Zs[((size_t)k * N + i) * n * n]; };
  auto zeta = [&](int k, int i) { return &zetas[((size_t)k * N + i) * n]; };

  for (int i = 0; i < N; i++) {  // :102-105
    std::memcpy(Z(T - 1, i), &in.Q[((size_t)(T - 1) * N + i) * n * n], sizeof(real) * n * n);
    std::memcpy(zeta(T - 1, i), &in.l[((size_t)(T - 1) * N + i) * n], sizeof(real) * n);
  }

  for (int kk = T - 2; kk >= 0; kk--) {
    const real* A = &in.A[(size_t)kk * n * n];
    const real* B = &in.B[(size_t)kk * n * M];
    const real* Rk = &in.R[(size_t)kk * pr.R_floats];
    const real* rk = &in.r[(size_t)kk * pr.r_floats];
    // :119-161 build S, Y
    for (int i = 0; i < N; i++) {
      const int mi = d.udim[i], ro = pr.uoff[i];
      const real* Zi = Z(kk + 1, i);
      const real* zi = zeta(kk + 1, i);
      // BiZi = B_i^T Z_i  (mi x n)
      for (int a = 0; a < mi; a++)
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < n; q++) acc += B[q * M + ro + a] * Zi[q * n + c];
          BiZi[a * n + c] = acc;
        }
      const int pii = pr.pair_of[i][i];
      for (int j = 0; j < N; j++) {
        const int mj = d.udim[j], co = pr.uoff[j];
        for (int a = 0; a < mi; a++)
          for (int c = 0; c < mj; c++) {
            real acc = 0;
            for (int q = 0; q < n; q++) acc += BiZi[a * n + q] * B[q * M + co + c];
            if (i == j) acc = acc + Rk[pr.pair_Roff[pii] + a * mi + c];
            S[(ro + a) * M + co + c] = acc;
          }
      }
      for (int a = 0; a < mi; a++) {
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < n; q++) acc += BiZi[a * n + q] * A[q * n + c];
          Y[(ro + a) * (n + 1) + c] = acc;
        }
        real acc = 0;
        for (int q = 0; q < n; q++) acc += B[q * M + ro + a] * zi[q];
        Y[(ro + a) * (n + 1) + n] = acc + rk[pr.pair_roff[pii] + a];
      }
    }
    // :163-176 Gershgorin (column-wise)
    if (pr.p.adaptive_regularization) {
      for (int c = 0; c < M; c++) {
        real col1 = 0;
        for (int a = 0; a < M; a++) col1 += std::abs(S[a * M + c]);
        const real radius = col1 - std::abs(S[c * M + c]);
        const real eval_lo = S[c * M + c] - radius;
        constexpr float min_eval = 1e-3;
        if (eval_lo < min_eval) S[c * M + c] += radius + min_eval;
      }
    }
    // :180 X = S.householderQr().solve(Y)
    HouseholderQrSolve(S.data(), M, Y.data(), n + 1);
    real* Pk = &in.lqPs[(size_t)kk * M * n];
    real* ak = &in.lqAlphas[(size_t)kk * M];
    for (int a = 0; a < M; a++) {
      for (int c = 0; c < n; c++) Pk[a * n + c] = Y[a * (n + 1) + c];
      ak[a] = Y[a * (n + 1) + n];
    }
    // :189-194 F = A - sum B_i P_i ; beta = - sum B_i alpha_i
    for (int a = 0; a < n; a++) {
      for (int c = 0; c < n; c++) F[a * n + c] = A[a * n + c];
      beta[a] = 0;
    }
    for (int i = 0; i < N; i++) {
      const int mi = d.udim[i], ro = pr.uoff[i];
      for (int a = 0; a < n; a++) {
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < mi; q++) acc += B[a * M + ro + q] * Pk[(ro + q) * n + c];
          F[a * n + c] -= acc;
        }
        real acc = 0;
        for (int q = 0; q < mi; q++) acc += B[a * M + ro + q] * ak[ro + q];
        beta[a] -= acc;
      }
    }
    // :197-213 update Z, zeta
    for (int i = 0; i < N; i++) {
      const real* Zn = Z(kk + 1, i);
      const real* zn = zeta(kk + 1, i);
      real* Zc = Z(kk, i);
      real* zc = zeta(kk, i);
      const real* Qi = &in.Q[((size_t)kk * N + i) * n * n];
      const real* li = &in.l[((size_t)kk * N + i) * n];
      // zeta = F^T (zeta_next + Z_next beta) + l
      for (int a = 0; a < n; a++) {
        real acc = 0;
        for (int q = 0; q < n; q++) acc += Zn[a * n + q] * beta[q];
        tmpv[a] = zn[a] + acc;
      }
      for (int a = 0; a < n; a++) {
        real acc = 0;
        for (int q = 0; q < n; q++) acc += F[q * n + a] * tmpv[q];
        zc[a] = acc + li[a];
      }
      // Z = (F^T Z_next) F + Q
      for (int a = 0; a < n; a++)
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < n; q++) acc += F[q * n + a] * Zn[q * n + c];
          tmp[a * n + c] = acc;
        }
      for (int a = 0; a < n; a++)
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < n; q++) acc += tmp[a * n + q] * F[q * n + c];
          Zc[a * n + c] = acc + Qi[a * n + c];
        }
      // :206-212 control terms for every R_ij present
      for (int j = 0; j < N; j++) {
        const int p = pr.pair_of[i][j];
        if (p < 0) continue;
        const int mj = d.udim[j], co = pr.uoff[j];
        const real* Rij = Rk + pr.pair_Roff[p];
        const real* rij = rk + pr.pair_roff[p];
        real v[ILQG_MAX_UDIM];
        for (int a = 0; a < mj; a++) {
          real acc = 0;
          for (int q = 0; q < mj; q++) acc += Rij[a * mj + q] * ak[co + q];
          v[a] = acc - rij[a];
        }
        for (int a = 0; a < n; a++) {
          real acc = 0;
          for (int q = 0; q < mj; q++) acc += Pk[(co + q) * n + a] * v[q];
          zc[a] += acc;
        }
        real PtR[ILQG_MAX_XDIM * ILQG_MAX_UDIM];
        for (int a = 0; a < n; a++)
          for (int c = 0; c < mj; c++) {
            real acc = 0;
            for (int q = 0; q < mj; q++) acc += Pk[(co + q) * n + a] * Rij[q * mj + c];
            PtR[a * mj + c] = acc;
          }
        for (int a = 0; a < n; a++)
          for (int c = 0; c < n; c++) {
            real acc = 0;
            for (int q = 0; q < mj; q++) acc += PtR[a * mj + q] * Pk[(co + q) * n + c];
            Zc[a * n + c] += acc;
          }
      }
    }
  }
  // :217-241 forward pass for delta_xs (costates are unused downstream)
  std::vector<real> xstar(x0arg, x0arg + n), last(n);
  for (int kk = 0; kk < T; kk++) {
    for (int a = 0; a < n; a++) in.dxs[(size_t)kk * n + a] = xstar[a];
    const real* A = &in.A[(size_t)kk * n * n];
    const real* B = &in.B[(size_t)kk * n * M];
    last = xstar;
    for (int a = 0; a < n; a++) {
      real acc = 0;
      for (int q = 0; q < n; q++) acc += A[a * n + q] * last[q];
      xstar[a] = acc;
    }
    for (int i = 0; i < N; i++) {
      const int mi = d.udim[i], ro = pr.uoff[i];
      for (int a = 0; a < n; a++) {
        real acc = 0;
        for (int q = 0; q < mi; q++) acc += B[a * M + ro + q] * in.lqAlphas[(size_t)kk * M + ro + q];
        xstar[a] -= acc;
      }
    }
  }
}

// Eigen::LDLT stand-in used for chol_Rs_ (src/lq_open_loop_solver.cpp:124-126): symmetric-pivoted
// LDL^T of the lower triangle (largest remaining |diagonal| first), then solve for c right-hand
// sides.  Same arithmetic as oracle/ref_shim/Eigen/Dense (LDLTImpl) so that the oracle and the
// shim build of the reference agree bit for bit; Eigen's own LDLT is pivoted the same way but
// blocked differently (rounding-level differences, unpinned like the QR).
// A is m x m (row-major, destroyed), Bmat is m x c (row-major, becomes the solution).
void LdltSolve(real* A, int m, real* Bmat, int c) {
  std::vector<real> w((size_t)m * m), L((size_t)m * m, 0), dvec(m);
  std::vector<int> perm(m);
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) w[i * m + j] = i >= j ? A[i * m + j] : A[j * m + i];
  for (int i = 0; i < m; i++) perm[i] = i;
  for (int k = 0; k < m; k++) {
    int piv = k;
    for (int i = k + 1; i < m; i++)
      if (std::abs(w[i * m + i]) > std::abs(w[piv * m + piv])) piv = i;
    if (piv != k) {
      std::swap(perm[k], perm[piv]);
      for (int j = 0; j < m; j++) std::swap(w[k * m + j], w[piv * m + j]);
      for (int i = 0; i < m; i++) std::swap(w[i * m + k], w[i * m + piv]);
    }
    const real dk = w[k * m + k];
    dvec[k] = dk;
    if (dk == 0) {
      for (int i = k + 1; i < m; i++) w[i * m + k] = 0;
      continue;
    }
    for (int i = k + 1; i < m; i++) w[i * m + k] /= dk;
    for (int j = k + 1; j < m; j++)
      for (int i = j; i < m; i++) {
        w[i * m + j] -= w[i * m + k] * dk * w[j * m + k];
        w[j * m + i] = w[i * m + j];
      }
  }
  for (int j = 0; j < m; j++) {
    L[j * m + j] = 1;
    for (int i = j + 1; i < m; i++) L[i * m + j] = w[i * m + j];
  }
  std::vector<real> y(m);
  for (int j = 0; j < c; j++) {
    for (int i = 0; i < m; i++) y[i] = Bmat[perm[i] * c + j];
    for (int i = 0; i < m; i++)
      for (int k = 0; k < i; k++) y[i] -= L[i * m + k] * y[k];
    for (int i = 0; i < m; i++) y[i] = dvec[i] != 0 ? y[i] / dvec[i] : (real)0;
    for (int i = m - 1; i >= 0; i--)
      for (int k = i + 1; k < m; k++) y[i] -= L[k * m + i] * y[k];
    for (int i = 0; i < m; i++) Bmat[perm[i] * c + j] = y[i];
  }
}

// LQOpenLoopSolver::Solve, src/lq_open_loop_solver.cpp:73-195 (selected by SolverParams::open_loop,
// include/ilqgames/solver/ilq_solver.h:76-81).  P stays zero; alphas hold the (sign-flipped)
// open-loop controls; delta_xs is the optimal state trajectory x*.
void LQOpenLoopSolve(const Problem& pr, Instance& in, const real* x0arg) {
  const int T = pr.T, n = pr.n, M = pr.M, N = pr.N;
  const ilqg_problem_desc& d = pr.d;
  std::vector<real> Ms((size_t)T * N * n * n), ms((size_t)T * N * n);
  // per step: Lambda (:119-127), the intermediate term (:134-139), warped B_i = R_ii^-1 B_i^T
  // and warped r_i = R_ii^-1 r_ii (:125-126)
  std::vector<real> Lam((size_t)T * n * n), inter((size_t)T * n), WB((size_t)T * M * n), Wr((size_t)T * M);
  std::vector<real> t1(n * n), t2(n * n), X(n * n), tmpv(n), tmpv2(n), Rcopy(ILQG_MAX_UDIM * ILQG_MAX_UDIM),
      rhs(ILQG_MAX_UDIM * (n + 1)), LamCopy(n * n);
  std::fill(in.lqPs.begin(), in.lqPs.end(), (real)0);
  std::fill(in.lqAlphas.begin(), in.lqAlphas.end(), (real)0);
  auto Mk = [&](int k, int i) { return &Ms[((size_t)k * N + i) * n * n]; };
  auto mk = [&](int k, int i) { return &ms[((size_t)k * N + i) * n]; };
  // Lambda^-1 applied to c right-hand sides (qr_capital_lambdas_[kk].solve, :130,143-151,170)
  auto lam_solve = [&](int kk, real* rhs_nc, int c) {
    std::memcpy(LamCopy.data(), &Lam[(size_t)kk * n * n], sizeof(real) * n * n);
    HouseholderQrSolve(LamCopy.data(), n, rhs_nc, c);
  };

  for (int i = 0; i < N; i++) {  // :112-115
    std::memcpy(mk(T - 1, i), &in.l[((size_t)(T - 1) * N + i) * n], sizeof(real) * n);
    std::memcpy(Mk(T - 1, i), &in.Q[((size_t)(T - 1) * N + i) * n * n], sizeof(real) * n * n);
  }
  for (int kk = T - 2; kk >= 0; kk--) {
    const real* A = &in.A[(size_t)kk * n * n];
    const real* B = &in.B[(size_t)kk * n * M];
    const real* Rk = &in.R[(size_t)kk * pr.R_floats];
    const real* rk = &in.r[(size_t)kk * pr.r_floats];
    real* L = &Lam[(size_t)kk * n * n];
    for (int a = 0; a < n; a++)
      for (int c = 0; c < n; c++) L[a * n + c] = a == c ? 1 : 0;
    for (int i = 0; i < N; i++) {
      const int mi = d.udim[i], ro = pr.uoff[i], pii = pr.pair_of[i][i];
      real* Wi = &WB[((size_t)kk * M + ro) * n];   // mi x n
      real* wri = &Wr[(size_t)kk * M + ro];
      // warped_Bs = chol(R_ii).solve(B_i^T), warped_rs = chol(R_ii).solve(r_ii)
      for (int a = 0; a < mi; a++)
        for (int c = 0; c < n; c++) rhs[a * (n + 1) + c] = B[c * M + ro + a];
      for (int a = 0; a < mi; a++) rhs[a * (n + 1) + n] = rk[pr.pair_roff[pii] + a];
      std::memcpy(Rcopy.data(), Rk + pr.pair_Roff[pii], sizeof(real) * mi * mi);
      LdltSolve(Rcopy.data(), mi, rhs.data(), n + 1);
      for (int a = 0; a < mi; a++) {
        for (int c = 0; c < n; c++) Wi[a * n + c] = rhs[a * (n + 1) + c];
        wri[a] = rhs[a * (n + 1) + n];
      }
      // Lambda += (B_i * warped_B_i) * M_{k+1,i}
      const real* Mn = Mk(kk + 1, i);
      for (int a = 0; a < n; a++)
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < mi; q++) acc += B[a * M + ro + q] * Wi[q * n + c];
          t1[a * n + c] = acc;
        }
      for (int a = 0; a < n; a++)
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < n; q++) acc += t1[a * n + q] * Mn[q * n + c];
          L[a * n + c] += acc;
        }
    }
    // intermediate = - sum_i B_i (warped_B_i m_{k+1,i} + warped_r_i)
    real* it = &inter[(size_t)kk * n];
    for (int a = 0; a < n; a++) it[a] = 0;
    for (int i = 0; i < N; i++) {
      const int mi = d.udim[i], ro = pr.uoff[i];
      const real* Wi = &WB[((size_t)kk * M + ro) * n];
      const real* mn = mk(kk + 1, i);
      real v[ILQG_MAX_UDIM];
      for (int a = 0; a < mi; a++) {
        real acc = 0;
        for (int q = 0; q < n; q++) acc += Wi[a * n + q] * mn[q];
        v[a] = acc + Wr[(size_t)kk * M + ro + a];
      }
      for (int a = 0; a < n; a++) {
        real acc = 0;
        for (int q = 0; q < mi; q++) acc += B[a * M + ro + q] * v[q];
        it[a] -= acc;
      }
    }
    // X = Lambda^-1 A ; s = Lambda^-1 intermediate
    std::memcpy(X.data(), A, sizeof(real) * n * n);
    lam_solve(kk, X.data(), n);
    for (int a = 0; a < n; a++) tmpv[a] = it[a];
    lam_solve(kk, tmpv.data(), 1);
    for (int i = 0; i < N; i++) {
      const real* Mn = Mk(kk + 1, i);
      const real* mn = mk(kk + 1, i);
      const real* Qi = &in.Q[((size_t)kk * N + i) * n * n];
      const real* li = &in.l[((size_t)kk * N + i) * n];
      // M_k = Q + (A^T M_{k+1}) X
      for (int a = 0; a < n; a++)
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < n; q++) acc += A[q * n + a] * Mn[q * n + c];
          t2[a * n + c] = acc;
        }
      real* Mc = Mk(kk, i);
      for (int a = 0; a < n; a++)
        for (int c = 0; c < n; c++) {
          real acc = 0;
          for (int q = 0; q < n; q++) acc += t2[a * n + q] * X[q * n + c];
          Mc[a * n + c] = Qi[a * n + c] + acc;
        }
      // m_k = l + A^T (m_{k+1} + M_{k+1} s)
      for (int a = 0; a < n; a++) {
        real acc = 0;
        for (int q = 0; q < n; q++) acc += Mn[a * n + q] * tmpv[q];
        tmpv2[a] = mn[a] + acc;
      }
      real* mc = mk(kk, i);
      for (int a = 0; a < n; a++) {
        real acc = 0;
        for (int q = 0; q < n; q++) acc += A[q * n + a] * tmpv2[q];
        mc[a] = li[a] + acc;
      }
    }
  }
  // :158-192 forward pass
  std::vector<real> xstar(x0arg, x0arg + n), last(n);
  for (int kk = 0; kk < T - 1; kk++) {
    for (int a = 0; a < n; a++) in.dxs[(size_t)kk * n + a] = xstar[a];
    const real* A = &in.A[(size_t)kk * n * n];
    last = xstar;
    for (int a = 0; a < n; a++) {
      real acc = 0;
      for (int q = 0; q < n; q++) acc += A[a * n + q] * last[q];
      xstar[a] = acc + inter[(size_t)kk * n + a];
    }
    lam_solve(kk, xstar.data(), 1);
    for (int i = 0; i < N; i++) {
      const int mi = d.udim[i], ro = pr.uoff[i];
      const real* Wi = &WB[((size_t)kk * M + ro) * n];
      const real* Mn = Mk(kk + 1, i);
      const real* mn = mk(kk + 1, i);
      for (int a = 0; a < n; a++) {
        real acc = 0;
        for (int q = 0; q < n; q++) acc += Mn[a * n + q] * xstar[q];
        tmpv[a] = acc + mn[a];
      }
      for (int a = 0; a < mi; a++) {
        real acc = 0;
        for (int q = 0; q < n; q++) acc += Wi[a * n + q] * tmpv[q];
        in.lqAlphas[(size_t)kk * M + ro + a] = acc + Wr[(size_t)kk * M + ro + a];
      }
    }
  }
  for (int a = 0; a < n; a++) in.dxs[(size_t)(T - 1) * n + a] = xstar[a];
}

// LQSolver::Solve through ILQSolver::lq_solver_ (ilq_solver.h:76-81)
void LQSolve(const Problem& pr, Instance& in, const real* x0arg) {
  if (pr.p.open_loop) LQOpenLoopSolve(pr, in, x0arg);
  else LQFeedbackSolve(pr, in, x0arg);
}

// ILQSolver::ExpectedDecrease, src/ilq_solver.cpp:364-398 (as written)
real ExpectedDecrease(const Problem& pr, const Instance& in) {
  const int T = pr.T, n = pr.n, M = pr.M, N = pr.N;
  real expected_decrease = 0.0;
  for (int kk = 0; kk < T; kk++)
    for (int i = 0; i < N; i++) {
      const int mi = pr.d.udim[i], ro = pr.uoff[i], p = pr.pair_of[i][i];
      const real* Rii = &in.R[(size_t)kk * pr.R_floats + pr.pair_Roff[p]];
      const real* rii = &in.r[(size_t)kk * pr.r_floats + pr.pair_roff[p]];
      const real* neg_ui = &in.lqAlphas[(size_t)kk * M + ro];
      real acc = 0;  // (neg_ui^T * hess) * dLdu
      for (int c = 0; c < mi; c++) {
        real row = 0;
        for (int a = 0; a < mi; a++) row += neg_ui[a] * Rii[a * mi + c];
        acc += row * rii[c];
      }
      expected_decrease -= acc;
      if (kk > 0) {
        const real* Qi = &in.Q[((size_t)kk * N + i) * n * n];
        const real* li = &in.l[((size_t)kk * N + i) * n];
        const real* dx = &in.dxs[(size_t)kk * n];
        real acc2 = 0;
        for (int c = 0; c < n; c++) {
          real row = 0;
          for (int a = 0; a < n; a++) row += dx[a] * Qi[a * n + c];
          acc2 += row * li[c];
        }
        expected_decrease -= acc2;
      }
    }
  return expected_decrease;
}

// ILQSolver::MeritFunction, src/ilq_solver.cpp:400-435: requadraticize
// at the candidate point, then 0.5 * sum_k sum_i (|r_ii|^2 + [k>0] |l_i|^2).
real MeritFunction(const Problem& pr, Instance& in) {
  const int T = pr.T, n = pr.n, N = pr.N;
  QuadraticizeAll(pr, in);
  real merit = 0.0;
  for (int kk = 0; kk < T; kk++)
    for (int i = 0; i < N; i++) {
      const int mi = pr.d.udim[i], p = pr.pair_of[i][i];
      const real* rii = &in.r[(size_t)kk * pr.r_floats + pr.pair_roff[p]];
      real sq = 0;
      for (int a = 0; a < mi; a++) sq += rii[a] * rii[a];
      merit += sq;
      if (kk > 0) {
        const real* li = &in.l[((size_t)kk * N + i) * n];
        real sq2 = 0;
        for (int a = 0; a < n; a++) sq2 += li[a] * li[a];
        merit += sq2;
      }
    }
  return 0.5 * merit;
}

// ILQSolver::ModifyLQStrategies, src/ilq_solver.cpp:289-348 (+ CheckArmijoCondition
// :350-362, HasConverged ilq_solver.h:126-130, ScaleAlphas :66-72).  Returns false
// on linesearch failure.  On success the scaled LQ strategies become current.
bool ModifyLQStrategies(const Problem& pr, Instance& in, bool* has_converged) {
  const ilqg_solver_params& sp = pr.p;
  in.expected_decrease = ExpectedDecrease(pr, in);
  std::vector<real> Ps = in.lqPs, alphas = in.lqAlphas;
  for (auto& a : alphas) a *= sp.initial_alpha_scaling;
  const std::vector<real> last_xs = in.xs, last_us = in.us;
  real current_stepsize = sp.initial_alpha_scaling;
  Rollout(pr, last_xs.data(), last_us.data(), Ps.data(), alphas.data(), in.xs.data(), in.us.data());
  in.backtracks++;
  if (!sp.linesearch) {
    in.Ps = Ps;
    in.alphas = alphas;
    in.step = current_stepsize;
    return true;
  }
  for (int ii = 0; ii < sp.max_backtracking_steps; ii++) {
    const real merit = MeritFunction(pr, in);
    const real scaled_expected_decrease =
        sp.expected_decrease_fraction * current_stepsize * in.expected_decrease;
    if (in.last_merit - merit >= scaled_expected_decrease) {
      *has_converged = (merit <= in.last_merit) &&
                       std::abs(in.last_merit - merit) < sp.convergence_tolerance;
      in.last_merit = merit;
      in.Ps = Ps;
      in.alphas = alphas;
      in.step = current_stepsize;
      return true;
    }
    for (auto& a : alphas) a *= sp.geometric_alpha_scaling;
    current_stepsize *= sp.geometric_alpha_scaling;
    Rollout(pr, last_xs.data(), last_us.data(), Ps.data(), alphas.data(), in.xs.data(),
            in.us.data());
    in.backtracks++;
  }
  // Failure: the log's final iterate is the last accepted one (ilq_solver.cpp:146-155).
  in.xs = last_xs;
  in.us = last_us;
  return false;
}

}  // namespace

// ============================== C ABI =======================================
struct ilqg_solver {
  Problem pr;
  int batch;
  double op_t0 = 0;  // OperatingPoint::t0 of the Problems (uniform over the batch)
  std::vector<Instance> inst;
  int al_max_iterates = 0;
  float al_tolerance = 0;
  bool al_first = true;
};

namespace {

void InitInstance(const Problem& pr, Instance& in) {
  const int T = pr.T, n = pr.n, M = pr.M, N = pr.N;
  in.x0.assign(n, 0);
  in.lq_x0.assign(n, 0);
  in.prob_xs.assign((size_t)T * n, 0);
  in.prob_us.assign((size_t)T * M, 0);
  in.prob_Ps.assign((size_t)T * M * n, 0);
  in.prob_alphas.assign((size_t)T * M, 0);
  in.xs.assign((size_t)T * n, 0);
  in.us.assign((size_t)T * M, 0);
  in.Ps.assign((size_t)T * M * n, 0);
  in.alphas.assign((size_t)T * M, 0);
  in.lqPs.assign((size_t)T * M * n, 0);
  in.lqAlphas.assign((size_t)T * M, 0);
  in.A.assign((size_t)T * n * n, 0);
  in.B.assign((size_t)T * n * M, 0);
  in.Q.assign((size_t)T * N * n * n, 0);
  in.l.assign((size_t)T * N * n, 0);
  in.R.assign((size_t)T * pr.R_floats, 0);
  in.r.assign((size_t)T * pr.r_floats, 0);
  in.dxs.assign((size_t)T * n, 0);
  in.lambdas.assign((size_t)pr.num_constraints * T, 0);  // kDefaultLambda
  in.mu = kDefaultMu;
  in.last_merit = kInfinity;
  in.expected_decrease = kInfinity;
  in.step = 0;
  in.total_costs.assign(N, 0);
  in.te_quad.assign(N, 0);
  in.te_new.assign(N, 0);
  in.status = ILQG_STATUS_IDLE;
  in.iters = 0;
  in.backtracks = 0;
  in.max_constraint_error = kInfinity;
  in.al_state = 0;
  in.al_iterates = 0;
  in.al_success = 1;
}

template <typename Src>
void CopyOut(const std::vector<Src>& v, float* dst) {
  for (size_t a = 0; a < v.size(); a++) dst[a] = (float)v[a];
}

// One pass of the while loop src/ilq_solver.cpp:123-166 for one instance.
void IterateOnce(const Problem& pr, Instance& in) {
  if (in.status != ILQG_STATUS_RUNNING) return;
  if (in.iters >= pr.p.max_solver_iters) {
    in.status = ILQG_STATUS_MAX_ITERS;
    return;
  }
  in.iters++;
  // Only the linearization is recomputed here; the quadraticization is the one the last
  // MeritFunction call (or the Solve prologue) left behind (src/ilq_solver.cpp:116,136-143).
  LinearizeAll(pr, in);
  in.te_quad = in.te_new;
  std::vector<real> zero(pr.n, 0);
  LQSolve(pr, in, zero.data());
  bool has_converged = false;
  if (!ModifyLQStrategies(pr, in, &has_converged)) {
    in.status = ILQG_STATUS_LINESEARCH_FAILED;
    return;
  }
  TotalCosts(pr, in);
  if (has_converged && !pr.p.disable_convergence_exit)
    in.status = ILQG_STATUS_CONVERGED;
  else if (in.iters >= pr.p.max_solver_iters)
    in.status = ILQG_STATUS_MAX_ITERS;
}

}  // namespace

extern "C" {

int ilqg_create(const ilqg_problem_desc* desc, const ilqg_solver_params* params, int batch,
                int /*device*/, ilqg_handle* out) {
  if (!desc || !params || !out || batch < 1) return ILQG_ERR_INVALID_ARGUMENT;
  ilqg_solver* s = new (std::nothrow) ilqg_solver;
  if (!s) return ILQG_ERR_OUT_OF_MEMORY;
  const int rc = BuildProblem(desc, params, &s->pr);
  if (rc != ILQG_OK) {
    delete s;
    return rc;
  }
  s->batch = batch;
  s->op_t0 = desc->initial_time;
  s->inst.resize(batch);
  for (auto& in : s->inst) InitInstance(s->pr, in);
  *out = s;
  return ILQG_OK;
}

// ilqg_create_multi (include/ilqg.h): the device list means nothing on the CPU
int ilqg_create_multi(const ilqg_problem_desc* desc, const ilqg_solver_params* params, int batch,
                      const int* devices, int num_devices, ilqg_handle* out) {
  if (!devices || num_devices < 1 || batch < num_devices) return ILQG_ERR_INVALID_ARGUMENT;
  return ilqg_create(desc, params, batch, 0, out);
}

int ilqg_destroy(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  delete h;
  return ILQG_OK;
}

const char* ilqg_strerror(int code) {
  switch (code) {
    case ILQG_OK: return "ok";
    case ILQG_ERR_INVALID_ARGUMENT: return "invalid argument";
    case ILQG_ERR_UNSUPPORTED: return "unsupported descriptor";
    case ILQG_ERR_CUDA: return "CUDA error";
    case ILQG_ERR_NO_DEVICE: return "no CUDA device";
    case ILQG_ERR_OUT_OF_MEMORY: return "out of memory";
    case ILQG_ERR_BAD_HANDLE: return "bad handle";
    case ILQG_ERR_SIZE_MISMATCH: return "host buffer size mismatch";
  }
  return "unknown error";
}

size_t ilqg_abi_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(ilqg_problem_desc);
    case 1: return sizeof(ilqg_solver_params);
    case 2: return sizeof(ilqg_layout);
    case 3: return sizeof(ilqg_cost_desc);
    case 4: return sizeof(ilqg_subsystem_desc);
  }
  return 0;
}

int ilqg_get_layout(ilqg_handle h, ilqg_layout* out) {
  if (!h || !out) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  std::memset(out, 0, sizeof(*out));
  out->batch = h->batch;
  out->num_time_steps = pr.T;
  out->num_players = pr.N;
  out->xdim = pr.n;
  out->total_udim = pr.M;
  for (int i = 0; i < pr.N; i++) {
    out->udim[i] = pr.d.udim[i];
    out->u_offset[i] = pr.uoff[i];
  }
  out->num_pairs = pr.num_pairs;
  for (int p = 0; p < pr.num_pairs; p++) {
    out->pair_player[p] = pr.pair_i[p];
    out->pair_arg[p] = pr.pair_j[p];
    out->pair_R_offset[p] = pr.pair_Roff[p];
    out->pair_r_offset[p] = pr.pair_roff[p];
  }
  out->R_floats = pr.R_floats;
  out->r_floats = pr.r_floats;
  out->num_constraints = pr.num_constraints;
  out->record_floats = 0;
  out->compact_record_floats = 0;
  for (int kk = 0; kk < pr.T; kk++) out->lambda_index[kk] = pr.lambda_index[kk];
  return ILQG_OK;
}

int ilqg_upload_x0(ilqg_handle h, const float* x0, size_t bytes) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  const int n = h->pr.n;
  if (bytes != sizeof(float) * (size_t)h->batch * n) return ILQG_ERR_SIZE_MISMATCH;
  for (int b = 0; b < h->batch; b++)
    for (int a = 0; a < n; a++) h->inst[b].x0[a] = x0[(size_t)b * n + a];
  return ILQG_OK;
}

int ilqg_upload_warmstart(ilqg_handle h, const float* xs, const float* us, const float* Ps,
                          const float* alphas) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  const size_t nx = (size_t)pr.T * pr.n, nu = (size_t)pr.T * pr.M, np = (size_t)pr.T * pr.M * pr.n;
  for (int b = 0; b < h->batch; b++) {
    Instance& in = h->inst[b];
    for (size_t a = 0; a < nx; a++) in.prob_xs[a] = xs ? xs[b * nx + a] : 0.f;
    for (size_t a = 0; a < nu; a++) in.prob_us[a] = us ? us[b * nu + a] : 0.f;
    for (size_t a = 0; a < np; a++) in.prob_Ps[a] = Ps ? Ps[b * np + a] : 0.f;
    for (size_t a = 0; a < nu; a++) in.prob_alphas[a] = alphas ? alphas[b * nu + a] : 0.f;
    // the working copies follow so single-stage calls (linearize_quadraticize,
    // lq_backward) see the uploaded point without a solve_begin
    in.xs = in.prob_xs; in.us = in.prob_us; in.Ps = in.prob_Ps; in.alphas = in.prob_alphas;
  }
  return ILQG_OK;
}

int ilqg_upload(ilqg_handle h, int what, const void* src, size_t bytes) {
  if (!h || !src) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  const int B = h->batch;
  const float* f = (const float*)src;
  const int32_t* iv = (const int32_t*)src;
  switch (what) {
    case ILQG_LAMBDAS: {
      const size_t per = (size_t)pr.num_constraints * pr.T;
      if (bytes != sizeof(float) * per * B) return ILQG_ERR_SIZE_MISMATCH;
      for (int b = 0; b < B; b++)
        for (size_t a = 0; a < per; a++) h->inst[b].lambdas[a] = f[b * per + a];
      return ILQG_OK;
    }
    case ILQG_MU:
      if (bytes != sizeof(float) * B) return ILQG_ERR_SIZE_MISMATCH;
      for (int b = 0; b < B; b++) h->inst[b].mu = f[b];
      return ILQG_OK;
    case ILQG_MERIT:
      if (bytes != sizeof(float) * B) return ILQG_ERR_SIZE_MISMATCH;
      for (int b = 0; b < B; b++) h->inst[b].last_merit = f[b];
      return ILQG_OK;
    case ILQG_TIME_OF_EXTREME:
      if (bytes != sizeof(int32_t) * B * pr.N) return ILQG_ERR_SIZE_MISMATCH;
      for (int b = 0; b < B; b++)
        for (int i = 0; i < pr.N; i++)
          h->inst[b].te_quad[i] = h->inst[b].te_new[i] = iv[b * pr.N + i];
      return ILQG_OK;
    case ILQG_X0:
      return ilqg_upload_x0(h, f, bytes);
    case ILQG_LQ_X0:
      if (bytes != sizeof(float) * B * pr.n) return ILQG_ERR_SIZE_MISMATCH;
      for (int b = 0; b < B; b++) h->inst[b].lq_x0.assign(f + (size_t)b * pr.n, f + (size_t)(b + 1) * pr.n);
      return ILQG_OK;
  }
  return ILQG_ERR_INVALID_ARGUMENT;
}

int ilqg_upload_lq(ilqg_handle h, const float* A, const float* Bs, const float* Q, const float* l,
                   const float* R, const float* r) {
  if (!h || !A || !Bs || !Q || !l || !R || !r) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  const size_t T = pr.T, n = pr.n, M = pr.M, N = pr.N;
  for (int b = 0; b < h->batch; b++) {
    Instance& in = h->inst[b];
    for (size_t a = 0; a < T * n * n; a++) in.A[a] = A[b * T * n * n + a];
    for (size_t a = 0; a < T * n * M; a++) in.B[a] = Bs[b * T * n * M + a];
    for (size_t a = 0; a < T * N * n * n; a++) in.Q[a] = Q[b * T * N * n * n + a];
    for (size_t a = 0; a < T * N * n; a++) in.l[a] = l[b * T * N * n + a];
    for (size_t a = 0; a < T * pr.R_floats; a++) in.R[a] = R[b * T * pr.R_floats + a];
    for (size_t a = 0; a < T * pr.r_floats; a++) in.r[a] = r[b * T * pr.r_floats + a];
  }
  return ILQG_OK;
}

int ilqg_solve_begin(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  for (auto& in : h->inst) {
    if (in.al_state == 2) continue;  // this game's AL loop has ended
    // src/ilq_solver.cpp:86-107
    std::vector<real> last_xs = in.prob_xs, last_us = in.prob_us;
    for (int a = 0; a < pr.n; a++) last_xs[a] = in.x0[a];
    in.Ps = in.prob_Ps;
    in.alphas = in.prob_alphas;
    Rollout(pr, last_xs.data(), last_us.data(), in.Ps.data(), in.alphas.data(), in.xs.data(),
            in.us.data());
    TotalCosts(pr, in);
    in.te_quad = in.te_new;
    QuadraticizeAll(pr, in);  // src/ilq_solver.cpp:116
    in.status = pr.p.max_solver_iters > 0 ? ILQG_STATUS_RUNNING : ILQG_STATUS_MAX_ITERS;
    in.iters = 0;
  }
  return ILQG_OK;
}

int ilqg_linearize_quadraticize(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (auto& in : h->inst) {
    LinearizeAll(h->pr, in);
    QuadraticizeAll(h->pr, in);
  }
  return ILQG_OK;
}

int ilqg_lq_backward(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (auto& in : h->inst) {
    in.te_quad = in.te_new;
    LQSolve(h->pr, in, in.lq_x0.data());
    in.expected_decrease = ExpectedDecrease(h->pr, in);
  }
  return ILQG_OK;
}

int ilqg_linesearch(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (auto& in : h->inst) {
    if (in.status != ILQG_STATUS_RUNNING) continue;
    in.iters++;  // the stage sequence LQ -> backward -> linesearch is one loop pass
    bool conv = false;
    if (!ModifyLQStrategies(h->pr, in, &conv)) {
      in.status = ILQG_STATUS_LINESEARCH_FAILED;
      continue;
    }
    TotalCosts(h->pr, in);
    if (conv && !h->pr.p.disable_convergence_exit)
      in.status = ILQG_STATUS_CONVERGED;
    else if (in.iters >= h->pr.p.max_solver_iters)
      in.status = ILQG_STATUS_MAX_ITERS;
  }
  return ILQG_OK;
}

int ilqg_iterate(ilqg_handle h, int max_iters, int* iters_done) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  int most = 0;
  for (auto& in : h->inst) {
    for (int it = 0; it < max_iters; it++) IterateOnce(h->pr, in);
    most = std::max(most, in.iters);
  }
  if (iters_done) *iters_done = most;
  return ILQG_OK;
}

}  // extern "C"

namespace {

// src/augmented_lagrangian_solver.cpp:113-143 for one game: multiplier sweep + ScaleMu.
void AlUpdateOne(const Problem& pr, Instance& in) {
  const ilqg_problem_desc& d = pr.d;
  real max_err = -kInfinity;
  // Outer loop over players then time (:116-139); constraints of one player in
  // record order, state constraints before control constraints.
  for (int i = 0; i < pr.N; i++)
    for (int kk = 0; kk < pr.T; kk++) {
      // t = op.t0 + kTimeStep * float(kk): same TimeIndex as RelativeTime(kk).
      const int li = pr.lambda_index[kk];
      for (int pass = 0; pass < 2; pass++)
        for (int c = 0; c < d.num_costs; c++) {
          const ilqg_cost_desc& cd = d.costs[c];
          const int slot = pr.constraint_slot[c];
          if (slot < 0 || cd.player != i) continue;
          if ((pass == 0) != (cd.arg < 0)) continue;
          const real g = cd.arg < 0
                             ? EvaluateRecord(pr, cd, &in.xs[(size_t)kk * pr.n], pr.n)
                             : EvaluateRecord(pr, cd, &in.us[(size_t)kk * pr.M + pr.uoff[cd.arg]],
                                              d.udim[cd.arg]);
          max_err = std::max(max_err, g);
          // Constraint::IncrementLambda, constraint.h:98-102
          real& lam = in.lambdas[(size_t)slot * pr.T + li];
          const real new_lambda = lam + in.mu * g;
          lam = cd.is_equality ? new_lambda : std::max((real)0.0f, new_lambda);
        }
    }
  in.mu *= pr.p.geometric_mu_scaling;  // :143
  in.max_constraint_error = max_err;
}

// Problem::OverwriteSolution(log->FinalOperatingPoint(), log->FinalStrategies()),
// src/problem.cpp:188-194 as called at augmented_lagrangian_solver.cpp:151-154.
void OverwriteOne(Instance& in) {
  in.prob_xs = in.xs;
  in.prob_us = in.us;
  in.prob_Ps = in.Ps;
  in.prob_alphas = in.alphas;
}

// src/augmented_lagrangian_solver.cpp:165-178
void DownscaleOne(const Problem& pr, Instance& in) {
  for (auto& lam : in.lambdas) lam *= pr.p.geometric_lambda_downscaling;
  in.mu *= pr.p.geometric_mu_downscaling;
}

}  // namespace

extern "C" {

int ilqg_al_update(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (auto& in : h->inst) AlUpdateOne(h->pr, in);
  return ILQG_OK;
}

int ilqg_overwrite_solution(ilqg_handle h, int only_successful) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (auto& in : h->inst) {
    if (only_successful && in.status == ILQG_STATUS_LINESEARCH_FAILED) continue;
    OverwriteOne(in);
  }
  return ILQG_OK;
}

// Problem::SetUpNextRecedingHorizon (src/problem.cpp:64-186) for every game of the batch: the
// warm start (Problem::operating_point_ / strategies_) is re-based to start planner_runtime after
// t0 from the measured state x0.  Times are shared by the batch, states are per game.
int ilqg_setup_next_receding_horizon(ilqg_handle h, const float* x0_in, double t0, double planner_runtime,
                                     double* new_t0) {
  if (!h || !x0_in) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  // a constrained problem CHECK-fails in the reference once initial_time_ > 0: quadraticization
  // asks Constraint::TimeIndex(RelativeTime(kk)) (relative_time_tracker.h:63-72, ilq_solver.cpp:475)
  if (pr.num_constraints > 0) return ILQG_ERR_UNSUPPORTED;
  const int T = pr.T, n = pr.n, M = pr.M, N = pr.N;
  const double kTimeStep = pr.d.time_step, kTimeHorizon = kTimeStep * T;
  // ---- SyncToExistingProblem :64-125, the scalar part (CHECKs become argument errors) ----
  if (planner_runtime < 0.0 || planner_runtime + t0 > h->op_t0 + kTimeHorizon || t0 < h->op_t0)
    return ILQG_ERR_INVALID_ARGUMENT;
  constexpr float kRoundingError = 0.9;
  const double relative_t0 = t0 - h->op_t0;
  size_t current_timestep = static_cast<size_t>(relative_t0 / kTimeStep);
  double remaining_time_this_step = (current_timestep + 1) * kTimeStep - relative_t0;
  if (remaining_time_this_step < kRoundingError * kTimeStep) {
    current_timestep += 1;
    remaining_time_this_step = kTimeStep - remaining_time_this_step;
  }
  if (!(remaining_time_this_step < kTimeStep)) return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_LT :87
  // IntegrateToNextTimeStep's own bookkeeping (src/multi_player_integrable_system.cpp:113-143)
  const size_t itn_timestep = static_cast<size_t>((relative_t0 + kSmallNumber) / kTimeStep);
  const double itn_remaining = kTimeStep * (itn_timestep + 1) - relative_t0;
  if (!(itn_remaining < kTimeStep + kSmallNumber) || itn_timestep >= (size_t)T)
    return ILQG_ERR_INVALID_ARGUMENT;
  const float frac = itn_remaining / kTimeStep;
  double op_t0 = t0 + remaining_time_this_step;
  size_t num_steps_to_integrate = 0;
  const bool integrate_more = remaining_time_this_step <= planner_runtime;
  if (integrate_more) {
    num_steps_to_integrate =
        static_cast<size_t>(kSmallNumber + (planner_runtime - remaining_time_this_step) / kTimeStep);
    op_t0 += kTimeStep * num_steps_to_integrate;
  }
  const size_t last_integration_timestep = current_timestep + num_steps_to_integrate;
  if (last_integration_timestep > (size_t)T) return ILQG_ERR_INVALID_ARGUMENT;
  if (!(std::abs(t0 + planner_runtime - op_t0) <= kTimeStep)) return ILQG_ERR_INVALID_ARGUMENT;  // :123

  const ilqg_subsystem_desc& ego = pr.d.subsystems[0];
  // DistanceBetween is positional for every dynamics class but a Dubins ego: ConcatenatedDynamicalSystem
  // (first subsystem), TwoPlayerUnicycle4D (two_player_unicycle_4d.h:141-147) and Air3D (head(2), air_3d.h:151-157)
  const bool concatenated = true;
  for (int b = 0; b < h->batch; b++) {
    Instance& in = h->inst[b];
    real x[ILQG_MAX_XDIM], u[ILQG_MAX_UDIM], ref[ILQG_MAX_XDIM], nx[ILQG_MAX_XDIM];
    // Strategy::operator() (strategy.h:73-76): u = u_ref - P dx - alpha
    auto controls = [&](size_t kk, const real* state, const real* state_ref) {
      for (int c = 0; c < M; c++) {
        real acc = 0;
        for (int a = 0; a < n; a++) acc += in.prob_Ps[(kk * M + c) * n + a] * (state[a] - state_ref[a]);
        u[c] = (in.prob_us[kk * M + c] - acc) - in.prob_alphas[kk * M + c];
      }
    };
    for (int a = 0; a < n; a++) x[a] = x0_in[(size_t)b * n + a];
    // x0_ref: interpolated reference state (:130-135)
    for (int a = 0; a < n; a++)
      ref[a] = itn_timestep + 1 < (size_t)T
                   ? frac * in.prob_xs[itn_timestep * n + a] + (real)(1.0 - frac) * in.prob_xs[(itn_timestep + 1) * n + a]
                   : in.prob_xs[(size_t)(T - 1) * n + a];
    controls(itn_timestep, x, ref);
    Integrate(pr, x, u, nx, itn_remaining);
    std::memcpy(x, nx, sizeof(real) * n);
    if (integrate_more) {
      for (size_t kk = current_timestep + 1; kk < last_integration_timestep; kk++) {  // :96-111
        controls(kk, x, &in.prob_xs[kk * n]);
        Integrate(pr, x, u, nx);
        std::memcpy(x, nx, sizeof(real) * n);
      }
    }
    // nearest state of the existing plan (:101-110); ConcatenatedDynamicalSystem::DistanceBetween
    // only looks at the first subsystem's position (src/concatenated_dynamical_system.cpp:109-113)
    auto distance = [&](const real* a) {
      if (ego.kind == ILQG_DYN_DUBINS) {
        // SinglePlayerDubinsCar has no DistanceBetween of its own: the base class's squared 2-norm
        // of the whole subsystem state, heading included (single_player_dynamical_system.h:68-71)
        real acc = 0;
        for (int q = 0; q < 3; q++) acc += (x[ego.x_offset + q] - a[ego.x_offset + q]) * (x[ego.x_offset + q] - a[ego.x_offset + q]);
        return acc;
      }
      if (concatenated) {
        const real dx = x[ego.x_offset] - a[ego.x_offset], dy = x[ego.x_offset + 1] - a[ego.x_offset + 1];
        return dx * dx + dy * dy;
      }
      real acc = 0;
      for (int q = 0; q < n; q++) acc += (x[q] - a[q]) * (x[q] - a[q]);
      return acc;
    };
    size_t first = 0;
    for (size_t kk = 1; kk < (size_t)T; kk++)
      if (distance(&in.prob_xs[kk * n]) < distance(&in.prob_xs[first * n])) first = kk;
    // x0_ = Stitch(nearest, x) (:117; concatenated_dynamical_system.h:75-85)
    const int ego_dim = ego.kind == ILQG_DYN_CAR6D ? 6 : ego.kind == ILQG_DYN_CAR5D ? 5 : (ego.kind == ILQG_DYN_UNICYCLE4D || ego.kind == ILQG_DYN_POINT_MASS_2D) ? 4 : ego.kind == ILQG_DYN_DUBINS ? 3 : n;
    for (int a = 0; a < n; a++) in.x0[a] = a < ego_dim ? in.prob_xs[first * n + a] : x[a];
    // ---- SetUpNextRecedingHorizon :127-186: shift the plan, extend it with zero controls ----
    const size_t kept = (size_t)T - first;
    for (size_t kk = 0; kk < kept; kk++) {
      const size_t src = kk + first;
      if (src == kk) continue;
      std::memcpy(&in.prob_xs[kk * n], &in.prob_xs[src * n], sizeof(real) * n);
      std::memcpy(&in.prob_us[kk * M], &in.prob_us[src * M], sizeof(real) * M);
      std::memcpy(&in.prob_Ps[kk * M * n], &in.prob_Ps[src * M * n], sizeof(real) * M * n);
      std::memcpy(&in.prob_alphas[kk * M], &in.prob_alphas[src * M], sizeof(real) * M);
    }
    for (size_t kk = kept; kk < (size_t)T; kk++) {
      std::fill(&in.prob_Ps[kk * M * n], &in.prob_Ps[(kk + 1) * M * n], (real)0);
      std::fill(&in.prob_alphas[kk * M], &in.prob_alphas[(kk + 1) * M], (real)0);
      std::fill(&in.prob_us[kk * M], &in.prob_us[(kk + 1) * M], (real)0);
      Integrate(pr, &in.prob_xs[(kk - 1) * n], &in.prob_us[(kk - 1) * M], &in.prob_xs[kk * n]);
    }
    (void)N;
  }
  h->op_t0 = op_t0;
  UpdateCostGates(&h->pr, op_t0);  // RelativeTimeTracker::ResetInitialTime(op.t0), src/problem.cpp:120
  if (new_t0) *new_t0 = op_t0;
  return ILQG_OK;
}

// MultiPlayerIntegrableSystem::Integrate(t0, t, x0, operating_point, strategies),
// src/multi_player_integrable_system.cpp:54-83, for every game under its warm start.
int ilqg_integrate_plan(ilqg_handle h, const float* x0_in, double t0, double t, float* x_out) {
  if (!h || !x0_in || !x_out) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  if (pr.d.num_subsystems <= 0 || pr.d.subsystems[0].kind == ILQG_DYN_NONE) return ILQG_ERR_UNSUPPORTED;
  const int T = pr.T, n = pr.n, M = pr.M;
  const double kTimeStep = pr.d.time_step, plan_t0 = h->op_t0;
  if (!(t >= t0) || !(t0 >= plan_t0)) return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_GE :57-58
  const double relative_t0 = t0 - plan_t0;                               // :64-70
  const size_t current_timestep = static_cast<size_t>(relative_t0 / kTimeStep);
  const double relative_t = t - plan_t0;
  const size_t final_timestep = static_cast<size_t>(relative_t / kTimeStep);
  // IntegrateToNextTimeStep :113-143, taken only when t0 is past the plan's start (:75)
  const bool to_next = t0 > plan_t0;
  const size_t itn_timestep = static_cast<size_t>((relative_t0 + kSmallNumber) / kTimeStep);
  const double itn_remaining = kTimeStep * (itn_timestep + 1) - relative_t0;
  if (to_next && (!(itn_remaining < kTimeStep + kSmallNumber) || itn_timestep >= (size_t)T))
    return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_LT :126-127
  const float frac = itn_remaining / kTimeStep;
  // IntegrateFromPriorTimeStep :145-171 (its time step is final_timestep again)
  const double remaining_until_t = relative_t - kTimeStep * final_timestep;
  if (final_timestep >= (size_t)T || !(remaining_until_t < kTimeStep)) return ILQG_ERR_INVALID_ARGUMENT;  // :154-155
  const int itn_substeps = Rk4Substeps(t0, itn_remaining);
  const int step_substeps = Rk4Substeps(0.0, kTimeStep);
  const int prior_substeps = Rk4Substeps(plan_t0 + kTimeStep * final_timestep, remaining_until_t);

  for (int b = 0; b < h->batch; b++) {
    const Instance& in = h->inst[b];
    real x[ILQG_MAX_XDIM], u[ILQG_MAX_UDIM], ref[ILQG_MAX_XDIM], nx[ILQG_MAX_XDIM];
    auto controls = [&](size_t kk, const real* state_ref) {  // Strategy::operator(), strategy.h:73-76
      for (int c = 0; c < M; c++) {
        real acc = 0;
        for (int a = 0; a < n; a++) acc += in.prob_Ps[(kk * M + c) * n + a] * (x[a] - state_ref[a]);
        u[c] = (in.prob_us[kk * M + c] - acc) - in.prob_alphas[kk * M + c];
      }
    };
    for (int a = 0; a < n; a++) x[a] = x0_in[(size_t)b * n + a];
    if (to_next) {
      for (int a = 0; a < n; a++)
        ref[a] = itn_timestep + 1 < (size_t)T
                     ? frac * in.prob_xs[itn_timestep * n + a] + (real)(1.0 - frac) * in.prob_xs[(itn_timestep + 1) * n + a]
                     : in.prob_xs[(size_t)(T - 1) * n + a];
      controls(itn_timestep, ref);
      Integrate(pr, x, u, nx, itn_remaining, itn_substeps);
      std::memcpy(x, nx, sizeof(real) * n);
    }
    for (size_t kk = current_timestep + 1; kk < final_timestep; kk++) {  // :85-111
      controls(kk, &in.prob_xs[kk * n]);
      Integrate(pr, x, u, nx, kTimeStep, step_substeps);
      std::memcpy(x, nx, sizeof(real) * n);
    }
    controls(final_timestep, &in.prob_xs[final_timestep * n]);
    Integrate(pr, x, u, nx, remaining_until_t, prior_substeps);
    for (int a = 0; a < n; a++) x_out[(size_t)b * n + a] = (float)nx[a];
  }
  return ILQG_OK;
}

int ilqg_al_post_solve(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (auto& in : h->inst) {
    if (in.status != ILQG_STATUS_LINESEARCH_FAILED) continue;
    DownscaleOne(h->pr, in);
  }
  return ILQG_OK;
}

// AugmentedLagrangianSolver::Solve, src/augmented_lagrangian_solver.cpp:72-210, as a per-game
// state machine around the inner solves the caller runs:
//   al_begin; do { solve_begin; iterate until nothing runs; al_advance(&active); } while (active);
int ilqg_al_begin(ilqg_handle h, int max_iterates, float constraint_error_tolerance) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  if (max_iterates < 0) return ILQG_ERR_INVALID_ARGUMENT;
  h->al_max_iterates = max_iterates;
  h->al_tolerance = constraint_error_tolerance;
  h->al_first = true;
  for (auto& in : h->inst) {
    in.al_state = 1;
    in.al_iterates = 0;
    in.al_success = 1;                     // :74
    in.max_constraint_error = kInfinity;   // :108
  }
  return ILQG_OK;
}

int ilqg_al_advance(ilqg_handle h, int* active) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  int n_active = 0;
  for (auto& in : h->inst) {
    if (in.al_state != 1) continue;
    if (in.status == ILQG_STATUS_RUNNING) return ILQG_ERR_INVALID_ARGUMENT;  // inner solve not finished
    const bool failed = in.status == ILQG_STATUS_LINESEARCH_FAILED;
    // the inner log holds the initial iterate plus one per completed iteration; the iteration
    // whose linesearch failed is not logged (src/ilq_solver.cpp:111,146-153,164); AddLog :94,180
    in.al_iterates += 1 + in.iters - (failed ? 1 : 0);
    if (!h->al_first && failed) DownscaleOne(pr, in);  // :165-178
    in.al_success = in.al_success && !failed;          // :100, :179
    const bool constrained = pr.num_constraints > 0;   // :103
    const bool again = constrained && in.al_iterates < h->al_max_iterates &&
                       in.max_constraint_error > (real)h->al_tolerance;  // :109-111 (time: SURVEY Q2)
    if (!again) {
      if (constrained && in.max_constraint_error > (real)h->al_tolerance) in.al_success = 0;  // :187-190
      in.al_state = 2;
      continue;
    }
    AlUpdateOne(pr, in);             // :113-143
    if (!failed) OverwriteOne(in);   // :151-154
    n_active++;
  }
  h->al_first = false;
  if (active) *active = n_active;
  return ILQG_OK;
}

int ilqg_download(ilqg_handle h, int what, void* dst, size_t bytes) {
  if (!h || !dst) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  const int B = h->batch;
  float* f = (float*)dst;
  int32_t* iv = (int32_t*)dst;
  auto vec = [&](std::vector<real> Instance::*m) -> int {
    const size_t per = (h->inst[0].*m).size();
    if (bytes != sizeof(float) * per * B) return ILQG_ERR_SIZE_MISMATCH;
    for (int b = 0; b < B; b++) CopyOut(h->inst[b].*m, f + b * per);
    return ILQG_OK;
  };
  auto scal = [&](real Instance::*m) -> int {
    if (bytes != sizeof(float) * B) return ILQG_ERR_SIZE_MISMATCH;
    for (int b = 0; b < B; b++) f[b] = (float)(h->inst[b].*m);
    return ILQG_OK;
  };
  auto iscal = [&](int Instance::*m) -> int {
    if (bytes != sizeof(int32_t) * B) return ILQG_ERR_SIZE_MISMATCH;
    for (int b = 0; b < B; b++) iv[b] = h->inst[b].*m;
    return ILQG_OK;
  };
  switch (what) {
    case ILQG_XS: return vec(&Instance::xs);
    case ILQG_WARM_XS: return vec(&Instance::prob_xs);
    case ILQG_WARM_US: return vec(&Instance::prob_us);
    case ILQG_WARM_PS: return vec(&Instance::prob_Ps);
    case ILQG_WARM_ALPHAS: return vec(&Instance::prob_alphas);
    case ILQG_US: return vec(&Instance::us);
    case ILQG_PS: return vec(&Instance::Ps);
    case ILQG_ALPHAS: return vec(&Instance::alphas);
    case ILQG_LQ_PS: return vec(&Instance::lqPs);
    case ILQG_LQ_ALPHAS: return vec(&Instance::lqAlphas);
    case ILQG_LIN_A: return vec(&Instance::A);
    case ILQG_LIN_B: return vec(&Instance::B);
    case ILQG_QUAD_Q: return vec(&Instance::Q);
    case ILQG_QUAD_L: return vec(&Instance::l);
    case ILQG_QUAD_R: return vec(&Instance::R);
    case ILQG_QUAD_RGRAD: return vec(&Instance::r);
    case ILQG_DELTA_XS: return vec(&Instance::dxs);
    case ILQG_LAMBDAS: return vec(&Instance::lambdas);
    case ILQG_TOTAL_COSTS: return vec(&Instance::total_costs);
    case ILQG_X0: return vec(&Instance::x0);
    case ILQG_LQ_X0: return vec(&Instance::lq_x0);
    case ILQG_MU: return scal(&Instance::mu);
    case ILQG_MERIT: return scal(&Instance::last_merit);
    case ILQG_EXPECTED_DECREASE: return scal(&Instance::expected_decrease);
    case ILQG_STEP: return scal(&Instance::step);
    case ILQG_MAX_CONSTRAINT_ERROR: return scal(&Instance::max_constraint_error);
    case ILQG_STATUS: return iscal(&Instance::status);
    case ILQG_ITERS: return iscal(&Instance::iters);
    case ILQG_BACKTRACKS: return iscal(&Instance::backtracks);
    case ILQG_AL_SUCCESS: return iscal(&Instance::al_success);
    case ILQG_AL_ITERATES: return iscal(&Instance::al_iterates);
    case ILQG_AL_STATE: return iscal(&Instance::al_state);
    case ILQG_TIME_OF_EXTREME:
      if (bytes != sizeof(int32_t) * B * pr.N) return ILQG_ERR_SIZE_MISMATCH;
      for (int b = 0; b < B; b++)
        for (int i = 0; i < pr.N; i++) iv[b * pr.N + i] = h->inst[b].te_new[i];
      return ILQG_OK;
  }
  return ILQG_ERR_INVALID_ARGUMENT;
}

int ilqg_synchronize(ilqg_handle h) { return h ? ILQG_OK : ILQG_ERR_BAD_HANDLE; }

int ilqg_reset(ilqg_handle h, int mask) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (auto& in : h->inst) {
    in.al_state = 0;  // any reset ends an augmented-Lagrangian solve
    if (mask & ILQG_RESET_SOLVER) {
      in.last_merit = kInfinity;
      in.expected_decrease = kInfinity;
    }
    if (mask & (ILQG_RESET_MULTIPLIERS | ILQG_RESET_LAMBDAS)) std::fill(in.lambdas.begin(), in.lambdas.end(), (real)0);
    if (mask & (ILQG_RESET_MULTIPLIERS | ILQG_RESET_MU)) in.mu = kDefaultMu;
    if (mask & ILQG_RESET_SOLUTION) {
      for (auto* v : {&in.prob_xs, &in.prob_us, &in.prob_Ps, &in.prob_alphas, &in.xs, &in.us, &in.Ps,
                      &in.alphas})
        std::fill(v->begin(), v->end(), (real)0);
    }
  }
  if (mask & ILQG_RESET_SOLUTION) {
    h->op_t0 = h->pr.d.initial_time;  // a fresh OperatingPoint's t0
    UpdateCostGates(&h->pr, h->op_t0);
  }
  return ILQG_OK;
}

int ilqg_count_running(ilqg_handle h, int* running) {
  if (!h || !running) return ILQG_ERR_BAD_HANDLE;
  int c = 0;
  for (auto& in : h->inst) c += in.status == ILQG_STATUS_RUNNING;
  *running = c;
  return ILQG_OK;
}

int ilqg_set_stream(ilqg_handle h, void*) { return h ? ILQG_OK : ILQG_ERR_BAD_HANDLE; }
int ilqg_profile(ilqg_handle h, int) { return h ? ILQG_OK : ILQG_ERR_BAD_HANDLE; }
int ilqg_profile_read(ilqg_handle h, int, double* total_ms, long long* launches) {
  if (!h || !total_ms || !launches) return ILQG_ERR_BAD_HANDLE;
  *total_ms = 0;
  *launches = 0;
  return ILQG_OK;
}

int ilqg_kernel_launches(ilqg_handle h, long long* out) {
  if (!h || !out) return ILQG_ERR_BAD_HANDLE;
  *out = 0;
  return ILQG_OK;
}

// ----- extra oracle-only probes used by tests/test_oracle_pins.py ------------
// Polyline2::ClosestPoint on polyline `p` of the handle's descriptor.
int ilqg_oracle_polyline_closest_point(ilqg_handle h, int p, float qx, float qy, float* closest,
                                       int* is_vertex, int* segment, float* signed_sq,
                                       int* is_endpoint) {
  if (!h || p < 0 || p >= (int)h->pr.polylines.size()) return ILQG_ERR_INVALID_ARGUMENT;
  bool v, e;
  int seg;
  real ssd;
  const Point2 c = PolylineClosestPoint(h->pr.polylines[p], {(real)qx, (real)qy}, &v, &seg, &ssd, &e);
  closest[0] = (float)c.x;
  closest[1] = (float)c.y;
  *is_vertex = v;
  *segment = seg;
  *signed_sq = (float)ssd;
  *is_endpoint = e;
  return ILQG_OK;
}

// LineSegment2::ClosestPoint for the segment (ax,ay)-(bx,by).
int ilqg_oracle_segment_closest_point(float ax, float ay, float bx, float by, float qx, float qy,
                                      float* closest, int* is_endpoint, float* signed_sq) {
  const Segment s = MakeSegment({(real)ax, (real)ay}, {(real)bx, (real)by});
  if (!(s.length > kSmallNumber)) return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_GT, line_segment2.h:61
  bool e;
  real ssd;
  const Point2 c = SegmentClosestPoint(s, {(real)qx, (real)qy}, &e, &ssd);
  closest[0] = (float)c.x;
  closest[1] = (float)c.y;
  *is_endpoint = e;
  *signed_sq = (float)ssd;
  return ILQG_OK;
}

// Evaluate record c (cost value, or g for constraints; with_al != 0 gives
// Constraint::EvaluateAugmentedLagrangian, constraint.h:80-84) at one input.
int ilqg_oracle_evaluate_record(ilqg_handle h, int c, const float* input, int dim, float lambda,
                                float mu, int with_al, float* value) {
  if (!h || c < 0 || c >= h->pr.d.num_costs) return ILQG_ERR_INVALID_ARGUMENT;
  std::vector<real> in(input, input + dim);
  const ilqg_cost_desc& cd = h->pr.d.costs[c];
  const real g = EvaluateRecord(h->pr, cd, in.data(), dim);
  if (with_al) {
    const real m = ConstraintMu(cd, mu, lambda, g);
    *value = (float)(lambda * g + 0.5 * m * g * g);
  } else {
    *value = (float)g;
  }
  return ILQG_OK;
}

// Quadraticize record c alone into zeroed hess/grad.
int ilqg_oracle_quadraticize_record(ilqg_handle h, int c, const float* input, int dim,
                                    float lambda, float mu, float* hess, float* grad) {
  if (!h || c < 0 || c >= h->pr.d.num_costs) return ILQG_ERR_INVALID_ARGUMENT;
  std::vector<real> in(input, input + dim), H((size_t)dim * dim, 0), g(dim, 0);
  QuadraticizeRecord(h->pr, h->pr.d.costs[c], in.data(), dim, lambda, mu, H.data(), g.data());
  for (int a = 0; a < dim * dim; a++) hess[a] = (float)H[a];
  for (int a = 0; a < dim; a++) grad[a] = (float)g[a];
  return ILQG_OK;
}

// xdot = f(x, u) and one Integrate() step, for the linearization FD checks.
int ilqg_oracle_dynamics(ilqg_handle h, const float* x, const float* u, float* xdot, float* xnext) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  const Problem& pr = h->pr;
  std::vector<real> xr(x, x + pr.n), ur(u, u + pr.M), xd(pr.n, 0), xn(pr.n, 0);
  EvaluateDynamics(pr, xr.data(), ur.data(), xd.data());
  Integrate(pr, xr.data(), ur.data(), xn.data());
  for (int a = 0; a < pr.n; a++) {
    xdot[a] = (float)xd[a];
    xnext[a] = (float)xn[a];
  }
  return ILQG_OK;
}

}  // extern "C"
