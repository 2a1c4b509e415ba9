/*
 * ilqg.h -- C ABI of the B200-native batched iLQ-games inner solver.
 *
 * This is the drop-in boundary for ONE hot path of HJReachability/ilqgames:
 *   forward rollout -> per-timestep linearization -> per-timestep/per-player
 *   quadraticization -> coupled backward Riccati Nash recursion (+ the Armijo
 *   linesearch glue that decides how often the first two run).
 *
 * The reference has no FFI of its own: its boundary is
 * the C++ virtual-class API.  Every entry point below therefore cites the
 * reference member function whose body it replaces; the re-authored host
 * classes (ILQSolver, LQFeedbackSolver, AugmentedLagrangianSolver) call these
 * and nothing else.  All arguments are PODs / plain pointers + sizes; no torch,
 * Eigen or C++ types cross this boundary; nothing throws or aborts across it.
 *
 * Two shared libraries implement this same header:
 *   ilqgames_b200/lib/libilqg_b200.so   the product: sm_100a CUDA kernels
 *   oracle/_build/libilqg_oracle.so     TEST INFRASTRUCTURE ONLY: the CPU
 *                                       restatement of the reference algorithm
 * (same symbols, loaded RTLD_LOCAL side by side by the parity tests).
 *
 * Host array layouts (all row-major, batch outermost, fp32 unless noted):
 *   x0      [B][n]
 *   xs      [B][T][n]             us     [B][T][M]   (M = sum of m_i, players
 *                                                     concatenated in order)
 *   Ps      [B][T][M][n]          alphas [B][T][M]   (rows of player i start at
 *                                                     u-offset of player i)
 *   A       [B][T][n][n]          Bs     [B][T][n][M]
 *   Q       [B][T][N][n][n]       l      [B][T][N][n]
 *   R       [B][T][rdim]          r      [B][T][udim_pairs]
 *           (control-cost pairs (i,j) sorted by (i,j); pair p holds an
 *            m_j x m_j row-major block at layout.pair_R_offset[p] and an m_j
 *            vector at layout.pair_r_offset[p])
 *   lambdas [B][num_constraints][T]      mu [B]
 */
#ifndef ILQG_H
#define ILQG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------- limits --------------------------------- */
#define ILQG_MAX_PLAYERS 4
#define ILQG_MAX_SUBSYSTEMS 4
#define ILQG_MAX_COSTS 64
#define ILQG_MAX_POLYLINES 8
#define ILQG_MAX_POLYLINE_POINTS 128
#define ILQG_MAX_PAIRS 16 /* ILQG_MAX_PLAYERS^2 */
#define ILQG_MAX_XDIM 24
#define ILQG_MAX_UDIM 8   /* total control dimension M */
#define ILQG_MAX_TIME_STEPS 512

/* ------------------------------ error codes ----------------------------- */
enum {
  ILQG_OK = 0,
  ILQG_ERR_INVALID_ARGUMENT = -1,
  ILQG_ERR_UNSUPPORTED = -2,      /* descriptor outside the compiled envelope */
  ILQG_ERR_CUDA = -3,             /* a CUDA runtime call failed               */
  ILQG_ERR_NO_DEVICE = -4,        /* no sm_100 device / driver                */
  ILQG_ERR_OUT_OF_MEMORY = -5,
  ILQG_ERR_BAD_HANDLE = -6,
  ILQG_ERR_SIZE_MISMATCH = -7     /* host buffer byte count is wrong          */
};

/* ----------------------------- descriptors ------------------------------ */

/* Subsystem kinds.  Reference: include/ilqgames/dynamics/
 *   single_player_car_6d.h:102-138, single_player_unicycle_4d.h:90-116,
 *   air_3d.h:114-149. */
enum {
  ILQG_DYN_NONE = 0,       /* LQ-only handle: lin/quad come from ilqg_upload_lq */
  ILQG_DYN_CAR6D = 1,      /* params[0] = inter-axle distance                  */
  ILQG_DYN_UNICYCLE4D = 2,
  ILQG_DYN_AIR3D = 3,      /* params[0] = evader speed, params[1] = pursuer    */
  ILQG_DYN_CAR5D = 4,      /* single_player_car_5d.h:102-147; params[0] = inter-axle distance */
  ILQG_DYN_DUBINS = 5,     /* single_player_dubins_car.h:56-118: (x, y, theta), control = turn rate,
                            * params[0] = constant speed */
  ILQG_DYN_TWO_PLAYER_UNICYCLE4D = 6, /* two_player_unicycle_4d.h:60-137: one coupled subsystem
                            * (x, y, theta, v); player first_player steers (omega, a), the next
                            * one pushes the position (dx, dy) */
  ILQG_DYN_POINT_MASS_2D = 7 /* single_player_point_mass_2d.h:56-118: (x, y, vx, vy), controls
                            * (ax, ay) */
};

typedef struct {
  int32_t kind;
  int32_t x_offset;      /* first state index of this subsystem             */
  int32_t first_player;  /* index of the player owning its first control    */
  int32_t reserved;
  float params[4];
} ilqg_subsystem_desc;

/* Cost / constraint record kinds.  Reference: src/quadratic_cost.cpp:51-94,
 * src/quadratic_polyline2_cost.cpp:52-126, src/proximity_cost.cpp:52-122,
 * src/semiquadratic_cost.cpp:51-85, src/semiquadratic_polyline2_cost.cpp:52-142,
 * src/polyline2_signed_distance_cost.cpp:52-121,
 * src/proximity_constraint.cpp:56-116,
 * include/ilqgames/constraint/single_dimension_constraint.h:68-96. */
enum {
  ILQG_COST_QUADRATIC = 1,                /* dim[0] (or -1 = all dims), weight, value=nominal */
  ILQG_COST_QUADRATIC_POLYLINE2 = 2,      /* dim[0..1]=x,y idx, polyline, weight              */
  ILQG_COST_PROXIMITY = 3,                /* dim[0..3]=x1,y1,x2,y2, weight, value=threshold   */
  ILQG_COST_SEMIQUADRATIC = 4,            /* dim[0], weight, value=threshold, flag=oriented_right */
  ILQG_COST_SEMIQUADRATIC_POLYLINE2 = 5,  /* dim[0..1], polyline, weight, value=threshold, flag=oriented_right */
  ILQG_COST_POLYLINE2_SIGNED_DISTANCE = 6,/* dim[0..1], polyline, value=nominal, flag=oriented_same_as_polyline */
  ILQG_CONSTRAINT_PROXIMITY = 7,          /* dim[0..3], value=threshold, flag=keep_within      */
  ILQG_CONSTRAINT_SINGLE_DIMENSION = 8,   /* dim[0], value=threshold, flag=keep_below          */
  ILQG_COST_SIGNED_DISTANCE = 9,          /* src/signed_distance_cost.cpp:50-112: dim[0..3]=x1,y1,x2,y2,
                                           * value=nominal, flag=less_is_positive (weight unused) */
  ILQG_COST_QUADRATIC_DIFFERENCE = 10,    /* src/quadratic_difference_cost.cpp:50-91: 0.5 w sum (in[a_k] - in[b_k])^2
                                           * over flag = 1 or 2 pairs, dim[0..1] = a, dim[2..3] = b */
  ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE = 11 /* src/polyline2_signed_distance_constraint.cpp:58-145:
                                           * g = +/-(signed distance to the polyline - value) <= 0, dim[0..1] = x, y,
                                           * polyline, value = threshold, flag = keep_left */
};

typedef struct {
  int32_t kind;
  int32_t player;      /* owner i: the PlayerCost this record was added to      */
  int32_t arg;         /* -1: function of the state; j >= 0: of player j's control */
  int32_t is_equality; /* constraints only (Constraint::is_equality_)          */
  int32_t dim[4];
  int32_t flag;
  int32_t polyline;    /* index into the polyline table, or -1                  */
  float weight;
  float value;
  /* FinalTimeCost(cost, threshold_time) (include/ilqgames/cost/final_time_cost.h:55-88): the
   * record counts only at time steps whose RelativeTime(kk) = kk * time_step is
   * >= RelativeTimeTracker::initial_time_ + active_from; 0 = always.  Costs only (not
   * constraints). */
  double active_from;
  /* ExtremeValueCost(costs, is_min) (include/ilqgames/cost/extreme_value_cost.h:54-80,
   * src/extreme_value_cost.cpp:50-84): consecutive records of one player with the same group > 0
   * are its sub-costs, in order; whenever the cost is evaluated or quadraticized only the member
   * with the largest (group_is_min: smallest) value counts, the first one on ties.  0 = a plain
   * record. */
  int32_t group;
  int32_t group_is_min;
} ilqg_cost_desc;

/* PlayerCost::CostStructure, include/ilqgames/cost/player_cost.h:105-111 */
enum { ILQG_COST_SUM = 0, ILQG_COST_MAX = 1, ILQG_COST_MIN = 2 };

typedef struct {
  /* horizon: time::kNumTimeSteps / kTimeStep made run-time parameters
   * (include/ilqgames/utils/types.h:133-143). */
  int32_t num_time_steps;
  int32_t num_players;
  int32_t xdim;
  int32_t udim[ILQG_MAX_PLAYERS];
  double time_step;    /* Time is double on Linux (types.h:85)                   */
  double initial_time; /* RelativeTimeTracker::initial_time_                    */

  int32_t num_subsystems;
  ilqg_subsystem_desc subsystems[ILQG_MAX_SUBSYSTEMS];

  /* per player: PlayerCost(name, state_regularization, control_regularization)
   * and its SUM/MAX/MIN structure (player_cost.h:63-70,105-111). */
  float state_regularization[ILQG_MAX_PLAYERS];
  float control_regularization[ILQG_MAX_PLAYERS];
  int32_t cost_structure[ILQG_MAX_PLAYERS];

  /* Cost records in the reference's accumulation order per player
   * (src/player_cost.cpp:194-215): state costs, control costs, state
   * constraints, control constraints. */
  int32_t num_costs;
  ilqg_cost_desc costs[ILQG_MAX_COSTS];

  /* Polyline table: polyline p owns points[polyline_start[p] .. polyline_start[p+1]). */
  int32_t num_polylines;
  int32_t polyline_start[ILQG_MAX_POLYLINES + 1];
  float polyline_points[ILQG_MAX_POLYLINE_POINTS][2];
} ilqg_problem_desc;

/* SolverParams, include/ilqgames/solver/solver_params.h:50-84 (the two dead
 * regularization fields are omitted). */
typedef struct {
  float convergence_tolerance;
  int32_t max_solver_iters;
  int32_t linesearch;
  float initial_alpha_scaling;
  float geometric_alpha_scaling;
  int32_t max_backtracking_steps;
  float expected_decrease_fraction;
  int32_t open_loop; /* nonzero: LQOpenLoopSolver (src/lq_open_loop_solver.cpp) instead of
                      * LQFeedbackSolver, as in ILQSolver's constructor (ilq_solver.h:76-81) */
  /* augmented Lagrangian outer loop */
  int32_t unconstrained_solver_max_iters;
  float geometric_mu_scaling;
  float geometric_mu_downscaling;
  float geometric_lambda_downscaling;
  float constraint_error_tolerance;
  /* additive (not in the reference): LQFeedbackSolver ctor's
   * adaptive_regularization (lq_feedback_solver.h:73-75) and a benchmark switch
   * that disables the HasConverged() exit so every instance runs max iters. */
  int32_t adaptive_regularization;
  int32_t disable_convergence_exit;
} ilqg_solver_params;

/* Shape / offset report so callers can size host buffers. */
typedef struct {
  int32_t batch, num_time_steps, num_players, xdim, total_udim;
  int32_t udim[ILQG_MAX_PLAYERS];
  int32_t u_offset[ILQG_MAX_PLAYERS];
  int32_t num_pairs;                    /* control-cost pairs (i,j)          */
  int32_t pair_player[ILQG_MAX_PAIRS];  /* i                                 */
  int32_t pair_arg[ILQG_MAX_PAIRS];     /* j                                 */
  int32_t pair_R_offset[ILQG_MAX_PAIRS];
  int32_t pair_r_offset[ILQG_MAX_PAIRS];
  int32_t R_floats;                     /* per (b,k): sum m_j^2 over pairs   */
  int32_t r_floats;                     /* per (b,k): sum m_j  over pairs    */
  int32_t num_constraints;
  int32_t record_floats;                /* device LQ record size per (b,k)   */
  int32_t lambda_index[ILQG_MAX_TIME_STEPS]; /* kk -> Constraint::TimeIndex */
  int32_t compact_record_floats;        /* floats per (b,k) of the compact LQ records the iLQ hot path
                                         * streams instead of dense ones (0: dense records only)   */
} ilqg_layout;

/* Per-instance status word (replaces `has_converged` / `*success`,
 * src/ilq_solver.cpp:95-96,146-155,168-171). */
enum {
  ILQG_STATUS_IDLE = 0,
  ILQG_STATUS_RUNNING = 1,
  ILQG_STATUS_CONVERGED = 2,         /* HasConverged() fired                    */
  ILQG_STATUS_MAX_ITERS = 3,         /* hit max_solver_iters (success = true)   */
  ILQG_STATUS_LINESEARCH_FAILED = 4, /* ModifyLQStrategies returned false       */
  ILQG_STATUS_NONFINITE = 5
};

/* what to download / upload */
enum {
  ILQG_XS = 1, ILQG_US = 2, ILQG_PS = 3, ILQG_ALPHAS = 4,
  ILQG_LIN_A = 5, ILQG_LIN_B = 6,
  ILQG_QUAD_Q = 7, ILQG_QUAD_L = 8, ILQG_QUAD_R = 9, ILQG_QUAD_RGRAD = 10,
  ILQG_DELTA_XS = 11,
  ILQG_STATUS = 12,            /* int32 [B]                                   */
  ILQG_ITERS = 13,             /* int32 [B] completed while-loop iterations   */
  ILQG_MERIT = 14,             /* float [B] last_merit_function_value_        */
  ILQG_TOTAL_COSTS = 15,       /* float [B][N]                                */
  ILQG_LAMBDAS = 16, ILQG_MU = 17,
  ILQG_EXPECTED_DECREASE = 18, /* float [B]                                   */
  ILQG_STEP = 19,              /* float [B] accepted step size                */
  ILQG_BACKTRACKS = 20,        /* int32 [B] cumulative rollouts in linesearch */
  ILQG_TIME_OF_EXTREME = 21,   /* int32 [B][N]                                */
  ILQG_X0 = 22,
  ILQG_LQ_PS = 23, ILQG_LQ_ALPHAS = 24, /* raw LQ solution (before linesearch scaling) */
  ILQG_MAX_CONSTRAINT_ERROR = 25, /* float [B], from ilqg_al_update / ilqg_al_advance */
  ILQG_AL_SUCCESS = 26,        /* int32 [B] AugmentedLagrangianSolver::Solve's *success */
  ILQG_AL_ITERATES = 27,       /* int32 [B] log->NumIterates() of the AL solve   */
  ILQG_AL_STATE = 28,          /* int32 [B] 0 = none, 1 = active, 2 = finished   */
  ILQG_WARM_XS = 30, ILQG_WARM_US = 31, ILQG_WARM_PS = 32, ILQG_WARM_ALPHAS = 33,
                               /* download: the Problem's warm start (Problem::CurrentOperatingPoint /
                                * CurrentStrategies, solver/problem.h:171-172), shapes of XS/US/PS/ALPHAS */
  ILQG_LQ_X0 = 29              /* float [B][n] the `x0` argument of LQSolver::Solve used by
                                * ilqg_lq_backward (upload; zeros until set).  ilqg_iterate
                                * always solves with x0 - xs[0] = 0 (ilq_solver.cpp:140-143) */
};

typedef struct ilqg_solver* ilqg_handle;

/* ------------------------------ entry points ---------------------------- */

/* Build the per-device slab for `batch` independent games described by `desc`.
 * Replaces ILQSolver::ILQSolver (include/ilqgames/solver/ilq_solver.h:69-94)
 * + LQFeedbackSolver::LQFeedbackSolver (lq_feedback_solver.h:73-115) +
 * Problem::Initialize (solver/problem.h:66-73): zero operating point, zero
 * strategies, lambda = 0, mu = 10, last merit = +inf.  `device` is the CUDA
 * ordinal (ignored by the oracle). */
int ilqg_create(const ilqg_problem_desc* desc, const ilqg_solver_params* params,
                int batch, int device, ilqg_handle* out);
/* Additive (the reference is single-threaded, one game per solver): the same handle with the batch
 * sharded over `num_devices` GPUs of one node -- device devices[k] owns a contiguous slice of the
 * games (sizes differ by at most one), every entry point fans out to the devices and joins them,
 * uploads / downloads address the global batch.  The games are independent (each has its own
 * ILQSolver state: src/ilq_solver.cpp:76-172), so there is no exchange between devices; the gather
 * of converged trajectories is the downloads.  The oracle ignores the device list. */
int ilqg_create_multi(const ilqg_problem_desc* desc, const ilqg_solver_params* params, int batch,
                      const int* devices, int num_devices, ilqg_handle* out);
int ilqg_destroy(ilqg_handle h);
const char* ilqg_strerror(int code);
/* sizeof() of the ABI structs as compiled (0: ilqg_problem_desc, 1:
 * ilqg_solver_params, 2: ilqg_layout, 3: ilqg_cost_desc, 4: ilqg_subsystem_desc)
 * so foreign-language bindings can verify their mirror of this header. */
size_t ilqg_abi_struct_size(int which);
int ilqg_get_layout(ilqg_handle h, ilqg_layout* out);

/* Problem::ResetInitialState (problem.h:82-85), one x0 per instance. */
int ilqg_upload_x0(ilqg_handle h, const float* x0, size_t bytes);

/* Problem::OverwriteSolution (src/problem.cpp:188-194) from host arrays: sets
 * the warm start every ilqg_solve_begin starts from (and the working iterate).
 * Any pointer may be NULL (= zeros). */
int ilqg_upload_warmstart(ilqg_handle h, const float* xs, const float* us,
                          const float* Ps, const float* alphas);

/* Generic upload of one per-instance array (ILQG_LAMBDAS, ILQG_MU, ILQG_MERIT,
 * ILQG_TIME_OF_EXTREME ...): the reference's mutable statics made per-instance
 * (Constraint::mu_ src/constraint.cpp:61, Constraint::lambdas_ constraint.h:138,
 * ILQSolver::last_merit_function_value_ ilq_solver.h:189). */
int ilqg_upload(ilqg_handle h, int what, const void* src, size_t bytes);

/* Host-supplied LQ game: the arguments of LQFeedbackSolver::Solve
 * (lq_feedback_solver.h:119-124).  Layouts as in the header comment. */
int ilqg_upload_lq(ilqg_handle h, const float* A, const float* Bs,
                   const float* Q, const float* l, const float* R,
                   const float* r);

/* ILQSolver::Solve prologue, src/ilq_solver.cpp:86-107: xs[0] <- x0, rollout
 * under the current strategies (ILQSolver::CurrentOperatingPoint :174-206 with
 * MultiPlayerDynamicalSystem::Integrate src/multi_player_dynamical_system.cpp:52-77),
 * TotalCosts (:220-257); marks every instance RUNNING, iteration count 0. */
int ilqg_solve_begin(ilqg_handle h);

/* ILQSolver::ComputeLinearization (:437-455) + ComputeCostQuadraticization
 * (:471-490) at the current operating point, fused; fills the LQ records. */
int ilqg_linearize_quadraticize(ilqg_handle h);

/* LQFeedbackSolver::Solve (src/lq_feedback_solver.cpp:71-244) -- or, with
 * ilqg_solver_params::open_loop, LQOpenLoopSolver::Solve
 * (src/lq_open_loop_solver.cpp:73-195) -- on the current LQ records with the
 * x0 argument ILQG_LQ_X0 (zeros unless uploaded; ILQSolver passes x0 - xs[0] = 0), plus
 * ILQSolver::ExpectedDecrease (:364-398).  Result: ILQG_LQ_PS/ILQG_LQ_ALPHAS,
 * ILQG_DELTA_XS, ILQG_EXPECTED_DECREASE. */
int ilqg_lq_backward(ilqg_handle h);

/* ILQSolver::ModifyLQStrategies (:289-348: alpha scaling, rollouts,
 * MeritFunction :400-435, Armijo :350-362, HasConverged ilq_solver.h:126-130)
 * followed by TotalCosts (:158).  Per instance, run to completion (all backtracking
 * candidates).  The LQ records are NOT refreshed by this call (the merit function only needs
 * gradients); ilqg_linearize_quadraticize / ilqg_iterate rebuild them. */
int ilqg_linesearch(ilqg_handle h);

/* Up to max_iters passes of the while loop src/ilq_solver.cpp:123-166 for every RUNNING
 * instance (linearize_quadraticize, lq_backward, linesearch), entirely device-side with no host
 * synchronisation.  iters_done (nullable; forces a sync) receives the largest per-instance
 * iteration count. */
int ilqg_iterate(ilqg_handle h, int max_iters, int* iters_done);

/* Number of instances still ILQG_STATUS_RUNNING (synchronises). */
int ilqg_count_running(ilqg_handle h, int* running);

/* One augmented-Lagrangian outer update, src/augmented_lagrangian_solver.cpp:
 * 113-178: lambda <- max(0, lambda + mu g) with the TimeIndex quirk, mu scaling,
 * per-instance failure down-scaling. max constraint error -> ILQG_MAX_CONSTRAINT_ERROR. */
int ilqg_al_update(ilqg_handle h);

/* Problem::OverwriteSolution(log->FinalOperatingPoint(), log->FinalStrategies())
 * (src/problem.cpp:188-194, called at augmented_lagrangian_solver.cpp:151-154):
 * the working iterate becomes the warm start of the next ilqg_solve_begin.
 * only_successful != 0 skips instances whose last solve failed its linesearch. */
int ilqg_overwrite_solution(ilqg_handle h, int only_successful);

/* Problem::SetUpNextRecedingHorizon(x0, t0, planner_runtime) (src/problem.cpp:64-186, with
 * SyncToExistingProblem :64-125 and MultiPlayerIntegrableSystem::IntegrateToNextTimeStep /
 * Integrate, src/multi_player_integrable_system.cpp:88-143) for every game: integrate the
 * measured state x0 [batch][n] (taken at time t0) along the warm start for ~planner_runtime,
 * find the nearest state of the plan, make it the new first time step, shift operating point and
 * strategies, and extend them to the horizon with zero controls.  Afterwards ILQG_X0 holds the
 * new initial states and the warm start the shifted plan; *new_t0 (may be NULL) receives the new
 * OperatingPoint::t0.  Times are shared by the batch.  Returns ILQG_ERR_INVALID_ARGUMENT where
 * the reference CHECK-fails on the times, ILQG_ERR_UNSUPPORTED for problems with constraints
 * (the reference aborts there once the initial time is nonzero: Constraint::TimeIndex is asked
 * for RelativeTime(kk) < initial_time_, include/ilqgames/utils/relative_time_tracker.h:63-72). */
int ilqg_setup_next_receding_horizon(ilqg_handle h, const float* x0, double t0,
                                     double planner_runtime, double* new_t0);

/* MultiPlayerIntegrableSystem::Integrate(Time t0, Time t, x0, operating_point, strategies)
 * (src/multi_player_integrable_system.cpp:54-83, with IntegrateToNextTimeStep :113-143, the time
 * step loop :85-111 and IntegrateFromPriorTimeStep :145-171) for every game: the state x0
 * [batch][n], taken at time t0, is carried to time t under the game's warm start (operating point
 * and strategies, u_i = u_ref - P_i (x - x_ref) - alpha_i), which is not modified.  x_out
 * [batch][n] receives the states; it may alias x0.  This is what a receding-horizon caller runs
 * between solves (src/receding_horizon_simulator.cpp:100-102,126-128).  Times are shared by the
 * batch.  ILQG_ERR_INVALID_ARGUMENT where the reference CHECK-fails: t < t0, t0 before the plan's
 * t0, or t at/after the plan's last time step. */
int ilqg_integrate_plan(ilqg_handle h, const float* x0, double t0, double t, float* x_out);

/* src/augmented_lagrangian_solver.cpp:165-178: for instances whose inner solve
 * failed, lambda *= geometric_lambda_downscaling, mu *= geometric_mu_downscaling. */
int ilqg_al_post_solve(ilqg_handle h);

/* AugmentedLagrangianSolver::Solve (src/augmented_lagrangian_solver.cpp:72-210) for a batch:
 * every game runs ITS OWN outer loop (own multipliers, own mu, own iterate count, own exit),
 * driven in rounds by the caller:
 *
 *     ilqg_al_begin(h, params.max_solver_iters of the AL solver, constraint_error_tolerance);
 *     do { ilqg_solve_begin(h); ilqg_iterate(h, ...) until ilqg_count_running == 0;
 *          ilqg_al_advance(h, &active); } while (active > 0);
 *
 * ilqg_al_advance does, per game still in its loop: account the inner solve's iterates and
 * success (:94-100, :179-180), down-scale multipliers after a failed inner solve (:165-178, not
 * after the first solve), test the loop condition (:109-111; the wall-clock term is the
 * caller's) and either finish the game (final success flag :187-190; later
 * ilqg_solve_begin calls skip it) or run the multiplier sweep + ScaleMu (:113-143) and
 * OverwriteSolution after a successful inner solve (:151-154).  `active` = games that need
 * another inner solve.  ilqg_solver_params.max_solver_iters is the INNER solver's cap
 * (params.unconstrained_solver_max_iters, solver_params.h).  ilqg_reset ends an AL solve. */
int ilqg_al_begin(ilqg_handle h, int max_iterates, float constraint_error_tolerance);
int ilqg_al_advance(ilqg_handle h, int* active);

int ilqg_download(ilqg_handle h, int what, void* dst, size_t bytes);
int ilqg_synchronize(ilqg_handle h);

/* Re-initialise per-instance state without reallocating.  `mask` bits:
 *   ILQG_RESET_SOLVER       last merit / expected decrease = +inf: a freshly constructed
 *                           ILQSolver (ilq_solver.h:69-74; see SURVEY Q8 for why this matters)
 *   ILQG_RESET_MULTIPLIERS  lambda = 0, mu = 10 (augmented_lagrangian_solver.cpp:196-207)
 *   ILQG_RESET_SOLUTION     zero operating point and strategies, t0 back to the descriptor's
 *                           initial time (Problem::Initialize) */
/*   ILQG_RESET_LAMBDAS / ILQG_RESET_MU  the two halves of ILQG_RESET_MULTIPLIERS on their own
 *                           (SolverParams::reset_lambdas / reset_mu are independent flags, :196-207) */
enum { ILQG_RESET_SOLVER = 1, ILQG_RESET_MULTIPLIERS = 2, ILQG_RESET_SOLUTION = 4, ILQG_RESET_LAMBDAS = 8, ILQG_RESET_MU = 16 };
int ilqg_reset(ilqg_handle h, int mask);

/* Run this handle's kernels and copies on the caller's CUDA stream (a cudaStream_t passed as
 * a plain pointer; NULL restores the handle's own stream).  Lets a host framework order and
 * time the work with its own events. */
int ilqg_set_stream(ilqg_handle h, void* cuda_stream);

/* Per-kernel device timing for the roofline report: when enabled, every hot-path launch is
 * bracketed by CUDA events on the launch stream.  ilqg_profile_read synchronizes and returns
 * the accumulated milliseconds and launch count of one kernel kind
 * (0 = linearize_quadraticize, 1 = lq_backward, 2 = linesearch [all of its launches],
 * 3 = solve_begin [all of its launches]; the launches inside 2 and 3 one by one:
 * 4 = rollout + merit kernels of the first linesearch window, 5 = of the queued window,
 * 6 = k_ls_decide, 7 = rollout + merit of solve_begin) since the last ilqg_profile(h, 1). */
int ilqg_profile(ilqg_handle h, int enable);
int ilqg_profile_read(ilqg_handle h, int kernel, double* total_ms, long long* launches);

/* Launch accounting for bench.py ("gpu_launches"): number of kernels this
 * handle has launched since creation (0 for the oracle). */
int ilqg_kernel_launches(ilqg_handle h, long long* out);

#ifdef __cplusplus
}
#endif
#endif /* ILQG_H */
