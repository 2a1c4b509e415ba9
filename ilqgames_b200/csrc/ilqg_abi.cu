// ilqg_abi.cu -- host side of the C ABI declared in include/ilqg.h: builds the device
// descriptor, owns the per-device slab and launches the sm_100a kernels.  No CPU
// implementation of the hot path lives here: without a CUDA device every entry point
// that would compute returns ILQG_ERR_NO_DEVICE / ILQG_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges around every launch site

#include "ilqg_backward_any.cuh"
#include "ilqg_backward_tc.cuh"
#include "ilqg_linesearch.cuh"
#include "ilqg_open_loop.cuh"
#include "ilqg_receding.cuh"

using namespace ilqg;

// One group of instances: its own slab, scratch and stream.  The public handle (ilqg_solver, at the
// end of this file) owns several groups and runs them on concurrent streams.
struct SubSolver {
  DevDesc d;
  DevParams p;
  ilqg_problem_desc host_desc;
  ilqg_layout layout;
  Slab s;
  LsScratch ls;
  RecordPattern pat;
  bool pat_ok;
  bool v3_ok = false;       // the dense-record K_lq v3 fits its shared-memory budget for this pattern
  // compact records (ilqg_records.cuh) and the tensor-core backward sweep's tables (ilqg_backward_tc.cuh)
  CompactPattern cp;
  TcTables tc;
  bool cp_ok = false;       // K_lq v4 + k_lq_backward_tc usable for this descriptor
  int tc_nxp = 0, tc_mup = 0;
  bool classic = true;      // only round 1's subsystem / record kinds, no gates, no groups: the lean kernel instances
  // ILQG_TRACE=<file>: a device timeline -- a one-thread kernel stamps %globaltimer after every launch site, in
  // stream order; written to <file> when the handle is destroyed (development aid, off by default)
  unsigned long long* trace = nullptr;
  int trace_n = 0, trace_cap = 0;
  std::vector<int> trace_tags;
  std::string trace_path;
  int rollout_mode = 0;     // ILQG_ROLLOUT: 0 auto (by window size, ilqg_linesearch.cuh: LsPick), 1 sp, 2 lanes (A/B runs, the bit-identity test)
  int cap_sp = 0, cap_lanes = 0;
  int tc_blocks_cap = 0;    // ILQG_TC_BLOCKS
  bool use_compact = true;  // ILQG_RECORDS=dense forces the round-1 dense-record kernels (A/B runs)
  // which representation of the LQ records is current: K_lq v4 writes the compact one,
  // ilqg_upload_lq / K_lq v3 the dense one; EnsureDense expands compact -> dense on demand
  bool compact_valid = false, dense_valid = false;
  int ls_blocks_max;
  int ls_cur;  // the open-linesearch queue the next continued window consumes (ilqg_linesearch.cuh: LsScratch::pend)
  std::vector<int> ls_tiers;  // candidates per continued window, in order (sum = max_backtracking_steps - JA)
  int device;
  int B;
  cudaStream_t stream;      // the stream work is issued on
  cudaStream_t own_stream;  // created with the handle
  // pipelined iteration: the open linesearches of pass i finish on `side` while `stream` already
  // linearizes / solves pass i + 1 for everyone else (IteratePipelined)
  cudaStream_t side;
  cudaEvent_t ev_fresh, ev_side;
  bool side_busy = false;
  // staggered groups (ILQG_STAGGER): group g starts its first pass when group g - 1 has finished its first
  // backward sweep, so that one group's latency-bound linesearch runs under the other's K_lq / K_bwd
  cudaEvent_t stagger_signal = nullptr;   // recorded after this group's first K_bwd of an iterate call
  cudaEvent_t stagger_wait = nullptr;     // armed per iterate call: the previous group's signal
  int sm_count = 148;
  // SolverParams::open_loop: LQOpenLoopSolver instead of LQFeedbackSolver (ilq_solver.h:76-81)
  bool open_loop = false;
  float* ol_scratch = nullptr;  // [B][T][OlLayout::srec], ilqg_open_loop.cuh
  int pipeline;  // 0 off, 1 = K_lq + K_bwd of the queue on the side stream, 2 = K_lq only
  std::vector<void*> allocs;
  // per-kernel event timing (ilqg_profile)
  bool profiling;
  struct Sample { cudaEvent_t a, b; int kind; };
  std::vector<Sample> samples;
  double prof_ms[8];
  long long prof_n[8];
  float* staging;
  size_t staging_floats;
  long long launches;
  int dims_key;  // index into the dispatch table
};

namespace {

#define CUDA_TRY(expr)                                                            \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      std::fprintf(stderr, "[ilqg_b200] %s failed: %s (%s:%d)\n", #expr,          \
                   cudaGetErrorString(_e), __FILE__, __LINE__);                   \
      return ILQG_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

template <typename T>
int DevAlloc(SubSolver* h, T** out, size_t count, bool zero = true);

inline bool IsConstraintKind(int kind) {
  return kind == ILQG_CONSTRAINT_PROXIMITY || kind == ILQG_CONSTRAINT_SINGLE_DIMENSION ||
         kind == ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE;
}

int SubsystemXdim(int kind) {
  switch (kind) {
    case ILQG_DYN_CAR6D: return 6;
    case ILQG_DYN_CAR5D: return 5;
    case ILQG_DYN_UNICYCLE4D:
    case ILQG_DYN_POINT_MASS_2D:
    case ILQG_DYN_TWO_PLAYER_UNICYCLE4D: return 4;
    case ILQG_DYN_DUBINS:
    case ILQG_DYN_AIR3D: return 3;
    default: return 0;
  }
}

// FinalTimeCost::Evaluate / Quadraticize (include/ilqgames/cost/final_time_cost.h:64-77): active iff
// t >= initial_time_ + threshold_time_ with t = RelativeTime(kk) = kk * kTimeStep in double
// (relative_time_tracker.h:63-65) and initial_time_ the tracker's static, which
// Problem::SyncToExistingProblem moves (src/problem.cpp:120).  -> first time step of every record.
void UpdateCostGates(DevDesc* d, double tracker_initial_time) {
  for (int c = 0; c < d->num_costs; c++) {
    const double threshold = d->cost[c].active_from;
    int first = 0;
    if (threshold != 0.0) {
      first = d->T;
      for (int kk = 0; kk < d->T; kk++)
        if (static_cast<double>(kk) * d->time_step >= tracker_initial_time + threshold) { first = kk; break; }
    }
    d->cost[c].first_step = first;
  }
}

// Host "problem compiler": ilqg_problem_desc -> DevDesc + ilqg_layout.
int BuildDeviceDesc(const ilqg_problem_desc& h, DevDesc* d, ilqg_layout* lo, std::vector<int>* lidx) {
  std::memset(d, 0, sizeof(*d));
  std::memset(lo, 0, sizeof(*lo));
  if (h.num_time_steps < 2 || h.num_time_steps > ILQG_MAX_TIME_STEPS) return ILQG_ERR_INVALID_ARGUMENT;
  if (h.num_players < 1 || h.num_players > ILQG_MAX_PLAYERS) return ILQG_ERR_INVALID_ARGUMENT;
  if (h.xdim < 1 || h.xdim > ILQG_MAX_XDIM) return ILQG_ERR_INVALID_ARGUMENT;
  if (h.num_costs < 0 || h.num_costs > ILQG_MAX_COSTS) return ILQG_ERR_INVALID_ARGUMENT;
  if (h.num_subsystems < 0 || h.num_subsystems > ILQG_MAX_SUBSYSTEMS) return ILQG_ERR_INVALID_ARGUMENT;
  if (h.num_polylines < 0 || h.num_polylines > ILQG_MAX_POLYLINES) return ILQG_ERR_INVALID_ARGUMENT;
  d->T = h.num_time_steps;
  d->N = h.num_players;
  d->n = h.xdim;
  d->time_step = h.time_step;
  d->uoff[0] = 0;
  for (int i = 0; i < d->N; i++) {
    if (h.udim[i] < 1) return ILQG_ERR_INVALID_ARGUMENT;
    d->udim[i] = h.udim[i];
    d->uoff[i + 1] = d->uoff[i] + h.udim[i];
    d->state_reg[i] = h.state_regularization[i];
    d->control_reg[i] = h.control_regularization[i];
    d->cost_structure[i] = h.cost_structure[i];
  }
  d->M = d->uoff[d->N];
  if (d->M > ILQG_MAX_UDIM) return ILQG_ERR_INVALID_ARGUMENT;

  d->num_subsystems = h.num_subsystems;
  for (int k = 0; k < h.num_subsystems; k++) {
    const ilqg_subsystem_desc& hs = h.subsystems[k];
    DevSubsystem& ds = d->sub[k];
    const int xd = SubsystemXdim(hs.kind);
    if (xd == 0 || hs.x_offset < 0 || hs.x_offset + xd > d->n) return ILQG_ERR_INVALID_ARGUMENT;
    const bool two = hs.kind == ILQG_DYN_AIR3D || hs.kind == ILQG_DYN_TWO_PLAYER_UNICYCLE4D;
    const int np = two ? 2 : 1;
    if (hs.first_player < 0 || hs.first_player + np > d->N) return ILQG_ERR_INVALID_ARGUMENT;
    // control dimension of each of its players: 1 for Air3D's turn rates and the Dubins car, else 2
    const int m_each = (hs.kind == ILQG_DYN_AIR3D || hs.kind == ILQG_DYN_DUBINS) ? 1 : 2;
    for (int q = 0; q < np; q++)
      if (d->udim[hs.first_player + q] != m_each) return ILQG_ERR_INVALID_ARGUMENT;
    ds.kind = hs.kind;
    ds.x_offset = hs.x_offset;
    ds.first_player = hs.first_player;
    ds.u_offset = d->uoff[hs.first_player];
    ds.u_offset2 = np == 2 ? d->uoff[hs.first_player + 1] : ds.u_offset + 1;
    ds.nu = np * m_each;
    for (int q = 0; q < 4; q++) ds.ucol[q] = 0;
    for (int pl = 0, q = 0; pl < np; pl++)
      for (int c = 0; c < m_each; c++) ds.ucol[q++] = d->uoff[hs.first_player + pl] + c;
    ds.p0 = hs.params[0];
    ds.p1 = hs.params[1];
  }

  // polylines -> segments (Polyline2 ctor src/polyline2.cpp:49-60; LineSegment2 ctor
  // line_segment2.h:55-62, fp32)
  d->num_polylines = h.num_polylines;
  int nseg = 0;
  for (int p = 0; p < h.num_polylines; p++) {
    d->seg_start[p] = nseg;
    const int a = h.polyline_start[p], b = h.polyline_start[p + 1];
    if (a < 0 || b > ILQG_MAX_POLYLINE_POINTS || b - a < 2) return ILQG_ERR_INVALID_ARGUMENT;
    // ranges must not overlap or run backwards, and all segments together must fit the table
    if (p > 0 && a < h.polyline_start[p]) return ILQG_ERR_INVALID_ARGUMENT;
    if (nseg + (b - a - 1) > ILQG_MAX_POLYLINE_POINTS) return ILQG_ERR_INVALID_ARGUMENT;
    for (int q = a + 1; q < b; q++) {
      DevSegment& sg = d->seg[nseg++];
      sg.p1x = h.polyline_points[q - 1][0];
      sg.p1y = h.polyline_points[q - 1][1];
      sg.p2x = h.polyline_points[q][0];
      sg.p2y = h.polyline_points[q][1];
      const float dx = sg.p1x - sg.p2x, dy = sg.p1y - sg.p2y;
      sg.length = std::sqrt(dx * dx + dy * dy);
      if (!(sg.length > kSmallNumber)) return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_GT line_segment2.h:61
      sg.ux = (sg.p2x - sg.p1x) / sg.length;
      sg.uy = (sg.p2y - sg.p1y) / sg.length;
    }
  }
  d->seg_start[h.num_polylines] = nseg;

  // control pairs (i,j) sorted by (i,j): (i,i) always (src/lq_feedback_solver.cpp:139-141)
  bool has[ILQG_MAX_PLAYERS][ILQG_MAX_PLAYERS] = {};
  for (int i = 0; i < d->N; i++) has[i][i] = true;
  for (int c = 0; c < h.num_costs; c++) {
    const ilqg_cost_desc& cd = h.costs[c];
    if (cd.player < 0 || cd.player >= d->N || cd.arg >= d->N || cd.arg < -1) return ILQG_ERR_INVALID_ARGUMENT;
    // a constraint cannot be time-gated; a member of an ExtremeValueCost group is a plain cost
    if (IsConstraintKind(cd.kind) && cd.active_from != 0.0) return ILQG_ERR_INVALID_ARGUMENT;
    if (cd.group < 0) return ILQG_ERR_INVALID_ARGUMENT;
    if (cd.group > 0 && (IsConstraintKind(cd.kind) || cd.active_from != 0.0)) return ILQG_ERR_INVALID_ARGUMENT;
    if (cd.arg >= 0) has[cd.player][cd.arg] = true;
  }
  d->num_pairs = 0;
  for (int i = 0; i < d->N; i++)
    for (int j = 0; j < d->N; j++) {
      d->pair_of[i][j] = -1;
      if (!has[i][j]) continue;
      const int p = d->num_pairs++;
      d->pair_of[i][j] = p;
      d->pair_i[p] = i;
      d->pair_j[p] = j;
      d->pair_Roff[p] = d->R_floats;
      d->pair_roff[p] = d->r_floats;
      d->R_floats += d->udim[j] * d->udim[j];
      d->r_floats += d->udim[j];
    }
  if (d->r_floats > ILQG_MAX_UDIM * ILQG_MAX_PLAYERS) return ILQG_ERR_UNSUPPORTED;

  // cost records, stably sorted by owner; lambda slots follow the host record order
  int slot_of[ILQG_MAX_COSTS];
  d->num_constraints = 0;
  for (int c = 0; c < h.num_costs; c++)
    slot_of[c] = IsConstraintKind(h.costs[c].kind) ? d->num_constraints++ : -1;
  d->num_costs = 0;
  for (int i = 0; i < d->N; i++) {
    d->cost_begin[i] = d->num_costs;
    for (int c = 0; c < h.num_costs; c++) {
      const ilqg_cost_desc& cd = h.costs[c];
      if (cd.player != i) continue;
      if (cd.kind < ILQG_COST_QUADRATIC || cd.kind > ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE)
        return ILQG_ERR_UNSUPPORTED;
      const int dim = cd.arg < 0 ? d->n : d->udim[cd.arg];
      const bool needs_poly = cd.kind == ILQG_COST_QUADRATIC_POLYLINE2 ||
                              cd.kind == ILQG_COST_SEMIQUADRATIC_POLYLINE2 ||
                              cd.kind == ILQG_COST_POLYLINE2_SIGNED_DISTANCE ||
                              cd.kind == ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE;
      if (needs_poly && (cd.polyline < 0 || cd.polyline >= h.num_polylines)) return ILQG_ERR_INVALID_ARGUMENT;
      const int ndims = (cd.kind == ILQG_COST_PROXIMITY || cd.kind == ILQG_CONSTRAINT_PROXIMITY ||
                         cd.kind == ILQG_COST_SIGNED_DISTANCE) ? 4
                        : needs_poly ? 2 : 1;
      if (cd.kind == ILQG_COST_QUADRATIC_DIFFERENCE) {  // flag = 1 or 2 pairs (dim[k], dim[2 + k])
        if (cd.flag < 1 || cd.flag > 2) return ILQG_ERR_INVALID_ARGUMENT;
        for (int q = 0; q < cd.flag; q++)
          if (cd.dim[q] < 0 || cd.dim[q] >= dim || cd.dim[2 + q] < 0 || cd.dim[2 + q] >= dim) return ILQG_ERR_INVALID_ARGUMENT;
      } else {
        for (int q = 0; q < ndims; q++) {
          const bool all_dims_ok = cd.kind == ILQG_COST_QUADRATIC && q == 0 && cd.dim[0] < 0;
          if (!all_dims_ok && (cd.dim[q] < 0 || cd.dim[q] >= dim)) return ILQG_ERR_INVALID_ARGUMENT;
        }
      }
      DevCost& dc = d->cost[d->num_costs++];
      dc.kind = cd.kind;
      dc.player = cd.player;
      dc.arg = cd.arg;
      dc.is_equality = cd.is_equality;
      dc.d0 = cd.dim[0];
      dc.d1 = cd.dim[1];
      dc.d2 = cd.dim[2];
      dc.d3 = cd.dim[3];
      dc.flag = cd.flag;
      dc.polyline = cd.polyline;
      dc.slot = slot_of[c];
      dc.pair = cd.arg >= 0 ? d->pair_of[cd.player][cd.arg] : -1;
      dc.weight = cd.weight;
      dc.value = cd.value;
      dc.first_step = 0;  // (UpdateCostGates below)
      dc.active_from = cd.active_from;
      dc.group = cd.group;
      dc.group_is_min = cd.group_is_min;
      dc.group_end = d->num_costs;
    }
  }
  d->cost_begin[d->N] = d->num_costs;
  // extent of every ExtremeValueCost group: consecutive records of one player and argument with one id
  for (int c = 0; c < d->num_costs;) {
    int e = c + 1;
    if (d->cost[c].group > 0)
      while (e < d->num_costs && d->cost[e].group == d->cost[c].group && d->cost[e].player == d->cost[c].player &&
             d->cost[e].arg == d->cost[c].arg)
        e++;
    for (int m = c; m < e; m++) d->cost[m].group_end = e;
    c = e;
  }
  UpdateCostGates(d, h.initial_time);

  // LQ record layout
  const int n = d->n, M = d->M, N = d->N;
  d->offA = 0;
  d->offB = d->offA + n * n;
  d->offQ = d->offB + n * M;
  d->offl = d->offQ + N * n * n;
  d->offR = d->offl + N * n;
  d->offr = d->offR + d->R_floats;
  d->rec = (d->offr + d->r_floats + 3) & ~3;

  // kk -> Constraint::TimeIndex(RelativeTime(kk)) in double
  // (include/ilqgames/utils/relative_time_tracker.h:63-72), SURVEY Q1
  lidx->resize(d->T);
  for (int kk = 0; kk < d->T; kk++) {
    const double t = static_cast<double>(kk) * h.time_step;
    long idx = static_cast<long>(static_cast<size_t>((t - h.initial_time) / h.time_step));
    idx = std::max(0L, std::min<long>(idx, d->T - 1));
    (*lidx)[kk] = (int)idx;
  }

  lo->num_time_steps = d->T;
  lo->num_players = N;
  lo->xdim = n;
  lo->total_udim = M;
  for (int i = 0; i < N; i++) {
    lo->udim[i] = d->udim[i];
    lo->u_offset[i] = d->uoff[i];
  }
  lo->num_pairs = d->num_pairs;
  for (int p = 0; p < d->num_pairs; p++) {
    lo->pair_player[p] = d->pair_i[p];
    lo->pair_arg[p] = d->pair_j[p];
    lo->pair_R_offset[p] = d->pair_Roff[p];
    lo->pair_r_offset[p] = d->pair_roff[p];
  }
  lo->R_floats = d->R_floats;
  lo->r_floats = d->r_floats;
  lo->num_constraints = d->num_constraints;
  lo->record_floats = d->rec;
  lo->compact_record_floats = 0;
  for (int kk = 0; kk < d->T; kk++) lo->lambda_index[kk] = (*lidx)[kk];
  return ILQG_OK;
}

// ------------------------- compiled dimension table -------------------------
// (n, M, N) envelopes with a specialised backward kernel.
struct DimsEntry {
  int n, M, N;
};
constexpr DimsEntry kDims[] = {
    {16, 6, 3},  // ThreePlayerIntersection: 2x Car6D + Unicycle4D
    {24, 8, 4},  // RoundaboutMerging: 4x Car6D
    {3, 2, 2},   // Air3D
    {2, 2, 2},   // test/test_lq_solver.cpp point-mass LQ game
    {-1, -1, -1},  // (was the n = 12 half-warp instance: never reached by a test, removed)
    {18, 6, 3},  // ThreePlayerOvertaking: 3x Car6D (warp-per-game kernel: 18 is not a multiple of 4)
};
constexpr int kNumDims = sizeof(kDims) / sizeof(kDims[0]);


// NVTX range names of the launch sites (the `kind` of a ProfScope)
static const char* const kRangeNames[8] = {"ilqg/K_lq linearize+quadraticize", "ilqg/K_bwd Riccati sweep", "ilqg/linesearch",
                                           "ilqg/solve_begin", "ilqg/K_ls first window", "ilqg/K_ls queued window",
                                           "ilqg/K_ls decide", "ilqg/K_ls prologue rollout"};

struct NvtxRange {  // the launch sites outside the hot loop
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

__global__ void k_stamp(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
// timeline stamp after a launch site (tag: site * 10 + 0 main / 1 side stream)
inline void Stamp(SubSolver* h, int site) {
  if (!h->trace || h->trace_n >= h->trace_cap) return;
  k_stamp<<<1, 1, 0, h->stream>>>(h->trace + h->trace_n++);
  h->trace_tags.push_back(site * 10 + (h->stream == h->side ? 1 : 0));
}

struct ProfScope {  // brackets one launch site: an NVTX range always, CUDA events when profiling is on
  SubSolver* h;
  cudaEvent_t a, b;
  int kind;
  bool on;
  ProfScope(SubSolver* hh, int k) : h(hh), a(nullptr), b(nullptr), kind(k), on(hh->profiling) {
    nvtxRangePushA(kRangeNames[k & 7]);
    if (on) {
      on = cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess;
      if (on) cudaEventRecord(a, h->stream);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(b, h->stream);
      h->samples.push_back({a, b, kind});
    }
    nvtxRangePop();
  }
};

template <typename K>
int SetSmem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return ILQG_OK;
}

// ---- compact records + tensor-core backward sweep (ilqg_records.cuh, ilqg_backward_tc.cuh) ----
// instantiations of k_lq_backward_tc: padded state dimension, padded stacked control dimension,
// scatter / Q-add table entries per lane, minimum resident blocks per SM
#define ILQG_TC_INSTANCES(X)                                                                        \
  X(8, 2, 2, 4, 8) X(8, 4, 2, 4, 8) X(16, 2, 3, 6, 4) X(16, 4, 3, 6, 4) X(16, 6, 3, 6, 4) X(16, 8, 3, 6, 4) \
  X(24, 4, 4, 10, 3) X(24, 6, 4, 10, 3) X(24, 8, 4, 10, 3)

// the per-lane table budgets of the instance for (nxp, mup), or false when there is none
bool TcBudget(int nxp, int mup, int* sc, int* qa) {
#define X(NXP, MUP, SC, QA, MINB) \
  if (nxp == NXP && mup == MUP) { *sc = SC; *qa = QA; return true; }
  ILQG_TC_INSTANCES(X)
#undef X
  return false;
}

int EnsureDenseAlloc(SubSolver* h) {
  if (h->s.rec) return ILQG_OK;
  return DevAlloc(h, &h->s.rec, (size_t)h->B * h->d.T * (size_t)h->d.rec);
}

// dense LQ records for the consumers that want them (downloads, the open-loop solver, the
// dense-record backward kernels): expanded from the compact ones when those are newer
int EnsureDense(SubSolver* h) {
  int rc = EnsureDenseAlloc(h);
  if (rc != ILQG_OK) return rc;
  if (h->dense_valid || !h->compact_valid) return ILQG_OK;
  NvtxRange nvtx("ilqg/expand compact records");
  const long long recs = (long long)h->B * h->d.T;
  k_expand_records<<<(int)((recs + 3) / 4), 128, 0, h->stream>>>(h->d, h->s, h->pat, h->cp);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  h->dense_valid = true;
  return ILQG_OK;
}

template <int NX, int MU, int NP>
int LaunchBackward(SubSolver* h, int only_running, Sel) {
  const size_t smem = sizeof(float) * KBWD_WARPS * (size_t)(BwdSmem<NX, MU, NP>::rec + h->d.rec);
  int rc = SetSmem(k_lq_backward<NX, MU, NP>, smem);
  if (rc != ILQG_OK) return rc;
  const int blocks = (h->B + KBWD_WARPS - 1) / KBWD_WARPS;
  ProfScope prof(h, 1);
  k_lq_backward<NX, MU, NP><<<blocks, KBWD_WARPS * 32, smem, h->stream>>>(h->d, h->p, h->s, only_running);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

template <int NX, int MU, int NP>
int LaunchBackwardHw(SubSolver* h, int only_running, Sel sel) {
  const size_t per_inst = HwSmem<NX, MU, NP>::lrr + (h->d.rec - h->d.offl);
  const size_t smem = sizeof(float) * KHW_WARPS * 2 * per_inst;
  int rc = SetSmem(k_lq_backward_hw<NX, MU, NP>, smem);
  if (rc != ILQG_OK) return rc;
  const int per_block = KHW_WARPS * 2;
  ProfScope prof(h, 1);
  k_lq_backward_hw<NX, MU, NP><<<(h->B + per_block - 1) / per_block, KHW_WARPS * 32, smem, h->stream>>>(
      h->d, h->p, h->s, only_running, sel);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

int LaunchOpenLoop(SubSolver* h, int only_running, bool use_lq_x0) {
  const OlLayout L = ol_layout(h->d.n, h->d.M, h->d.N, h->d.rec - h->d.offl);
  const size_t smem = sizeof(float) * KOL_WARPS * (size_t)L.total;
  int rc = SetSmem(k_lq_open_loop, smem);
  if (rc != ILQG_OK) return rc;
  ProfScope prof(h, 1);
  k_lq_open_loop<<<(h->B + KOL_WARPS - 1) / KOL_WARPS, KOL_WARPS * 32, smem, h->stream>>>(
      h->d, h->s, h->ol_scratch, only_running, use_lq_x0 ? h->s.lq_x0 : nullptr);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

// with_dxs: also produce ILQG_DELTA_XS (an optional output of LQFeedbackSolver::Solve that the
// iLQ loop itself never reads once ExpectedDecrease is fused into the backward sweep)
template <int NXP, int MUP, int SC, int QA, int MINB>
int LaunchBackwardTc(SubSolver* h, int only_running, Sel sel, const float* x0arg) {
  size_t smem = sizeof(float) * (((size_t)h->tc.total_words + 3) / 4 * 4 + 2 * KTC_WARPS + (size_t)KTC_WARPS * h->tc.per_game);
  // ILQG_TC_BLOCKS = k caps the sweep at k resident blocks per SM (by asking for more shared memory than it
  // needs), which leaves registers and shared memory for the side stream's linesearch blocks to run under it
  if (h->tc_blocks_cap > 0)  // the least request with which cap + 1 blocks no longer fit (1 KB per block is reserved)
    smem = std::max(smem, ((size_t)(228 * 1024 / (h->tc_blocks_cap + 1) - 1024 + 16) + 15) & ~(size_t)15);
  int rc = SetSmem(k_lq_backward_tc<NXP, MUP, SC, QA, MINB>, smem);
  if (rc != ILQG_OK) return rc;
  ProfScope prof(h, 1);
  k_lq_backward_tc<NXP, MUP, SC, QA, MINB><<<(h->B + KTC_WARPS - 1) / KTC_WARPS, KTC_WARPS * 32, smem, h->stream>>>(
      h->d, h->p, h->s, h->tc, only_running, sel, x0arg);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

int DispatchBackwardTc(SubSolver* h, int only_running, Sel sel, const float* x0arg) {
#define X(NXP, MUP, SC, QA, MINB) \
  if (h->tc_nxp == NXP && h->tc_mup == MUP) return LaunchBackwardTc<NXP, MUP, SC, QA, MINB>(h, only_running, sel, x0arg);
  ILQG_TC_INSTANCES(X)
#undef X
  return ILQG_ERR_UNSUPPORTED;
}

int DispatchBackward(SubSolver* h, int only_running, bool with_dxs, Sel sel = Sel{SEL_ALL, nullptr, nullptr}) {
  int rc = ILQG_ERR_UNSUPPORTED;
  if (!h->open_loop && h->cp_ok && h->use_compact && h->compact_valid) {
    // the hot path: tensor-core sweep over the compact records
    if ((rc = DispatchBackwardTc(h, only_running, sel, with_dxs ? h->s.lq_x0 : nullptr)) != ILQG_OK) return rc;
    if (with_dxs) {
      if ((rc = EnsureDense(h)) != ILQG_OK) return rc;
      k_delta_xs<<<(h->B + 3) / 4, 128, sizeof(float) * 4 * 2 * ILQG_MAX_XDIM, h->stream>>>(h->d, h->s, h->s.lq_x0);
      h->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    return ILQG_OK;
  }
  if ((rc = EnsureDense(h)) != ILQG_OK) return rc;
  // with_dxs is the stand-alone ilqg_lq_backward: the only caller with a nonzero x0 argument
  if (h->open_loop) return LaunchOpenLoop(h, only_running, with_dxs);  // writes delta_xs itself
  rc = ILQG_ERR_UNSUPPORTED;
  bool hw = true;
  int key = h->dims_key;
  // the half-warp kernel hard-codes a uniform control dimension m = M / N: anything
  // else must not reach it
  if (key == 0 || key == 1)
    for (int i = 0; i < h->d.N; i++)
      if (h->d.udim[i] * h->d.N != h->d.M) key = -1;
  // a stand-alone LQ solve (with_dxs: it may carry a non-zero x0 argument) on the half-warp kernels' shapes also
  // goes to the run-time-dimension kernel: the half-warp kernel's adjoint form of ExpectedDecrease assumes
  // delta_x_0 = 0, the forward sweep of k_lq_backward_any does not
  if (with_dxs && (key == 0 || key == 1)) key = -1;
  if (key < 0 || (sel.mode != SEL_ALL && (key == 2 || key == 3 || key == 5))) {
    // any other shape (and list / main selections on the warp kernels' shapes never occur: those
    // handles do not pipeline): the run-time-dimension kernel, which writes delta_xs itself
    if (sel.mode != SEL_ALL) return ILQG_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float) * KANY_WARPS * (size_t)any_smem_floats(h->d.n, h->d.M, h->d.N, h->d.rec);
    if ((rc = SetSmem(k_lq_backward_any, smem)) != ILQG_OK) return rc;
    ProfScope prof(h, 1);
    k_lq_backward_any<<<(h->B + KANY_WARPS - 1) / KANY_WARPS, KANY_WARPS * 32, smem, h->stream>>>(
        h->d, h->p, h->s, only_running, with_dxs ? h->s.lq_x0 : nullptr);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return ILQG_OK;
  }
  switch (key) {
    case 0: rc = LaunchBackwardHw<16, 6, 3>(h, only_running, sel); break;
    case 1: rc = LaunchBackwardHw<24, 8, 4>(h, only_running, sel); break;
    case 2: rc = LaunchBackward<3, 2, 2>(h, only_running, sel); hw = false; break;
    case 3: rc = LaunchBackward<2, 2, 2>(h, only_running, sel); hw = false; break;
    case 5: rc = LaunchBackward<18, 6, 3>(h, only_running, sel); hw = false; break;
  }
  if (rc == ILQG_OK && hw && with_dxs) {
    k_delta_xs<<<(h->B + 3) / 4, 128, sizeof(float) * 4 * 2 * ILQG_MAX_XDIM, h->stream>>>(h->d, h->s, h->s.lq_x0);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return rc;
}

// upper bound on the `+= v` updates one role emits for one record (sizes the K_lq value lists)
int MaxRoleEntries(const DevDesc& d) {
  int best = 0;
  for (int i = 0; i < d.N; i++) {
    int e = 0;
    for (int c = d.cost_begin[i]; c < d.cost_begin[i + 1]; c++) {
      const DevCost& cd = d.cost[c];
      const int dim = cd.arg < 0 ? d.n : d.udim[cd.arg];
      switch (cd.kind) {
        case ILQG_COST_QUADRATIC: e += cd.d0 >= 0 ? 2 : 2 * dim; break;
        case ILQG_COST_PROXIMITY:
        case ILQG_CONSTRAINT_PROXIMITY:
        case ILQG_COST_SIGNED_DISTANCE: e += 20; break;
        case ILQG_COST_QUADRATIC_DIFFERENCE: e += 6 * cd.flag; break;
        case ILQG_COST_SEMIQUADRATIC:
        case ILQG_CONSTRAINT_SINGLE_DIMENSION: e += 2; break;
        default: e += 6; break;  // polyline kinds
      }
    }
    best = std::max(best, e);
  }
  int lin = 0;
  for (int k = 0; k < d.num_subsystems; k++) lin += 9;
  return std::max(best, lin);
}

template <typename T>
int UploadVector(SubSolver* h, const std::vector<T>& v, const T** out) {
  T* dptr = nullptr;
  int rc = DevAlloc(h, &dptr, v.size());
  if (rc != ILQG_OK) return rc;
  if (!v.empty())
    CUDA_TRY(cudaMemcpyAsync(dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *out = dptr;
  return ILQG_OK;
}

// From the gather items (one per touched record offset): the compact record layout, the g_k table of
// K_lq v4 and the index tables of the tensor-core sweep.  Leaves h->cp_ok = false (dense-record
// kernels are used) when the descriptor does not fit the compact kernels' envelope.
int BuildCompactTables(SubSolver* h, const std::vector<GatherItem>& items) {
  const DevDesc& d = h->d;
  h->cp_ok = false;
  const int n = d.n, M = d.M, N = d.N;
  const int NI = (int)items.size();
  if (NI == 0 || NI >= (int)kItemReg || d.num_subsystems == 0) return ILQG_OK;
  const int NIp = (NI + n + N + 3) & ~3;  // items | g_k | the N state regularisers | pad (16-byte multiple: bulk copies)
  const int NXP = (n + 7) & ~7, MUP = (M + 1) & ~1;
  int sc_budget = 0, qa_budget = 0;
  if (!TcBudget(NXP, MUP, &sc_budget, &qa_budget)) return ILQG_OK;
  std::vector<int> item_at(d.rec, -1);
  for (int g = 0; g < NI; g++) {
    if (item_at[items[g].off] >= 0) return ILQG_OK;  // two roles on one offset: not a compact pattern
    item_at[items[g].off] = g;
  }
  // g_k[a] = sum_i sum_c Q_i[a][c] l_i[c] over the terms that can be non-zero
  std::vector<uint2> gk;
  std::vector<int> gk_start(n + 1, 0);
  for (int a = 0; a < n; a++) {
    gk_start[a] = (int)gk.size();
    for (int i = 0; i < N; i++)
      for (int c = 0; c < n; c++) {
        const int li = item_at[d.offl + i * n + c];
        if (li < 0) continue;  // l_i[c] is the template's zero
        const int qi = item_at[d.offQ + (i * n + a) * n + c];
        unsigned qcode;
        if (qi >= 0) qcode = (unsigned)qi;
        else if (a == c && d.state_reg[i] != 0.f) qcode = kItemReg + i;
        else continue;
        gk.push_back(make_uint2(qcode | ((unsigned)li << 16), (unsigned)i));
      }
  }
  gk_start[n] = (int)gk.size();

  const TcFixed L = tc_fixed(NXP, MUP);
  TcTables& tc = h->tc;
  std::memset(&tc, 0, sizeof(tc));
  tc.NI = NI;
  tc.NIp = NIp;
  const int MM = (MUP * MUP + 3) & ~3;
  tc.off_zeta = L.fixed;
  tc.off_l = tc.off_zeta + N * NXP;
  tc.off_Om = tc.off_l + N * NXP;
  tc.off_rho = tc.off_Om + N * MM;
  tc.off_om = tc.off_rho + N * 8;
  tc.off_edr = tc.off_om + N * 8;
  tc.off_vals = tc.off_edr + N * 8;
  tc.off_Z = tc.off_vals + 2 * NIp;  // two record buffers: the next record lands while this one is in use
  // control-cost pair that owns offset `rel` of the R (mode 0) / r (mode 1) block
  auto pair_at = [&](int rel, int mode) {
    for (int pp = 0; pp < d.num_pairs; pp++) {
      const int mj = d.udim[d.pair_j[pp]];
      const int lo = mode == 0 ? d.pair_Roff[pp] : d.pair_roff[pp], len = mode == 0 ? mj * mj : mj;
      if (rel >= lo && rel < lo + len) return pp;
    }
    return -1;
  };
  tc.per_game = tc.off_Z + N * NXP * L.LD + 4;  // + a scratch word the padding entries of the tables hit
  if (tc.per_game >= 65536) return ILQG_OK;
  std::vector<unsigned> scat, bnz, brow_rows, brow, qadd;
  struct BEntry { int q, c, item, player; };
  std::vector<BEntry> bs;
  auto owner_of = [&](int c) {
    int o = 0;
    for (int i = 1; i < N; i++)
      if (c >= d.uoff[i]) o = i;
    return o;
  };
  for (int g = 0; g < NI; g++) {
    const int off = items[g].off;
    if (off < d.offB) {  // A
      const int idx = off - d.offA, q = idx / n, c = idx % n;
      scat.push_back((unsigned)g | ((unsigned)(L.F + q * L.LD + c) << 16));
    } else if (off < d.offQ) {  // B
      const int idx = off - d.offB, q = idx / M, c = idx % M;
      bs.push_back({q, c, g, owner_of(c)});
    } else if (off < d.offl) {  // Q_i
      const int idx = off - d.offQ, i = idx / (n * n), r = (idx / n) % n, c = idx % n;
      qadd.push_back((unsigned)g | ((unsigned)(i * NXP * L.LD + r * L.LD + c) << 16));
    } else if (off < d.offR) {  // l_i
      const int idx = off - d.offl, i = idx / n, a = idx % n;
      scat.push_back((unsigned)g | ((unsigned)(tc.off_l + i * NXP + a) << 16));
    } else if (off < d.offr) {  // R_ij -> Omega_i[co_j + a][co_j + c]
      const int pp = pair_at(off - d.offR, 0);
      if (pp < 0) return ILQG_OK;
      const int i = d.pair_i[pp], j = d.pair_j[pp], mj = d.udim[j], co = d.uoff[j];
      const int idx = off - d.offR - d.pair_Roff[pp], a = idx / mj, c = idx % mj;
      scat.push_back((unsigned)g | ((unsigned)(tc.off_Om + i * MM + (co + a) * MUP + co + c) << 16));
    } else {  // r_ij -> rho_i[co_j + c]
      const int pp = pair_at(off - d.offr, 1);
      if (pp < 0) return ILQG_OK;
      const int i = d.pair_i[pp], j = d.pair_j[pp], co = d.uoff[j];
      scat.push_back((unsigned)g | ((unsigned)(tc.off_rho + i * 8 + co + off - d.offr - d.pair_roff[pp]) << 16));
    }
  }
  for (int i = 0; i < N; i++)  // the template's state regulariser on the diagonal of Q_i
    for (int a = 0; a < n; a++)
      if (d.state_reg[i] != 0.f && item_at[d.offQ + (i * n + a) * n + a] < 0)
        qadd.push_back((unsigned)(NI + n + i) | ((unsigned)(i * NXP * L.LD + a * L.LD + a) << 16));
  std::sort(bs.begin(), bs.end(), [](const BEntry& x, const BEntry& y) { return x.c != y.c ? x.c < y.c : x.q < y.q; });
  for (int c = 0, e = 0; c <= M; c++) {
    while (e < (int)bs.size() && bs[e].c < c) e++;
    tc.bnz_start[c] = e;
  }
  tc.bnz_start[M] = (int)bs.size();
  for (const BEntry& b : bs) bnz.push_back((unsigned)b.q | ((unsigned)b.c << 8) | ((unsigned)b.item << 16));
  std::sort(bs.begin(), bs.end(), [](const BEntry& x, const BEntry& y) {
    return x.q != y.q ? x.q < y.q : (x.player != y.player ? x.player < y.player : x.c < y.c);
  });
  for (size_t e = 0; e < bs.size();) {
    size_t e1 = e;
    while (e1 < bs.size() && bs[e1].q == bs[e].q) e1++;
    brow_rows.push_back((unsigned)bs[e].q | ((unsigned)e << 8) | ((unsigned)(e1 - e) << 20));
    e = e1;
  }
  for (const BEntry& b : bs) brow.push_back((unsigned)b.c | ((unsigned)b.player << 8) | ((unsigned)b.item << 16));
  if (bs.size() >= 4096 || (int)scat.size() > 32 * sc_budget || (int)qadd.size() > 32 * qa_budget) return ILQG_OK;
  tc.nscat = (int)scat.size();
  tc.nqadd = (int)qadd.size();
  // the kernel walks exactly 32 x budget entries: pad with item 0 -> the scratch word after Z
  scat.resize((size_t)32 * sc_budget, (unsigned)(tc.off_Z + N * NXP * L.LD) << 16);
  qadd.resize((size_t)32 * qa_budget, (unsigned)(N * NXP * L.LD) << 16);
  std::vector<unsigned> words;
  auto append = [&](const std::vector<unsigned>& v) {
    const int at = (int)words.size();
    words.insert(words.end(), v.begin(), v.end());
    return at;
  };
  tc.scat = append(scat);
  tc.bnz = append(bnz); tc.nbnz = (int)bnz.size();
  tc.brow_rows = append(brow_rows); tc.nbrow_rows = (int)brow_rows.size();
  tc.brow = append(brow);
  tc.qadd = append(qadd);
  tc.total_words = (int)words.size();
  const size_t smem = sizeof(float) * (((size_t)tc.total_words + 3) / 4 * 4 + 2 * KTC_WARPS + (size_t)KTC_WARPS * tc.per_game);
  if (smem > 200 * 1024) return ILQG_OK;
  int rc;
  if ((rc = UploadVector(h, words, &tc.words)) != ILQG_OK) return rc;
  if ((rc = UploadVector(h, gk, &h->cp.gk)) != ILQG_OK) return rc;
  if ((rc = UploadVector(h, gk_start, &h->cp.gk_start)) != ILQG_OK) return rc;
  h->cp.NI = NI;
  h->cp.NIp = NIp;
  if ((rc = DevAlloc(h, &h->s.crec, (size_t)h->B * d.T * NIp)) != ILQG_OK) return rc;
  h->tc_nxp = NXP;
  h->tc_mup = MUP;
  h->cp_ok = true;
  if (h->use_compact) h->layout.compact_record_floats = NIp;
  return ILQG_OK;
}

// Discover the static update pattern of the records on the device and build the gather table
// (ilqg_records.cuh).  Leaves h->pat_ok = false when the shape does not fit (K_lq v1 is used).
int BuildRecordPattern(SubSolver* h) {
  const DevDesc& d = h->d;
  h->pat_ok = false;
  const int NR = d.N + 1;
  const int E = (MaxRoleEntries(d) + 3) & ~3;
  if (NR > 5 || d.rec >= 65536 || E <= 0) return ILQG_OK;
  int *d_off = nullptr, *d_cnt = nullptr;
  int rc;
  if ((rc = DevAlloc(h, &d_off, (size_t)NR * E)) != ILQG_OK) return rc;
  if ((rc = DevAlloc(h, &d_cnt, (size_t)NR)) != ILQG_OK) return rc;
  k_record_pattern<<<1, 32, 0, h->stream>>>(h->d, d_off, d_cnt, E);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  std::vector<int> off((size_t)NR * E), cnt(NR);
  CUDA_TRY(cudaMemcpyAsync(off.data(), d_off, off.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  std::vector<GatherItem> items;
  std::vector<unsigned short> idx;
  for (int r = 0; r < NR; r++) {
    if (cnt[r] > E) return ILQG_OK;  // bound was wrong: fall back
    // distinct offsets of this role in first-appearance order, entries kept in emission order
    std::vector<int> seen;
    for (int e = 0; e < cnt[r]; e++) {
      const int o = off[(size_t)r * E + e];
      if (o < 0 || o >= d.rec) return ILQG_ERR_INVALID_ARGUMENT;
      if (std::find(seen.begin(), seen.end(), o) == seen.end()) seen.push_back(o);
    }
    for (int o : seen) {
      GatherItem it;
      std::memset(&it, 0, sizeof(it));
      it.off = o;
      it.role = r;
      it.start = (int)idx.size();
      for (int e = 0; e < cnt[r]; e++)
        if (off[(size_t)r * E + e] == o) {
          idx.push_back((unsigned short)e);
          if (it.count < 4) it.e[it.count] = (unsigned short)e;
          it.count++;
        }
      items.push_back(it);
    }
  }
  // record template: A = I, Q_i = state_reg I, R_p = control_reg I (owner of the pair)
  std::vector<float> tmpl(d.rec, 0.f);
  for (int a = 0; a < d.n; a++) {
    tmpl[d.offA + a * d.n + a] = 1.f;
    for (int i = 0; i < d.N; i++) tmpl[d.offQ + (i * d.n + a) * d.n + a] = d.state_reg[i];
  }
  for (int p = 0; p < d.num_pairs; p++) {
    const int mj = d.udim[d.pair_j[p]];
    for (int a = 0; a < mj; a++) tmpl[d.offR + d.pair_Roff[p] + a * mj + a] = d.control_reg[d.pair_i[p]];
  }
  for (GatherItem& it : items) it.base = tmpl[it.off];
  // (K_lq v3 stages NR dense records per block: too much shared memory for the n = 24 problems, which
  // then produce dense records by expanding compact ones, EnsureDense)
  h->v3_ok = klq_smem_bytes(d.n, d.M, d.N, E, d.rec, (int)items.size(), (int)idx.size()) <= 110 * 1024;
  GatherItem* d_items = nullptr;
  unsigned short* d_idx = nullptr;
  float* d_tmpl = nullptr;
  if ((rc = DevAlloc(h, &d_items, items.size())) != ILQG_OK) return rc;
  if ((rc = DevAlloc(h, &d_idx, idx.size())) != ILQG_OK) return rc;
  if ((rc = DevAlloc(h, &d_tmpl, tmpl.size())) != ILQG_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(GatherItem), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(d_tmpl, tmpl.data(), tmpl.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->pat.items = d_items;
  h->pat.idx = d_idx;
  h->pat.tmpl = d_tmpl;
  h->pat.num_items = (int)items.size();
  h->pat.num_idx = (int)idx.size();
  h->pat.E = E;
  h->pat_ok = true;
  return BuildCompactTables(h, items);
}

int LaunchLqRecords(SubSolver* h, int only_running, Sel sel = Sel{SEL_ALL, nullptr, nullptr}) {
  const DevDesc& d = h->d;
  if (h->pat_ok && h->cp_ok && h->use_compact && !h->open_loop) {
    const size_t smem4 = klq4_smem_bytes(d.n, d.M, d.N, h->pat.E, h->cp.NIp, h->pat.num_items, h->pat.num_idx);
    auto kernel = h->classic ? k_linearize_quadraticize_v4<false> : k_linearize_quadraticize_v4<true>;
    int rc4 = SetSmem(kernel, smem4);
    if (rc4 != ILQG_OK) return rc4;
    const long long recs = (long long)h->B * d.T;
    ProfScope prof(h, 0);
    kernel<<<(int)((recs + 31) / 32), (d.N + 1) * 32, smem4, h->stream>>>(h->d, h->s, h->pat, h->cp, only_running, sel,
                                                                        h->p.linesearch);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->compact_valid = true;
    h->dense_valid = false;
    return ILQG_OK;
  }
  if (h->pat_ok && h->cp_ok && (!h->v3_ok || !h->p.linesearch)) {  // (v3 has no keep-the-quadraticization mode)
    // dense records wanted (open-loop solver, ILQG_RECORDS=dense) but K_lq v3 does not fit: v4, then expand
    const size_t smem4 = klq4_smem_bytes(d.n, d.M, d.N, h->pat.E, h->cp.NIp, h->pat.num_items, h->pat.num_idx);
    auto kernel = h->classic ? k_linearize_quadraticize_v4<false> : k_linearize_quadraticize_v4<true>;
    int rc4 = SetSmem(kernel, smem4);
    if (rc4 != ILQG_OK) return rc4;
    const long long recs = (long long)h->B * d.T;
    {
      ProfScope prof(h, 0);
      kernel<<<(int)((recs + 31) / 32), (d.N + 1) * 32, smem4, h->stream>>>(h->d, h->s, h->pat, h->cp, only_running, sel,
                                                                          h->p.linesearch);
      h->launches++;
      CUDA_TRY(cudaGetLastError());
    }
    h->compact_valid = true;
    h->dense_valid = false;
    return EnsureDense(h);
  }
  {
    int rca = EnsureDense(h);  // partial launches (only_running / lists) keep the other games' records
    if (rca != ILQG_OK) return rca;
    h->dense_valid = true;
    h->compact_valid = false;
  }
  if (h->pat_ok && h->v3_ok) {
    const size_t smem3 = klq_smem_bytes(d.n, d.M, d.N, h->pat.E, d.rec, h->pat.num_items, h->pat.num_idx);
    int rc3 = SetSmem(k_linearize_quadraticize_v3, smem3);
    if (rc3 != ILQG_OK) return rc3;
    const long long recs = (long long)h->B * d.T;
    ProfScope prof(h, 0);
    k_linearize_quadraticize_v3<<<(int)((recs + 31) / 32), (d.N + 1) * 32, smem3, h->stream>>>(h->d, h->s, h->pat, only_running, sel);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return ILQG_OK;
  }
  return ILQG_ERR_UNSUPPORTED;  // no record pattern (never for a descriptor inside the envelope)
}

// rollout + merit of one linesearch window as two kernels (ilqg_linesearch.cuh, "Split evaluation")
int LaunchLsSplit(SubSolver* h, int mode, int blocks, int q_offset) {
  const DevDesc& d = h->d;
  const int S = d.num_subsystems;
  const size_t smem_r = sizeof(float) * (size_t)ls_rollout_smem_floats(d.n, S);
  const size_t smem_m = sizeof(float) * (size_t)ls_merit_smem_floats(d.n, d.M, d.N);
  int rc = ILQG_ERR_UNSUPPORTED;
  int nuq = 2;
  bool classic = true;  // only the headline examples' subsystem kinds: the lean rollout instance
  for (int k = 0; k < S; k++) {
    nuq = std::max(nuq, d.sub[k].nu);
    classic = classic && (d.sub[k].kind == ILQG_DYN_CAR6D || d.sub[k].kind == ILQG_DYN_UNICYCLE4D || d.sub[k].kind == ILQG_DYN_AIR3D);
  }
  classic = classic && nuq <= 2;
  // the two rollout kernels of this descriptor (ilqg_linesearch.cuh: LsPick)
  using SpKernel = void (*)(const DevDesc, const DevParams, Slab, LsScratch, int, int, int, int, LsPick);
  SpKernel k_sp = nullptr, k_lanes = nullptr;
  if (classic && d.n == 16) k_sp = k_ls_rollout_sp<2, false, 4>;
  else if (classic && d.n == 24) k_sp = k_ls_rollout_sp<2, false, 6>;
  else if (classic) k_sp = k_ls_rollout_sp<2, false, 0>;
  else if (nuq <= 2) k_sp = k_ls_rollout_sp<2, true, 0>;
  else k_sp = k_ls_rollout_sp<4, true, 0>;
#define LS_ROLL(SS)                                                                                          \
  case SS:                                                                                                   \
    k_lanes = classic ? k_ls_rollout<SS, 2, false> : nuq <= 2 ? k_ls_rollout<SS, 2, true> : k_ls_rollout<SS, 4, true>; \
    break;
  switch (S) {
    LS_ROLL(1) LS_ROLL(2) LS_ROLL(3) LS_ROLL(4)
    default: return ILQG_ERR_UNSUPPORTED;
  }
#undef LS_ROLL
  if ((rc = SetSmem(k_lanes, smem_r)) != ILQG_OK) return rc;
  if (h->cap_sp <= 0) {  // items one resident wave of each kernel holds on this device
    int occ_sp = 0, occ_lanes = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sp, k_sp, S * 32, 0));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lanes, k_lanes, S * 32, smem_r));
    h->cap_sp = std::max(1, occ_sp) * h->sm_count * RSP_ITEMS;
    h->cap_lanes = std::max(1, occ_lanes) * h->sm_count * 32;
  }
  // ILQG_ROLLOUT = sp | lanes forces one kernel; auto (default): the first window's size is known here, a
  // continued window's only on the device -- both kernels are launched and one of them returns at once
  const long long nblk = ((long long)blocks * h->ls.lpw + RSP_ITEMS - 1) / RSP_ITEMS;
  const int grid_sp = (int)std::min<long long>(nblk, (long long)h->sm_count * 16);
  // Measured (profiles/r02_summary.md): with several subsystem warps per block the lane-per-item kernel loses at
  // every window size (C3: 71 vs 80 k instance-iterations/s); with one (Air3D: a block is a single
  // warp, so resident blocks per SM, not registers, cap the stage-parallel kernel) it wins on the large
  // continued windows (C4: 2.23 vs 1.79 M).  auto therefore only considers it when S = 1.
  int run_sp = h->rollout_mode == 1 || (h->rollout_mode == 0 && S > 1) ? 1 : h->rollout_mode == 2 ? 0 : -1;
  int run_lanes = run_sp < 0 ? -1 : 1 - run_sp;
  if (run_sp < 0 && mode != LS_MODE_QUEUED) {
    run_lanes = ls_prefers_lanes((long long)blocks * h->ls.lpw, h->cap_sp, h->cap_lanes) ? 1 : 0;
    run_sp = 1 - run_lanes;
  }
  if (run_sp != 0) {
    const LsPick pick{run_sp < 0 ? 1 : 0, h->cap_sp, h->cap_lanes};
    k_sp<<<grid_sp, S * 32, 0, h->stream>>>(h->d, h->p, h->s, h->ls, mode, h->ls_cur, q_offset, blocks, pick);
  }
  if (run_lanes != 0) {
    const LsPick pick{run_lanes < 0 ? 2 : 0, h->cap_sp, h->cap_lanes};
    k_lanes<<<blocks, S * 32, smem_r, h->stream>>>(h->d, h->p, h->s, h->ls, mode, h->ls_cur, q_offset, blocks, pick);
    h->launches += run_sp != 0 ? 1 : 0;
  }
  Stamp(h, 3 + mode * 10);  // rollout done
  auto merit = h->classic ? k_ls_merit<false> : k_ls_merit<true>;
  if ((rc = SetSmem(merit, smem_m)) != ILQG_OK) return rc;
  // the merit kernels stride over the item blocks that hold work (known on the device only):
  // grids sized for the machine, a few resident blocks per SM
  const int chunks = (d.T + KLS_MERIT_CHUNK - 1) / KLS_MERIT_CHUNK;
  const dim3 grid_m(std::min(blocks, std::max(1, h->sm_count * 8 / chunks)), chunks);
  merit<<<grid_m, d.N * 32, smem_m, h->stream>>>(h->d, h->p, h->s, h->ls, mode, h->ls_cur, q_offset, blocks);
  Stamp(h, 4 + mode * 10);  // merit done
  h->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

int LaunchLsEval(SubSolver* h, int mode, int blocks, int q_offset) {
  if (blocks <= 0) return ILQG_OK;
  if (blocks > h->ls_blocks_max) return ILQG_ERR_INVALID_ARGUMENT;
  ProfScope prof(h, mode == LS_MODE_FRESH ? 4 : mode == LS_MODE_QUEUED ? 5 : 7);
  return LaunchLsSplit(h, mode, blocks, q_offset);
}

// ILQSolver::ModifyLQStrategies for the whole batch (ilqg_linesearch.cuh): the first window for
// everyone, then the remaining candidates for the instances that rejected it, chunk by chunk.
// first window: candidate j = 0 for every running instance; rejections are appended to a queue
int LaunchLinesearchFresh(SubSolver* h) {
  const int B = h->B;
  int rc;
  const int dec_blocks = (B + KDEC_WARPS - 1) / KDEC_WARPS;
  CUDA_TRY(cudaMemsetAsync(h->ls.counts, 0, sizeof(int), h->stream));
  h->ls_cur = 0;
  if ((rc = LaunchLsEval(h, LS_MODE_FRESH, h->ls.nA_blocks, 0)) != ILQG_OK) return rc;
  {
    ProfScope pd(h, 6);
    k_ls_decide<<<dec_blocks, KDEC_WARPS * 32, 0, h->stream>>>(h->d, h->p, h->s, h->ls, LS_MODE_FRESH, 0, 0, 0);
  }
  Stamp(h, 15);  // first-window decide done
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

// the remaining candidates of the queued instances, window after window
int LaunchLinesearchQueued(SubSolver* h) {
  const int B = h->B;
  int rc;
  if (h->p.linesearch && h->ls.JB > 0) {
    const int cap = h->ls.cap;
    const int widest = h->ls.JB;
    // Tier after tier (ilqg_linesearch.cuh): a tier evaluates its candidates for every game still in
    // the queue; games whose tier ended without an accept AND whose last rollout was not "absorbed"
    // are handed on to the next tier's queue.  Launches over an empty queue exit at once.
    for (int jb : h->ls_tiers) {
      const int dst = h->ls_cur == 1 ? 2 : 1;
      h->ls.JB = jb;  // (kernel arguments are copied at launch)
      const int qblocks = (int)(((long long)cap * jb + h->ls.lpw - 1) / h->ls.lpw);
      CUDA_TRY(cudaMemsetAsync(h->ls.counts + dst, 0, sizeof(int), h->stream));
      for (int q0 = 0; q0 < B; q0 += cap) {
        if ((rc = LaunchLsEval(h, LS_MODE_QUEUED, qblocks, q0)) != ILQG_OK) {
          h->ls.JB = widest;
          return rc;
        }
        {
          ProfScope pd(h, 6);
          k_ls_decide<<<(cap + KDEC_WARPS - 1) / KDEC_WARPS, KDEC_WARPS * 32, 0, h->stream>>>(
              h->d, h->p, h->s, h->ls, LS_MODE_QUEUED, h->ls_cur, q0, dst);
        }
        Stamp(h, 25);  // tier decide done
        h->launches++;
      }
      h->ls_cur = dst;
    }
    h->ls.JB = widest;
  }
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

// ILQSolver::ModifyLQStrategies for the whole batch (ilqg_linesearch.cuh)
int LaunchLinesearch(SubSolver* h) {
  int rc;
  ProfScope prof(h, 2);
  if ((rc = LaunchLinesearchFresh(h)) != ILQG_OK) return rc;
  return LaunchLinesearchQueued(h);
}

// One pass of the pipelined schedule (pass `it` of `n`).
int PipelinedStep(SubSolver* h, int it, int n) {
  int rc;
  cudaStream_t main = h->stream;
  if (it == 0) h->side_busy = false;
  const Sel all{SEL_ALL, nullptr, nullptr};
  const Sel sel = it > 0 ? Sel{SEL_MAIN, nullptr, nullptr} : all;
  const bool bwd_on_side = h->pipeline == 1;
  if (it == 0 && h->stagger_wait) CUDA_TRY(cudaStreamWaitEvent(main, h->stagger_wait, 0));
  Stamp(h, 0);  // pass begins
  if ((rc = LaunchLqRecords(h, 1, sel)) != ILQG_OK) return rc;
  Stamp(h, 1);  // K_lq done
  if (!bwd_on_side && h->side_busy) CUDA_TRY(cudaStreamWaitEvent(main, h->ev_side, 0));
  if ((rc = DispatchBackward(h, 1, false, bwd_on_side ? sel : all)) != ILQG_OK) return rc;
  Stamp(h, 2);  // K_bwd done
  if (it == 0 && h->stagger_signal) CUDA_TRY(cudaEventRecord(h->stagger_signal, main));
  if (bwd_on_side && h->side_busy) CUDA_TRY(cudaStreamWaitEvent(main, h->ev_side, 0));
  if ((rc = LaunchLinesearchFresh(h)) != ILQG_OK) return rc;
  const int q_first = 0;  // the queue the first window just filled (left intact by the tiers)
  CUDA_TRY(cudaEventRecord(h->ev_fresh, main));
  CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fresh, 0));
  h->stream = h->side;
  rc = LaunchLinesearchQueued(h);
  if (rc == ILQG_OK && it + 1 < n) {
    const Sel list{SEL_LIST, h->ls.pend[q_first], h->ls.counts + q_first};
    rc = LaunchLqRecords(h, 1, list);
    Stamp(h, 6);  // K_lq over the queue list done
    if (rc == ILQG_OK && h->pipeline == 1) rc = DispatchBackward(h, 1, false, list);
  }
  h->stream = main;
  if (rc != ILQG_OK) return rc;
  CUDA_TRY(cudaEventRecord(h->ev_side, h->side));
  h->side_busy = true;
  return ILQG_OK;
}

int PipelinedEnd(SubSolver* h) {
  if (h->side_busy) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_side, 0));
  h->side_busy = false;
  return ILQG_OK;
}

// n iterations with the open linesearches of pass i overlapped with K_lq / K_bwd of pass i + 1:
//   main: K_lq, K_bwd (everyone but the instances queued in pass i-1) | wait side | first window
//   side:                         wait first window | remaining windows, K_lq + K_bwd over the queue
// Every instance still completes exactly one iteration per pass; only the order in which the two
// groups are processed inside a pass changes.
int IteratePipelined(SubSolver* h, int n) {
  int rc;
  for (int it = 0; it < n; it++)
    if ((rc = PipelinedStep(h, it, n)) != ILQG_OK) return rc;
  return PipelinedEnd(h);
}

// the pipelined schedule needs the list-capable kernels (static K_lq, half-warp K_bwd) and a
// single queued window; per-kernel profiling wants every kernel alone on the device
bool CanPipeline(const SubSolver* h, int max_iters) {
  const bool hw = h->dims_key == 0 || h->dims_key == 1 || (h->cp_ok && h->use_compact);
  return h->pipeline && !h->open_loop && max_iters > 1 && h->pat_ok && hw && h->p.linesearch && !h->profiling;
}

int LaunchSolveBegin(SubSolver* h) {
  int rc;
  ProfScope prof(h, 3);
  CUDA_TRY(cudaMemsetAsync(h->ls.counts, 0, 3 * sizeof(int), h->stream));
  h->ls_cur = 0;
  rc = LaunchLsEval(h, LS_MODE_BEGIN, (h->B + h->ls.lpw - 1) / h->ls.lpw, 0);
  if (rc == ILQG_OK) {
    k_begin_finalize<<<(h->B + KDEC_WARPS - 1) / KDEC_WARPS, KDEC_WARPS * 32, 0, h->stream>>>(h->d, h->p, h->s, h->ls);
    h->launches++;
    if (cudaGetLastError() != cudaSuccess) rc = ILQG_ERR_CUDA;
  }
  return rc;
}

template <typename T>
int DevAlloc(SubSolver* h, T** out, size_t count, bool zero) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    std::fprintf(stderr, "[ilqg_b200] cudaMalloc(%zu) failed: %s\n", bytes, cudaGetErrorString(e));
    return ILQG_ERR_OUT_OF_MEMORY;
  }
  h->allocs.push_back(p);
  if (zero) CUDA_TRY(cudaMemsetAsync(p, 0, bytes, h->stream));
  *out = (T*)p;
  return ILQG_OK;
}

int Fill(SubSolver* h, float* p, float v, size_t count) {
  const int blocks = (int)std::min<size_t>((count + 255) / 256, 1024);
  k_fill<<<blocks, 256, 0, h->stream>>>(p, v, count);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

int EnsureStaging(SubSolver* h, size_t floats) {
  if (h->staging_floats >= floats) return ILQG_OK;
  if (h->staging) cudaFree(h->staging);
  h->staging = nullptr;
  h->staging_floats = 0;
  cudaError_t e = cudaMalloc((void**)&h->staging, floats * sizeof(float));
  if (e != cudaSuccess) return ILQG_ERR_OUT_OF_MEMORY;
  h->staging_floats = floats;
  return ILQG_OK;
}

struct Guard {  // select the handle's device for the duration of a call
  int prev;
  bool ok;
  explicit Guard(int dev) : prev(0), ok(true) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~Guard() {
    if (ok && prev != target) cudaSetDevice(prev);
  }
  int target;
};

#define ENTER(h)                              \
  if (!(h)) return ILQG_ERR_BAD_HANDLE;       \
  Guard _guard((h)->device);                  \
  if (!_guard.ok) return ILQG_ERR_CUDA

// download one parity-selected per-instance array through the staging buffer
int DownloadParity(SubSolver* h, float* const src[2], const int* sel, int flip, size_t per, void* dst,
                   size_t bytes) {
  const size_t total = per * (size_t)h->B;
  if (bytes != total * sizeof(float)) return ILQG_ERR_SIZE_MISMATCH;
  int rc = EnsureStaging(h, total);
  if (rc != ILQG_OK) return rc;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 4096);
  k_gather_parity<<<blocks, 256, 0, h->stream>>>(h->staging, src[0], src[1], sel, flip, per, h->B);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(dst, h->staging, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

int DownloadFlat(SubSolver* h, const void* src, size_t elem_bytes, size_t per, void* dst, size_t bytes) {
  if (bytes != per * (size_t)h->B * elem_bytes) return ILQG_ERR_SIZE_MISMATCH;
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

// one field of the LQ records, de-interleaved by the copy engine
int DownloadRecordField(SubSolver* h, int off, int floats, void* dst, size_t bytes) {
  const size_t rows = (size_t)h->B * h->d.T;
  if (bytes != rows * floats * sizeof(float)) return ILQG_ERR_SIZE_MISMATCH;
  if (floats == 0) return ILQG_OK;
  int rcd = EnsureDense(h);
  if (rcd != ILQG_OK) return rcd;
  CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)floats * sizeof(float), h->s.rec + off,
                             (size_t)h->d.rec * sizeof(float), (size_t)floats * sizeof(float), rows,
                             cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

int UploadRecordField(SubSolver* h, int off, int floats, const float* src) {
  const size_t rows = (size_t)h->B * h->d.T;
  if (floats == 0) return ILQG_OK;
  int rcd = EnsureDense(h);  // fields not uploaded keep what the records held
  if (rcd != ILQG_OK) return rcd;
  h->dense_valid = true;
  h->compact_valid = false;
  CUDA_TRY(cudaMemcpy2DAsync(h->s.rec + off, (size_t)h->d.rec * sizeof(float), src,
                             (size_t)floats * sizeof(float), (size_t)floats * sizeof(float), rows,
                             cudaMemcpyHostToDevice, h->stream));
  return ILQG_OK;
}

}  // namespace

typedef SubSolver* SubHandle;

namespace sub {

int ilqg_destroy(SubHandle h);




int ilqg_create(const ilqg_problem_desc* desc, const ilqg_solver_params* params, int batch, int device,
                SubHandle* out) {
  if (!desc || !params || !out || batch < 1) return ILQG_ERR_INVALID_ARGUMENT;
  SubSolver* h = new (std::nothrow) SubSolver();
  if (!h) return ILQG_ERR_OUT_OF_MEMORY;
  std::vector<int> lidx;
  int rc = BuildDeviceDesc(*desc, &h->d, &h->layout, &lidx);  // descriptor errors first
  if (rc != ILQG_OK) {
    delete h;
    return rc;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    delete h;
    return ILQG_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    delete h;
    return ILQG_ERR_INVALID_ARGUMENT;
  }
  if (cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || h->sm_count < 1)
    h->sm_count = 148;
  h->dims_key = -1;
  for (int k = 0; k < kNumDims; k++)
    if (kDims[k].n == h->d.n && kDims[k].M == h->d.M && kDims[k].N == h->d.N) h->dims_key = k;
  // the open-loop kernel and the tensor-core sweep over compact records take run-time dimensions;
  // the dense-record feedback kernels are instantiated per shape (checked after the pattern is built)
  if (const char* e = std::getenv("ILQG_RECORDS")) h->use_compact = std::strcmp(e, "dense") != 0;
  if (const char* e = std::getenv("ILQG_TC_BLOCKS")) h->tc_blocks_cap = std::max(0, std::atoi(e));
  h->host_desc = *desc;
  h->open_loop = params->open_loop != 0;
  h->B = batch;
  h->device = device;
  h->layout.batch = batch;
  h->staging = nullptr;
  h->staging_floats = 0;
  h->launches = 0;
  h->profiling = false;
  h->stream = nullptr;
  h->own_stream = nullptr;
  h->side = nullptr;
  h->ev_fresh = h->ev_side = nullptr;
  h->pipeline = 0;
  for (int k = 0; k < 8; k++) { h->prof_ms[k] = 0; h->prof_n[k] = 0; }
  DevParams& p = h->p;
  p.convergence_tolerance = params->convergence_tolerance;
  p.max_solver_iters = params->max_solver_iters;
  p.linesearch = params->linesearch;
  p.initial_alpha_scaling = params->initial_alpha_scaling;
  p.geometric_alpha_scaling = params->geometric_alpha_scaling;
  p.max_backtracking_steps = params->max_backtracking_steps;
  p.expected_decrease_fraction = params->expected_decrease_fraction;
  p.geometric_mu_scaling = params->geometric_mu_scaling;
  p.geometric_mu_downscaling = params->geometric_mu_downscaling;
  p.geometric_lambda_downscaling = params->geometric_lambda_downscaling;
  p.adaptive_regularization = params->adaptive_regularization;
  p.disable_convergence_exit = params->disable_convergence_exit;

  Guard guard(device);
  if (!guard.ok) {
    delete h;
    return ILQG_ERR_CUDA;
  }
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return ILQG_ERR_CUDA;
  }
  h->stream = h->own_stream;
  const size_t B = batch, T = h->d.T, n = h->d.n, M = h->d.M, N = h->d.N;
  Slab& s = h->s;
  std::memset(&s, 0, sizeof(s));
  s.B = batch;
  auto fail = [&](int code) {
    ilqg_destroy(h);
    return code;
  };
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  // the side stream carries the few open linesearches, which are latency chains: its blocks must
  // not queue behind the thousands of K_lq blocks of the main stream
  const bool side_high = !std::getenv("ILQG_SIDE_PRIORITY") || std::atoi(std::getenv("ILQG_SIDE_PRIORITY")) != 0;
  if (cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, side_high ? prio_hi : prio_lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fresh, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->stagger_signal, cudaEventDisableTiming) != cudaSuccess)
    return fail(ILQG_ERR_CUDA);
#define ALLOC(ptr, count)                                          \
  if ((rc = DevAlloc(h, &(ptr), (count))) != ILQG_OK) return fail(rc)
  ALLOC(s.x0, B * n);
  ALLOC(s.lq_x0, B * n);
  for (int k = 0; k < 2; k++) {
    ALLOC(s.op_xs[k], B * T * n);
    ALLOC(s.op_us[k], B * T * M);
    ALLOC(s.st_P[k], B * T * M * n);
    ALLOC(s.st_a[k], B * T * M);
  }
  ALLOC(s.prob_xs, B * T * n);
  ALLOC(s.prob_us, B * T * M);
  ALLOC(s.prob_P, B * T * M * n);
  ALLOC(s.prob_a, B * T * M);
  ALLOC(s.dxs, B * T * n);
  if (h->open_loop)
    ALLOC(h->ol_scratch, B * T * (size_t)ol_layout(h->d.n, h->d.M, h->d.N, h->d.rec - h->d.offl).srec);
  ALLOC(s.lambdas, B * (size_t)h->d.num_constraints * T);
  ALLOC(s.mu, B);
  ALLOC(s.last_merit, B);
  ALLOC(s.expected_decrease, B);
  ALLOC(s.step, B);
  ALLOC(s.total_costs, B * N);
  ALLOC(s.max_con_err, B);
  ALLOC(s.status, B);
  ALLOC(s.iters, B);
  ALLOC(s.al_state, B);
  ALLOC(s.al_iterates, B);
  ALLOC(s.al_success, B);
  ALLOC(s.al_flags, B);
  ALLOC(s.queued_flag, B);
  ALLOC(s.backtracks, B);
  ALLOC(s.te_quad, B * N);
  ALLOC(s.te_new, B * N);
  ALLOC(s.op_cur, B);
  ALLOC(s.st_cur, B);
  int* lidx_dev = nullptr;
  ALLOC(lidx_dev, T);
  {
    // linesearch scratch (ilqg_linesearch.cuh): a fresh linesearch evaluates JA candidates per
    // pass, a continued one JB
    LsScratch& ls = h->ls;
    std::memset(&ls, 0, sizeof(ls));
    const int max_bt = std::max(1, params->max_backtracking_steps);
    int JA = 1;
    if (const char* e = std::getenv("ILQG_LS_JA")) JA = std::max(1, std::atoi(e));  // tuning knobs
    if (!params->linesearch) JA = 1;
    JA = std::min(JA, max_bt);
    // continued linesearches go on in windows of JB candidates; by default one window covers every
    // remaining candidate (measured best: a second window costs a full rollout latency, and the
    // "absorbed" shortcut of k_ls_decide rarely triggers before j ~ 40)
    // Tiers of the continued linesearch (round 2): by default candidates JA .. JA + 38, then the rest.
    // Measured with the oracle on the three headline problems: a linesearch that rejects j = 0
    // either accepts by j ~ 36 or fails, and by then alpha * s0 * rho^j has vanished in rounding
    // (the "absorbed" shortcut of k_ls_decide), so the second tier almost never has work -- the
    // first tier is 2.5 x narrower than one window over all candidates.  ILQG_LS_TIERS="7,32"
    // gives tiers of 7, 32 and the remainder; ILQG_LS_JB=n is round 1's equal windows of n.
    std::vector<int> tiers;
    {
      const int remaining = max_bt - JA;
      const char* e = std::getenv("ILQG_LS_TIERS");
      std::string spec = e ? e : "39";
      if (const char* jb = std::getenv("ILQG_LS_JB")) {
        const int w = std::max(1, std::atoi(jb));
        spec.clear();
        for (int done = w; done < remaining; done += w) spec += std::to_string(w) + ",";
      }
      int used = 0;
      for (size_t pos = 0; pos < spec.size() && used < remaining;) {
        const int w = std::atoi(spec.c_str() + pos);
        if (w > 0) {
          tiers.push_back(std::min(w, remaining - used));
          used += tiers.back();
        }
        const size_t comma = spec.find(',', pos);
        if (comma == std::string::npos) break;
        pos = comma + 1;
      }
      if (used < remaining) tiers.push_back(remaining - used);
    }
    h->ls_tiers = tiers;
    int JB = 0;
    for (int w : tiers) JB = std::max(JB, w);
    // queue slots one continued-window launch holds candidate trajectories for: the whole batch
    // (one launch, no empty second chunk) while that scratch stays under 8 GiB, else chunks
    const size_t per_slot = (size_t)std::max(1, JB) * T * (n + M) * sizeof(float);
    int cap = (int)std::max<size_t>(1, std::min<size_t>(B, std::max<size_t>((B + 1) / 2, ((size_t)8 << 30) / per_slot)));
    if (const char* e = std::getenv("ILQG_LS_CAP")) cap = std::max(1, std::min<int>((int)B, std::atoi(e)));
    // measured with the split linesearch: 2 with a high-priority side stream > 0 > 2 > 1
    // (profiles/r01_schedule_experiments.md)
    h->pipeline = 2;
    if (const char* e = std::getenv("ILQG_PIPELINE")) h->pipeline = std::atoi(e);
    if (const char* e = std::getenv("ILQG_ROLLOUT")) h->rollout_mode = !std::strcmp(e, "lanes") ? 2 : !std::strcmp(e, "sp") ? 1 : 0;
    if (const char* e = std::getenv("ILQG_TRACE")) {
      h->trace_path = e;
      h->trace_cap = 1 << 16;
      if ((rc = DevAlloc(h, &h->trace, (size_t)h->trace_cap, true)) != ILQG_OK) return fail(rc);
    }
    ls.JA = JA;
    ls.JB = JB;
    ls.cap = cap;
    int lpw = 32;
    if (const char* e = std::getenv("ILQG_LS_LPW")) lpw = std::atoi(e);
    if (lpw != 8 && lpw != 16 && lpw != 32 && lpw != 4) lpw = 32;
    ls.lpw = lpw;
    ls.nA_blocks = (int)((B * JA + lpw - 1) / lpw);
    const size_t blocks_max = std::max<size_t>(std::max<size_t>(ls.nA_blocks, (B + lpw - 1) / lpw),
                                               ((size_t)cap * JB + lpw - 1) / lpw);
    h->ls_blocks_max = (int)blocks_max;
    h->ls_cur = 0;
#define ALLOCNZ(ptr, count)                                               \
  if ((rc = DevAlloc(h, &(ptr), (count), false)) != ILQG_OK) return fail(rc)
    ALLOCNZ(ls.traj_xs, blocks_max * lpw * T * n);
    ALLOCNZ(ls.traj_us, blocks_max * lpw * T * M);
    ALLOCNZ(ls.terms, blocks_max * T * 2 * N * 32);
    ALLOCNZ(ls.vals, blocks_max * T * N * 32);
    ALLOCNZ(ls.absorbed, blocks_max * lpw);
#undef ALLOCNZ
    ALLOC(ls.pend[0], B);
    ALLOC(ls.pend[1], B);
    ALLOC(ls.pend[2], B);
    ALLOC(ls.slot, B);
    ALLOC(ls.counts, 3);
    ALLOC(s.ls_next_j, B);
  }
#undef ALLOC
  s.lambda_index = lidx_dev;
  if (cudaMemcpyAsync(lidx_dev, lidx.data(), T * sizeof(int), cudaMemcpyHostToDevice, h->stream) !=
      cudaSuccess)
    return fail(ILQG_ERR_CUDA);
  // kDefaultMu = 10, last merit / expected decrease = +inf (types.h:125-126, ilq_solver.h:73-74)
  if ((rc = Fill(h, s.mu, 10.0f, B)) != ILQG_OK) return fail(rc);
  if ((rc = Fill(h, s.last_merit, INFINITY, B)) != ILQG_OK) return fail(rc);
  if ((rc = Fill(h, s.expected_decrease, INFINITY, B)) != ILQG_OK) return fail(rc);
  if ((rc = Fill(h, s.max_con_err, INFINITY, B)) != ILQG_OK) return fail(rc);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail(ILQG_ERR_CUDA);
  h->classic = true;
  for (int k = 0; k < h->d.num_subsystems; k++) {
    const int kind = h->d.sub[k].kind;
    h->classic = h->classic && (kind == ILQG_DYN_CAR6D || kind == ILQG_DYN_UNICYCLE4D || kind == ILQG_DYN_AIR3D);
  }
  for (int c = 0; c < h->d.num_costs; c++) {
    const DevCost& cd = h->d.cost[c];
    h->classic = h->classic && cd.kind <= ILQG_CONSTRAINT_SINGLE_DIMENSION && cd.active_from == 0.0 && cd.group == 0;
  }
  if ((rc = BuildRecordPattern(h)) != ILQG_OK) return fail(rc);
  *out = h;
  return ILQG_OK;
}

int ilqg_destroy(SubHandle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  Guard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->trace && h->trace_n > 0) {
    std::vector<unsigned long long> t(h->trace_n);
    cudaDeviceSynchronize();
    if (cudaMemcpy(t.data(), h->trace, sizeof(unsigned long long) * h->trace_n, cudaMemcpyDeviceToHost) == cudaSuccess)
      if (FILE* f = std::fopen(h->trace_path.c_str(), "w")) {
        for (int i = 0; i < h->trace_n; i++) std::fprintf(f, "%d %llu\n", h->trace_tags[i], t[i] - t[0]);
        std::fclose(f);
      }
  }
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fresh) cudaEventDestroy(h->ev_fresh);
  if (h->ev_side) cudaEventDestroy(h->ev_side);
  if (h->stagger_signal) cudaEventDestroy(h->stagger_signal);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  for (auto& sm : h->samples) {
    cudaEventDestroy(sm.a);
    cudaEventDestroy(sm.b);
  }
  for (void* p : h->allocs) cudaFree(p);
  if (h->staging) cudaFree(h->staging);
  delete h;
  return ILQG_OK;
}

int ilqg_get_layout(SubHandle h, ilqg_layout* out) {
  if (!h || !out) return ILQG_ERR_BAD_HANDLE;
  *out = h->layout;
  return ILQG_OK;
}

int ilqg_upload_x0(SubHandle h, const float* x0, size_t bytes) {
  ENTER(h);
  if (!x0) return ILQG_ERR_INVALID_ARGUMENT;
  if (bytes != sizeof(float) * (size_t)h->B * h->d.n) return ILQG_ERR_SIZE_MISMATCH;
  CUDA_TRY(cudaMemcpyAsync(h->s.x0, x0, bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));  // the host buffer may be pageable / reused
  return ILQG_OK;
}

int ilqg_upload_warmstart(SubHandle h, const float* xs, const float* us, const float* Ps,
                          const float* alphas) {
  ENTER(h);
  NvtxRange nvtx("ilqg/upload warm start");
  const size_t B = h->B, T = h->d.T, n = h->d.n, M = h->d.M;
  struct Item {
    float* dst;
    const float* src;
    size_t count;
  } items[4] = {{h->s.prob_xs, xs, B * T * n}, {h->s.prob_us, us, B * T * M},
                {h->s.prob_P, Ps, B * T * M * n}, {h->s.prob_a, alphas, B * T * M}};
  for (const Item& it : items) {
    if (it.src)
      CUDA_TRY(cudaMemcpyAsync(it.dst, it.src, it.count * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    else
      CUDA_TRY(cudaMemsetAsync(it.dst, 0, it.count * sizeof(float), h->stream));
  }
  k_prob_to_working<<<h->B, 256, 0, h->stream>>>(h->d, h->s);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

int ilqg_upload(SubHandle h, int what, const void* src, size_t bytes) {
  ENTER(h);
  if (!src) return ILQG_ERR_INVALID_ARGUMENT;
  const size_t B = h->B, T = h->d.T, N = h->d.N;
  void* dst = nullptr;
  size_t want = 0;
  switch (what) {
    case ILQG_LAMBDAS: dst = h->s.lambdas; want = sizeof(float) * B * h->d.num_constraints * T; break;
    case ILQG_MU: dst = h->s.mu; want = sizeof(float) * B; break;
    case ILQG_MERIT: dst = h->s.last_merit; want = sizeof(float) * B; break;
    case ILQG_X0: dst = h->s.x0; want = sizeof(float) * B * h->d.n; break;
    case ILQG_LQ_X0: dst = h->s.lq_x0; want = sizeof(float) * B * h->d.n; break;
    case ILQG_TIME_OF_EXTREME: {
      if (bytes != sizeof(int) * B * N) return ILQG_ERR_SIZE_MISMATCH;
      CUDA_TRY(cudaMemcpyAsync(h->s.te_new, src, bytes, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(cudaMemcpyAsync(h->s.te_quad, src, bytes, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      return ILQG_OK;
    }
    default: return ILQG_ERR_INVALID_ARGUMENT;
  }
  if (bytes != want) return ILQG_ERR_SIZE_MISMATCH;
  if (want) CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

int ilqg_upload_lq(SubHandle h, const float* A, const float* Bs, const float* Q, const float* l,
                   const float* R, const float* r) {
  ENTER(h);
  if (!A || !Bs || !Q || !l || !R || !r) return ILQG_ERR_INVALID_ARGUMENT;
  const DevDesc& d = h->d;
  int rc;
  if ((rc = UploadRecordField(h, d.offA, d.n * d.n, A)) != ILQG_OK) return rc;
  if ((rc = UploadRecordField(h, d.offB, d.n * d.M, Bs)) != ILQG_OK) return rc;
  if ((rc = UploadRecordField(h, d.offQ, d.N * d.n * d.n, Q)) != ILQG_OK) return rc;
  if ((rc = UploadRecordField(h, d.offl, d.N * d.n, l)) != ILQG_OK) return rc;
  if ((rc = UploadRecordField(h, d.offR, d.R_floats, R)) != ILQG_OK) return rc;
  if ((rc = UploadRecordField(h, d.offr, d.r_floats, r)) != ILQG_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

int ilqg_solve_begin(SubHandle h) {
  ENTER(h);
  return LaunchSolveBegin(h);
}

int ilqg_linearize_quadraticize(SubHandle h) {
  ENTER(h);
  return LaunchLqRecords(h, 0);
}

int ilqg_lq_backward(SubHandle h) {
  ENTER(h);
  return DispatchBackward(h, 0, true);
}

int ilqg_linesearch(SubHandle h) {
  ENTER(h);
  return LaunchLinesearch(h);
}

int ilqg_iterate(SubHandle h, int max_iters, int* iters_done) {
  ENTER(h);
  int rc;
  if (CanPipeline(h, max_iters)) {
    if ((rc = IteratePipelined(h, max_iters)) != ILQG_OK) return rc;
  } else {
    for (int it = 0; it < max_iters; it++) {
      Stamp(h, 0);
      if ((rc = LaunchLqRecords(h, 1)) != ILQG_OK) return rc;
      Stamp(h, 1);
      if ((rc = DispatchBackward(h, 1, false)) != ILQG_OK) return rc;
      Stamp(h, 2);
      if ((rc = LaunchLinesearch(h)) != ILQG_OK) return rc;
    }
  }
  if (iters_done) {
    std::vector<int> iters(h->B);
    CUDA_TRY(cudaMemcpyAsync(iters.data(), h->s.iters, sizeof(int) * h->B, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *iters_done = *std::max_element(iters.begin(), iters.end());
  }
  return ILQG_OK;
}

int ilqg_count_running(SubHandle h, int* running) {
  ENTER(h);
  if (!running) return ILQG_ERR_INVALID_ARGUMENT;
  std::vector<int> st(h->B);
  CUDA_TRY(cudaMemcpyAsync(st.data(), h->s.status, sizeof(int) * h->B, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *running = (int)std::count(st.begin(), st.end(), (int)ILQG_STATUS_RUNNING);
  return ILQG_OK;
}

int ilqg_al_update(SubHandle h) {
  ENTER(h);
  NvtxRange nvtx("ilqg/AL multiplier sweep");
  k_al_update<<<h->B, ILQG_MAX_COSTS, 0, h->stream>>>(h->d, h->p, h->s, 0);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

int ilqg_overwrite_solution(SubHandle h, int only_successful) {
  ENTER(h);
  NvtxRange nvtx("ilqg/overwrite solution");
  k_overwrite_solution<<<h->B, 256, 0, h->stream>>>(h->d, h->s, only_successful, 0);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

// Problem::SetUpNextRecedingHorizon for this group's games (ilqg_receding.cuh); x_meas is host memory
int ilqg_setup_next_receding_horizon(SubHandle h, const float* x_meas, const RhTimes& times) {
  ENTER(h);
  NvtxRange nvtx("ilqg/receding horizon");
  const size_t floats = (size_t)h->B * h->d.n;
  int rc = EnsureStaging(h, floats);
  if (rc != ILQG_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->staging, x_meas, sizeof(float) * floats, cudaMemcpyHostToDevice, h->stream));
  k_receding_horizon<<<(h->B + KRH_WARPS - 1) / KRH_WARPS, KRH_WARPS * 32, 0, h->stream>>>(h->d, h->s, h->staging, times);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->stream));  // the host buffer may be pageable / reused
  return ILQG_OK;
}

int ilqg_integrate_plan(SubHandle h, const float* x_in, const IpTimes& times, float* x_out) {
  ENTER(h);
  NvtxRange nvtx("ilqg/integrate plan");
  const size_t floats = (size_t)h->B * h->d.n;
  int rc = EnsureStaging(h, floats);
  if (rc != ILQG_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->staging, x_in, sizeof(float) * floats, cudaMemcpyHostToDevice, h->stream));
  k_integrate_plan<<<(h->B + KRH_WARPS - 1) / KRH_WARPS, KRH_WARPS * 32, 0, h->stream>>>(h->d, h->s, h->staging, times);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(x_out, h->staging, sizeof(float) * floats, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));  // the host buffers may be pageable / reused
  return ILQG_OK;
}

int ilqg_al_post_solve(SubHandle h) {
  ENTER(h);
  NvtxRange nvtx("ilqg/AL post solve");
  k_al_post_solve<<<h->B, 128, 0, h->stream>>>(h->d, h->p, h->s, 0);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

int ilqg_al_begin(SubHandle h) {
  ENTER(h);
  NvtxRange nvtx("ilqg/AL begin");
  k_al_begin<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->s);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

// one AugmentedLagrangianSolver::Solve round; *active accumulates (device counter, read by the caller)
int ilqg_al_advance(SubHandle h, int first, int max_iterates, float tolerance, int* d_active) {
  ENTER(h);
  NvtxRange nvtx("ilqg/AL advance");
  k_al_account<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->d, h->s, first, max_iterates, tolerance, d_active);
  k_al_post_solve<<<h->B, 128, 0, h->stream>>>(h->d, h->p, h->s, AL_DO_DOWNSCALE);
  k_al_update<<<h->B, ILQG_MAX_COSTS, 0, h->stream>>>(h->d, h->p, h->s, AL_DO_UPDATE);
  k_overwrite_solution<<<h->B, 256, 0, h->stream>>>(h->d, h->s, 0, AL_DO_OVERWRITE);
  h->launches += 4;
  CUDA_TRY(cudaGetLastError());
  return ILQG_OK;
}

int ilqg_download(SubHandle h, int what, void* dst, size_t bytes) {
  ENTER(h);
  if (!dst) return ILQG_ERR_INVALID_ARGUMENT;
  const DevDesc& d = h->d;
  const Slab& s = h->s;
  const size_t T = d.T, n = d.n, M = d.M, N = d.N;
  switch (what) {
    case ILQG_XS: return DownloadParity(h, s.op_xs, s.op_cur, 0, T * n, dst, bytes);
    case ILQG_US: return DownloadParity(h, s.op_us, s.op_cur, 0, T * M, dst, bytes);
    case ILQG_PS: return DownloadParity(h, s.st_P, s.st_cur, 0, T * M * n, dst, bytes);
    case ILQG_ALPHAS: return DownloadParity(h, s.st_a, s.st_cur, 0, T * M, dst, bytes);
    case ILQG_LQ_PS: return DownloadParity(h, s.st_P, s.st_cur, 1, T * M * n, dst, bytes);
    case ILQG_WARM_XS: return DownloadFlat(h, s.prob_xs, 4, T * n, dst, bytes);
    case ILQG_WARM_US: return DownloadFlat(h, s.prob_us, 4, T * M, dst, bytes);
    case ILQG_WARM_PS: return DownloadFlat(h, s.prob_P, 4, T * M * n, dst, bytes);
    case ILQG_WARM_ALPHAS: return DownloadFlat(h, s.prob_a, 4, T * M, dst, bytes);
    case ILQG_LQ_ALPHAS: return DownloadParity(h, s.st_a, s.st_cur, 1, T * M, dst, bytes);
    case ILQG_LIN_A: return DownloadRecordField(h, d.offA, d.n * d.n, dst, bytes);
    case ILQG_LIN_B: return DownloadRecordField(h, d.offB, d.n * d.M, dst, bytes);
    case ILQG_QUAD_Q: return DownloadRecordField(h, d.offQ, d.N * d.n * d.n, dst, bytes);
    case ILQG_QUAD_L: return DownloadRecordField(h, d.offl, d.N * d.n, dst, bytes);
    case ILQG_QUAD_R: return DownloadRecordField(h, d.offR, d.R_floats, dst, bytes);
    case ILQG_QUAD_RGRAD: return DownloadRecordField(h, d.offr, d.r_floats, dst, bytes);
    case ILQG_DELTA_XS: return DownloadFlat(h, s.dxs, 4, T * n, dst, bytes);
    case ILQG_LAMBDAS: return DownloadFlat(h, s.lambdas, 4, (size_t)d.num_constraints * T, dst, bytes);
    case ILQG_TOTAL_COSTS: return DownloadFlat(h, s.total_costs, 4, N, dst, bytes);
    case ILQG_X0: return DownloadFlat(h, s.x0, 4, n, dst, bytes);
    case ILQG_LQ_X0: return DownloadFlat(h, s.lq_x0, 4, n, dst, bytes);
    case ILQG_MU: return DownloadFlat(h, s.mu, 4, 1, dst, bytes);
    case ILQG_MERIT: return DownloadFlat(h, s.last_merit, 4, 1, dst, bytes);
    case ILQG_EXPECTED_DECREASE: return DownloadFlat(h, s.expected_decrease, 4, 1, dst, bytes);
    case ILQG_STEP: return DownloadFlat(h, s.step, 4, 1, dst, bytes);
    case ILQG_MAX_CONSTRAINT_ERROR: return DownloadFlat(h, s.max_con_err, 4, 1, dst, bytes);
    case ILQG_STATUS: return DownloadFlat(h, s.status, 4, 1, dst, bytes);
    case ILQG_ITERS: return DownloadFlat(h, s.iters, 4, 1, dst, bytes);
    case ILQG_BACKTRACKS: return DownloadFlat(h, s.backtracks, 4, 1, dst, bytes);
    case ILQG_AL_SUCCESS: return DownloadFlat(h, s.al_success, 4, 1, dst, bytes);
    case ILQG_AL_ITERATES: return DownloadFlat(h, s.al_iterates, 4, 1, dst, bytes);
    case ILQG_AL_STATE: return DownloadFlat(h, s.al_state, 4, 1, dst, bytes);
    case ILQG_TIME_OF_EXTREME: return DownloadFlat(h, s.te_new, 4, N, dst, bytes);
  }
  return ILQG_ERR_INVALID_ARGUMENT;
}

int ilqg_synchronize(SubHandle h) {
  ENTER(h);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

int ilqg_reset(SubHandle h, int mask) {
  ENTER(h);
  NvtxRange nvtx("ilqg/reset");
  int rc;
  const size_t B = h->B, T = h->d.T, n = h->d.n, M = h->d.M;
  Slab& s = h->s;
  CUDA_TRY(cudaMemsetAsync(s.al_state, 0, sizeof(int) * B, h->stream));  // any reset ends an AL solve
  if (mask & ILQG_RESET_SOLVER) {
    if ((rc = Fill(h, s.last_merit, INFINITY, B)) != ILQG_OK) return rc;
    if ((rc = Fill(h, s.expected_decrease, INFINITY, B)) != ILQG_OK) return rc;
  }
  if (mask & (ILQG_RESET_MULTIPLIERS | ILQG_RESET_LAMBDAS))
    CUDA_TRY(cudaMemsetAsync(s.lambdas, 0, sizeof(float) * std::max<size_t>(B * h->d.num_constraints * T, 1), h->stream));
  if (mask & (ILQG_RESET_MULTIPLIERS | ILQG_RESET_MU)) {
    if ((rc = Fill(h, s.mu, 10.0f, B)) != ILQG_OK) return rc;
  }
  if (mask & ILQG_RESET_SOLUTION) {
    CUDA_TRY(cudaMemsetAsync(s.prob_xs, 0, sizeof(float) * B * T * n, h->stream));
    CUDA_TRY(cudaMemsetAsync(s.prob_us, 0, sizeof(float) * B * T * M, h->stream));
    CUDA_TRY(cudaMemsetAsync(s.prob_P, 0, sizeof(float) * B * T * M * n, h->stream));
    CUDA_TRY(cudaMemsetAsync(s.prob_a, 0, sizeof(float) * B * T * M, h->stream));
    k_prob_to_working<<<h->B, 256, 0, h->stream>>>(h->d, h->s);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return ILQG_OK;
}

int ilqg_set_stream(SubHandle h, void* cuda_stream) {
  ENTER(h);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return ILQG_OK;
}

static int DrainSamples(SubSolver* h) {
  for (auto& sm : h->samples) {
    CUDA_TRY(cudaEventSynchronize(sm.b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, sm.a, sm.b));
    h->prof_ms[sm.kind] += ms;
    h->prof_n[sm.kind] += 1;
    cudaEventDestroy(sm.a);
    cudaEventDestroy(sm.b);
  }
  h->samples.clear();
  return ILQG_OK;
}

int ilqg_profile(SubHandle h, int enable) {
  ENTER(h);
  int rc = DrainSamples(h);
  if (rc != ILQG_OK) return rc;
  if (enable)
    for (int k = 0; k < 8; k++) { h->prof_ms[k] = 0; h->prof_n[k] = 0; }
  h->profiling = enable != 0;
  return ILQG_OK;
}

int ilqg_profile_read(SubHandle h, int kernel, double* total_ms, long long* launches) {
  ENTER(h);
  if (kernel < 0 || kernel > 7 || !total_ms || !launches) return ILQG_ERR_INVALID_ARGUMENT;
  int rc = DrainSamples(h);
  if (rc != ILQG_OK) return rc;
  *total_ms = h->prof_ms[kernel];
  *launches = h->prof_n[kernel];
  return ILQG_OK;
}

int ilqg_kernel_launches(SubHandle h, long long* out) {
  if (!h || !out) return ILQG_ERR_BAD_HANDLE;
  *out = h->launches;
  return ILQG_OK;
}

}  // namespace sub


// ===========================================================================
// Public handle: G groups of instances on concurrent streams.
//
// Every kernel of the hot path is latency-bound at the benchmark batch (a sequential sweep over
// T = 100 timesteps with a few thousand instances leaves most issue slots empty), so the batch is
// split into groups whose kernel sequences run on separate CUDA streams: while one group's
// backward sweep waits on shared-memory latency another group's rollout or record assembly
// issues.  Instances are independent, so the split changes no result.  ILQG_GROUPS overrides the
// group count (1 = a single stream, used for per-kernel profiling).
// ===========================================================================
struct ilqg_solver {
  std::vector<SubSolver*> subs;
  std::vector<int> first;  // first instance of each group
  std::vector<cudaEvent_t> done;
  cudaEvent_t fork;
  cudaStream_t stream, own_stream;
  ilqg_layout layout;
  int B, device;
  double op_t0 = 0.0;  // OperatingPoint::t0 of the Problems (shared by the batch)
  // AugmentedLagrangianSolver::Solve driver state (ilqg_al_begin / ilqg_al_advance)
  int al_max_iterates = 0;
  float al_tolerance = 0.f;
  bool al_first = true;
  std::vector<int*> al_active_dev;  // one counter per group, on the group's device
};

namespace {

int Fork(ilqg_solver* h) {
  if (h->subs.size() == 1 && h->subs[0]->stream == h->stream) return ILQG_OK;
  CUDA_TRY(cudaEventRecord(h->fork, h->stream));
  for (SubSolver* g : h->subs) CUDA_TRY(cudaStreamWaitEvent(g->stream, h->fork, 0));
  return ILQG_OK;
}

int Join(ilqg_solver* h) {
  if (h->subs.size() == 1 && h->subs[0]->stream == h->stream) return ILQG_OK;
  for (size_t k = 0; k < h->subs.size(); k++) {
    {
      Guard gd(h->subs[k]->device);  // (a multi-device handle: the event lives on the group's device)
      CUDA_TRY(cudaEventRecord(h->done[k], h->subs[k]->stream));
    }
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->done[k], 0));
  }
  return ILQG_OK;
}

template <typename F>
int ForEach(ilqg_solver* h, F&& f) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  Guard guard(h->device);
  if (!guard.ok) return ILQG_ERR_CUDA;
  int rc = Fork(h);
  if (rc != ILQG_OK) return rc;
  for (size_t k = 0; k < h->subs.size(); k++)
    if ((rc = f(h->subs[k], h->first[k])) != ILQG_OK) return rc;
  return Join(h);
}

}  // namespace

extern "C" {

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

const char* ilqg_strerror(int code) {
  switch (code) {
    case ILQG_OK: return "ok";
    case ILQG_ERR_INVALID_ARGUMENT: return "invalid argument";
    case ILQG_ERR_UNSUPPORTED: return "unsupported descriptor (no kernel compiled for these dimensions/kinds)";
    case ILQG_ERR_CUDA: return "CUDA error";
    case ILQG_ERR_NO_DEVICE: return "no CUDA device";
    case ILQG_ERR_OUT_OF_MEMORY: return "out of memory";
    case ILQG_ERR_BAD_HANDLE: return "bad handle";
    case ILQG_ERR_SIZE_MISMATCH: return "host buffer size mismatch";
  }
  return "unknown error";
}

// groups[k] = (device, instances): the public handle is a list of per-device / per-stream groups over
// contiguous slices of the batch (independent games, no hot-path exchange)
static int CreateGroups(const ilqg_problem_desc* desc, const ilqg_solver_params* params, int batch,
                        const std::vector<std::pair<int, int>>& groups, ilqg_handle* out) {
  ilqg_solver* h = new (std::nothrow) ilqg_solver();
  if (!h) return ILQG_ERR_OUT_OF_MEMORY;
  h->B = batch;
  h->device = groups[0].first;
  h->stream = nullptr;
  h->own_stream = nullptr;
  h->fork = nullptr;
  int rc = ILQG_OK, lo = 0;
  for (size_t g = 0; g < groups.size() && rc == ILQG_OK; g++) {
    SubSolver* sh = nullptr;
    rc = sub::ilqg_create(desc, params, groups[g].second, groups[g].first, &sh);
    if (rc == ILQG_OK) {
      h->subs.push_back(sh);
      h->first.push_back(lo);
      lo += groups[g].second;
    }
  }
  if (rc == ILQG_OK) {
    Guard guard(h->device);
    if (!guard.ok || cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming) != cudaSuccess)
      rc = ILQG_ERR_CUDA;
    h->stream = h->own_stream;
    for (size_t k = 0; k < h->subs.size() && rc == ILQG_OK; k++) {
      Guard gd(h->subs[k]->device);
      cudaEvent_t ev;
      if (!gd.ok || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) rc = ILQG_ERR_CUDA;
      else h->done.push_back(ev);
    }
  }
  if (rc != ILQG_OK) {
    ilqg_destroy(h);
    return rc;
  }
  h->layout = h->subs[0]->layout;
  h->layout.batch = batch;
  h->op_t0 = desc->initial_time;
  *out = h;
  return ILQG_OK;
}

int ilqg_create(const ilqg_problem_desc* desc, const ilqg_solver_params* params, int batch, int device,
                ilqg_handle* out) {
  if (!desc || !params || !out || batch < 1) return ILQG_ERR_INVALID_ARGUMENT;
  // one group: with the linesearch split into rollout and merit kernels the plain sequence on one
  // stream is the fastest schedule measured (profiles/r01_schedule_experiments.md)
  int groups = 1;
  if (const char* e = std::getenv("ILQG_GROUPS")) groups = std::max(1, std::atoi(e));
  groups = std::min(groups, batch);
  std::vector<std::pair<int, int>> g;
  for (int k = 0; k < groups; k++)
    g.push_back({device, (int)((long long)batch * (k + 1) / groups) - (int)((long long)batch * k / groups)});
  return CreateGroups(desc, params, batch, g, out);
}

// The batch sharded over several GPUs of one node behind ONE handle: device k owns a contiguous slice
// (sizes differ by at most one game), every entry point fans out to the devices' streams and joins
// them, uploads / downloads address the global batch.  No data-path collective: the games are
// independent; the "gather" of converged trajectories is the downloads' device-to-host copies.
int ilqg_create_multi(const ilqg_problem_desc* desc, const ilqg_solver_params* params, int batch,
                      const int* devices, int num_devices, ilqg_handle* out) {
  if (!desc || !params || !out || !devices || num_devices < 1 || batch < num_devices) return ILQG_ERR_INVALID_ARGUMENT;
  for (int a = 0; a < num_devices; a++)
    for (int b = a + 1; b < num_devices; b++)
      if (devices[a] == devices[b]) return ILQG_ERR_INVALID_ARGUMENT;
  std::vector<std::pair<int, int>> g;
  const int base = batch / num_devices, rem = batch % num_devices;
  for (int k = 0; k < num_devices; k++) g.push_back({devices[k], base + (k < rem ? 1 : 0)});
  return CreateGroups(desc, params, batch, g, out);
}

int ilqg_destroy(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  Guard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (SubSolver* g : h->subs) sub::ilqg_destroy(g);
  for (cudaEvent_t ev : h->done) cudaEventDestroy(ev);
  if (h->fork) cudaEventDestroy(h->fork);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return ILQG_OK;
}

int ilqg_get_layout(ilqg_handle h, ilqg_layout* out) {
  if (!h || !out) return ILQG_ERR_BAD_HANDLE;
  *out = h->layout;
  return ILQG_OK;
}

// per-instance element counts of every uploadable / downloadable array (floats or ints)
static size_t PerInstance(const ilqg_solver* h, int what) {
  const ilqg_layout& lo = h->layout;
  const size_t T = lo.num_time_steps, n = lo.xdim, M = lo.total_udim, N = lo.num_players;
  switch (what) {
    case ILQG_XS: case ILQG_DELTA_XS: case ILQG_WARM_XS: return T * n;
    case ILQG_US: case ILQG_ALPHAS: case ILQG_LQ_ALPHAS: case ILQG_WARM_US: case ILQG_WARM_ALPHAS: return T * M;
    case ILQG_PS: case ILQG_LQ_PS: case ILQG_WARM_PS: return T * M * n;
    case ILQG_LIN_A: return T * n * n;
    case ILQG_LIN_B: return T * n * M;
    case ILQG_QUAD_Q: return T * N * n * n;
    case ILQG_QUAD_L: return T * N * n;
    case ILQG_QUAD_R: return T * (size_t)lo.R_floats;
    case ILQG_QUAD_RGRAD: return T * (size_t)lo.r_floats;
    case ILQG_LAMBDAS: return (size_t)lo.num_constraints * T;
    case ILQG_TOTAL_COSTS: case ILQG_TIME_OF_EXTREME: return N;
    case ILQG_X0: case ILQG_LQ_X0: return n;
    case ILQG_MU: case ILQG_MERIT: case ILQG_EXPECTED_DECREASE: case ILQG_STEP: case ILQG_MAX_CONSTRAINT_ERROR:
    case ILQG_STATUS: case ILQG_ITERS: case ILQG_BACKTRACKS:
    case ILQG_AL_SUCCESS: case ILQG_AL_ITERATES: case ILQG_AL_STATE: return 1;
  }
  return 0;
}

int ilqg_upload_x0(ilqg_handle h, const float* x0, size_t bytes) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  if (!x0) return ILQG_ERR_INVALID_ARGUMENT;
  const size_t n = h->layout.xdim;
  if (bytes != sizeof(float) * (size_t)h->B * n) return ILQG_ERR_SIZE_MISMATCH;
  return ForEach(h, [&](SubSolver* g, int lo) { return sub::ilqg_upload_x0(g, x0 + (size_t)lo * n, sizeof(float) * (size_t)g->B * n); });
}

int ilqg_upload_warmstart(ilqg_handle h, const float* xs, const float* us, const float* Ps, const float* alphas) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  const size_t px = PerInstance(h, ILQG_XS), pu = PerInstance(h, ILQG_US), pp = PerInstance(h, ILQG_PS);
  return ForEach(h, [&](SubSolver* g, int lo) {
    return sub::ilqg_upload_warmstart(g, xs ? xs + lo * px : nullptr, us ? us + lo * pu : nullptr,
                                      Ps ? Ps + lo * pp : nullptr, alphas ? alphas + lo * pu : nullptr);
  });
}

int ilqg_upload(ilqg_handle h, int what, const void* src, size_t bytes) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  if (!src) return ILQG_ERR_INVALID_ARGUMENT;
  const size_t per = PerInstance(h, what) * 4;
  if (per == 0 && what != ILQG_LAMBDAS) return ILQG_ERR_INVALID_ARGUMENT;
  if (bytes != per * (size_t)h->B) return ILQG_ERR_SIZE_MISMATCH;
  return ForEach(h, [&](SubSolver* g, int lo) {
    return sub::ilqg_upload(g, what, (const char*)src + (size_t)lo * per, per * (size_t)g->B);
  });
}

int ilqg_upload_lq(ilqg_handle h, const float* A, const float* Bs, const float* Q, const float* l, const float* R,
                   const float* r) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  if (!A || !Bs || !Q || !l || !R || !r) return ILQG_ERR_INVALID_ARGUMENT;
  return ForEach(h, [&](SubSolver* g, int lo) {
    return sub::ilqg_upload_lq(g, A + lo * PerInstance(h, ILQG_LIN_A), Bs + lo * PerInstance(h, ILQG_LIN_B),
                               Q + lo * PerInstance(h, ILQG_QUAD_Q), l + lo * PerInstance(h, ILQG_QUAD_L),
                               R + lo * PerInstance(h, ILQG_QUAD_R), r + lo * PerInstance(h, ILQG_QUAD_RGRAD));
  });
}

int ilqg_solve_begin(ilqg_handle h) { return ForEach(h, [](SubSolver* g, int) { return sub::ilqg_solve_begin(g); }); }
int ilqg_linearize_quadraticize(ilqg_handle h) {
  return ForEach(h, [](SubSolver* g, int) { return sub::ilqg_linearize_quadraticize(g); });
}
int ilqg_lq_backward(ilqg_handle h) { return ForEach(h, [](SubSolver* g, int) { return sub::ilqg_lq_backward(g); }); }
int ilqg_linesearch(ilqg_handle h) { return ForEach(h, [](SubSolver* g, int) { return sub::ilqg_linesearch(g); }); }
int ilqg_al_update(ilqg_handle h) { return ForEach(h, [](SubSolver* g, int) { return sub::ilqg_al_update(g); }); }
int ilqg_al_post_solve(ilqg_handle h) { return ForEach(h, [](SubSolver* g, int) { return sub::ilqg_al_post_solve(g); }); }
int ilqg_overwrite_solution(ilqg_handle h, int only_successful) {
  return ForEach(h, [&](SubSolver* g, int) { return sub::ilqg_overwrite_solution(g, only_successful); });
}
int ilqg_reset(ilqg_handle h, int mask) {
  const int rc = ForEach(h, [&](SubSolver* g, int) { return sub::ilqg_reset(g, mask); });
  if (rc == ILQG_OK && (mask & ILQG_RESET_SOLUTION)) {
    h->op_t0 = h->subs[0]->host_desc.initial_time;  // a fresh OperatingPoint's t0
    for (SubSolver* g : h->subs) UpdateCostGates(&g->d, h->op_t0);
  }
  return rc;
}

int ilqg_setup_next_receding_horizon(ilqg_handle h, const float* x0, double t0, double planner_runtime,
                                     double* new_t0) {
  if (!h || !x0) return ILQG_ERR_BAD_HANDLE;
  const ilqg_layout& lo = h->layout;
  // a constrained problem CHECK-fails in the reference once initial_time_ > 0 (ilqg.h)
  if (lo.num_constraints > 0) return ILQG_ERR_UNSUPPORTED;
  const DevDesc& d = h->subs[0]->d;
  const int T = lo.num_time_steps;
  const double kTimeStep = d.time_step, kTimeHorizon = kTimeStep * T;
  constexpr float kSmallNumber = 1e-4;  // constants::kSmallNumber, utils/types.h:118
  // ---- SyncToExistingProblem, src/problem.cpp:64-125: the time bookkeeping, in double ----
  if (planner_runtime < 0.0 || planner_runtime + t0 > h->op_t0 + kTimeHorizon || t0 < h->op_t0)
    return ILQG_ERR_INVALID_ARGUMENT;  // CHECKs :68-71
  constexpr float kRoundingError = 0.9;
  const double relative_t0 = t0 - h->op_t0;
  size_t current_timestep = static_cast<size_t>(relative_t0 / kTimeStep);
  double remaining_time_this_step = (current_timestep + 1) * kTimeStep - relative_t0;
  if (remaining_time_this_step < kRoundingError * kTimeStep) {
    current_timestep += 1;
    remaining_time_this_step = kTimeStep - remaining_time_this_step;
  }
  if (!(remaining_time_this_step < kTimeStep)) return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_LT :87
  // IntegrateToNextTimeStep, src/multi_player_integrable_system.cpp:113-143
  const size_t itn_timestep = static_cast<size_t>((relative_t0 + kSmallNumber) / kTimeStep);
  const double itn_remaining = kTimeStep * (itn_timestep + 1) - relative_t0;
  if (!(itn_remaining < kTimeStep + kSmallNumber) || itn_timestep >= (size_t)T) return ILQG_ERR_INVALID_ARGUMENT;
  double op_t0 = t0 + remaining_time_this_step;
  size_t num_steps_to_integrate = 0;
  if (remaining_time_this_step <= planner_runtime) {
    num_steps_to_integrate =
        static_cast<size_t>(kSmallNumber + (planner_runtime - remaining_time_this_step) / kTimeStep);
    op_t0 += kTimeStep * num_steps_to_integrate;
  }
  const size_t last_integration_timestep = current_timestep + num_steps_to_integrate;
  if (last_integration_timestep > (size_t)T) return ILQG_ERR_INVALID_ARGUMENT;
  if (!(std::abs(t0 + planner_runtime - op_t0) <= kTimeStep)) return ILQG_ERR_INVALID_ARGUMENT;  // :123

  RhTimes times;
  times.itn_timestep = (int)itn_timestep;
  times.frac = (float)(itn_remaining / kTimeStep);
  times.itn_dt_half = (float)(itn_remaining / 2.0);
  times.integrate_from = (int)current_timestep + 1;
  times.integrate_to = remaining_time_this_step <= planner_runtime ? (int)last_integration_timestep : 0;
  times.dt_half = (float)(kTimeStep / 2.0);
  const int ego_kind = d.sub[0].kind;
  // MultiPlayerIntegrableSystem::DistanceBetween of the problem's dynamics is positional everywhere: the
  // concatenated systems look at the first subsystem's (x, y) (src/concatenated_dynamical_system.cpp:109-113;
  // a Dubins ego has no override: its whole state), Air3D and TwoPlayerUnicycle4D at their first two states
  // (air_3d.h:151-157, two_player_unicycle_4d.h:141-147)
  times.position_distance = ego_kind == ILQG_DYN_DUBINS ? 2 : 1;
  times.ego_dim = (ego_kind == ILQG_DYN_AIR3D || ego_kind == ILQG_DYN_TWO_PLAYER_UNICYCLE4D) ? lo.xdim
                                                                                            : SubsystemXdim(ego_kind);
  const size_t n = lo.xdim;
  const int rc = ForEach(h, [&](SubSolver* g, int first) {
    return sub::ilqg_setup_next_receding_horizon(g, x0 + (size_t)first * n, times);
  });
  if (rc != ILQG_OK) return rc;
  h->op_t0 = op_t0;
  for (SubSolver* g : h->subs) UpdateCostGates(&g->d, op_t0);  // RelativeTimeTracker::ResetInitialTime(op.t0), src/problem.cpp:120
  if (new_t0) *new_t0 = op_t0;
  return ILQG_OK;
}

// The trip count of the RK4 loop of MultiPlayerDynamicalSystem::Integrate
// (src/multi_player_dynamical_system.cpp:61-65) for one (t0, interval), in double as there.
static int Rk4Substeps(double t0, double time_interval) {
  const double dt = time_interval / static_cast<double>(2);
  int count = 0;
  for (double t = t0; t < t0 + time_interval - 0.5 * dt && count < 4; t += dt) count++;
  return count;
}

int ilqg_integrate_plan(ilqg_handle h, const float* x0, double t0, double t, float* x_out) {
  if (!h || !x0 || !x_out) return ILQG_ERR_BAD_HANDLE;
  const ilqg_layout& lo = h->layout;
  const DevDesc& d = h->subs[0]->d;
  if (d.num_subsystems <= 0 || d.sub[0].kind == ILQG_DYN_NONE) return ILQG_ERR_UNSUPPORTED;
  const int T = lo.num_time_steps;
  const double kTimeStep = d.time_step, plan_t0 = h->op_t0;
  constexpr float kSmallNumber = 1e-4;  // constants::kSmallNumber, utils/types.h:118
  // ---- src/multi_player_integrable_system.cpp:54-83: the time bookkeeping, in double ----
  if (!(t >= t0) || !(t0 >= plan_t0)) return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_GE :57-58
  const double relative_t0 = t0 - plan_t0;
  const size_t current_timestep = static_cast<size_t>(relative_t0 / kTimeStep);
  const double relative_t = t - plan_t0;
  const size_t final_timestep = static_cast<size_t>(relative_t / kTimeStep);
  // IntegrateToNextTimeStep :113-143
  const bool to_next = t0 > plan_t0;
  const size_t itn_timestep = static_cast<size_t>((relative_t0 + kSmallNumber) / kTimeStep);
  const double itn_remaining = kTimeStep * (itn_timestep + 1) - relative_t0;
  if (to_next && (!(itn_remaining < kTimeStep + kSmallNumber) || itn_timestep >= (size_t)T))
    return ILQG_ERR_INVALID_ARGUMENT;  // CHECK_LT :126-127
  // IntegrateFromPriorTimeStep :145-171
  const double remaining_until_t = relative_t - kTimeStep * final_timestep;
  if (final_timestep >= (size_t)T || !(remaining_until_t < kTimeStep)) return ILQG_ERR_INVALID_ARGUMENT;  // :154-155

  IpTimes times;
  times.to_next = to_next ? 1 : 0;
  times.itn_timestep = (int)itn_timestep;
  times.frac = (float)(itn_remaining / kTimeStep);
  times.itn_dt_half = (float)(itn_remaining / 2.0);
  times.itn_substeps = Rk4Substeps(t0, itn_remaining);
  times.integrate_from = (int)current_timestep + 1;
  times.integrate_to = (int)final_timestep;
  times.dt_half = (float)(kTimeStep / 2.0);
  times.step_substeps = Rk4Substeps(0.0, kTimeStep);
  times.prior_timestep = (int)final_timestep;
  times.prior_dt_half = (float)(remaining_until_t / 2.0);
  times.prior_substeps = Rk4Substeps(plan_t0 + kTimeStep * final_timestep, remaining_until_t);
  const size_t n = lo.xdim;
  return ForEach(h, [&](SubSolver* g, int first) {
    return sub::ilqg_integrate_plan(g, x0 + (size_t)first * n, times, x_out + (size_t)first * n);
  });
}

int ilqg_iterate(ilqg_handle h, int max_iters, int* iters_done) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  Guard guard(h->device);
  if (!guard.ok) return ILQG_ERR_CUDA;
  int rc = Fork(h);
  if (rc != ILQG_OK) return rc;
  if (h->subs[0]->pipeline && max_iters > 1) {
    // each group runs its own two-stream pipeline (sub::IteratePipelined); with ILQG_STAGGER the groups
    // start half a pass apart
    static const bool stagger = std::getenv("ILQG_STAGGER") && std::atoi(std::getenv("ILQG_STAGGER")) != 0;
    for (size_t k = 0; k < h->subs.size(); k++) {
      h->subs[k]->stagger_wait = (stagger && k > 0) ? h->subs[k - 1]->stagger_signal : nullptr;
      if ((rc = sub::ilqg_iterate(h->subs[k], max_iters, nullptr)) != ILQG_OK) return rc;
    }
  } else {
    // iteration-major, group-minor issue order so the groups' kernels interleave in the hardware queues
    for (int it = 0; it < max_iters; it++)
      for (SubSolver* g : h->subs)
        if ((rc = sub::ilqg_iterate(g, 1, nullptr)) != ILQG_OK) return rc;
  }
  if ((rc = Join(h)) != ILQG_OK) return rc;
  if (iters_done) {
    int most = 0;
    for (SubSolver* g : h->subs) {
      int v = 0;
      if ((rc = sub::ilqg_iterate(g, 0, &v)) != ILQG_OK) return rc;
      most = std::max(most, v);
    }
    *iters_done = most;
  }
  return ILQG_OK;
}

int ilqg_al_begin(ilqg_handle h, int max_iterates, float constraint_error_tolerance) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  if (max_iterates < 0) return ILQG_ERR_INVALID_ARGUMENT;
  h->al_max_iterates = max_iterates;
  h->al_tolerance = constraint_error_tolerance;
  h->al_first = true;
  return ForEach(h, [](SubSolver* g, int) { return sub::ilqg_al_begin(g); });
}

int ilqg_al_advance(ilqg_handle h, int* active) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  int running = 0;
  int rc = ilqg_count_running(h, &running);
  if (rc != ILQG_OK) return rc;
  if (running > 0) return ILQG_ERR_INVALID_ARGUMENT;  // the inner solve has not finished
  Guard guard(h->device);
  if (!guard.ok) return ILQG_ERR_CUDA;
  if (h->al_active_dev.empty()) {
    for (SubSolver* g : h->subs) {
      Guard gd(g->device);
      int* ctr = nullptr;
      if ((rc = DevAlloc(g, &ctr, 1)) != ILQG_OK) return rc;
      h->al_active_dev.push_back(ctr);
    }
  }
  for (size_t k = 0; k < h->subs.size(); k++) {
    Guard gd(h->subs[k]->device);
    CUDA_TRY(cudaMemsetAsync(h->al_active_dev[k], 0, sizeof(int), h->subs[k]->stream));
    CUDA_TRY(cudaStreamSynchronize(h->subs[k]->stream));
  }
  const int first = h->al_first ? 1 : 0;
  size_t which = 0;
  rc = ForEach(h, [&](SubSolver* g, int) {
    return sub::ilqg_al_advance(g, first, h->al_max_iterates, h->al_tolerance, h->al_active_dev[which++]);
  });
  if (rc != ILQG_OK) return rc;
  h->al_first = false;
  if ((rc = ilqg_synchronize(h)) != ILQG_OK) return rc;
  int n = 0;
  for (size_t k = 0; k < h->subs.size(); k++) {
    Guard gd(h->subs[k]->device);
    int v = 0;
    CUDA_TRY(cudaMemcpy(&v, h->al_active_dev[k], sizeof(int), cudaMemcpyDeviceToHost));
    n += v;
  }
  if (active) *active = n;
  return ILQG_OK;
}

int ilqg_count_running(ilqg_handle h, int* running) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  if (!running) return ILQG_ERR_INVALID_ARGUMENT;
  int total = 0;
  for (SubSolver* g : h->subs) {
    int v = 0;
    const int rc = sub::ilqg_count_running(g, &v);
    if (rc != ILQG_OK) return rc;
    total += v;
  }
  *running = total;
  return ILQG_OK;
}

int ilqg_download(ilqg_handle h, int what, void* dst, size_t bytes) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  if (!dst) return ILQG_ERR_INVALID_ARGUMENT;
  const size_t per = PerInstance(h, what) * 4;
  if (per == 0 && what != ILQG_LAMBDAS && what != ILQG_QUAD_R && what != ILQG_QUAD_RGRAD) return ILQG_ERR_INVALID_ARGUMENT;
  if (bytes != per * (size_t)h->B) return ILQG_ERR_SIZE_MISMATCH;
  return ForEach(h, [&](SubSolver* g, int lo) {
    return sub::ilqg_download(g, what, (char*)dst + (size_t)lo * per, per * (size_t)g->B);
  });
}

int ilqg_synchronize(ilqg_handle h) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  Guard guard(h->device);
  for (SubSolver* g : h->subs) CUDA_TRY(cudaStreamSynchronize(g->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ILQG_OK;
}

int ilqg_set_stream(ilqg_handle h, void* cuda_stream) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  Guard guard(h->device);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return ILQG_OK;
}

int ilqg_profile(ilqg_handle h, int enable) {
  if (!h) return ILQG_ERR_BAD_HANDLE;
  for (SubSolver* g : h->subs) {
    const int rc = sub::ilqg_profile(g, enable);
    if (rc != ILQG_OK) return rc;
  }
  return ILQG_OK;
}

int ilqg_profile_read(ilqg_handle h, int kernel, double* total_ms, long long* launches) {
  if (!h || !total_ms || !launches) return ILQG_ERR_BAD_HANDLE;
  *total_ms = 0;
  *launches = 0;
  for (SubSolver* g : h->subs) {
    double ms = 0;
    long long n = 0;
    const int rc = sub::ilqg_profile_read(g, kernel, &ms, &n);
    if (rc != ILQG_OK) return rc;
    *total_ms += ms;
    *launches += n;
  }
  return ILQG_OK;
}

int ilqg_kernel_launches(ilqg_handle h, long long* out) {
  if (!h || !out) return ILQG_ERR_BAD_HANDLE;
  *out = 0;
  for (SubSolver* g : h->subs) *out += g->launches;
  return ILQG_OK;
}

}  // extern "C"
