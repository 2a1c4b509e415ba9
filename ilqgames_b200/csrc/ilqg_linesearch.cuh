// ilqg_linesearch.cuh -- K_ls: ILQSolver::ModifyLQStrategies (src/ilq_solver.cpp:289-348) as a
// speculative, asynchronous pipeline of small kernels, plus the Solve() prologue built from the
// same kernels.
//
// The reference backtracks sequentially: roll out with alpha * s0 * rho^j, evaluate the merit,
// test Armijo, repeat.  Candidate j's trajectory and merit depend only on j, never on the
// outcome of candidates < j, so evaluating a WINDOW of candidates at once and taking the FIRST
// one that passes Armijo gives exactly the sequential result.  On the benchmark workload 97.8 %
// of all linesearches accept j = 0 and most of the rest backtrack >= 8 times (measured with the
// oracle), hence two windows per linesearch:
//   window [0, JA) for every running instance (JA = 1 by default); instances that reject it are
//                  queued (per-instance Armijo state machine: ls_next_j)
//   window [JA, max_backtracking_steps) for the queued instances, all candidates at once, in chunks
//                  of `cap` queue slots so the candidate trajectories fit the scratch; first passing
//                  candidate is accepted, none -> LINESEARCH_FAILED
// Launches whose queue range is empty exit immediately; the host never synchronises.
//
// A window is k_ls_rollout (one (instance, candidate) ITEM per lane, one subsystem per warp:
// u = u_ref - P dx - alpha, then RK4 x 2 substeps), k_ls_merit (one player per warp: the cost
// gradients and values of the stored trajectory, all time steps in parallel) and k_ls_decide.
#pragma once
#include "ilqg_backward.cuh"

namespace ilqg {

enum { LS_MODE_BEGIN = 0 };

struct LsScratch {
  float* traj_xs;   // [items][T][n]   candidate trajectories (item = global lane index of the launch)
  float* traj_us;   // [items][T][M]
  float* terms;     // [blocks][T][2N][32]  merit terms, item = lane of its block
  float* vals;      // [blocks][T][N][32]   per-player cost values
  int* absorbed;    // [items] 1 = every alpha term of this candidate vanished in rounding (see k_ls_rollout)
  int* pend[3];     // queues of instances with an open linesearch: [0] is filled by the first window and
                    // stays intact for the rest of the pass (the pipelined schedule runs K_lq over it), the
                    // tiers of the continued linesearch hand the still-open games on between [1] and [2]
  int* counts;      // [3] queue lengths
  int* slot;        // [B] position of an instance in the queue it is in
  int JA, JB;       // window sizes: fresh linesearch / continued linesearch
  int nA_blocks;    // item blocks of a first-window launch
  int lpw;          // items per block (active lanes per warp): 32, 16 or 8.  The rollout is a
                    // latency chain, so fewer items per warp = more warps per SM to hide it
  int cap;          // queue slots one continued-window launch can hold trajectories for
};

struct LsItem {
  int b, j;
  bool valid;
};

// cur: which queue this pass consumes
enum { LS_MODE_FRESH = 1, LS_MODE_QUEUED = 2 };

__device__ __forceinline__ LsItem ls_decode(const DevParams& p, const Slab& s, const LsScratch& ls, int mode,
                                            int cur, int q_offset, int block, int lane) {
  LsItem it;
  it.b = 0;
  it.j = 0;
  it.valid = false;
  if (lane >= ls.lpw) return it;
  if (mode == LS_MODE_BEGIN) {
    it.b = block * ls.lpw + lane;
    it.valid = it.b < s.B && s.al_state[it.b] != 2;  // a finished AL game is left alone
  } else if (mode == LS_MODE_FRESH) {
    const int item = block * ls.lpw + lane;
    it.b = item / ls.JA;
    it.j = item % ls.JA;
    it.valid = it.b < s.B && s.status[it.b] == ILQG_STATUS_RUNNING && s.ls_next_j[it.b] == 0 &&
               it.j < p.max_backtracking_steps;
  } else {
    const int item = block * ls.lpw + lane;
    const int q = q_offset + item / ls.JB;
    if (item / ls.JB < ls.cap && q < ls.counts[cur]) {
      it.b = ls.pend[cur][q];
      it.j = s.ls_next_j[it.b] + item % ls.JB;
      it.valid = it.j < p.max_backtracking_steps;
    }
  }
  return it;
}

// Which rollout kernel a window takes.  The stage-parallel kernel spends eight lanes per (item, subsystem) to cut
// the latency chain of a time step; the lane-per-item kernel spends one and takes ~2.8 x as long per wave,
// but a wave of it holds eight times the items.  A window of a few thousand items is latency bound
// (stage-parallel wins), one of a hundred thousand is throughput bound (lane-per-item wins: half the
// instructions per item).  The size of a continued window exists only on the device, so both kernels
// are launched for it and each asks this function whether it is the one to run.
struct LsPick {
  int which;      // 0: run unconditionally; 1: "I am the stage-parallel kernel"; 2: "I am the lane-per-item kernel"
  int cap_sp;     // items one resident wave of the stage-parallel kernel holds on this device
  int cap_lanes;  // ... of the lane-per-item kernel
};
__host__ __device__ inline bool ls_prefers_lanes(long long items, int cap_sp, int cap_lanes) {
  if (items <= 0 || cap_sp <= 0 || cap_lanes <= 0) return false;
  const long long w_sp = (items + cap_sp - 1) / cap_sp, w_lanes = (items + cap_lanes - 1) / cap_lanes;
  return 14 * w_lanes < 5 * w_sp;
}
__device__ __forceinline__ bool ls_pick_runs(const LsScratch& ls, int mode, int cur_q, int q_offset, int blocks,
                                             const LsPick& pick) {
  if (pick.which == 0) return true;
  long long items = (long long)blocks * ls.lpw;
  if (mode == 2 /* LS_MODE_QUEUED */)
    items = min(items, (long long)min(max(ls.counts[cur_q] - q_offset, 0), ls.cap) * ls.JB);
  const bool lanes = ls_prefers_lanes(items, pick.cap_sp, pick.cap_lanes);
  return pick.which == 2 ? lanes : !lanes;
}

// gradient accumulator in the lane-minor shared tile; `on` = false keeps only the cost value
struct GatedSink {
  float* grad;  // element idx at grad[idx * 32]
  bool on;
  __device__ __forceinline__ void H(int, int, float) const {}
  __device__ __forceinline__ void G(int i, float v) const {
    if (on) grad[i * 32] += v;
  }
};


// ===========================================================================
// Split evaluation: rollout and merit as two kernels.
//
// Only the rollout is inherently sequential in k.  The merit terms (per-step gradients of the
// players' costs) depend on (x_k, u_k) alone, so once a candidate trajectory is in global memory
// -- where it has to go anyway to be promoted -- they can be evaluated for all time steps at
// once.  (Until session 2 of round 1 one fused kernel did both, with cost warps one step behind
// the dynamics warps; the cost warps sat on the rollout's latency chain and paced it: ncu source
// view, profiles/r01_schedule_experiments.md.)  Split:
//   k_ls_rollout    S warps per block (one per subsystem), 32 items per block: the dynamics role
//                   alone, one named barrier per step, ~30 KB shared memory and ~100 threads per
//                   block, so a whole queued window is resident at once;
//   k_ls_merit      N warps per block (one per player), grid (item block, chunk of time steps):
//                   the cost role over a chunk of the stored trajectory, all chunks in parallel;
//   (k_ls_decide then adds up each candidate's terms in the reference's (k, i) order, one lane
//   per candidate.)
// Same arithmetic on the same fp32 values in the same order as the fused kernel had: merits, cost
// values and therefore every Armijo decision are bit-identical to it.
// ===========================================================================
struct LsIo {
  const float *last_xs, *last_us, *P, *alpha, *x_start;
  float *out_xs, *out_us;
  bool scaled;
};

// where an item reads its strategy / reference from and rolls its trajectory to
__device__ __forceinline__ LsIo ls_io(const Slab& s, const LsScratch& ls, int mode, const LsItem& it, int item,
                                      int T, int n, int M) {
  LsIo io;
  const int b = it.b;
  const size_t ox = (size_t)b * T * n, ou = (size_t)b * T * M, oP = (size_t)b * T * M * n;
  io.scaled = true;
  if (mode == LS_MODE_BEGIN) {
    io.last_xs = s.prob_xs + ox;
    io.last_us = s.prob_us + ou;
    io.P = s.prob_P + oP;
    io.alpha = s.prob_a + ou;
    io.x_start = s.x0 + (size_t)b * n;
    io.scaled = false;
    io.out_xs = s.op_xs[0] + ox;
    io.out_us = s.op_us[0] + ou;
  } else {
    const int cur = it.valid ? s.op_cur[b] : 0, scur = it.valid ? s.st_cur[b] : 0;
    io.last_xs = s.op_xs[cur] + ox;
    io.last_us = s.op_us[cur] + ou;
    io.P = s.st_P[1 - scur] + oP;
    io.alpha = s.st_a[1 - scur] + ou;
    io.x_start = io.last_xs;
    if (mode == LS_MODE_FRESH && ls.JA == 1) {
      // the lone first-window candidate is accepted 98 % of the time: roll it straight into the
      // candidate operating-point buffer (k_ls_decide then has nothing to copy)
      io.out_xs = s.op_xs[1 - cur] + ox;
      io.out_us = s.op_us[1 - cur] + ou;
    } else {
      io.out_xs = ls.traj_xs + (size_t)item * T * n;
      io.out_us = ls.traj_us + (size_t)item * T * M;
    }
  }
  return io;
}

// shared memory of k_ls_rollout (floats): dx[2][n][32] + abs[S][32] + pbuf[S][2][32][2n+4]
__host__ __device__ inline int ls_rollout_smem_floats(int n, int S) {
  return 2 * n * 32 + S * 32 + S * 2 * 32 * (2 * n + 4);
}

// NUQ: control inputs a subsystem can have in this instance (2; 4 only with a TwoPlayerUnicycle4D)
// WIDE: subsystem kinds beyond Car6D / Unicycle4D / Air3D are compiled in (subsystem_xdot)
template <int S, int NUQ, bool WIDE>
__global__ void __launch_bounds__(S * 32)
k_ls_rollout(const __grid_constant__ DevDesc d, const DevParams p, Slab s, LsScratch ls, int mode, int cur_q,
             int q_offset, int blocks, LsPick pick) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = d.n, M = d.M, T = d.T;
  if (!ls_pick_runs(ls, mode, cur_q, q_offset, blocks, pick)) return;
  const int item = blockIdx.x * ls.lpw + lane;
  const LsItem it = ls_decode(p, s, ls, mode, cur_q, q_offset, blockIdx.x, lane);
  if (!__syncthreads_or(it.valid)) return;
  const bool valid = it.valid;
  const LsIo io = ls_io(s, ls, mode, it, item, T, n, M);
  float* dxs2 = smem;                          // [2][n][32]
  float* absf = dxs2 + 2 * n * 32;             // [S][32]
  float* pbase = absf + S * 32;                // [S][2][32][2n+4]  prefetched feedback rows
  const float s0 = p.initial_alpha_scaling, rho = p.geometric_alpha_scaling;
  const float dt_half = (float)(d.time_step / 2.0);
  // ScaleAlphas multiplies alpha by rho once per backtrack (src/ilq_solver.cpp:66-72, 331).  When
  // rho is a power of two every one of those products is exact, so rho^j can be formed once.
  int rho_e;
  const bool rho_exact = fabsf(frexpf(rho, &rho_e)) == 0.5f;
  float rho_j = 1.0f;
  for (int jj = 0; jj < it.j; jj++) rho_j *= rho;

  const DevSubsystem& sub = d.sub[warp];
  const int xd = subsystem_xdim(sub.kind);
  const int nu = sub.nu;  // control inputs of this subsystem (rows of P it evaluates)
  float x[6];
#pragma unroll
  for (int a = 0; a < 6; a++) x[a] = (valid && a < xd) ? io.x_start[sub.x_offset + a] : 0.f;
  // Software pipeline: the feedback rows P[k][own rows][:] of the NEXT step are copied
  // asynchronously (cp.async, 16 B granules) into a per-lane double buffer while this step
  // integrates; the small reference values ride in registers.  Lane stride PST = 2n + 4
  // floats keeps the 128-bit reads of 8 consecutive lanes on distinct banks.
  const bool vecP = (n & 3) == 0 && nu <= 2;
  const int PST = 2 * n + 4;
  float* pbuf = pbase + (size_t)warp * 2 * 32 * PST;
  float nref[6], nuref[NUQ], nal[NUQ];
#pragma unroll
  for (int q = 0; q < NUQ; q++) nuref[q] = nal[q] = 0.f;
  bool absorbed = true;
  auto prefetch = [&](int k) {
    if (!valid) return;
#pragma unroll
    for (int a = 0; a < 6; a++)
      if (a < xd)
        nref[a] = k > 0 ? io.last_xs[(size_t)k * n + sub.x_offset + a] : io.x_start[sub.x_offset + a];
#pragma unroll
    for (int q = 0; q < NUQ; q++) {
      if (q >= nu) break;
      const int c = sub.ucol[q];
      nuref[q] = io.last_us[(size_t)k * M + c];
      nal[q] = io.alpha[(size_t)k * M + c];
      if (vecP) {
        const float* src = io.P + ((size_t)k * M + c) * n;
        float* dst = pbuf + ((size_t)(k & 1) * 32 + lane) * PST + q * n;
        for (int a4 = 0; a4 < n / 4; a4++) {
          const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + 4 * a4);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src + 4 * a4) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int a = 0; a < 6; a++) nref[a] = 0.f;
  prefetch(0);
  for (int k = 0; k < T; k++) {
    float* dxs = dxs2 + (k & 1) * n * 32;
    float ref[6], uref[NUQ], al[NUQ];
#pragma unroll
    for (int a = 0; a < 6; a++) ref[a] = nref[a];
#pragma unroll
    for (int q = 0; q < NUQ; q++) { uref[q] = nuref[q]; al[q] = nal[q]; }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (k + 1 < T) prefetch(k + 1);
#pragma unroll
    for (int a = 0; a < 6; a++)
      if (a < xd) {
        // last_operating_point.xs[0] is the start state (src/ilq_solver.cpp:88-89)
        dxs[(sub.x_offset + a) * 32 + lane] = x[a] - ref[a];
        if (valid) io.out_xs[(size_t)k * n + sub.x_offset + a] = x[a];
      }
    // the one barrier of a step: dx is double-buffered, so the warp that races ahead writes the
    // other half while the others may still be reading this one
    if (S > 1) __syncthreads();
    else __syncwarp();
    float uu[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < NUQ; q++) {
      if (q >= nu) break;
      const int c = sub.ucol[q];
      float uv = 0.f;
      if (valid) {
        float acc = 0.f;
        if (vecP) {
          const float4* Prow = reinterpret_cast<const float4*>(pbuf + ((size_t)(k & 1) * 32 + lane) * PST + q * n);
          for (int a4 = 0; a4 < n / 4; a4++) {
            const float4 pv = Prow[a4];
            acc = fmaf(pv.x, dxs[(4 * a4 + 0) * 32 + lane], acc);
            acc = fmaf(pv.y, dxs[(4 * a4 + 1) * 32 + lane], acc);
            acc = fmaf(pv.z, dxs[(4 * a4 + 2) * 32 + lane], acc);
            acc = fmaf(pv.w, dxs[(4 * a4 + 3) * 32 + lane], acc);
          }
        } else {
          const float* Pr = io.P + ((size_t)k * M + c) * n;
          for (int a = 0; a < n; a++) acc = fmaf(__ldg(Pr + a), dxs[a * 32 + lane], acc);
        }
        float alv = al[q];
        if (io.scaled) {
          alv *= s0;  // ScaleAlphas(initial_alpha_scaling), then geometric_alpha_scaling^j
          if (rho_exact) {
            alv *= rho_j;
          } else {
            for (int jj = 0; jj < it.j; jj++) alv *= rho;
          }
        }
        const float t = uref[q] - acc;
        uv = t - alv;  // Strategy::operator(), strategy.h:73-76
        absorbed = absorbed && (uv == t);
        io.out_us[(size_t)k * M + c] = uv;
      }
      uu[q] = uv;
    }
    if (k < T - 1) subsystem_integrate<WIDE>(sub, dt_half, x, uu);
  }
  // If u_k = (u_ref - P dx) - alpha_k s0 rho^j rounded to (u_ref - P dx) at every step, every
  // deeper candidate (smaller alpha) reproduces this rollout bit for bit: k_ls_decide can run
  // the rest of the Armijo loop on this merit without another rollout.
  absf[warp * 32 + lane] = absorbed ? 1.f : 0.f;
  __syncthreads();
  if (warp == 0 && valid) {
    bool all_absorbed = true;
    for (int w2 = 0; w2 < S; w2++) all_absorbed = all_absorbed && absf[w2 * 32 + lane] != 0.f;
    ls.absorbed[item] = all_absorbed ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------
// The stage-parallel rollout (round 2): the same trajectories, bit for bit, from a different
// decomposition.  k_ls_rollout above gives one lane a whole (item, subsystem): ~1,400 dependent
// warp instructions per time step on a warp that is alone on its scheduler -- 3.9 us per step,
// 0.39 ms for the first window of the benchmark with the chip 95 % idle (profiles/r02d).  Here
// eight lanes share one (item, subsystem): one lane per RK4 stage of the two substeps
// (subsystem_integrate_sp, ilqg_device.cuh), one lane per control row for u = u_ref - P dx - alpha,
// one lane per state component for dx and the stores.  A warp is (subsystem, 4 items); a block is
// the S subsystem warps of those 4 items, exchanging dx through a double-buffered shared tile
// with one barrier per step as before.  Feedback rows ride in registers, loaded one step ahead.
// ---------------------------------------------------------------------------
constexpr int RSP_ITEMS = 4;       // items per block (eight lanes each in every subsystem warp)
constexpr int RSP_DX_STRIDE = 28;  // floats per item row of the dx tile: n <= 24, and 28 g mod 32 keeps the four
                                   // items' 128-bit reads on distinct banks

// N4T: n / 4 known at compile time (n = 4 N4T: the feedback row is N4T 128-bit loads, the dot product
// an unrolled chain), or 0 for any n at run time.
#ifndef ILQG_RSP_MINB
#define ILQG_RSP_MINB 6
#endif
template <int NUQ, bool WIDE, int N4T>
__global__ void __launch_bounds__(ILQG_MAX_SUBSYSTEMS * 32, ILQG_RSP_MINB)
k_ls_rollout_sp(const __grid_constant__ DevDesc d, const DevParams p, Slab s, LsScratch ls, int mode, int cur_q,
                int q_offset, int blocks /* item blocks of ls.lpw items the window spans */, LsPick pick) {
  __shared__ __align__(16) float dxs2[2][RSP_ITEMS][RSP_DX_STRIDE];
  __shared__ int absf[ILQG_MAX_SUBSYSTEMS][RSP_ITEMS];
  if (!ls_pick_runs(ls, mode, cur_q, q_offset, blocks, pick)) return;
  const unsigned full = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, t = lane & 7;  // item of the block, stage slot / component / control row
  const int n = N4T ? 4 * N4T : d.n, M = d.M, T = d.T, S = d.num_subsystems;
  for (int e = threadIdx.x; e < 2 * RSP_ITEMS * RSP_DX_STRIDE; e += blockDim.x) (&dxs2[0][0][0])[e] = 0.f;
  long long total_items = (long long)blocks * ls.lpw;
  if (mode == LS_MODE_QUEUED) {
    const int live = min(max(ls.counts[cur_q] - q_offset, 0), ls.cap);
    total_items = min(total_items, (long long)live * ls.JB);
  }
  const int nblk = (int)((total_items + RSP_ITEMS - 1) / RSP_ITEMS);
  const float s0 = p.initial_alpha_scaling, rho = p.geometric_alpha_scaling;
  const float dt_half = (float)(d.time_step / 2.0);
  int rho_e;
  const bool rho_exact = fabsf(frexpf(rho, &rho_e)) == 0.5f;  // (see k_ls_rollout)
  // the subsystem's constants in registers: the loop below is a latency chain, and an indexed read
  // of the descriptor (constant bank, register offset) costs more than the arithmetic around it
  const int kind = d.sub[warp].kind, xoff = d.sub[warp].x_offset, nu = d.sub[warp].nu;
  const float sp0 = d.sub[warp].p0, sp1 = d.sub[warp].p1;
  const int xd = subsystem_xdim(kind);
  const bool vecP = N4T || (n & 3) == 0;
  const int n4 = N4T ? N4T : (n + 3) >> 2;
  constexpr int NV = N4T ? N4T : ILQG_MAX_XDIM / 4;

  for (int ib = blockIdx.x; ib < nblk; ib += gridDim.x) {
    const int item = ib * RSP_ITEMS + g;
    const LsItem it = ls_decode(p, s, ls, mode, cur_q, q_offset, item / ls.lpw, item % ls.lpw);
    if (!__syncthreads_or(it.valid)) continue;  // (also fences the tiles against the previous trip)
    const bool valid = it.valid;
    const LsIo io = ls_io(s, ls, mode, it, item, T, n, M);
    float rho_j = 1.0f;
    for (int jj = 0; jj < it.j; jj++) rho_j *= rho;
    const float m_s0 = io.scaled ? s0 : 1.0f, m_rho = (io.scaled && rho_exact) ? rho_j : 1.0f;
    const bool slow_rho = io.scaled && !rho_exact;
    const bool rowlane = valid && t < nu && t < NUQ;  // this lane evaluates control row t of the subsystem
    const int c = rowlane ? d.sub[warp].ucol[t & 3] : 0;
    const bool complane = valid && t < xd;            // ... and owns state component t
    float x[6];
#pragma unroll
    for (int a = 0; a < 6; a++) x[a] = (valid && a < xd) ? io.x_start[xoff + a] : 0.f;
    // running pointers (advanced by one time step per trip)
    const float* ref_p = io.last_xs + xoff + t;
    float* outx_p = io.out_xs + xoff + t;
    const float* uref_p = io.last_us + c;
    const float* al_p = io.alpha + c;
    const float* P_p = io.P + (size_t)c * n;
    float* outu_p = io.out_us + c;
    float* dx_w = &dxs2[0][g][xoff + t];
    const float* dx_r = &dxs2[0][g][0];
    const int dx_flip = RSP_ITEMS * RSP_DX_STRIDE;
    // values of the step ahead: the reference state component and control row this lane owns
    float ref = complane ? io.x_start[xoff + t] : 0.f;  // last_operating_point.xs[0] is the start state
    float uref = 0.f, al = 0.f;
    float pr[4 * NV];
#pragma unroll
    for (int a = 0; a < 4 * NV; a++) pr[a] = 0.f;
    auto fetch_row = [&]() __attribute__((always_inline)) {
      if (rowlane) {
        uref = *uref_p;
        al = *al_p;
        if (vecP) {
#pragma unroll
          for (int a4 = 0; a4 < NV; a4++)
            if (N4T || a4 < n4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(P_p) + a4);
              pr[4 * a4] = v.x; pr[4 * a4 + 1] = v.y; pr[4 * a4 + 2] = v.z; pr[4 * a4 + 3] = v.w;
            }
        } else {
#pragma unroll
          for (int a = 0; a < 4 * NV; a++)
            if (a < n) pr[a] = __ldg(P_p + a);
        }
      }
      uref_p += M;
      al_p += M;
      P_p += (size_t)M * n;
    };
    fetch_row();
    bool absorbed = true;
    for (int k = 0; k < T; k++) {
      {  // one lane per state component: dx into the tile, x_k into the trajectory
        const float x01 = (t & 1) ? x[1] : x[0], x23 = (t & 1) ? x[3] : x[2], x45 = (t & 1) ? x[5] : x[4];
        const float xa = (t & 4) ? x45 : ((t & 2) ? x23 : x01);
        if (t < xd) *dx_w = xa - ref;
        if (complane) *outx_p = xa;
        outx_p += n;
      }
      if (S > 1) __syncthreads();
      else __syncwarp();
      float acc = 0.f;
#pragma unroll
      for (int a4 = 0; a4 < NV; a4++)
        if (N4T || a4 < n4) {
          const float4 dv = *reinterpret_cast<const float4*>(dx_r + 4 * a4);
          acc = fmaf(pr[4 * a4], dv.x, acc);
          if (N4T || 4 * a4 + 1 < n) acc = fmaf(pr[4 * a4 + 1], dv.y, acc);
          if (N4T || 4 * a4 + 2 < n) acc = fmaf(pr[4 * a4 + 2], dv.z, acc);
          if (N4T || 4 * a4 + 3 < n) acc = fmaf(pr[4 * a4 + 3], dv.w, acc);
        }
      // ScaleAlphas(initial_alpha_scaling), then geometric_alpha_scaling^j; an unscaled strategy (the
      // Solve() prologue) multiplies by 1, which is exact
      float alv = (al * m_s0) * m_rho;
      if (slow_rho) {
#pragma unroll 1
        for (int jj = 0; jj < it.j; jj++) alv *= rho;
      }
      const float tt = uref - acc;
      const float uv = tt - alv;  // Strategy::operator(), strategy.h:73-76
      if (rowlane) {
        absorbed = absorbed && (uv == tt);
        *outu_p = uv;
      }
      outu_p += M;
      dx_w += (k & 1) ? -dx_flip : dx_flip;
      dx_r += (k & 1) ? -dx_flip : dx_flip;
      if (k + 1 < T) {
        ref_p += n;
        if (complane) ref = *ref_p;
        fetch_row();
      }
      float uu[4] = {0.f, 0.f, 0.f, 0.f};
      const float umine = rowlane ? uv : 0.f;
#pragma unroll
      for (int qq = 0; qq < NUQ; qq++) uu[qq] = __shfl_sync(full, umine, (lane & 24) | qq);
      if (k < T - 1) subsystem_integrate_sp<WIDE>(kind, sp0, sp1, dt_half, x, uu, lane);
    }
    // (see k_ls_rollout: a candidate whose alpha terms all vanished in rounding repeats for every deeper j)
    const unsigned live_rows = __ballot_sync(full, !absorbed);
    if (t == 0) absf[warp][g] = ((live_rows >> (8 * g)) & 0xffu) == 0u;
    __syncthreads();
    if (warp == 0 && t == 0 && valid) {
      bool all_absorbed = true;
      for (int w2 = 0; w2 < S; w2++) all_absorbed = all_absorbed && absf[w2][g] != 0;
      ls.absorbed[item] = all_absorbed ? 1 : 0;
    }
  }
}

// Item blocks of a window launch that can hold valid items.  A queued window is launched for `cap`
// queue slots without the host knowing how many are filled; the filled ones are a prefix.
__device__ __forceinline__ int ls_live_blocks(const LsScratch& ls, int mode, int cur_q, int q_offset, int blocks) {
  if (mode != LS_MODE_QUEUED) return blocks;
  const int live = min(max(ls.counts[cur_q] - q_offset, 0), ls.cap);
  return min(blocks, (int)(((long long)live * ls.JB + ls.lpw - 1) / ls.lpw));
}

// time steps one k_ls_merit block covers
#ifndef ILQG_MERIT_CHUNK
#define ILQG_MERIT_CHUNK 3
#endif
constexpr int KLS_MERIT_CHUNK = ILQG_MERIT_CHUNK;

// shared memory of k_ls_merit (floats): per player warp xu[n + M][32] + acc[n + M][32]
__host__ __device__ inline int ls_merit_smem_floats(int n, int M, int N) { return N * 2 * (n + M) * 32; }

template <bool WIDE>  // false: round 1's record kinds only, no FinalTimeCost gates, no ExtremeValueCost groups
__global__ void __launch_bounds__(ILQG_MAX_PLAYERS * 32, 6)  // (<= 85 registers: eight 3-warp blocks per SM)
k_ls_merit(const __grid_constant__ DevDesc d, const DevParams p, Slab s, LsScratch ls, int mode, int cur_q,
           int q_offset, int blocks) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = d.n, M = d.M, N = d.N, T = d.T;
  const int live_blocks = ls_live_blocks(ls, mode, cur_q, q_offset, blocks);
  // grid-stride over the item blocks (the grid is sized for the machine, not for `blocks`)
  for (int ib = blockIdx.x; ib < live_blocks; ib += gridDim.x) {
  const int item = ib * ls.lpw + lane;
  const LsItem it = ls_decode(p, s, ls, mode, cur_q, q_offset, ib, lane);
  // (the player warps of a block share nothing -- each has its own tiles -- so there is no block barrier)
  if (!__any_sync(0xffffffffu, it.valid)) continue;
  const bool valid = it.valid;
  const int b = it.b;
  const LsIo io = ls_io(s, ls, mode, it, item, T, n, M);
  const int i = warp;  // player
  float* slot = smem + (size_t)i * 2 * (n + M) * 32;   // [n + M][32]  this step's state and controls
  float* acc = slot + (n + M) * 32;                     // [n + M][32]  gradient accumulators
  float* terms = ls.terms + (size_t)ib * T * 2 * N * 32;
  float* vals = ls.vals + (size_t)ib * T * N * 32;
  const float mu = valid ? s.mu[b] : 0.f;
  const int te = valid ? s.te_quad[(size_t)b * N + i] : 0;
  const float* lam_base = s.lambdas + (size_t)(valid ? b : 0) * d.num_constraints * T;  // this game's multipliers [slot][T]
  const bool additive = d.cost_structure[i] == ILQG_COST_SUM;
  const int mi = d.udim[i];
  const int k0 = blockIdx.y * KLS_MERIT_CHUNK, k1 = min(T, k0 + KLS_MERIT_CHUNK);
  for (int kk = k0; kk < k1; kk++) {
    // the stored candidate trajectory (item-major in global memory) -> lane-minor tile
    if (valid) {
      const float* xs = io.out_xs + (size_t)kk * n;
      const float* us = io.out_us + (size_t)kk * M;
      if ((n & 3) == 0 && (((size_t)T * n) & 3) == 0) {
        for (int a4 = 0; a4 < n / 4; a4++) {
          const float4 v = *reinterpret_cast<const float4*>(xs + 4 * a4);
          slot[(4 * a4 + 0) * 32 + lane] = v.x;
          slot[(4 * a4 + 1) * 32 + lane] = v.y;
          slot[(4 * a4 + 2) * 32 + lane] = v.z;
          slot[(4 * a4 + 3) * 32 + lane] = v.w;
        }
      } else {
        for (int a = 0; a < n; a++) slot[a * 32 + lane] = xs[a];
      }
      for (int c = 0; c < M; c++) slot[(n + c) * 32 + lane] = us[c];
    } else {
      for (int a = 0; a < n + M; a++) slot[a * 32 + lane] = 0.f;
    }
    for (int a = 0; a < n + M; a++) acc[a * 32 + lane] = 0.f;
    const bool full = additive || te == kk;
    const int lidx = s.lambda_index[kk];
    float value = 0.f;
    for (int c0 = d.cost_begin[i]; c0 < d.cost_begin[i + 1];) {
      // an ExtremeValueCost (group > 0) stands for its extreme member (src/extreme_value_cost.cpp:50-62)
      const bool grouped = WIDE && d.cost[c0].group > 0;
      const int c = grouped ? extreme_member<32>(d, c0, slot + lane, slot + n * 32 + lane) : c0;
      c0 = grouped ? d.cost[c0].group_end : c0 + 1;
      const DevCost& cd = d.cost[c];
      if (WIDE && kk < cd.first_step) continue;  // FinalTimeCost: zero value and derivatives before its threshold
      const bool is_con = cd.slot >= 0;
      // PlayerCost::Quadraticize vs QuadraticizeControlCosts (src/ilq_solver.cpp:483-487): off the
      // extreme timestep of a MAX/MIN player only control COSTS enter the gradient; the cost
      // VALUE (PlayerCost::Evaluate) always counts every state and control cost.
      const bool in_quad = full || (cd.arg >= 0 && !is_con);
      if (!in_quad && is_con) continue;
      const float lambda = (is_con && valid) ? lam_base[cd.slot * T + lidx] : 0.f;
      const int aoff = (cd.arg < 0 ? 0 : (n + d.uoff[cd.arg]) * 32) + lane;  // the record's argument in the tiles
      GatedSink sink{acc + aoff, in_quad};
      const float* in = slot + aoff;
      float v = 0.f;
      quadraticize_record_sink<false, 32, true, GatedSink, WIDE>(d, cd, in, cd.arg < 0 ? n : d.udim[cd.arg], lambda, mu, sink, &v);
      if (!is_con) value += v;  // PlayerCost::Evaluate: costs only
    }
    // ILQSolver::MeritFunction terms (src/ilq_solver.cpp:416-430)
    float sq = 0.f;
    for (int a = 0; a < mi; a++) {
      const float rv = acc[(n + d.uoff[i] + a) * 32 + lane];
      sq = fmaf(rv, rv, sq);
    }
    float sq2 = 0.f;
    if (kk > 0)
      for (int a = 0; a < n; a++) {
        const float lv = acc[a * 32 + lane];
        sq2 = fmaf(lv, lv, sq2);
      }
    terms[((size_t)kk * 2 * N + 2 * i) * 32 + lane] = sq;
    terms[((size_t)kk * 2 * N + 2 * i + 1) * 32 + lane] = sq2;
    vals[((size_t)kk * N + i) * 32 + lane] = value;
  }
  __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// decide: one warp per instance
// ---------------------------------------------------------------------------
// acc + v[0] + v[1] + ... + v[m - 1] in that order (one fp32 accumulator, as the reference's loops),
// v[l] held by lane l.  The shuffles do not depend on the running sum: unrolled, they issue back to
// back and only the additions form the chain (a rolled loop pays a shuffle latency per element).
__device__ __forceinline__ float ordered_add(float acc, float v, int m) {
  if (m == 32) {
#pragma unroll
    for (int l = 0; l < 32; l++) acc += __shfl_sync(0xffffffffu, v, l);
  } else {
#pragma unroll
    for (int l = 0; l < 32; l++) {
      const float cur = __shfl_sync(0xffffffffu, v, l);
      if (l < m) acc += cur;
    }
  }
  return acc;
}

// ILQSolver::TotalCosts (src/ilq_solver.cpp:220-257) from the per-step values an eval block left
// behind: ordered over k, first extreme wins.
__device__ __forceinline__ void ls_total_costs(const DevDesc& d, const Slab& s, int b, const float* vals_block,
                                               int item_lane, int lane) {
  // All lanes of the (converged) warp fetch 32 time steps of a player's values at once; the ordered
  // scan (single fp32 accumulator / strict first-extreme compare, as the reference) then runs over
  // shuffled registers instead of a chain of dependent global loads.
  const int N = d.N;
  for (int i = 0; i < N; i++) {
    const int cs = d.cost_structure[i];
    float total = cs == ILQG_COST_SUM ? 0.f : cs == ILQG_COST_MAX ? -INFINITY : INFINITY;
    int te = s.te_new[(size_t)b * N + i];
    for (int k0 = 0; k0 < d.T; k0 += 128) {
      // four rounds of loads in flight ahead of the scan
      float v[4];
#pragma unroll
      for (int r = 0; r < 4; r++)
        v[r] = k0 + 32 * r + lane < d.T ? vals_block[((size_t)(k0 + 32 * r + lane) * N + i) * 32 + item_lane] : 0.f;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int m = min(32, d.T - k0 - 32 * r);
        if (m <= 0) break;
        if (cs == ILQG_COST_SUM) {
          total = ordered_add(total, v[r], m);
        } else {
#pragma unroll
          for (int l = 0; l < 32; l++) {
            const float cur = __shfl_sync(0xffffffffu, v[r], l);
            if (l < m && ((cs == ILQG_COST_MAX && cur > total) || (cs == ILQG_COST_MIN && cur < total))) {
              total = cur;
              te = k0 + 32 * r + l;
            }
          }
        }
      }
    }
    if (lane == 0) {
      s.total_costs[(size_t)b * N + i] = total;
      s.te_new[(size_t)b * N + i] = te;
    }
  }
}

// CheckArmijoCondition, src/ilq_solver.cpp:350-362
__device__ __forceinline__ bool ls_armijo(const DevParams& p, float last_merit, float merit, float ed, int j) {
  float step = p.initial_alpha_scaling;
  for (int jj = 0; jj < j; jj++) step *= p.geometric_alpha_scaling;
  const float scaled_expected_decrease = p.expected_decrease_fraction * step * ed;
  return last_merit - merit >= scaled_expected_decrease;
}

constexpr int KDEC_WARPS = 4;

__global__ void __launch_bounds__(KDEC_WARPS * 32)
k_ls_decide(const __grid_constant__ DevDesc d, const DevParams p, Slab s, LsScratch ls, int mode, int cur_q,
            int q_offset, int dst_q /* the queue still-open games are appended to */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * KDEC_WARPS + warp;
  int b;
  if (mode == LS_MODE_FRESH) {
    b = w;
    if (b >= s.B || s.status[b] != ILQG_STATUS_RUNNING || s.ls_next_j[b] != 0) return;
  } else {
    if (w >= ls.cap || q_offset + w >= ls.counts[cur_q]) return;
    b = ls.pend[cur_q][q_offset + w];
  }
  const int T = d.T, n = d.n, M = d.M, max_bt = p.max_backtracking_steps;
  const int j0 = s.ls_next_j[b];
  const bool fresh = mode == LS_MODE_FRESH;
  const int W = fresh ? ls.JA : ls.JB;
  const size_t base = fresh ? (size_t)b * ls.JA : (size_t)w * ls.JB;
  const float lm = s.last_merit[b], ed = s.expected_decrease[b];
  bool exhausted = false;
  int acc_jj = -1, acc_j = -1;
  float acc_merit = 0.f;
  if (!p.linesearch) {
    acc_jj = 0;  // ModifyLQStrategies returns after the first rollout (:322)
    acc_j = j0;
  } else {
    // The reference's loop looks at the candidates one after the other; here each lane takes one:
    // its merit (the ordered (k, i) sum of the terms k_ls_merit left -- a single
    // running fp32 accumulator as in src/ilq_solver.cpp:416-430; consecutive candidates sit in
    // consecutive lanes of the term tiles, so the loads of a round coalesce), its Armijo test, and
    // a ballot finds the first candidate that passes.
    const int ncand = max(0, min(W, max_bt - j0));
    const int cnt = T * 2 * d.N;
    float last_merit = 0.f;  // merit of candidate ncand - 1 (for the absorbed shortcut)
    for (int r0 = 0; r0 < ncand && acc_jj < 0; r0 += 32) {
      const int c = r0 + lane;
      float merit = 0.f;
      if (ncand == 1) {
        // a lone candidate (the first window): the lanes fetch its terms together, 32 per round,
        // and the ordered sum runs over shuffled values instead of dependent loads
        const size_t item = base;
        const float* terms = ls.terms + (item / ls.lpw) * (size_t)cnt * 32 + item % ls.lpw;
        float acc = 0.f;
        for (int e0 = 0; e0 < cnt; e0 += 256) {
          float v[8];  // eight rounds of loads in flight ahead of the ordered sum
#pragma unroll
          for (int r = 0; r < 8; r++) v[r] = e0 + 32 * r + lane < cnt ? terms[(size_t)(e0 + 32 * r + lane) * 32] : 0.f;
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int m = min(32, cnt - e0 - 32 * r);
            if (m > 0) acc = ordered_add(acc, v[r], m);
          }
        }
        merit = 0.5 * acc;
      } else if (c < ncand) {
        const size_t item = base + c;
        const float* terms = ls.terms + (item / ls.lpw) * (size_t)cnt * 32 + item % ls.lpw;
        // 24 loads in flight per trip: the adds are a dependent chain, the loads are not
        float acc = 0.f;
        int e = 0;
        for (; e + 24 <= cnt; e += 24) {
          float v[24];
#pragma unroll
          for (int u = 0; u < 24; u++) v[u] = terms[(size_t)(e + u) * 32];
#pragma unroll
          for (int u = 0; u < 24; u++) acc += v[u];
        }
        for (; e < cnt; e++) acc += terms[(size_t)e * 32];
        merit = 0.5 * acc;
      }
      const bool pass = c < ncand && ls_armijo(p, lm, merit, ed, j0 + c);
      const unsigned hits = __ballot_sync(0xffffffffu, pass);
      if (hits) {
        const int first = __ffs(hits) - 1;
        acc_jj = r0 + first;
        acc_j = j0 + acc_jj;
        acc_merit = __shfl_sync(0xffffffffu, merit, first);
      }
      if (ncand - 1 >= r0 && ncand - 1 < r0 + 32) last_merit = __shfl_sync(0xffffffffu, merit, ncand - 1 - r0);
    }
    const int jj = ncand;  // candidates examined when none passed
    if (acc_jj < 0 && jj > 0 && j0 + jj < max_bt && ls.absorbed[base + jj - 1]) {
      // the deeper candidates repeat rollout jj-1 exactly: finish the backtracking loop on its merit
      const float merit = last_merit;
      for (int j = j0 + jj; j < max_bt; j++)
        if (ls_armijo(p, lm, merit, ed, j)) {
          acc_jj = jj - 1;
          acc_j = j;
          acc_merit = merit;
          break;
        }
      if (acc_jj < 0) exhausted = true;
    }
  }
  if (acc_jj >= 0) {
    const int j = acc_j;
    const size_t item = base + acc_jj;
    const int cur = s.op_cur[b], scur = s.st_cur[b];
    // accepted candidate -> operating point (already there for the lone fresh candidate)
    if (!(fresh && ls.JA == 1)) {
      float* dxs = s.op_xs[1 - cur] + (size_t)b * T * n;
      const float* sxs = ls.traj_xs + item * T * n;
      float* dus = s.op_us[1 - cur] + (size_t)b * T * M;
      const float* sus = ls.traj_us + item * T * M;
      if ((T * n) % 4 == 0) {
#pragma unroll 8
        for (int e = lane; e < T * n / 4; e += 32)
          reinterpret_cast<float4*>(dxs)[e] = reinterpret_cast<const float4*>(sxs)[e];
      } else {
#pragma unroll 8
        for (int e = lane; e < T * n; e += 32) dxs[e] = sxs[e];
      }
#pragma unroll 8
      for (int e = lane; e < T * M; e += 32) dus[e] = sus[e];
    }
    ls_total_costs(d, s, b, ls.vals + (item / ls.lpw) * T * d.N * 32, (int)(item % ls.lpw), lane);
    // the scaled LQ strategies become current (ScaleAlphas, src/ilq_solver.cpp:66-72,314,339)
    float* alpha = s.st_a[1 - scur] + (size_t)b * T * M;
    const float s0 = p.initial_alpha_scaling, rho = p.geometric_alpha_scaling;
    for (int e0 = 0; e0 < T * M; e0 += 256) {
      float al[8];
#pragma unroll
      for (int r = 0; r < 8; r++) al[r] = e0 + 32 * r + lane < T * M ? alpha[e0 + 32 * r + lane] * s0 : 0.f;
      for (int jj = 0; jj < j; jj++) {
#pragma unroll
        for (int r = 0; r < 8; r++) al[r] *= rho;
      }
#pragma unroll
      for (int r = 0; r < 8; r++)
        if (e0 + 32 * r + lane < T * M) alpha[e0 + 32 * r + lane] = al[r];
    }
    __syncwarp();
    if (lane == 0) {
      float step = s0;
      for (int jj = 0; jj < j; jj++) step *= rho;
      const bool with_merit = p.linesearch != 0;
      // HasConverged, ilq_solver.h:126-130
      const bool converged = with_merit && (acc_merit <= lm) && fabsf(lm - acc_merit) < p.convergence_tolerance;
      const int itn = s.iters[b] + 1;
      s.iters[b] = itn;
      s.backtracks[b] += j + 1;
      s.op_cur[b] = 1 - cur;
      s.st_cur[b] = 1 - scur;
      if (with_merit) s.last_merit[b] = acc_merit;
      s.step[b] = step;
      s.ls_next_j[b] = 0;
      if (fresh) s.queued_flag[b] = 0;
      if (converged && !p.disable_convergence_exit)
        s.status[b] = ILQG_STATUS_CONVERGED;
      else if (itn >= p.max_solver_iters)
        s.status[b] = ILQG_STATUS_MAX_ITERS;
    }
  } else if (lane == 0) {
    const int jn = j0 + W;
    if (jn >= max_bt || exhausted) {
      // ModifyLQStrategies returns false (:345-347); the log's final iterate stays current
      s.iters[b] += 1;
      s.backtracks[b] += max_bt + 1;  // the reference's last rollout is never evaluated
      s.status[b] = ILQG_STATUS_LINESEARCH_FAILED;
      s.ls_next_j[b] = 0;
    } else {
      s.ls_next_j[b] = jn;
      s.queued_flag[b] = 1;
      const int q = atomicAdd(&ls.counts[dst_q], 1);
      ls.pend[dst_q][q] = b;
      ls.slot[b] = q;
    }
  }
}

// Solve() prologue bookkeeping (src/ilq_solver.cpp:86-107) after the LS_MODE_BEGIN rollout
__global__ void __launch_bounds__(KDEC_WARPS * 32)
k_begin_finalize(const __grid_constant__ DevDesc d, const DevParams p, Slab s, LsScratch ls) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * KDEC_WARPS + warp;
  if (b >= s.B || s.al_state[b] == 2) return;
  const int T = d.T, n = d.n, M = d.M, N = d.N;
  // current strategies <- problem strategies
  const float4* pP = reinterpret_cast<const float4*>(s.prob_P + (size_t)b * T * M * n);
  float4* P0 = reinterpret_cast<float4*>(s.st_P[0] + (size_t)b * T * M * n);
  if ((T * M * n) % 4 == 0) {
    for (int e = lane; e < T * M * n / 4; e += 32) P0[e] = pP[e];
  } else {
    for (int e = lane; e < T * M * n; e += 32)
      (s.st_P[0] + (size_t)b * T * M * n)[e] = (s.prob_P + (size_t)b * T * M * n)[e];
  }
  for (int e = lane; e < T * M; e += 32) (s.st_a[0] + (size_t)b * T * M)[e] = (s.prob_a + (size_t)b * T * M)[e];
  ls_total_costs(d, s, b, ls.vals + (size_t)(b / ls.lpw) * T * N * 32, b % ls.lpw, lane);
  __syncwarp();
  if (lane < N) s.te_quad[(size_t)b * N + lane] = s.te_new[(size_t)b * N + lane];
  if (lane == 0) {
    s.op_cur[b] = 0;
    s.st_cur[b] = 0;
    s.iters[b] = 0;
    s.ls_next_j[b] = 0;
    s.status[b] = p.max_solver_iters > 0 ? ILQG_STATUS_RUNNING : ILQG_STATUS_MAX_ITERS;
  }
}

}  // namespace ilqg
