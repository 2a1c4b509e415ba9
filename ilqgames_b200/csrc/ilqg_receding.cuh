// ilqg_receding.cuh -- Problem::SetUpNextRecedingHorizon (src/problem.cpp:64-186) on the device:
// the warm start of every game (Problem::operating_point_ / strategies_) is re-based in place, so
// a receding-horizon caller (src/receding_horizon_simulator.cpp:65-137) never moves plans across
// the PCIe bus -- only the measured states come in.
//
// The time bookkeeping of SyncToExistingProblem (:64-125) and IntegrateToNextTimeStep
// (src/multi_player_integrable_system.cpp:113-143) is double arithmetic on values shared by the
// whole batch; ilqg_setup_next_receding_horizon (ilqg_abi.cu) does it once on the host and hands
// the results over in RhTimes.  One warp per game does the per-state part.
#pragma once
#include "ilqg_kernels.cuh"

namespace ilqg {

struct RhTimes {
  int itn_timestep;        // IntegrateToNextTimeStep's current_timestep
  float frac;              // interpolation weight of xs[itn_timestep]
  float itn_dt_half;       // (remaining time of this step) / 2, RK4 substep
  int integrate_from, integrate_to;  // Integrate(current_timestep + 1, last_integration_timestep, ...)
  float dt_half;           // kTimeStep / 2
  int ego_dim;             // Stitch: leading states taken from the nearest plan state
  int position_distance;   // 1: DistanceBetween = first subsystem's (x, y) (ConcatenatedDynamicalSystem, Air3D::DistanceBetween
                           // air_3d.h:151-157); 2: a Dubins ego's whole state; 0: squared 2-norm of the full state
};

constexpr int KRH_WARPS = 4;

__global__ void __launch_bounds__(KRH_WARPS * 32)
k_receding_horizon(const __grid_constant__ DevDesc d, Slab s, const float* __restrict__ x_meas, RhTimes r) {
  __shared__ float sx[KRH_WARPS][ILQG_MAX_XDIM], sref[KRH_WARPS][ILQG_MAX_XDIM], su[KRH_WARPS][ILQG_MAX_UDIM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * KRH_WARPS + warp;
  if (b >= s.B) return;
  const int T = d.T, n = d.n, M = d.M;
  float* x = sx[warp];
  float* ref = sref[warp];
  float* u = su[warp];
  float* pxs = s.prob_xs + (size_t)b * T * n;
  float* pus = s.prob_us + (size_t)b * T * M;
  float* pP = s.prob_P + (size_t)b * T * M * n;
  float* pa = s.prob_a + (size_t)b * T * M;

  // Strategy::operator() (strategy.h:73-76): u = (u_ref - P (x - x_ref)) - alpha, at time step kk
  auto controls = [&](int kk) {
    for (int c = lane; c < M; c += 32) {
      float acc = 0.f;
      for (int a = 0; a < n; a++) acc += pP[((size_t)kk * M + c) * n + a] * (x[a] - ref[a]);
      u[c] = (pus[(size_t)kk * M + c] - acc) - pa[(size_t)kk * M + c];
    }
    __syncwarp();
  };
  // MultiPlayerDynamicalSystem::Integrate: one lane per subsystem (block-separable dynamics)
  auto integrate = [&](const float* from, const float* uu, float* to, float dt_half) {
    if (lane < d.num_subsystems) {
      const DevSubsystem& sub = d.sub[lane];
      const int xd = subsystem_xdim(sub.kind);
      float xl[6];
#pragma unroll
      for (int a = 0; a < 6; a++) xl[a] = a < xd ? from[sub.x_offset + a] : 0.f;
      float us[4];
#pragma unroll
      for (int q = 0; q < 4; q++) us[q] = q < sub.nu ? uu[sub.ucol[q]] : 0.f;
      subsystem_integrate(sub, dt_half, xl, us);
#pragma unroll
      for (int a = 0; a < 6; a++)
        if (a < xd) to[sub.x_offset + a] = xl[a];
    }
    __syncwarp();
  };

  // ---- IntegrateToNextTimeStep (:113-143) ----
  for (int a = lane; a < n; a += 32) {
    x[a] = x_meas[(size_t)b * n + a];
    ref[a] = r.itn_timestep + 1 < T
                 ? r.frac * pxs[(size_t)r.itn_timestep * n + a] + (float)(1.0 - r.frac) * pxs[(size_t)(r.itn_timestep + 1) * n + a]
                 : pxs[(size_t)(T - 1) * n + a];
  }
  __syncwarp();
  controls(r.itn_timestep);
  integrate(x, u, x, r.itn_dt_half);
  // ---- Integrate(current_timestep + 1, last_integration_timestep, ...) (:96-111) ----
  for (int kk = r.integrate_from; kk < r.integrate_to; kk++) {
    for (int a = lane; a < n; a += 32) ref[a] = pxs[(size_t)kk * n + a];
    __syncwarp();
    controls(kk);
    integrate(x, u, x, r.dt_half);
  }
  // ---- nearest state of the existing plan (:101-110), first minimum wins ----
  float best = 3.4e38f;
  int first = T;
  for (int kk = lane; kk < T; kk += 32) {
    const float* a = pxs + (size_t)kk * n;
    float dist = 0.f;
    if (r.position_distance == 1) {
      const int o = d.sub[0].x_offset;
      const float dx = x[o] - a[o], dy = x[o + 1] - a[o + 1];
      dist = dx * dx + dy * dy;
    } else if (r.position_distance == 2) {
      // SinglePlayerDubinsCar has no DistanceBetween of its own: the base class's squared 2-norm of the
      // whole subsystem state, heading included (single_player_dynamical_system.h:68-71)
      const int o = d.sub[0].x_offset;
      for (int q = 0; q < 3; q++) dist += (x[o + q] - a[o + q]) * (x[o + q] - a[o + q]);
    } else {
      for (int q = 0; q < n; q++) dist += (x[q] - a[q]) * (x[q] - a[q]);
    }
    if (dist < best) { best = dist; first = kk; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int of = __shfl_xor_sync(0xffffffffu, first, o);
    if (ob < best || (ob == best && of < first)) { best = ob; first = of; }
  }
  if (first >= T) first = 0;  // every distance was NaN: min_element returns begin()
  // ---- x0_ = Stitch(nearest, x) (:117; concatenated_dynamical_system.h:75-85) ----
  for (int a = lane; a < n; a += 32) s.x0[(size_t)b * n + a] = a < r.ego_dim ? pxs[(size_t)first * n + a] : x[a];
  __syncwarp();
  // ---- SetUpNextRecedingHorizon (:127-186): shift the plan to start at `first` ... ----
  const int kept = T - first;
  if (first > 0) {
    for (int kk = 0; kk < kept; kk++) {  // ascending: a source row is always ahead of every row written so far
      const int src = kk + first;
      for (int e = lane; e < n; e += 32) pxs[(size_t)kk * n + e] = pxs[(size_t)src * n + e];
      for (int e = lane; e < M; e += 32) {
        pus[(size_t)kk * M + e] = pus[(size_t)src * M + e];
        pa[(size_t)kk * M + e] = pa[(size_t)src * M + e];
      }
      for (int e = lane; e < M * n; e += 32) pP[(size_t)kk * M * n + e] = pP[(size_t)src * M * n + e];
      __syncwarp();
    }
  }
  // ---- ... and extend it to the horizon with zero controls and strategies ----
  for (int kk = kept; kk < T; kk++) {
    for (int e = lane; e < M; e += 32) {
      pus[(size_t)kk * M + e] = 0.f;
      pa[(size_t)kk * M + e] = 0.f;
    }
    for (int e = lane; e < M * n; e += 32) pP[(size_t)kk * M * n + e] = 0.f;
    for (int e = lane; e < M; e += 32) u[e] = pus[(size_t)(kk - 1) * M + e];
    __syncwarp();
    integrate(pxs + (size_t)(kk - 1) * n, u, pxs + (size_t)kk * n, r.dt_half);
  }
}

// ---------------------------------------------------------------------------------------------
// MultiPlayerIntegrableSystem::Integrate(Time t0, Time t, x0, operating_point, strategies)
// (src/multi_player_integrable_system.cpp:54-83): carry a measured state along the stored plan --
// what src/receding_horizon_simulator.cpp:100-102,126-128 runs between solves.  The plan is read
// only.  As above, the double time arithmetic is done once on the host (IpTimes); one warp per game.
struct IpTimes {
  int to_next;                       // t0 is past the plan's start: IntegrateToNextTimeStep first (:75-76)
  int itn_timestep;                  // its time step, interpolation weight, RK4 substep and trip count
  float frac, itn_dt_half;
  int itn_substeps;
  int integrate_from, integrate_to;  // whole time steps (:79-80)
  float dt_half;
  int step_substeps;
  int prior_timestep;                // IntegrateFromPriorTimeStep (:145-171)
  float prior_dt_half;
  int prior_substeps;
};

__global__ void __launch_bounds__(KRH_WARPS * 32)
k_integrate_plan(const __grid_constant__ DevDesc d, Slab s, float* __restrict__ x_io, IpTimes r) {
  __shared__ float sx[KRH_WARPS][ILQG_MAX_XDIM], sref[KRH_WARPS][ILQG_MAX_XDIM], su[KRH_WARPS][ILQG_MAX_UDIM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * KRH_WARPS + warp;
  if (b >= s.B) return;
  const int T = d.T, n = d.n, M = d.M;
  float* x = sx[warp];
  float* ref = sref[warp];
  float* u = su[warp];
  const float* pxs = s.prob_xs + (size_t)b * T * n;
  const float* pus = s.prob_us + (size_t)b * T * M;
  const float* pP = s.prob_P + (size_t)b * T * M * n;
  const float* pa = s.prob_a + (size_t)b * T * M;

  // Strategy::operator() (strategy.h:73-76) at time step kk, then one Integrate over 2 * dt_half
  auto advance = [&](int kk, float dt_half, int substeps) {
    for (int c = lane; c < M; c += 32) {
      float acc = 0.f;
      for (int a = 0; a < n; a++) acc += pP[((size_t)kk * M + c) * n + a] * (x[a] - ref[a]);
      u[c] = (pus[(size_t)kk * M + c] - acc) - pa[(size_t)kk * M + c];
    }
    __syncwarp();
    if (lane < d.num_subsystems) {
      const DevSubsystem& sub = d.sub[lane];
      const int xd = subsystem_xdim(sub.kind);
      float xl[6];
#pragma unroll
      for (int a = 0; a < 6; a++) xl[a] = a < xd ? x[sub.x_offset + a] : 0.f;
      float us[4];
#pragma unroll
      for (int q = 0; q < 4; q++) us[q] = q < sub.nu ? u[sub.ucol[q]] : 0.f;
      subsystem_integrate(sub, dt_half, xl, us, substeps);
#pragma unroll
      for (int a = 0; a < 6; a++)
        if (a < xd) x[sub.x_offset + a] = xl[a];
    }
    __syncwarp();
  };
  auto reference_row = [&](int kk) {
    for (int a = lane; a < n; a += 32) ref[a] = pxs[(size_t)kk * n + a];
    __syncwarp();
  };

  for (int a = lane; a < n; a += 32) x[a] = x_io[(size_t)b * n + a];
  __syncwarp();
  if (r.to_next) {  // :113-143, against the interpolated reference state
    for (int a = lane; a < n; a += 32)
      ref[a] = r.itn_timestep + 1 < T
                   ? r.frac * pxs[(size_t)r.itn_timestep * n + a] + (float)(1.0 - r.frac) * pxs[(size_t)(r.itn_timestep + 1) * n + a]
                   : pxs[(size_t)(T - 1) * n + a];
    __syncwarp();
    advance(r.itn_timestep, r.itn_dt_half, r.itn_substeps);
  }
  for (int kk = r.integrate_from; kk < r.integrate_to; kk++) {  // :85-111
    reference_row(kk);
    advance(kk, r.dt_half, r.step_substeps);
  }
  reference_row(r.prior_timestep);
  advance(r.prior_timestep, r.prior_dt_half, r.prior_substeps);
  for (int a = lane; a < n; a += 32) x_io[(size_t)b * n + a] = x[a];
}

}  // namespace ilqg
