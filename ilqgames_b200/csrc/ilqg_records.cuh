// ilqg_records.cuh -- K_lq v2: fused ComputeLinearization + ComputeCostQuadraticization
// (src/ilq_solver.cpp:437-455, 471-490) producing the dense LQ records.
//
// A block owns 32 consecutive (instance, timestep) records.
//   phase 1  one ROLE per warp, one record per lane: warps [0, N) run player i's cost /
//            constraint records, warp N runs every subsystem's Jacobian.  Instead of touching a
//            dense Hessian each role appends its `+= v` updates (offset in the record, value) to
//            a lane-minor list in shared memory, in the reference's accumulation order.  All lanes
//            of a warp execute the same record kinds, so the scalar cost code runs at full SIMT
//            width (v1 ran it on 3 of 32 lanes).
//   phase 2  one record at a time per warp: write the record template (zeros, A = I, Q_i =
//            state_reg * I, R = control_reg * I) into a shared staging buffer, let lane `role`
//            replay its list in order, then stream the record to HBM with 128-bit stores.
#pragma once
#include "ilqg_kernels.cuh"

namespace ilqg {

struct ListSink {
  unsigned short* off;  // [E][32] lane-minor
  float* val;           // [E][32]
  int lane, cnt, cap;
  int base_H, ld, base_G;
  __device__ __forceinline__ void push(int o, float v) {
    if (cnt < cap) {
      off[cnt * 32 + lane] = (unsigned short)o;
      val[cnt * 32 + lane] = v;
    }
    cnt++;
  }
  __device__ __forceinline__ void H(int r, int c, float v) { push(base_H + r * ld + c, v); }
  __device__ __forceinline__ void G(int i, float v) { push(base_G + i, v); }
};

struct LinListSink {
  ListSink* ls;
  int offA, offB, n, M;
  __device__ __forceinline__ void addA(int r, int c, float v) { ls->push(offA + r * n + c, v); }
  __device__ __forceinline__ void addB(int r, int c, float v) { ls->push(offB + r * M + c, v); }
};

// shared memory (bytes): xu[n+M][32] floats | cnt[NR][32] ints | val[NR][E][32] floats |
//                        off[NR][E][32] u16 | rec[NR][rec] floats      (NR = N + 1 roles = warps)
__host__ __device__ inline size_t klq2_smem_bytes(int n, int M, int N, int E, int rec) {
  const int NR = N + 1;
  return sizeof(float) * ((size_t)(n + M) * 32 + (size_t)NR * 32 + (size_t)NR * E * 32 + (size_t)NR * rec) +
         sizeof(unsigned short) * (size_t)NR * E * 32;
}

__global__ void __launch_bounds__(160)
k_linearize_quadraticize_v2(const __grid_constant__ DevDesc d, Slab s, int only_running, int E) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = d.n, M = d.M, N = d.N, T = d.T, NR = N + 1;
  float* xu = smem;                                   // [n+M][32]
  int* cnts = reinterpret_cast<int*>(xu + (n + M) * 32);  // [NR][32]
  float* vals = reinterpret_cast<float*>(cnts + NR * 32);  // [NR][E][32]
  float* recs = vals + (size_t)NR * E * 32;           // [NR][rec]
  unsigned short* offs = reinterpret_cast<unsigned short*>(recs + (size_t)NR * d.rec);  // [NR][E][32]

  const long long first = (long long)blockIdx.x * 32;
  const long long total = (long long)s.B * T;
  const long long w = first + lane;
  const bool in_range = w < total;
  const int b = in_range ? (int)(w / T) : 0, k = in_range ? (int)(w % T) : 0;
  const bool live = in_range && (!only_running || instance_iterates(s, b));
  if (!__syncthreads_or(live)) return;

  // ---- stage x, u of the 32 records (each role reads all of it) ----
  {
    const int cur = in_range ? s.op_cur[b] : 0;
    const float* xs = s.op_xs[cur] + ((size_t)b * T + k) * n;
    const float* us = s.op_us[cur] + ((size_t)b * T + k) * M;
    for (int e = warp; e < n + M; e += NR) xu[e * 32 + lane] = in_range ? (e < n ? xs[e] : us[e - n]) : 0.f;
  }
  __syncthreads();

  // ---- phase 1: role = warp, record = lane ----
  {
    ListSink sink;
    sink.off = offs + (size_t)warp * E * 32;
    sink.val = vals + (size_t)warp * E * 32;
    sink.lane = lane;
    sink.cnt = 0;
    sink.cap = E;
    const float* x = xu + lane;            // element idx at x[idx * 32]
    const float* u = xu + n * 32 + lane;
    if (warp < N) {
      const int i = warp;
      const bool full = d.cost_structure[i] == ILQG_COST_SUM || (live && s.te_quad[(size_t)b * N + i] == k);
      const float mu = live ? s.mu[b] : 0.f;
      for (int c = d.cost_begin[i]; c < d.cost_begin[i + 1]; c++) {
        const DevCost& cd = d.cost[c];
        const bool is_con = cd.slot >= 0;
        if (!full && (cd.arg < 0 || is_con)) continue;  // QuadraticizeControlCosts
        const float lambda =
            (is_con && live) ? s.lambdas[((size_t)b * d.num_constraints + cd.slot) * T + s.lambda_index[k]] : 0.f;
        if (cd.arg < 0) {
          sink.base_H = d.offQ + i * n * n;
          sink.ld = n;
          sink.base_G = d.offl + i * n;
          quadraticize_record_sink<true, 32, false>(d, cd, x, n, lambda, mu, sink);
        } else {
          const int mj = d.udim[cd.arg];
          sink.base_H = d.offR + d.pair_Roff[cd.pair];
          sink.ld = mj;
          sink.base_G = d.offr + d.pair_roff[cd.pair];
          quadraticize_record_sink<true, 32, false>(d, cd, u + d.uoff[cd.arg] * 32, mj, lambda, mu, sink);
        }
      }
    } else {
      LinListSink lin{&sink, d.offA, d.offB, n, M};
      for (int sidx = 0; sidx < d.num_subsystems; sidx++) subsystem_linearize_sink<32>(d, d.sub[sidx], x, u, lin);
    }
    cnts[warp * 32 + lane] = sink.cnt < E ? sink.cnt : E;
  }
  __syncthreads();

  // ---- phase 2: warp w assembles records w, w + NR, ... ----
  float* rec = recs + (size_t)warp * d.rec;
  for (int r = warp; r < 32; r += NR) {
    const long long wr = first + r;
    if (wr >= total) break;
    const int br = (int)(wr / T);
    if (only_running && !instance_iterates(s, br)) continue;
    // template: LinearDynamicsApproximation ctor A = I, B = 0 (linear_dynamics_approximation.h:65-70);
    // QuadraticCostApproximation(xdim, state_reg) / SingleCostApproximation(udim, control_reg)
    for (int e = lane; e < d.rec / 4; e += 32) reinterpret_cast<float4*>(rec)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    for (int a = lane; a < n; a += 32) {
      rec[d.offA + a * n + a] = 1.f;
      for (int i = 0; i < N; i++) rec[d.offQ + (i * n + a) * n + a] = d.state_reg[i];
    }
    if (lane < d.num_pairs) {
      const int mj = d.udim[d.pair_j[lane]];
      for (int a = 0; a < mj; a++) rec[d.offR + d.pair_Roff[lane] + a * mj + a] = d.control_reg[d.pair_i[lane]];
    }
    __syncwarp();
    if (lane < NR) {
      const int cnt = cnts[lane * 32 + r];
      const unsigned short* o = offs + (size_t)lane * E * 32 + r;
      const float* v = vals + (size_t)lane * E * 32 + r;
      for (int e = 0; e < cnt; e++) rec[o[e * 32]] += v[e * 32];
    }
    __syncwarp();
    float4* dst = reinterpret_cast<float4*>(s.rec + (size_t)wr * d.rec);
    const float4* src = reinterpret_cast<const float4*>(rec);
    for (int e = lane; e < d.rec / 4; e += 32) dst[e] = src[e];
    __syncwarp();
  }
}

}  // namespace ilqg
