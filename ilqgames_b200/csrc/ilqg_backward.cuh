// ilqg_backward.cuh -- K_bwd v2: LQFeedbackSolver::Solve (src/lq_feedback_solver.cpp:71-244)
// with one HALF-WARP (16 lanes) per game instance, for state dimensions that are a multiple
// of 4 and a uniform control dimension m = M / N (ThreePlayerIntersection 16/6/3,
// RoundaboutMerging 24/8/4, the n = 12 variant).  Other shapes use k_lq_backward.
//
//  * Z_i (N x n x n), F, A / W^T and the small operands live in shared memory for the whole
//    sweep; every n x n product is an X^T Y form so both operands are read as contiguous row
//    segments (128-bit, conflict-free) and each lane keeps an (n/4) x (n/4) accumulator tile.
//  * The LQ record of a timestep is consumed piecewise: [A|B] and [l|R|r] go to shared memory,
//    each lane pulls its own tile of Q_i straight from global memory into registers.
//  * S X = Y: lane c owns right-hand-side column c (a column of P); all lanes eliminate S
//    redundantly in registers (LU, partial pivoting, one reciprocal per pivot).
//  * ExpectedDecrease (src/ilq_solver.cpp:364-398) is accumulated in the same backward sweep
//    through the adjoint identity  sum_k dx_k . g_k = sum_k beta_k . p_{k+1},
//    p_k = g_k + A_k^T p_{k+1},  g_k = sum_i Q_i l_i  (dx_0 = 0, dx_{k+1} = A_k dx_k + beta_k),
//    so the records are read exactly once.  delta_xs themselves (an optional output of
//    LQFeedbackSolver::Solve) are produced by k_delta_xs when asked for.
#pragma once
#include "ilqg_records.cuh"

#ifndef ILQG_MM_UNROLL
#define ILQG_MM_UNROLL 4
#endif

namespace ilqg {

constexpr int kMmUnroll = ILQG_MM_UNROLL;

__host__ __device__ constexpr int r4(int v) { return (v + 3) & ~3; }

template <int K>
__device__ __forceinline__ void ldvec(const float* p, float (&out)[K]) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int i = 0; i < K / 4; i++) {
      const float4 v = reinterpret_cast<const float4*>(p)[i];
      out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
    }
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int i = 0; i < K / 2; i++) {
      const float2 v = reinterpret_cast<const float2*>(p)[i];
      out[2 * i] = v.x; out[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; i++) out[i] = p[i];
  }
}

template <int K>
__device__ __forceinline__ void ldgvec(const float* p, float (&out)[K]) {  // read-only global
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int i = 0; i < K / 4; i++) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
    }
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int i = 0; i < K / 2; i++) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p) + i);
      out[2 * i] = v.x; out[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; i++) out[i] = __ldg(p + i);
  }
}

template <int K>
__device__ __forceinline__ void stvec(float* p, const float (&in)[K]) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int i = 0; i < K / 4; i++)
      reinterpret_cast<float4*>(p)[i] = make_float4(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]);
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int i = 0; i < K / 2; i++) reinterpret_cast<float2*>(p)[i] = make_float2(in[2 * i], in[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < K; i++) p[i] = in[i];
  }
}

// acc[TR][TC] = sum_q X[q][a0 + .] * Y[q][c0 + .]   (X: ldx floats per row, Y: ldy)
//
// Even TC: the inner products run on FFMA2 (`fma.rn.f32x2`, new on sm_100): one instruction does
// acc[i][j], acc[i][j+1] += x[i] * (y[j], y[j+1]) -- the SASS form takes the scalar x[i] as a
// broadcast operand, so no packing moves are needed and each half is an IEEE fma (results are bit
// for bit those of two FFMAs).  K_bwd is bound by instruction issue / per-warp latency, not by
// the FMA pipe (profiles/r01_schedule_experiments.md), and these loops were 38 % of its
// instructions.
template <int QN, int TR, int TC>
__device__ __forceinline__ void mm_tn(const float* __restrict__ X, int ldx, const float* __restrict__ Y,
                                      int ldy, float (&acc)[TR][TC]) {
  if constexpr (TC % 2 == 0) {
    float2 acc2[TR][TC / 2];
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TC / 2; j++) acc2[i][j] = make_float2(0.f, 0.f);
#pragma unroll kMmUnroll
    for (int q = 0; q < QN; q++) {
      float xr[TR], yr[TC];
      ldvec<TR>(X + q * ldx, xr);
      ldvec<TC>(Y + q * ldy, yr);
#pragma unroll
      for (int i = 0; i < TR; i++)
#pragma unroll
        for (int j = 0; j < TC / 2; j++)
          acc2[i][j] = __ffma2_rn(make_float2(xr[i], xr[i]), make_float2(yr[2 * j], yr[2 * j + 1]), acc2[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TC / 2; j++) {
        acc[i][2 * j] = acc2[i][j].x;
        acc[i][2 * j + 1] = acc2[i][j].y;
      }
  } else {
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TC; j++) acc[i][j] = 0.f;
    // partially unrolled: the kernel is instruction-fetch sensitive (stall_no_instruction in
    // profiles/r01_lq_backward.md); ILQG_MM_UNROLL q-steps per loop trip keep 2 x ILQG_MM_UNROLL
    // 128-bit loads in flight
#pragma unroll kMmUnroll
    for (int q = 0; q < QN; q++) {
      float xr[TR], yr[TC];
      ldvec<TR>(X + q * ldx, xr);
      ldvec<TC>(Y + q * ldy, yr);
#pragma unroll
      for (int i = 0; i < TR; i++)
#pragma unroll
        for (int j = 0; j < TC; j++) acc[i][j] = fmaf(xr[i], yr[j], acc[i][j]);
    }
  }
}

template <int NX, int MU, int NP>
struct HwSmem {
  static constexpr int m = MU / NP;
  static constexpr int Z = 0;                      // [NP][NX][NX]
  static constexpr int F = Z + NP * NX * NX;       // [NX][NX]
  static constexpr int AW = F + NX * NX;           // A, later W^T
  static constexpr int Bm = AW + NX * NX;          // [NX][MU]
  // (B_i^T Z_i)^T [NX][MU] lives in the head of the F buffer: it is dead once S X = Y is
  // solved, F is only formed after that.  Together with reading B in its stored layout when F is
  // formed (no transposed copy) this keeps an instance under 7 KB, i.e. 8 blocks per SM -- the
  // occupancy the 128-register budget allows; K_bwd's time follows its resident warps.
  static constexpr int BZt = F;
  static constexpr int P = Bm + r4(NX * MU);       // [MU][NX]   Y, then the solution P
  static constexpr int zeta = P + r4(MU * NX);     // [NP][NX]
  static constexpr int beta = zeta + r4(NP * NX);  // [NX]
  static constexpr int pv = beta + r4(NX);         // [NX] adjoint p_{k+1}
  static constexpr int pn = pv + r4(NX);           // [NX] scratch for p_k
  static constexpr int S = pn + r4(NX);            // [MU][MU]
  static constexpr int tv = S;                     // [NX] shares S (dead once the solve has loaded it)
  static_assert(r4(MU * MU) >= NX, "tv is overlaid on S");
  static constexpr int ya = S + r4(MU * MU);       // [MU] alpha column
  static constexpr int lrr = ya + r4(MU);          // [l | R | r] copied from the record (run-time size)
  // resident blocks per SM the shared-memory footprint allows (own-control pairs only), used as
  // the register cap in __launch_bounds__
  static constexpr int lrr_typ = r4(NP * NX + NP * m * m + NP * m);
  static constexpr int block_bytes = 4 * 4 * (lrr + lrr_typ) + 1024;  // + the per-block reservation
  static constexpr int fit = 228 * 1024 / block_bytes;               // 228 KB of shared memory per SM
  static constexpr int min_blocks = fit < 1 ? 1 : (fit > 12 ? 12 : fit);
};

constexpr int KHW_WARPS = 2;  // 4 instances per 64-thread block

// LU solve of S X = [Y | y_alpha] with S held redundantly in registers.  G column groups of
// right-hand sides per lane plus the alpha column.  PIVOT = false is the fast path used when
// the Gershgorin step has made S strictly column diagonally dominant (LU without pivoting is
// then stable, growth factor <= 2); PIVOT = true does partial pivoting with row swaps.
template <int MU, int G, bool PIVOT>
__device__ __forceinline__ void lu_solve(float (&Sm)[MU][MU], float (&y)[G][MU], float (&yal)[MU]) {
#pragma unroll
  for (int k = 0; k < MU; k++) {
    if constexpr (PIVOT) {
      int piv = k;
      float best = fabsf(Sm[k][k]);
#pragma unroll
      for (int r = k + 1; r < MU; r++) {
        const float v = fabsf(Sm[r][k]);
        if (v > best) { best = v; piv = r; }
      }
#pragma unroll
      for (int r = k + 1; r < MU; r++) {
        if (piv == r) {
#pragma unroll
          for (int c = 0; c < MU; c++) { const float t = Sm[k][c]; Sm[k][c] = Sm[r][c]; Sm[r][c] = t; }
#pragma unroll
          for (int g = 0; g < G; g++) { const float t = y[g][k]; y[g][k] = y[g][r]; y[g][r] = t; }
          const float t = yal[k]; yal[k] = yal[r]; yal[r] = t;
        }
      }
    }
    const float inv = 1.0f / Sm[k][k];
    Sm[k][k] = inv;
#pragma unroll
    for (int r = k + 1; r < MU; r++) {
      const float f = Sm[r][k] * inv;
#pragma unroll
      for (int c = k + 1; c < MU; c++) Sm[r][c] = fmaf(-f, Sm[k][c], Sm[r][c]);
#pragma unroll
      for (int g = 0; g < G; g++) y[g][r] = fmaf(-f, y[g][k], y[g][r]);
      yal[r] = fmaf(-f, yal[k], yal[r]);
    }
  }
#pragma unroll
  for (int r = MU - 1; r >= 0; r--) {
#pragma unroll
    for (int g = 0; g < G; g++) {
      float acc = y[g][r];
#pragma unroll
      for (int c = r + 1; c < MU; c++) acc = fmaf(-Sm[r][c], y[g][c], acc);
      y[g][r] = acc * Sm[r][r];
    }
    float acc = yal[r];
#pragma unroll
    for (int c = r + 1; c < MU; c++) acc = fmaf(-Sm[r][c], yal[c], acc);
    yal[r] = acc * Sm[r][r];
  }
}

template <int NX, int MU, int NP>
__global__ void __launch_bounds__(KHW_WARPS * 32, HwSmem<NX, MU, NP>::min_blocks)
k_lq_backward_hw(const __grid_constant__ DevDesc d, const DevParams p, Slab s, int only_running, Sel sel) {
  using L = HwSmem<NX, MU, NP>;
  constexpr int m = L::m;
  constexpr int TR = NX / 4, TC = NX / 4;   // Z tile per lane (16 lanes cover NX x NX)
  constexpr int G = (NX + 15) / 16;         // right-hand-side columns per lane in the solve
  constexpr int AB4 = (NX * NX + NX * MU) / 4, A4 = NX * NX / 4;
  constexpr int AB_PER = (AB4 + 15) / 16;
  constexpr int LRR4_MAX = (NP * NX + NP * NP * m * m + NP * NP * m + 3) / 4;
  constexpr int LRR_PER = (LRR4_MAX + 15) / 16;
  static_assert(NX % 4 == 0 && MU % NP == 0 && (NX * MU) % 4 == 0, "shape not supported by the half-warp kernel");
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int T = d.T;
  const int lrr_floats = d.rec - d.offl, lrr4 = lrr_floats / 4;
  const int per_inst = L::lrr + lrr_floats;

  bool active = false;
  const int b_own = sel_instance(s, sel, (blockIdx.x * KHW_WARPS + warp) * 2 + half, only_running, &active);
  const unsigned act_mask = __ballot_sync(0xffffffffu, active);
  if (act_mask == 0) return;
  // an inactive half shadows its partner (valid loads, no stores) so the warp stays convergent
  const int b_partner = __shfl_xor_sync(0xffffffffu, b_own, 16);
  const int b = active ? b_own : b_partner;

  float* sm = smem + (size_t)((warp * 2 + half)) * per_inst;
  float* Z = sm + L::Z;
  float* F = sm + L::F;
  float* AW = sm + L::AW;
  float* Bm = sm + L::Bm;
  float* BZt = sm + L::BZt;
  float* P = sm + L::P;
  float* zeta = sm + L::zeta;
  float* tv = sm + L::tv;
  float* beta = sm + L::beta;
  float* pv = sm + L::pv;
  float* pn = sm + L::pn;
  float* S = sm + L::S;
  float* ya = sm + L::ya;
  float* lrr = sm + L::lrr;
  const float* lvec = lrr;                          // [NP][NX]
  const float* Rk = lrr + (d.offR - d.offl);
  const float* rk = lrr + (d.offr - d.offl);

  if (active && l16 < NP) s.te_quad[(size_t)b * NP + l16] = s.te_new[(size_t)b * NP + l16];

  const int cand = 1 - s.st_cur[b];
  float* outP = s.st_P[cand] + (size_t)b * T * MU * NX;
  float* outa = s.st_a[cand] + (size_t)b * T * MU;
  const float* recb = s.rec + (size_t)b * T * d.rec;

  const int a0 = (l16 >> 2) * TR, c0 = (l16 & 3) * TC;

  // register prefetch of the next record's [A|B] and [l|R|r]
  float4 preAB[AB_PER], preL[LRR_PER];
  auto prefetch = [&](int k) {
    const float4* ab = reinterpret_cast<const float4*>(recb + (size_t)k * d.rec + d.offA);
    const float4* lr = reinterpret_cast<const float4*>(recb + (size_t)k * d.rec + d.offl);
#pragma unroll
    for (int t = 0; t < AB_PER; t++) {
      const int e = l16 + 16 * t;
      if (e < AB4) preAB[t] = __ldg(ab + e);
    }
#pragma unroll
    for (int t = 0; t < LRR_PER; t++) {
      const int e = l16 + 16 * t;
      if (e < lrr4) preL[t] = __ldg(lr + e);
    }
  };

  // ---- terminal condition (:102-105): Z_i = Q_i[T-1], zeta_i = l_i[T-1]; p_{T-1} = g_{T-1} ----
  prefetch(T - 2);
  {
    const float* last = recb + (size_t)(T - 1) * d.rec;
    for (int e = l16; e < NP * NX * NX / 4; e += 16)
      reinterpret_cast<float4*>(Z)[e] = __ldg(reinterpret_cast<const float4*>(last + d.offQ) + e);
    for (int e = l16; e < NP * NX; e += 16) zeta[e] = __ldg(last + d.offl + e);
    if (active) {
      for (int e = l16; e < MU * NX; e += 16) outP[(size_t)(T - 1) * MU * NX + e] = 0.f;
      for (int e = l16; e < MU; e += 16) outa[(size_t)(T - 1) * MU + e] = 0.f;
    }
    __syncwarp();
    for (int a = l16; a < NX; a += 16) {  // g_{T-1}[a] = sum_i sum_c Q_i[a][c] l_i[c]
      float g = 0.f;
      for (int i = 0; i < NP; i++) {
        float gi = 0.f;
#pragma unroll 4
        for (int c = 0; c < NX; c++) gi = fmaf(Z[(i * NX + a) * NX + c], zeta[i * NX + c], gi);
        g += gi;
      }
      pv[a] = g;
    }
    __syncwarp();
  }

  float expected_decrease = 0.f;  // accumulated redundantly by every lane of the half-warp

  for (int kk = T - 2; kk >= 0; kk--) {
    const float* rec = recb + (size_t)kk * d.rec;
    // ---- stage [A|B] and [l|R|r] from the prefetch registers ----
#pragma unroll
    for (int t = 0; t < AB_PER; t++) {
      const int e = l16 + 16 * t;
      if (e < A4) {
        reinterpret_cast<float4*>(AW)[e] = preAB[t];
      } else if (e < AB4) {
        reinterpret_cast<float4*>(Bm)[e - A4] = preAB[t];
      }
    }
#pragma unroll
    for (int t = 0; t < LRR_PER; t++) {
      const int e = l16 + 16 * t;
      if (e < lrr4) reinterpret_cast<float4*>(lrr)[e] = preL[t];
    }
    __syncwarp();

    // ---- BZ_i = B_i^T Z_i (:128), stored transposed: BZt[col][c] ----
    for (int t = l16; t < NP * 4; t += 16) {
      const int i = t >> 2, cg = (t & 3) * TC;
      float acc[m][TC];
      mm_tn<NX, m, TC>(Bm + i * m, MU, Z + i * NX * NX + cg, NX, acc);
#pragma unroll
      for (int r = 0; r < m; r++)
#pragma unroll
        for (int j = 0; j < TC; j++) BZt[(cg + j) * MU + i * m + r] = acc[r][j];
    }
    __syncwarp();
    // ---- Y = BZ A (:152-153) into P[c][col]; S = BZ B (+R_ii) (:131-149); y_alpha (:154-157) ----
    for (int t = l16; t < NP * 4; t += 16) {
      const int i = t >> 2, cg = (t & 3) * TC;
      float acc[m][TC];
      mm_tn<NX, m, TC>(BZt + i * m, MU, AW + cg, NX, acc);
#pragma unroll
      for (int r = 0; r < m; r++) stvec<TC>(P + (i * m + r) * NX + cg, acc[r]);
    }
    for (int t = l16; t < NP * NP; t += 16) {
      const int i = t / NP, j = t % NP;
      float acc[m][m];
      mm_tn<NX, m, m>(BZt + i * m, MU, Bm + j * m, MU, acc);
      if (i == j) {
        const float* Rii = Rk + d.pair_Roff[d.pair_of[i][i]];
#pragma unroll
        for (int r = 0; r < m; r++)
#pragma unroll
          for (int c = 0; c < m; c++) acc[r][c] = acc[r][c] + Rii[r * m + c];
      }
#pragma unroll
      for (int r = 0; r < m; r++)
#pragma unroll
        for (int c = 0; c < m; c++) S[(i * m + r) * MU + j * m + c] = acc[r][c];
    }
    for (int c = l16; c < MU; c += 16) {
      const int i = c / m;
      float acc = 0.f;
#pragma unroll 4
      for (int q = 0; q < NX; q++) acc = fmaf(Bm[q * MU + c], zeta[i * NX + q], acc);
      ya[c] = acc + rk[d.pair_roff[d.pair_of[i][i]] + (c - i * m)];
    }
    __syncwarp();

    // ---- Gershgorin (:163-176) + S X = Y (:180): lane col owns P[:, col]; every lane carries alpha ----
    {
      float Sm[MU][MU], y[G][MU], yal[MU];
      {
        float sflat[r4(MU * MU)];
        ldvec<r4(MU * MU)>(S, sflat);  // broadcast 128-bit loads
#pragma unroll
        for (int r = 0; r < MU; r++)
#pragma unroll
          for (int c = 0; c < MU; c++) Sm[r][c] = sflat[r * MU + c];
        float yflat[r4(MU)];
        ldvec<r4(MU)>(ya, yflat);
#pragma unroll
        for (int r = 0; r < MU; r++) {
          yal[r] = yflat[r];
#pragma unroll
          for (int g = 0; g < G; g++) {
            const int col = l16 + 16 * g;
            y[g][r] = col < NX ? P[r * NX + col] : 0.f;
          }
        }
      }
      bool dominant = p.adaptive_regularization != 0;
      if (p.adaptive_regularization) {
#pragma unroll
        for (int c = 0; c < MU; c++) {
          float col1 = 0.f;
#pragma unroll
          for (int r = 0; r < MU; r++) col1 += fabsf(Sm[r][c]);
          const float radius = col1 - fabsf(Sm[c][c]);
          const float eval_lo = Sm[c][c] - radius;
          constexpr float min_eval = 1e-3;
          if (eval_lo < min_eval) Sm[c][c] += radius + min_eval;
          dominant = dominant && (Sm[c][c] > radius);
        }
      }
      if (dominant)
        lu_solve<MU, G, false>(Sm, y, yal);
      else
        lu_solve<MU, G, true>(Sm, y, yal);
      __syncwarp();  // every lane has read S, P(=Y) and ya
#pragma unroll
      for (int r = 0; r < MU; r++) {
#pragma unroll
        for (int g = 0; g < G; g++) {
          const int col = l16 + 16 * g;
          if (col < NX) {
            P[r * NX + col] = y[g][r];
            if (active) outP[((size_t)kk * MU + r) * NX + col] = y[g][r];
          }
        }
        if (l16 == 0) {
          ya[r] = yal[r];
          if (active) outa[(size_t)kk * MU + r] = yal[r];
        }
      }
      // own-control part of ExpectedDecrease: (alpha_i^T R_ii) r_ii (src/ilq_solver.cpp:384-386)
#pragma unroll
      for (int i = 0; i < NP; i++) {
        const int pii = d.pair_of[i][i];
        const float* Rii = Rk + d.pair_Roff[pii];
        const float* rii = rk + d.pair_roff[pii];
        float t1 = 0.f;
#pragma unroll
        for (int c = 0; c < m; c++) {
          float row = 0.f;
#pragma unroll
          for (int a = 0; a < m; a++) row = fmaf(yal[i * m + a], Rii[a * m + c], row);
          t1 = fmaf(row, rii[c], t1);
        }
        expected_decrease -= t1;
      }
    }
    __syncwarp();
    // the next record's [A|B], [l|R|r] travel while this step's products run (issued here, after
    // the solve, so the prefetch registers are not live across it)
    if (kk > 0) prefetch(kk - 1);

    // ---- F = A - sum_i B_i P_i ; beta = - sum_i B_i alpha_i (:189-194) ----
    {
      float f[TR][TC];
#pragma unroll
      for (int r = 0; r < TR; r++) ldvec<TC>(AW + (a0 + r) * NX + c0, f[r]);
#pragma unroll
      for (int i = 0; i < NP; i++) {
        // F[a][c] -= sum_q B_i[a][q] P_i[q][c], B read in its stored [NX][MU] layout
        float acc[TR][TC];
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
          for (int j = 0; j < TC; j++) acc[r][j] = 0.f;
#pragma unroll
        for (int q = 0; q < m; q++) {
          float pr[TC];
          ldvec<TC>(P + (i * m + q) * NX + c0, pr);
#pragma unroll
          for (int r = 0; r < TR; r++) {
            const float bv = Bm[(a0 + r) * MU + i * m + q];
#pragma unroll
            for (int j = 0; j < TC; j++) acc[r][j] = fmaf(bv, pr[j], acc[r][j]);
          }
        }
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
          for (int j = 0; j < TC; j++) f[r][j] -= acc[r][j];
      }
#pragma unroll
      for (int r = 0; r < TR; r++) stvec<TC>(F + (a0 + r) * NX + c0, f[r]);
      for (int a = l16; a < NX; a += 16) {
        float bsum = 0.f;
#pragma unroll
        for (int i = 0; i < NP; i++) {
          float acc = 0.f;
#pragma unroll
          for (int q = 0; q < m; q++) acc = fmaf(Bm[a * MU + i * m + q], ya[i * m + q], acc);
          bsum -= acc;
        }
        beta[a] = bsum;
      }
    }
    __syncwarp();
    // state part of ExpectedDecrease via the adjoint: beta_k . p_{k+1}
    {
      float part = 0.f;
      for (int a = l16; a < NX; a += 16) part = fmaf(beta[a], pv[a], part);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      expected_decrease -= part;
    }
    // p_k = A_k^T p_{k+1} (+ g_k added per player below); AW still holds A here
    {
      float pvr[NX];
      ldvec<NX>(pv, pvr);
      for (int a = l16; a < NX; a += 16) {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int q = 0; q < NX; q += 2) {
          acc0 = fmaf(AW[q * NX + a], pvr[q], acc0);
          acc1 = fmaf(AW[(q + 1) * NX + a], pvr[q + 1], acc1);
        }
        pn[a] = acc0 + acc1;
      }
    }
    __syncwarp();

    // ---- per player: zeta_i, Z_i (:197-213) ----
#pragma unroll 1
    for (int i = 0; i < NP; i++) {
      float* Zi = Z + i * NX * NX;
      float* zi = zeta + i * NX;
      const float* li = lvec + i * NX;
      // this lane's tile of Q_i straight from global memory (latency covered by the products)
      float q[TR][TC];
#pragma unroll
      for (int r = 0; r < TR; r++) ldgvec<TC>(rec + d.offQ + (i * NX + a0 + r) * NX + c0, q[r]);
      // tv = zeta_next + Z_next beta: each lane reduces its tile's columns, then the 4 column
      // groups of a row are combined with two shuffles
      {
        float bseg[TC];
        ldvec<TC>(beta + c0, bseg);
        // rows (l16 >> 2) + 4 r: consecutive lanes' rows differ in parity, so the 128-bit row
        // segments of a quarter-warp fall on distinct banks
#pragma unroll
        for (int r = 0; r < TR; r++) {
          const int row = (l16 >> 2) + 4 * r;
          float zr[TC];
          ldvec<TC>(Zi + row * NX + c0, zr);
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < TC; j++) part = fmaf(zr[j], bseg[j], part);
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          if ((l16 & 3) == 0) tv[row] = zi[row] + part;
        }
      }
      __syncwarp();
      // zeta = F^T tv + l (+ P_j^T (R_ij alpha_j - r_ij))
      float tvr[NX];
      ldvec<NX>(tv, tvr);
      for (int a = l16; a < NX; a += 16) {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int c = 0; c < NX; c += 2) {
          acc0 = fmaf(F[c * NX + a], tvr[c], acc0);
          acc1 = fmaf(F[(c + 1) * NX + a], tvr[c + 1], acc1);
        }
        float znew = (acc0 + acc1) + li[a];
        for (int j = 0; j < NP; j++) {
          const int pr = d.pair_of[i][j];
          if (pr < 0) continue;
          const float* Rij = Rk + d.pair_Roff[pr];
          const float* rij = rk + d.pair_roff[pr];
          float t = 0.f;
#pragma unroll
          for (int c = 0; c < m; c++) {
            float v = 0.f;
#pragma unroll
            for (int c2 = 0; c2 < m; c2++) v = fmaf(Rij[c * m + c2], ya[j * m + c2], v);
            v -= rij[c];
            t = fmaf(P[(j * m + c) * NX + a], v, t);
          }
          znew += t;
        }
        zi[a] = znew;
      }
      // W^T = Z_next^T F computed directly (tile rows = columns of Z), over the dead A buffer
      float acc[TR][TC];
      mm_tn<NX, TR, TC>(Zi + a0, NX, F + c0, NX, acc);
#pragma unroll
      for (int r = 0; r < TR; r++) stvec<TC>(AW + (a0 + r) * NX + c0, acc[r]);
      __syncwarp();
      // Z = W F + Q (+ P_j^T R_ij P_j)
      mm_tn<NX, TR, TC>(AW + a0, NX, F + c0, NX, acc);
#pragma unroll
      for (int r = 0; r < TR; r++)
#pragma unroll
        for (int j = 0; j < TC; j++) acc[r][j] = acc[r][j] + q[r][j];
      for (int j2 = 0; j2 < NP; j2++) {
        const int pr = d.pair_of[i][j2];
        if (pr < 0) continue;
        const float* Rij = Rk + d.pair_Roff[pr];
        float prow[m][TR], pcol[m][TC];  // P_j[:, a0..] and P_j[:, c0..]
#pragma unroll
        for (int c = 0; c < m; c++) {
          ldvec<TR>(P + (j2 * m + c) * NX + a0, prow[c]);
          ldvec<TC>(P + (j2 * m + c) * NX + c0, pcol[c]);
        }
        float ptr[TR][m];  // (P_j^T R_ij)[a0 + r][c2]
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
          for (int c2 = 0; c2 < m; c2++) {
            float v = 0.f;
#pragma unroll
            for (int c = 0; c < m; c++) v = fmaf(prow[c][r], Rij[c * m + c2], v);
            ptr[r][c2] = v;
          }
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
          for (int j = 0; j < TC; j++) {
            float t = 0.f;
#pragma unroll
            for (int c2 = 0; c2 < m; c2++) t = fmaf(ptr[r][c2], pcol[c2][j], t);
            acc[r][j] += t;
          }
      }
      // g_k contribution of this player: sum_c Q_i[a][c] l_i[c], reduced over the 4 column groups
      {
        float gpart[TR], lseg[TC];
        ldvec<TC>(li + c0, lseg);
#pragma unroll
        for (int r = 0; r < TR; r++) {
          float g = 0.f;
#pragma unroll
          for (int j = 0; j < TC; j++) g = fmaf(q[r][j], lseg[j], g);
          g += __shfl_xor_sync(0xffffffffu, g, 1);
          g += __shfl_xor_sync(0xffffffffu, g, 2);
          gpart[r] = g;
        }
        if ((l16 & 3) == 0) {
#pragma unroll
          for (int r = 0; r < TR; r++) pn[a0 + r] += gpart[r];
        }
      }
      __syncwarp();  // all lanes finished reading Z_i and W^T
#pragma unroll
      for (int r = 0; r < TR; r++) stvec<TC>(Zi + (a0 + r) * NX + c0, acc[r]);
      __syncwarp();
    }
    for (int a = l16; a < NX; a += 16) pv[a] = pn[a];
    __syncwarp();
  }
  if (active && l16 == 0) s.expected_decrease[b] = expected_decrease;
}

// delta_xs of LQFeedbackSolver::Solve (src/lq_feedback_solver.cpp:217-241): dx_0 = the x0 argument
// (ILQG_LQ_X0), dx_{k+1} = A_k dx_k - sum_i B_i alpha_i[k].  One warp per instance; only
// launched when delta_xs are requested (ilqg_lq_backward / ILQG_DELTA_XS), never by ilqg_iterate.
__global__ void __launch_bounds__(128)
k_delta_xs(const __grid_constant__ DevDesc d, Slab s, const float* x0arg) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= s.B) return;
  const int n = d.n, M = d.M, T = d.T;
  float* dx = smem + warp * 2 * ILQG_MAX_XDIM;
  float* dxn = dx + ILQG_MAX_XDIM;
  const float* alpha = s.st_a[1 - s.st_cur[b]] + (size_t)b * T * M;
  float* out = s.dxs + (size_t)b * T * n;
  for (int a = lane; a < n; a += 32) dx[a] = x0arg ? x0arg[(size_t)b * n + a] : 0.f;
  __syncwarp();
  for (int kk = 0; kk < T; kk++) {
    const float* rec = s.rec + ((size_t)b * T + kk) * d.rec;
    for (int a = lane; a < n; a += 32) {
      out[(size_t)kk * n + a] = dx[a];
      float acc = 0.f;
      for (int q = 0; q < n; q++) acc = fmaf(__ldg(rec + d.offA + a * n + q), dx[q], acc);
      for (int i = 0; i < d.N; i++) {
        float pacc = 0.f;
        for (int q = d.uoff[i]; q < d.uoff[i + 1]; q++)
          pacc = fmaf(__ldg(rec + d.offB + a * M + q), alpha[(size_t)kk * M + q], pacc);
        acc -= pacc;
      }
      dxn[a] = acc;
    }
    __syncwarp();
    for (int a = lane; a < n; a += 32) dx[a] = dxn[a];
    __syncwarp();
  }
}

}  // namespace ilqg
