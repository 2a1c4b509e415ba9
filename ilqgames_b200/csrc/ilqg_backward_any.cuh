// ilqg_backward_any.cuh -- K_bwd for ANY shape inside the compiled envelope (n <= ILQG_MAX_XDIM,
// M <= ILQG_MAX_UDIM, N <= ILQG_MAX_PLAYERS, any m_i), over dense LQ records: the fallback behind
// LQFeedbackSolver::Solve (src/lq_feedback_solver.cpp:71-244), which takes any dimensions.  The hot
// shapes never come here (pattern-based handles run k_lq_backward_tc, ilqg_backward_tc.cuh; a few
// dense-record shapes keep their specialised kernels); this one serves stand-alone LQ solves on
// uploaded matrices (ILQG_DYN_NONE handles, the host class LQFeedbackSolver) of arbitrary shape.
// One warp per game, everything in shared memory, run-time loops, Gaussian elimination with partial
// pivoting on the augmented matrix [S | Y | y_alpha] (Householder QR in the reference: same solution
// up to rounding).  delta_xs and ExpectedDecrease by the reference's forward sweep (:217-241,
// src/ilq_solver.cpp:364-398).  Not tuned.
#pragma once
#include "ilqg_kernels.cuh"

namespace ilqg {

constexpr int KANY_WARPS = 2;

// per-game shared memory (floats)
__host__ __device__ inline int any_smem_floats(int n, int M, int N, int rec) {
  const int r = 3;
  return ((N * n * n + r) & ~r) + ((N * n + r) & ~r) + 2 * ((n * n + r) & ~r) + ((M * n + r) & ~r) +
         ((M * (M + n + 1) + r) & ~r) + 3 * ((n + r) & ~r) + ((rec + r) & ~r);
}

__global__ void __launch_bounds__(KANY_WARPS * 32)
k_lq_backward_any(const __grid_constant__ DevDesc d, const DevParams p, Slab s, int only_running, const float* x0arg) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * KANY_WARPS + warp;
  if (b >= s.B) return;
  if (only_running && !instance_iterates(s, b)) return;
  const int n = d.n, M = d.M, NP = d.N, T = d.T, AW = M + n + 1;
  auto up4 = [](int v) { return (v + 3) & ~3; };
  float* sm = smem + (size_t)warp * any_smem_floats(n, M, NP, d.rec);
  float* Z = sm;                          // [NP][n][n]
  float* zeta = Z + up4(NP * n * n);      // [NP][n]
  float* F = zeta + up4(NP * n);          // [n][n]
  float* W = F + up4(n * n);              // [n][n]
  float* BZ = W + up4(n * n);             // [M][n]
  float* Aug = BZ + up4(M * n);           // [M][M + n + 1]  [S | Y | y_alpha], then [. | P | alpha]
  float* beta = Aug + up4(M * AW);        // [n]
  float* tv = beta + up4(n);              // [n]
  float* tv2 = tv + up4(n);               // [n]
  float* rec = tv2 + up4(n);              // [rec]
  const float* A = rec + d.offA;
  const float* Bm = rec + d.offB;
  const float* Rk = rec + d.offR;
  const float* rk = rec + d.offr;
  auto owner = [&](int c) {
    int o = 0;
    for (int i = 1; i < NP; i++)
      if (c >= d.uoff[i]) o = i;
    return o;
  };
  if (lane < NP) s.te_quad[(size_t)b * NP + lane] = s.te_new[(size_t)b * NP + lane];
  const int cand = 1 - s.st_cur[b];
  float* outP = s.st_P[cand] + (size_t)b * T * M * n;
  float* outa = s.st_a[cand] + (size_t)b * T * M;
  const float* recb = s.rec + (size_t)b * T * d.rec;
  {
    const float* last = recb + (size_t)(T - 1) * d.rec;
    for (int e = lane; e < NP * n * n; e += 32) Z[e] = __ldg(last + d.offQ + e);
    for (int e = lane; e < NP * n; e += 32) zeta[e] = __ldg(last + d.offl + e);
    for (int e = lane; e < M * n; e += 32) outP[(size_t)(T - 1) * M * n + e] = 0.f;
    for (int e = lane; e < M; e += 32) outa[(size_t)(T - 1) * M + e] = 0.f;
  }
  __syncwarp();
  for (int kk = T - 2; kk >= 0; kk--) {
    for (int e = lane; e < d.rec; e += 32) rec[e] = __ldg(recb + (size_t)kk * d.rec + e);
    __syncwarp();
    // BZ[c][:] = B[:, c]' Z_owner(c) (:128)
    for (int e = lane; e < M * n; e += 32) {
      const int c = e / n, col = e % n;
      const float* Zi = Z + owner(c) * n * n;
      float acc = 0.f;
      for (int q = 0; q < n; q++) acc = fmaf(Bm[q * M + c], Zi[q * n + col], acc);
      BZ[e] = acc;
    }
    __syncwarp();
    // [S | Y | y_alpha] (:131-157)
    for (int e = lane; e < M * AW; e += 32) {
      const int c = e / AW, col = e % AW, i = owner(c), r0 = d.uoff[i];
      float acc = 0.f;
      if (col < M) {
        for (int q = 0; q < n; q++) acc = fmaf(BZ[c * n + q], Bm[q * M + col], acc);
        if (owner(col) == i) acc = acc + Rk[d.pair_Roff[d.pair_of[i][i]] + (c - r0) * d.udim[i] + (col - r0)];
      } else if (col < M + n) {
        for (int q = 0; q < n; q++) acc = fmaf(BZ[c * n + q], A[q * n + col - M], acc);
      } else {
        const float* zi = zeta + i * n;
        for (int q = 0; q < n; q++) acc = fmaf(Bm[q * M + c], zi[q], acc);
        acc = acc + rk[d.pair_roff[d.pair_of[i][i]] + (c - r0)];
      }
      Aug[e] = acc;
    }
    __syncwarp();
    // Gershgorin, column by column (:163-176)
    if (p.adaptive_regularization && lane < M) {
      float col1 = 0.f;
      for (int r = 0; r < M; r++) col1 += fabsf(Aug[r * AW + lane]);
      const float diag = Aug[lane * AW + lane];
      const float radius = col1 - fabsf(diag);
      const float eval_lo = diag - radius;
      constexpr float min_eval = 1e-3;
      if (eval_lo < min_eval) Aug[lane * AW + lane] = diag + (radius + min_eval);
    }
    __syncwarp();
    // S X = Y (:180): elimination with partial pivoting, lanes over the columns of the augmented matrix
    for (int k = 0; k < M; k++) {
      int piv = k;
      float best = fabsf(Aug[k * AW + k]);
      for (int r = k + 1; r < M; r++) {
        const float v = fabsf(Aug[r * AW + k]);
        if (v > best) { best = v; piv = r; }
      }
      __syncwarp();  // everyone has chosen the same pivot before rows move
      if (piv != k)
        for (int c = lane; c < AW; c += 32) {
          const float tmp = Aug[k * AW + c];
          Aug[k * AW + c] = Aug[piv * AW + c];
          Aug[piv * AW + c] = tmp;
        }
      __syncwarp();
      const float inv = 1.0f / Aug[k * AW + k];
      __syncwarp();
      for (int r = k + 1; r < M; r++) {
        const float f = Aug[r * AW + k] * inv;
        __syncwarp();
        for (int c = lane; c < AW; c += 32)
          if (c > k) Aug[r * AW + c] = fmaf(-f, Aug[k * AW + c], Aug[r * AW + c]);
        __syncwarp();
      }
    }
    for (int c = lane; c < n + 1; c += 32) {  // back substitution, one right-hand side per lane
      const int col = M + c;
      for (int r = M - 1; r >= 0; r--) {
        float acc = Aug[r * AW + col];
        for (int q = r + 1; q < M; q++) acc = fmaf(-Aug[r * AW + q], Aug[q * AW + col], acc);
        Aug[r * AW + col] = acc * (1.0f / Aug[r * AW + r]);
      }
      for (int r = 0; r < M; r++) {
        if (c < n) outP[((size_t)kk * M + r) * n + c] = Aug[r * AW + col];
        else outa[(size_t)kk * M + r] = Aug[r * AW + col];
      }
    }
    __syncwarp();
    auto Pe = [&](int r, int c) { return Aug[r * AW + M + c]; };   // P[r][c]
    auto al = [&](int r) { return Aug[r * AW + M + n]; };          // alpha[r]
    // F = A - sum_i B_i P_i, beta = - sum_i B_i alpha_i (:189-194), per player as the reference
    for (int e = lane; e < n * n + n; e += 32) {
      const int a = e < n * n ? e / n : e - n * n, c = e % n;
      float out = e < n * n ? A[e] : 0.f;
      for (int i = 0; i < NP; i++) {
        float acc = 0.f;
        for (int q = d.uoff[i]; q < d.uoff[i + 1]; q++) acc = fmaf(Bm[a * M + q], e < n * n ? Pe(q, c) : al(q), acc);
        out -= acc;
      }
      if (e < n * n) F[e] = out; else beta[a] = out;
    }
    __syncwarp();
    for (int i = 0; i < NP; i++) {  // (:197-213)
      float* Zi = Z + i * n * n;
      float* zi = zeta + i * n;
      const float* Qi = rec + d.offQ + i * n * n;
      const float* li = rec + d.offl + i * n;
      for (int a = lane; a < n; a += 32) {
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(Zi[a * n + q], beta[q], acc);
        tv[a] = zi[a] + acc;
      }
      __syncwarp();
      for (int a = lane; a < n; a += 32) {
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(F[q * n + a], tv[q], acc);
        float znew = acc + li[a];
        for (int j = 0; j < NP; j++) {
          const int pr = d.pair_of[i][j];
          if (pr < 0) continue;
          const int mj = d.udim[j], co = d.uoff[j];
          float t = 0.f;
          for (int q = 0; q < mj; q++) {
            float v = 0.f;
            for (int q2 = 0; q2 < mj; q2++) v = fmaf(Rk[d.pair_Roff[pr] + q * mj + q2], al(co + q2), v);
            v -= rk[d.pair_roff[pr] + q];
            t = fmaf(Pe(co + q, a), v, t);
          }
          znew += t;
        }
        tv2[a] = znew;
      }
      for (int e = lane; e < n * n; e += 32) {  // W = F' Z_i
        const int a = e / n, c = e % n;
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(F[q * n + a], Zi[q * n + c], acc);
        W[e] = acc;
      }
      __syncwarp();
      for (int a = lane; a < n; a += 32) zi[a] = tv2[a];
      for (int e = lane; e < n * n; e += 32) {  // Z_i = W F + Q_i + sum_j P_j' R_ij P_j
        const int a = e / n, c = e % n;
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(W[a * n + q], F[q * n + c], acc);
        acc = acc + Qi[e];
        for (int j = 0; j < NP; j++) {
          const int pr = d.pair_of[i][j];
          if (pr < 0) continue;
          const int mj = d.udim[j], co = d.uoff[j];
          float t = 0.f;
          for (int q2 = 0; q2 < mj; q2++) {
            float ptr = 0.f;
            for (int q = 0; q < mj; q++) ptr = fmaf(Pe(co + q, a), Rk[d.pair_Roff[pr] + q * mj + q2], ptr);
            t = fmaf(ptr, Pe(co + q2, c), t);
          }
          acc += t;
        }
        Zi[e] = acc;  // (element e of Z_i is read only by this lane's own W row pass above: W is complete)
      }
      __syncwarp();
    }
  }
  // forward sweep: delta_xs (x0 argument) and ExpectedDecrease
  float* dx = beta;
  float* dxn = tv;
  for (int a = lane; a < n; a += 32) dx[a] = x0arg ? x0arg[(size_t)b * n + a] : 0.f;
  float expected_decrease = 0.f;
  float* dxs = s.dxs + (size_t)b * T * n;
  __syncwarp();
  for (int kk = 0; kk < T; kk++) {
    for (int e = lane; e < d.rec; e += 32) rec[e] = __ldg(recb + (size_t)kk * d.rec + e);
    for (int e = lane; e < M; e += 32) tv2[e] = outa[(size_t)kk * M + e];
    for (int a = lane; a < n; a += 32) dxs[(size_t)kk * n + a] = dx[a];
    __syncwarp();
    for (int i = 0; i < NP; i++) {
      const int mi = d.udim[i], r0 = d.uoff[i], pii = d.pair_of[i][i];
      float t1 = 0.f;
      for (int c = 0; c < mi; c++) {
        float row = 0.f;
        for (int a = 0; a < mi; a++) row = fmaf(tv2[r0 + a], Rk[d.pair_Roff[pii] + a * mi + c], row);
        t1 = fmaf(row, rk[d.pair_roff[pii] + c], t1);
      }
      expected_decrease -= t1;
      if (kk > 0) {
        const float* Qi = rec + d.offQ + i * n * n;
        const float* li = rec + d.offl + i * n;
        float part = 0.f;
        for (int c = lane; c < n; c += 32) {
          float row = 0.f;
          for (int a = 0; a < n; a++) row = fmaf(dx[a], Qi[a * n + c], row);
          part = fmaf(row, li[c], part);
        }
        expected_decrease -= warp_sum(part);
      }
    }
    for (int a = lane; a < n; a += 32) {
      float acc = 0.f;
      for (int q = 0; q < n; q++) acc = fmaf(A[a * n + q], dx[q], acc);
      for (int i = 0; i < NP; i++) {
        float pacc = 0.f;
        for (int q = d.uoff[i]; q < d.uoff[i + 1]; q++) pacc = fmaf(Bm[a * M + q], tv2[q], pacc);
        acc -= pacc;
      }
      dxn[a] = acc;
    }
    __syncwarp();
    for (int a = lane; a < n; a += 32) dx[a] = dxn[a];
    __syncwarp();
  }
  if (lane == 0) s.expected_decrease[b] = expected_decrease;
}

}  // namespace ilqg
