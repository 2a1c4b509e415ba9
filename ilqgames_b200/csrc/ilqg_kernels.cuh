// ilqg_kernels.cuh -- the sm_100a kernels of the batched iLQ hot path.
//
//   (K_lq, linearize + quadraticize, lives in ilqg_records.cuh)
//   k_lq_backward             K_bwd: one warp per instance, coupled Riccati sweep with the
//                                    running Z_i, zeta_i resident in shared memory, register-tiled
//                                    F^T Z F, per-lane LU of the stacked S X = Y system,
//                                    then the delta-x / expected-decrease forward sweep
//   (K_ls, the linesearch / rollout pipeline, lives in ilqg_linesearch.cuh)
//
// Data layout: instance-major slab; the LQ records [B][T][rec] hold
// [A | B | Q_0..Q_{N-1} | l | R | r] contiguously so one (instance, timestep) is a single
// 16-byte aligned chunk that a warp streams with 128-bit accesses.
#pragma once
#include <cooperative_groups.h>

#include "ilqg_device.cuh"

namespace ilqg {
namespace cg = cooperative_groups;

struct Slab {
  int B;
  float* x0;                        // [B][n]
  float* lq_x0;                     // [B][n]  ILQG_LQ_X0: x0 argument of a stand-alone LQ solve
  float* op_xs[2];                  // [B][T][n]   operating point double buffer
  float* op_us[2];                  // [B][T][M]
  float* st_P[2];                   // [B][T][M][n] strategies double buffer
  float* st_a[2];                   // [B][T][M]
  float *prob_xs, *prob_us, *prob_P, *prob_a;  // Problem's warm start
  float* rec;                       // [B][T][rec]  dense LQ records (allocated on first use, ilqg_abi.cu:EnsureDense)
  float* crec;                      // [B][T][NIp]  compact LQ records (ilqg_records.cuh)
  float* dxs;                       // [B][T][n]
  float* lambdas;                   // [B][ncon][T]
  float *mu, *last_merit, *expected_decrease, *step, *total_costs, *max_con_err;
  int *status, *iters, *backtracks, *te_quad, *te_new, *op_cur, *st_cur;
  int* ls_next_j;                   // [B] next linesearch candidate; 0 = no linesearch open
  // AugmentedLagrangianSolver::Solve per game (ilqg_al_begin / ilqg_al_advance):
  int *al_state, *al_iterates, *al_success;  // state 0 none, 1 active, 2 finished (solve_begin skips it)
  int* al_flags;                    // [B] scratch of one ilqg_al_advance: AL_DO_* bits
  int* queued_flag;                 // [B] 1 = this pass's first-window candidate was rejected (set by k_ls_decide)
  const int* lambda_index;          // [T]  kk -> Constraint::TimeIndex
};

__device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// an instance takes part in this pass's linearize / quadraticize / LQ solve if it is running and
// is not in the middle of a linesearch (ilqg_linesearch.cuh)
__device__ __forceinline__ bool instance_iterates(const Slab& s, int b) {
  return s.status[b] == ILQG_STATUS_RUNNING && s.ls_next_j[b] == 0;
}

// Which instances a K_lq / K_bwd launch covers.  The pipelined iteration (ilqg_abi.cu:
// IteratePipelined) splits a pass in two streams: the instances whose linesearch is still open
// (the "queued" ones, a few percent) finish it, then get their own K_lq / K_bwd over the queue
// list, while the rest of the batch is already linearized / solved for the next iteration.
enum { SEL_ALL = 0, SEL_MAIN = 1, SEL_LIST = 2 };
struct Sel {
  int mode;
  const int* list;   // SEL_LIST: instance ids
  const int* count;  // SEL_LIST: valid entries
};

// slot -> instance id (or -1), and whether that instance takes part in this launch
__device__ __forceinline__ int sel_instance(const Slab& s, const Sel& sel, int slot, int only_running, bool* live) {
  int b = -1;
  *live = false;
  if (sel.mode == SEL_LIST) {
    if (slot < *sel.count) {
      b = sel.list[slot];
      *live = instance_iterates(s, b);
    }
  } else if (slot < s.B) {
    b = slot;
    *live = !only_running || instance_iterates(s, b);
    if (sel.mode == SEL_MAIN && s.queued_flag[b]) *live = false;
  }
  return b;
}

// ===========================================================================
// small warp-level helpers
// ===========================================================================
__device__ __forceinline__ void warp_copy_f4(float* dst, const float* src, int floats, int lane) {
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (int e = lane; e < floats / 4; e += 32) d4[e] = __ldg(s4 + e);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// C tile of X^T Y for row-major NX x NX operands in shared memory.  Lane owns rows
// [a0, a0+TR) x cols [c0, c0+TC) with TR = NX/8, TC = NX/4 (32 lanes cover the matrix).
template <int NX>
struct Tile {
  static constexpr bool kTiled = (NX % 8 == 0) && NX >= 8;
  static constexpr int TR = kTiled ? NX / 8 : 1;
  static constexpr int TC = kTiled ? NX / 4 : 1;
};

template <int NX>
__device__ __forceinline__ void mm_tn_tile(const float* __restrict__ X, const float* __restrict__ Y,
                                           int a0, int c0, float (&acc)[Tile<NX>::TR][Tile<NX>::TC]) {
  constexpr int TR = Tile<NX>::TR, TC = Tile<NX>::TC;
#pragma unroll
  for (int i = 0; i < TR; i++)
#pragma unroll
    for (int j = 0; j < TC; j++) acc[i][j] = 0.f;
#pragma unroll 4
  for (int q = 0; q < NX; q++) {
    float xr[TR], yr[TC];
#pragma unroll
    for (int i = 0; i < TR; i++) xr[i] = X[q * NX + a0 + i];
#pragma unroll
    for (int j = 0; j < TC; j++) yr[j] = Y[q * NX + c0 + j];
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TC; j++) acc[i][j] = fmaf(xr[i], yr[j], acc[i][j]);
  }
}

// ===========================================================================
// K_bwd: LQFeedbackSolver::Solve (src/lq_feedback_solver.cpp:71-244) + ExpectedDecrease
// (src/ilq_solver.cpp:364-398).  One warp per instance.
// ===========================================================================
constexpr int KBWD_WARPS = 4;

template <int NX, int MU, int NP>
struct BwdSmem {
  static constexpr int Z = 0;
  static constexpr int zeta = Z + ((NP * NX * NX + 3) & ~3);
  static constexpr int F = zeta + ((NP * NX + 3) & ~3);
  static constexpr int W = F + ((NX * NX + 3) & ~3);
  static constexpr int BZ = W + ((NX * NX + 3) & ~3);
  static constexpr int S = BZ + ((MU * NX + 3) & ~3);
  static constexpr int Y = S + ((MU * MU + 3) & ~3);
  static constexpr int beta = Y + ((MU * (NX + 1) + 3) & ~3);
  static constexpr int tv = beta + ((NX + 3) & ~3);
  static constexpr int rec = tv + ((NX + 3) & ~3);  // + d.rec (run time)
};

template <int NX, int MU, int NP>
__global__ void __launch_bounds__(KBWD_WARPS * 32)
k_lq_backward(const __grid_constant__ DevDesc d, const DevParams p, Slab s, int only_running) {
  using L = BwdSmem<NX, MU, NP>;
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * KBWD_WARPS + warp;
  if (b >= s.B) return;
  if (only_running && !instance_iterates(s, b)) return;
  const int T = d.T;
  float* sm = smem + (size_t)warp * (L::rec + d.rec);
  float* Z = sm + L::Z;
  float* zeta = sm + L::zeta;
  float* F = sm + L::F;
  float* W = sm + L::W;
  float* BZ = sm + L::BZ;
  float* S = sm + L::S;
  float* Y = sm + L::Y;
  float* beta = sm + L::beta;
  float* tv = sm + L::tv;
  float* rec = sm + L::rec;
  const float* A = rec + d.offA;
  const float* Bm = rec + d.offB;
  const float* Rk = rec + d.offR;
  const float* rk = rec + d.offr;

  // the quadraticization that follows this solve sees the newest extreme-time index
  if (lane < NP) s.te_quad[(size_t)b * NP + lane] = s.te_new[(size_t)b * NP + lane];

  const int cand = 1 - s.st_cur[b];
  float* outP = s.st_P[cand] + (size_t)b * T * MU * NX;
  float* outa = s.st_a[cand] + (size_t)b * T * MU;
  const float* recb = s.rec + (size_t)b * T * d.rec;

  // owner player and row offset of each stacked control row
  int owner[MU], ro[MU];
#pragma unroll
  for (int c = 0; c < MU; c++) {
    int o = 0;
    for (int i = 1; i < NP; i++)
      if (c >= d.uoff[i]) o = i;
    owner[c] = o;
    ro[c] = d.uoff[o];
  }

  // Z_i[T-1] = Q_i[T-1], zeta_i[T-1] = l_i[T-1]  (:102-105); strategy at T-1 stays zero
  {
    const float* last = recb + (size_t)(T - 1) * d.rec;
    for (int e = lane; e < NP * NX * NX; e += 32) Z[e] = __ldg(last + d.offQ + e);
    for (int e = lane; e < NP * NX; e += 32) zeta[e] = __ldg(last + d.offl + e);
    for (int e = lane; e < MU * NX; e += 32) outP[(size_t)(T - 1) * MU * NX + e] = 0.f;
    for (int e = lane; e < MU; e += 32) outa[(size_t)(T - 1) * MU + e] = 0.f;
  }
  __syncwarp();

  constexpr int TR = Tile<NX>::TR, TC = Tile<NX>::TC;
  const int a0 = (lane >> 2) * TR, c0 = (lane & 3) * TC;  // tiled path only

  for (int kk = T - 2; kk >= 0; kk--) {
    warp_copy_f4(rec, recb + (size_t)kk * d.rec, d.rec, lane);
    __syncwarp();
    // ---- BZ = B_i^T Z_i (:128) ----
    for (int e = lane; e < MU * NX; e += 32) {
      const int c = e / NX, col = e % NX;
      int i = 0;
#pragma unroll
      for (int cc = 0; cc < MU; cc++)
        if (cc == c) i = owner[cc];
      const float* Zi = Z + i * NX * NX;
      float acc = 0.f;
#pragma unroll 4
      for (int q = 0; q < NX; q++) acc = fmaf(Bm[q * MU + c], Zi[q * NX + col], acc);
      BZ[e] = acc;
    }
    __syncwarp();
    // ---- S (:131-149), Y (:152-157) ----
    for (int e = lane; e < MU * MU + MU * (NX + 1); e += 32) {
      if (e < MU * MU) {
        const int c = e / MU, c2 = e % MU;
        int i = 0, r0 = 0, i2 = 0;
#pragma unroll
        for (int cc = 0; cc < MU; cc++) {
          if (cc == c) { i = owner[cc]; r0 = ro[cc]; }
          if (cc == c2) i2 = owner[cc];
        }
        float acc = 0.f;
#pragma unroll 4
        for (int q = 0; q < NX; q++) acc = fmaf(BZ[c * NX + q], Bm[q * MU + c2], acc);
        if (i == i2) {
          const int pii = d.pair_of[i][i], mi = d.udim[i];
          acc = acc + Rk[d.pair_Roff[pii] + (c - r0) * mi + (c2 - r0)];
        }
        S[e] = acc;
      } else {
        const int f = e - MU * MU;
        const int c = f / (NX + 1), col = f % (NX + 1);
        float acc = 0.f;
        if (col < NX) {
#pragma unroll 4
          for (int q = 0; q < NX; q++) acc = fmaf(BZ[c * NX + q], A[q * NX + col], acc);
        } else {
          int i = 0, r0 = 0;
#pragma unroll
          for (int cc = 0; cc < MU; cc++)
            if (cc == c) { i = owner[cc]; r0 = ro[cc]; }
          const float* zi = zeta + i * NX;
#pragma unroll 4
          for (int q = 0; q < NX; q++) acc = fmaf(Bm[q * MU + c], zi[q], acc);
          acc = acc + rk[d.pair_roff[d.pair_of[i][i]] + (c - r0)];
        }
        Y[f] = acc;
      }
    }
    __syncwarp();
    // ---- Gershgorin (:163-176) + solve S X = Y (:180) ----
    // Every lane holds S in registers and eliminates redundantly; lane `col` owns RHS column
    // `col` (NX state columns of P and the alpha column).  LU with partial pivoting replaces
    // the reference's Householder QR (same solution up to rounding).
    for (int col = lane; col < NX + 1; col += 32) {
      float Sm[MU][MU], y[MU];
#pragma unroll
      for (int r = 0; r < MU; r++) {
#pragma unroll
        for (int c = 0; c < MU; c++) Sm[r][c] = S[r * MU + c];
        y[r] = Y[r * (NX + 1) + col];
      }
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
        }
      }
#pragma unroll
      for (int k = 0; k < MU; k++) {
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
            const float t = y[k]; y[k] = y[r]; y[r] = t;
          }
        }
#pragma unroll
        for (int r = k + 1; r < MU; r++) {
          const float f = Sm[r][k] / Sm[k][k];
#pragma unroll
          for (int c = k + 1; c < MU; c++) Sm[r][c] = fmaf(-f, Sm[k][c], Sm[r][c]);
          y[r] = fmaf(-f, y[k], y[r]);
        }
      }
#pragma unroll
      for (int r = MU - 1; r >= 0; r--) {
        float acc = y[r];
#pragma unroll
        for (int c = r + 1; c < MU; c++) acc = fmaf(-Sm[r][c], y[c], acc);
        y[r] = acc / Sm[r][r];
      }
#pragma unroll
      for (int r = 0; r < MU; r++) {
        Y[r * (NX + 1) + col] = y[r];  // X overwrites Y: rows = [P | alpha]
        if (col < NX)
          outP[((size_t)kk * MU + r) * NX + col] = y[r];
        else
          outa[(size_t)kk * MU + r] = y[r];
      }
    }
    __syncwarp();
    const float* X = Y;  // X[r*(NX+1)+c] = P[r][c]; X[r*(NX+1)+NX] = alpha[r]
    // ---- F = A - sum B_i P_i ; beta = - sum B_i alpha_i (:189-194) ----
    for (int e = lane; e < NX * NX + NX; e += 32) {
      if (e < NX * NX) {
        const int a = e / NX, c = e % NX;
        float f = A[e];
        // per-player subtraction, as F_ -= B_i * P_i
        int i = 0;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < MU; q++) {
          if (owner[q] != i) { f -= acc; acc = 0.f; i = owner[q]; }
          acc = fmaf(Bm[a * MU + q], X[q * (NX + 1) + c], acc);
        }
        f -= acc;
        F[e] = f;
      } else {
        const int a = e - NX * NX;
        float bsum = 0.f, acc = 0.f;
        int i = 0;
#pragma unroll
        for (int q = 0; q < MU; q++) {
          if (owner[q] != i) { bsum -= acc; acc = 0.f; i = owner[q]; }
          acc = fmaf(Bm[a * MU + q], X[q * (NX + 1) + NX], acc);
        }
        bsum -= acc;
        beta[a] = bsum;
      }
    }
    __syncwarp();
    // ---- Z_i, zeta_i update (:197-213) ----
#pragma unroll 1
    for (int i = 0; i < NP; i++) {
      float* Zi = Z + i * NX * NX;
      float* zi = zeta + i * NX;
      const float* Qi = rec + d.offQ + i * NX * NX;
      const float* li = rec + d.offl + i * NX;
      // tv = zeta_next + Z_next beta
      for (int a = lane; a < NX; a += 32) {
        float acc = 0.f;
#pragma unroll 4
        for (int q = 0; q < NX; q++) acc = fmaf(Zi[a * NX + q], beta[q], acc);
        tv[a] = zi[a] + acc;
      }
      __syncwarp();
      // zeta = F^T tv + l  (+ control terms)
      for (int a = lane; a < NX; a += 32) {
        float acc = 0.f;
#pragma unroll 4
        for (int q = 0; q < NX; q++) acc = fmaf(F[q * NX + a], tv[q], acc);
        float znew = acc + li[a];
        for (int j = 0; j < NP; j++) {
          const int pr = d.pair_of[i][j];
          if (pr < 0) continue;
          const int mj = d.udim[j], co = d.uoff[j];
          const float* Rij = Rk + d.pair_Roff[pr];
          const float* rij = rk + d.pair_roff[pr];
          float t = 0.f;
          for (int q = 0; q < mj; q++) {
            float v = 0.f;
            for (int q2 = 0; q2 < mj; q2++) v = fmaf(Rij[q * mj + q2], X[(co + q2) * (NX + 1) + NX], v);
            v -= rij[q];
            t = fmaf(X[(co + q) * (NX + 1) + a], v, t);
          }
          znew += t;
        }
        zi[a] = znew;
      }
      // Z = (F^T Z_next) F + Q (+ P_j^T R_ij P_j)
      if constexpr (Tile<NX>::kTiled) {
        float acc[TR][TC];
        mm_tn_tile<NX>(F, Zi, a0, c0, acc);  // W = F^T Z   (tile rows a0.., cols c0..)
        // store W transposed so the second product is again an X^T Y form
#pragma unroll
        for (int ii = 0; ii < TR; ii++)
#pragma unroll
          for (int jj = 0; jj < TC; jj++) W[(c0 + jj) * NX + a0 + ii] = acc[ii][jj];
        __syncwarp();
        mm_tn_tile<NX>(W, F, a0, c0, acc);   // (W^T)^T F = W F
#pragma unroll
        for (int ii = 0; ii < TR; ii++)
#pragma unroll
          for (int jj = 0; jj < TC; jj++) acc[ii][jj] = acc[ii][jj] + Qi[(a0 + ii) * NX + c0 + jj];
        for (int j = 0; j < NP; j++) {
          const int pr = d.pair_of[i][j];
          if (pr < 0) continue;
          const int mj = d.udim[j], co = d.uoff[j];
          const float* Rij = Rk + d.pair_Roff[pr];
#pragma unroll
          for (int ii = 0; ii < TR; ii++)
#pragma unroll
            for (int jj = 0; jj < TC; jj++) {
              float t = 0.f;
              for (int q2 = 0; q2 < mj; q2++) {
                float ptr = 0.f;  // (P_j^T R_ij)[a][q2]
                for (int q = 0; q < mj; q++)
                  ptr = fmaf(X[(co + q) * (NX + 1) + a0 + ii], Rij[q * mj + q2], ptr);
                t = fmaf(ptr, X[(co + q2) * (NX + 1) + c0 + jj], t);
              }
              acc[ii][jj] += t;
            }
        }
        __syncwarp();  // all lanes finished reading Zi (first product) before overwrite
#pragma unroll
        for (int ii = 0; ii < TR; ii++)
#pragma unroll
          for (int jj = 0; jj < TC; jj++) Zi[(a0 + ii) * NX + c0 + jj] = acc[ii][jj];
      } else {
        for (int e = lane; e < NX * NX; e += 32) {
          const int a = e / NX, c = e % NX;
          float acc = 0.f;
          for (int q = 0; q < NX; q++) acc = fmaf(F[q * NX + a], Zi[q * NX + c], acc);
          W[e] = acc;
        }
        __syncwarp();
        float znew[(NX * NX + 31) / 32];
        int cnt = 0;
        for (int e = lane; e < NX * NX; e += 32, cnt++) {
          const int a = e / NX, c = e % NX;
          float acc = 0.f;
          for (int q = 0; q < NX; q++) acc = fmaf(W[a * NX + q], F[q * NX + c], acc);
          acc = acc + Qi[e];
          for (int j = 0; j < NP; j++) {
            const int pr = d.pair_of[i][j];
            if (pr < 0) continue;
            const int mj = d.udim[j], co = d.uoff[j];
            const float* Rij = Rk + d.pair_Roff[pr];
            float t = 0.f;
            for (int q2 = 0; q2 < mj; q2++) {
              float ptr = 0.f;
              for (int q = 0; q < mj; q++) ptr = fmaf(X[(co + q) * (NX + 1) + a], Rij[q * mj + q2], ptr);
              t = fmaf(ptr, X[(co + q2) * (NX + 1) + c], t);
            }
            acc += t;
          }
          znew[cnt] = acc;
        }
        __syncwarp();
        cnt = 0;
        for (int e = lane; e < NX * NX; e += 32, cnt++) Zi[e] = znew[cnt];
      }
      __syncwarp();
    }
  }

  // ---- forward sweep: delta_xs (:217-241, x0 argument = 0) and ExpectedDecrease ----
  float* dx = beta;  // reuse
  float* dxn = tv;
  for (int a = lane; a < NX; a += 32) dx[a] = 0.f;
  float expected_decrease = 0.f;
  float* dxs = s.dxs + (size_t)b * T * NX;
  __syncwarp();
  for (int kk = 0; kk < T; kk++) {
    warp_copy_f4(rec, recb + (size_t)kk * d.rec, d.rec, lane);
    for (int e = lane; e < MU; e += 32) S[e] = outa[(size_t)kk * MU + e];  // alpha_k
    for (int a = lane; a < NX; a += 32) dxs[(size_t)kk * NX + a] = dx[a];
    __syncwarp();
    for (int i = 0; i < NP; i++) {
      const int mi = d.udim[i], r0 = d.uoff[i], pii = d.pair_of[i][i];
      const float* Rii = Rk + d.pair_Roff[pii];
      const float* rii = rk + d.pair_roff[pii];
      float t1 = 0.f;  // (alpha^T R_ii) r_ii, tiny: every lane computes it
      for (int c = 0; c < mi; c++) {
        float row = 0.f;
        for (int a = 0; a < mi; a++) row = fmaf(S[r0 + a], Rii[a * mi + c], row);
        t1 = fmaf(row, rii[c], t1);
      }
      expected_decrease -= t1;
      if (kk > 0) {
        const float* Qi = rec + d.offQ + i * NX * NX;
        const float* li = rec + d.offl + i * NX;
        float part = 0.f;
        for (int c = lane; c < NX; c += 32) {
          float row = 0.f;
#pragma unroll 4
          for (int a = 0; a < NX; a++) row = fmaf(dx[a], Qi[a * NX + c], row);
          part = fmaf(row, li[c], part);
        }
        expected_decrease -= warp_sum(part);
      }
    }
    // dx_next = A dx - sum_i B_i alpha_i   (no feedback term)
    for (int a = lane; a < NX; a += 32) {
      float acc = 0.f;
#pragma unroll 4
      for (int q = 0; q < NX; q++) acc = fmaf(A[a * NX + q], dx[q], acc);
      int i = 0;
      float pacc = 0.f;
#pragma unroll
      for (int q = 0; q < MU; q++) {
        if (owner[q] != i) { acc -= pacc; pacc = 0.f; i = owner[q]; }
        pacc = fmaf(Bm[a * MU + q], S[q], pacc);
      }
      acc -= pacc;
      dxn[a] = acc;
    }
    __syncwarp();
    for (int a = lane; a < NX; a += 32) dx[a] = dxn[a];
    __syncwarp();
  }
  if (lane == 0) s.expected_decrease[b] = expected_decrease;
}

// ===========================================================================
// small utility kernels
// ===========================================================================
// dst[b][e] = (sel[b] ^ flip ? src1 : src0)[b][e]
__global__ void k_gather_parity(float* dst, const float* src0, const float* src1, const int* sel,
                                int flip, size_t per, int B) {
  const size_t total = per * (size_t)B;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / per);
    dst[e] = ((sel[b] ^ flip) ? src1 : src0)[e];
  }
}

// Problem::OverwriteSolution (src/problem.cpp:188-194): prob <- current, optionally only
// for instances whose last solve did not fail.
__global__ void k_overwrite_solution(const __grid_constant__ DevDesc d, Slab s, int only_successful, int flag_bit) {
  const int b = blockIdx.x;
  if (flag_bit && !(s.al_flags[b] & flag_bit)) return;
  if (only_successful && s.status[b] == ILQG_STATUS_LINESEARCH_FAILED) return;
  const int T = d.T, n = d.n, M = d.M, cur = s.op_cur[b], scur = s.st_cur[b];
  const size_t ox = (size_t)b * T * n, ou = (size_t)b * T * M, oP = (size_t)b * T * M * n;
  for (int e = threadIdx.x; e < T * n; e += blockDim.x) s.prob_xs[ox + e] = s.op_xs[cur][ox + e];
  for (int e = threadIdx.x; e < T * M; e += blockDim.x) {
    s.prob_us[ou + e] = s.op_us[cur][ou + e];
    s.prob_a[ou + e] = s.st_a[scur][ou + e];
  }
  for (int e = threadIdx.x; e < T * M * n; e += blockDim.x) s.prob_P[oP + e] = s.st_P[scur][oP + e];
}

// upload_warmstart mirrors the problem solution into the working buffers
__global__ void k_prob_to_working(const __grid_constant__ DevDesc d, Slab s) {
  const int b = blockIdx.x;
  const int T = d.T, n = d.n, M = d.M;
  const size_t ox = (size_t)b * T * n, ou = (size_t)b * T * M, oP = (size_t)b * T * M * n;
  for (int e = threadIdx.x; e < T * n; e += blockDim.x) s.op_xs[0][ox + e] = s.prob_xs[ox + e];
  for (int e = threadIdx.x; e < T * M; e += blockDim.x) {
    s.op_us[0][ou + e] = s.prob_us[ou + e];
    s.st_a[0][ou + e] = s.prob_a[ou + e];
  }
  for (int e = threadIdx.x; e < T * M * n; e += blockDim.x) s.st_P[0][oP + e] = s.prob_P[oP + e];
  if (threadIdx.x == 0) {
    s.op_cur[b] = 0;
    s.st_cur[b] = 0;
  }
}

// One augmented-Lagrangian multiplier sweep, src/augmented_lagrangian_solver.cpp:113-143,
// one thread per (instance, constraint): lambda <- max(0, lambda + mu g) visiting kk in order
// so the duplicated TimeIndex slots are incremented twice, exactly as the reference.
__global__ void k_al_update(const __grid_constant__ DevDesc d, const DevParams p, Slab s, int flag_bit) {
  const int b = blockIdx.x;
  if (flag_bit && !(s.al_flags[b] & flag_bit)) return;
  __shared__ float smax[ILQG_MAX_COSTS];
  const int cidx = threadIdx.x;
  float max_err = -INFINITY;
  if (cidx < d.num_costs && d.cost[cidx].slot >= 0) {
    const DevCost& cd = d.cost[cidx];
    const int cur = s.op_cur[b];
    const float mu = s.mu[b];
    float* lam = s.lambdas + ((size_t)b * d.num_constraints + cd.slot) * d.T;
    for (int kk = 0; kk < d.T; kk++) {
      const float* x = s.op_xs[cur] + ((size_t)b * d.T + kk) * d.n;
      const float* u = s.op_us[cur] + ((size_t)b * d.T + kk) * d.M;
      const float g = cd.arg < 0 ? evaluate_record(d, cd, x, d.n)
                                 : evaluate_record(d, cd, u + d.uoff[cd.arg], d.udim[cd.arg]);
      max_err = fmaxf(max_err, g);
      const int li = s.lambda_index[kk];
      const float new_lambda = lam[li] + mu * g;  // Constraint::IncrementLambda constraint.h:98-102
      lam[li] = cd.is_equality ? new_lambda : fmaxf(0.0f, new_lambda);
    }
  }
  if (cidx < ILQG_MAX_COSTS) smax[cidx] = max_err;
  __syncthreads();
  if (cidx == 0) {
    float m = -INFINITY;
    for (int c = 0; c < d.num_costs; c++) m = fmaxf(m, smax[c]);
    s.max_con_err[b] = m;
    s.mu[b] = s.mu[b] * p.geometric_mu_scaling;  // Constraint::ScaleMu :143
  }
}

// src/augmented_lagrangian_solver.cpp:165-178
__global__ void k_al_post_solve(const __grid_constant__ DevDesc d, const DevParams p, Slab s, int flag_bit) {
  const int b = blockIdx.x;
  if (flag_bit && !(s.al_flags[b] & flag_bit)) return;
  if (s.status[b] != ILQG_STATUS_LINESEARCH_FAILED) return;
  const int cnt = d.num_constraints * d.T;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x)
    s.lambdas[(size_t)b * cnt + e] *= p.geometric_lambda_downscaling;
  if (threadIdx.x == 0) s.mu[b] *= p.geometric_mu_downscaling;
}

// Bookkeeping of one AugmentedLagrangianSolver::Solve round per game
// (src/augmented_lagrangian_solver.cpp:94-111, 165-190); the multiplier work it decides on is
// done by k_al_post_solve / k_al_update / k_overwrite_solution under the AL_DO_* flags.
enum { AL_DO_DOWNSCALE = 1, AL_DO_UPDATE = 2, AL_DO_OVERWRITE = 4 };

__global__ void k_al_account(const __grid_constant__ DevDesc d, Slab s, int first, int max_iterates, float tolerance,
                             int* active) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  s.al_flags[b] = 0;
  if (s.al_state[b] != 1) return;
  const bool failed = s.status[b] == ILQG_STATUS_LINESEARCH_FAILED;
  // the inner log holds the initial iterate plus one per completed iteration; the iteration whose
  // linesearch failed is not logged (src/ilq_solver.cpp:111,146-153,164)
  const int iterates = s.al_iterates[b] + 1 + s.iters[b] - (failed ? 1 : 0);
  s.al_iterates[b] = iterates;
  int flags = (!first && failed) ? AL_DO_DOWNSCALE : 0;
  if (failed) s.al_success[b] = 0;
  const bool constrained = d.num_constraints > 0;
  const float err = s.max_con_err[b];
  const bool again = constrained && iterates < max_iterates && err > tolerance;
  if (!again) {
    if (constrained && err > tolerance) s.al_success[b] = 0;
    s.al_state[b] = 2;
  } else {
    flags |= AL_DO_UPDATE | (failed ? 0 : AL_DO_OVERWRITE);
    atomicAdd(active, 1);
  }
  s.al_flags[b] = flags;
}

__global__ void k_al_begin(Slab s) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.B) return;
  s.al_state[b] = 1;
  s.al_iterates[b] = 0;
  s.al_success[b] = 1;
  s.max_con_err[b] = INFINITY;
}

__global__ void k_fill(float* p, float v, size_t count) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < count;
       e += (size_t)gridDim.x * blockDim.x)
    p[e] = v;
}

}  // namespace ilqg
