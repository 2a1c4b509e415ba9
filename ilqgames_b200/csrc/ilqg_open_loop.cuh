// ilqg_open_loop.cuh -- LQOpenLoopSolver::Solve (src/lq_open_loop_solver.cpp:73-195), the LQ
// solver ILQSolver uses under SolverParams::open_loop (include/ilqgames/solver/ilq_solver.h:76-81).
//
// One warp per game instance, run-time dimensions (n <= 24, M <= 8, N <= 4, any m_i).  Backward
// sweep (:117-155): per time step
//     W_i   = R_ii^-1 B_i^T,  w_i = R_ii^-1 r_ii              (LDLT in the reference; R_ii is tiny)
//     K_i   = W_i M_i,        kappa_i = W_i m_i + w_i
//     Lam   = I + sum_i B_i K_i,   nu = - sum_i B_i kappa_i
//     [G g] = Lam^-1 [A nu]                                   (Householder QR in the reference;
//                                                              warp-cooperative LU with partial
//                                                              pivoting here)
//     M_i  <- Q_i + A^T (M_i G),   m_i <- l_i + A^T (m_i + M_i g)
// with M_i, m_i resident in shared memory.  What the forward sweep (:158-192) needs --
// x*_{k+1} = G_k x*_k + g_k,  alpha_i[k] = K_i x*_{k+1} + kappa_i -- is G, g, K, kappa: one small
// record per step in a scratch buffer, instead of the reference's M, m, QR factors for all k.
// ILQSolver::ExpectedDecrease (src/ilq_solver.cpp:364-398) is accumulated in the same forward
// sweep from two more per-step vectors (sum_i Q_i l_i and R_ii r_ii).  P stays zero.
// The association of the products differs from the reference's (e.g. (B_i W_i) M_i there), so
// parity is at tolerance level, like the feedback kernel's.
#pragma once
#include "ilqg_kernels.cuh"

namespace ilqg {

constexpr int KOL_WARPS = 2;

struct OlLayout {
  int Mi, mi, Lam, X, T1, K, kap, W, wr, A, B, lrr, Rinv, tv, total;  // shared memory, floats per warp
  int sG, sg, sK, skap, sgq, sw, srec;                               // scratch record, floats per step
};

__host__ __device__ inline OlLayout ol_layout(int n, int M, int N, int lrr_floats) {
  OlLayout L;
  auto r4 = [](int v) { return (v + 3) & ~3; };
  int o = 0;
  L.Mi = o; o += N * n * n;
  L.mi = o; o += r4(N * n);
  L.Lam = o; o += n * n;
  L.X = o; o += r4(n * (n + 1));
  L.T1 = o; o += n * n;
  L.K = o; o += r4(M * n);
  L.kap = o; o += r4(M);
  L.W = o; o += r4(M * n);
  L.wr = o; o += r4(M);
  L.A = o; o += n * n;
  L.B = o; o += r4(n * M);
  L.lrr = o; o += r4(lrr_floats);
  L.Rinv = o; o += r4(N * ILQG_MAX_UDIM * ILQG_MAX_UDIM);
  L.tv = o; o += r4(n);
  L.total = o;
  int s = 0;
  L.sG = s; s += n * n;
  L.sg = s; s += n;
  L.sK = s; s += M * n;
  L.skap = s; s += M;
  L.sgq = s; s += n;
  L.sw = s; s += M;
  L.srec = r4(s);
  return L;
}

// owner player of stacked control row c
__device__ __forceinline__ int ol_owner(const DevDesc& d, int c) {
  int o = 0;
  for (int i = 1; i < d.N; i++)
    if (c >= d.uoff[i]) o = i;
  return o;
}

__global__ void __launch_bounds__(KOL_WARPS * 32)
k_lq_open_loop(const __grid_constant__ DevDesc d, Slab s, float* scratch, int only_running,
               const float* x0arg) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * KOL_WARPS + warp;
  if (b >= s.B) return;
  if (only_running && !instance_iterates(s, b)) return;
  const int T = d.T, n = d.n, M = d.M, N = d.N;
  const int lrr_floats = d.rec - d.offl;
  const OlLayout L = ol_layout(n, M, N, lrr_floats);
  float* sm = smem + (size_t)warp * L.total;
  float* Mi = sm + L.Mi;
  float* mi = sm + L.mi;
  float* Lam = sm + L.Lam;
  float* X = sm + L.X;      // [n][n + 1]
  float* T1 = sm + L.T1;
  float* K = sm + L.K;      // [M][n]
  float* kap = sm + L.kap;
  float* W = sm + L.W;      // [M][n]
  float* wr = sm + L.wr;
  float* A = sm + L.A;
  float* Bm = sm + L.B;     // [n][M]
  float* lrr = sm + L.lrr;  // [l | R | r]
  float* Rinv = sm + L.Rinv;
  float* tv = sm + L.tv;
  const float* lvec = lrr;
  const float* Rk = lrr + (d.offR - d.offl);
  const float* rk = lrr + (d.offr - d.offl);
  const int nc = n + 1;

  if (lane < N) s.te_quad[(size_t)b * N + lane] = s.te_new[(size_t)b * N + lane];

  const int cand = 1 - s.st_cur[b];
  float* outP = s.st_P[cand] + (size_t)b * T * M * n;
  float* outa = s.st_a[cand] + (size_t)b * T * M;
  float* dxs = s.dxs + (size_t)b * T * n;
  const float* recb = s.rec + (size_t)b * T * d.rec;
  float* scr = scratch + (size_t)b * T * L.srec;

  // Strategy ctor zero-fills P and alpha (strategy.h:64-70); P is never touched again (:101-107)
  for (int e = lane; e < T * M * n; e += 32) outP[e] = 0.f;
  for (int e = lane; e < M; e += 32) outa[(size_t)(T - 1) * M + e] = 0.f;

  // sum_i Q_i l_i of a record (for ExpectedDecrease), from global memory
  auto store_gq = [&](int kk) {
    const float* rec = recb + (size_t)kk * d.rec;
    for (int a = lane; a < n; a += 32) {
      float g = 0.f;
      for (int i = 0; i < N; i++) {
        float gi = 0.f;
        for (int c = 0; c < n; c++)
          gi = fmaf(__ldg(rec + d.offQ + (i * n + a) * n + c), __ldg(rec + d.offl + i * n + c), gi);
        g += gi;
      }
      scr[(size_t)kk * L.srec + L.sgq + a] = g;
    }
  };

  // ---- terminal condition (:112-115) ----
  {
    const float* last = recb + (size_t)(T - 1) * d.rec;
    for (int e = lane; e < N * n * n; e += 32) Mi[e] = __ldg(last + d.offQ + e);
    for (int e = lane; e < N * n; e += 32) mi[e] = __ldg(last + d.offl + e);
    store_gq(T - 1);
  }
  __syncwarp();

  for (int kk = T - 2; kk >= 0; kk--) {
    const float* rec = recb + (size_t)kk * d.rec;
    float* out = scr + (size_t)kk * L.srec;
    for (int e = lane; e < n * n; e += 32) A[e] = __ldg(rec + d.offA + e);
    for (int e = lane; e < n * M; e += 32) Bm[e] = __ldg(rec + d.offB + e);
    for (int e = lane; e < lrr_floats; e += 32) lrr[e] = __ldg(rec + d.offl + e);
    __syncwarp();
    // ---- R_ii^-1 (Gauss-Jordan with partial pivoting, one lane per player) ----
    if (lane < N) {
      const int i = lane, m = d.udim[i], pii = d.pair_of[i][i];
      float a[ILQG_MAX_UDIM][2 * ILQG_MAX_UDIM];
      for (int r = 0; r < m; r++)
        for (int c = 0; c < m; c++) {
          a[r][c] = Rk[d.pair_Roff[pii] + r * m + c];
          a[r][m + c] = r == c ? 1.f : 0.f;
        }
      for (int k = 0; k < m; k++) {
        int piv = k;
        for (int r = k + 1; r < m; r++)
          if (fabsf(a[r][k]) > fabsf(a[piv][k])) piv = r;
        if (piv != k)
          for (int c = 0; c < 2 * m; c++) { const float t = a[k][c]; a[k][c] = a[piv][c]; a[piv][c] = t; }
        const float inv = 1.0f / a[k][k];
        for (int c = 0; c < 2 * m; c++) a[k][c] *= inv;
        for (int r = 0; r < m; r++) {
          if (r == k) continue;
          const float f = a[r][k];
          for (int c = 0; c < 2 * m; c++) a[r][c] = fmaf(-f, a[k][c], a[r][c]);
        }
      }
      for (int r = 0; r < m; r++)
        for (int c = 0; c < m; c++) Rinv[(i * ILQG_MAX_UDIM + r) * ILQG_MAX_UDIM + c] = a[r][m + c];
    }
    __syncwarp();
    // ---- W = R^-1 B^T (:125), w = R^-1 r (:126); ExpectedDecrease's R_ii r_ii ----
    for (int e = lane; e < M * n; e += 32) {
      const int c = e / n, col = e % n;
      const int i = ol_owner(d, c), ro = d.uoff[i], m = d.udim[i];
      float acc = 0.f;
      for (int q = 0; q < m; q++) acc = fmaf(Rinv[(i * ILQG_MAX_UDIM + (c - ro)) * ILQG_MAX_UDIM + q], Bm[col * M + ro + q], acc);
      W[e] = acc;
    }
    for (int c = lane; c < M; c += 32) {
      const int i = ol_owner(d, c), ro = d.uoff[i], m = d.udim[i], pii = d.pair_of[i][i];
      float acc = 0.f, rr = 0.f;
      for (int q = 0; q < m; q++) {
        acc = fmaf(Rinv[(i * ILQG_MAX_UDIM + (c - ro)) * ILQG_MAX_UDIM + q], rk[d.pair_roff[pii] + q], acc);
        rr = fmaf(Rk[d.pair_Roff[pii] + (c - ro) * m + q], rk[d.pair_roff[pii] + q], rr);
      }
      wr[c] = acc;
      out[L.sw + c] = rr;
    }
    __syncwarp();
    // ---- K = W M_i, kappa = W m_i + w ----
    for (int e = lane; e < M * n; e += 32) {
      const int c = e / n, col = e % n;
      const float* Mo = Mi + ol_owner(d, c) * n * n;
      float acc = 0.f;
      for (int q = 0; q < n; q++) acc = fmaf(W[c * n + q], Mo[q * n + col], acc);
      K[e] = acc;
      out[L.sK + e] = acc;
    }
    for (int c = lane; c < M; c += 32) {
      const float* mo = mi + ol_owner(d, c) * n;
      float acc = 0.f;
      for (int q = 0; q < n; q++) acc = fmaf(W[c * n + q], mo[q], acc);
      kap[c] = acc + wr[c];
      out[L.skap + c] = acc + wr[c];
    }
    __syncwarp();
    // ---- Lam = I + B K (:119-127); right-hand sides [A | nu], nu = -B kappa (:134-139) ----
    for (int e = lane; e < n * n; e += 32) {
      const int a = e / n, col = e % n;
      float acc = a == col ? 1.f : 0.f;
      for (int c = 0; c < M; c++) acc = fmaf(Bm[a * M + c], K[c * n + col], acc);
      Lam[e] = acc;
      X[a * nc + col] = A[e];
    }
    for (int a = lane; a < n; a += 32) {
      float acc = 0.f;
      for (int c = 0; c < M; c++) acc = fmaf(Bm[a * M + c], kap[c], acc);
      X[a * nc + n] = -acc;
    }
    __syncwarp();
    // ---- [G | g] = Lam^-1 [A | nu]: LU with partial pivoting, then back substitution ----
    for (int k = 0; k < n; k++) {
      float best = -1.f;
      int piv = k;
      for (int r = k + lane; r < n; r += 32) {
        const float v = fabsf(Lam[r * n + k]);
        if (v > best) { best = v; piv = r; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, piv, o);
        if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
      }
      if (piv != k) {
        for (int c = lane; c < n + nc; c += 32) {
          float* row_k = c < n ? Lam + k * n + c : X + k * nc + (c - n);
          float* row_p = c < n ? Lam + piv * n + c : X + piv * nc + (c - n);
          const float t = *row_k; *row_k = *row_p; *row_p = t;
        }
      }
      __syncwarp();
      const float inv = 1.0f / Lam[k * n + k];
      const int rows = n - k - 1, cols = (n - k - 1) + nc;  // trailing Lam columns, then all of X
      for (int e = lane; e < rows * cols; e += 32) {
        const int r = k + 1 + e / cols, cc = e % cols;
        const float f = Lam[r * n + k] * inv;
        if (cc < n - k - 1) {
          const int c = k + 1 + cc;
          Lam[r * n + c] = fmaf(-f, Lam[k * n + c], Lam[r * n + c]);
        } else {
          const int c = cc - (n - k - 1);
          X[r * nc + c] = fmaf(-f, X[k * nc + c], X[r * nc + c]);
        }
      }
      __syncwarp();
    }
    for (int c = lane; c < nc; c += 32) {
      for (int r = n - 1; r >= 0; r--) {
        float acc = X[r * nc + c];
        for (int q = r + 1; q < n; q++) acc = fmaf(-Lam[r * n + q], X[q * nc + c], acc);
        X[r * nc + c] = acc / Lam[r * n + r];
      }
    }
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) out[L.sG + e] = X[(e / n) * nc + e % n];
    for (int a = lane; a < n; a += 32) out[L.sg + a] = X[a * nc + n];
    store_gq(kk);
    // ---- M_i <- Q_i + A^T (M_i G) (:143-146);  m_i <- l_i + A^T (m_i + M_i g) (:147-151) ----
    for (int i = 0; i < N; i++) {
      float* Mo = Mi + i * n * n;
      float* mo = mi + i * n;
      for (int e = lane; e < n * n; e += 32) {
        const int a = e / n, col = e % n;
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(Mo[a * n + q], X[q * nc + col], acc);
        T1[e] = acc;
      }
      for (int a = lane; a < n; a += 32) {
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(Mo[a * n + q], X[q * nc + n], acc);
        tv[a] = mo[a] + acc;
      }
      __syncwarp();
      for (int e = lane; e < n * n; e += 32) {
        const int a = e / n, col = e % n;
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(A[q * n + a], T1[q * n + col], acc);
        Mo[e] = __ldg(rec + d.offQ + i * n * n + e) + acc;
      }
      for (int a = lane; a < n; a += 32) {
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc = fmaf(A[q * n + a], tv[q], acc);
        mo[a] = lvec[i * n + a] + acc;
      }
      __syncwarp();
    }
  }

  // ---- forward sweep (:158-192) + ExpectedDecrease; x* lives in tv, the next one in T1 ----
  // x0 argument: x0 - xs[0] = 0 from ILQSolver (ilq_solver.cpp:140-143), ILQG_LQ_X0 stand-alone
  for (int a = lane; a < n; a += 32) tv[a] = x0arg ? x0arg[(size_t)b * n + a] : 0.f;
  __syncwarp();
  float ed = 0.f;  // per-lane partial sums, reduced at the end
  for (int kk = 0; kk < T - 1; kk++) {
    const float* in = scr + (size_t)kk * L.srec;
    for (int a = lane; a < n; a += 32) {
      dxs[(size_t)kk * n + a] = tv[a];
      if (kk > 0) ed -= tv[a] * in[L.sgq + a];
      float acc = in[L.sg + a];
      for (int q = 0; q < n; q++) acc = fmaf(in[L.sG + a * n + q], tv[q], acc);
      T1[a] = acc;
    }
    __syncwarp();
    for (int c = lane; c < M; c += 32) {
      float acc = in[L.skap + c];
      for (int q = 0; q < n; q++) acc = fmaf(in[L.sK + c * n + q], T1[q], acc);
      outa[(size_t)kk * M + c] = acc;
      ed -= acc * in[L.sw + c];
    }
    for (int a = lane; a < n; a += 32) tv[a] = T1[a];
    __syncwarp();
  }
  for (int a = lane; a < n; a += 32) {
    dxs[(size_t)(T - 1) * n + a] = tv[a];
    ed -= tv[a] * scr[(size_t)(T - 1) * L.srec + L.sgq + a];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ed += __shfl_xor_sync(0xffffffffu, ed, o);
  if (lane == 0) s.expected_decrease[b] = ed;
}

}  // namespace ilqg
