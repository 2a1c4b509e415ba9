// ilqg_backward_tc.cuh -- K_bwd v3: LQFeedbackSolver::Solve (src/lq_feedback_solver.cpp:71-244) with one
// WARP per game, the two n x n x n products of  Z_i <- F' Z_i F  (80 % of the sweep's flops) on
// the tensor cores, and COMPACT records in (ilqg_records.cuh): the sweep never sees a dense record.
//
//  * Tensor cores: `mma.sync.aligned.m16n8k8 ... tf32` with the three-product split
//    x = hi + lo, x y ~ lo_x hi_y + hi_x lo_y + hi_x hi_y  (fp32 accumulate), which
//    profiles/r01_tf32_study.md shows is indistinguishable from fp32 products for this recursion
//    (plain TF32 is 15 x outside the parity tolerance).  X_i = F' Z_i is chained into Z_i' = X_i F
//    in registers: with the contraction slots of a k-step mapped to rows (8 ks + 2 t, 8 ks + 2 t + 1)
//    the accumulator fragment of X_i IS the A fragment of the second product, and one set of F
//    fragments (element F[8 ks + 2 t + e][8 j + g]) serves as A operand of the first product and B
//    operand of the second for every player.
//  * Structure: B_i enters only through its static non-zero list (discovered with the record
//    pattern at ilqg_create; two entries per player for the car / unicycle systems), so
//    B_i' Z_i, (B'Z) B, B' zeta, A - B P and B alpha cost a handful of FMAs instead of dense
//    products; skipped terms are exact zeros, so these sums equal the dense ones bit for bit.
//  * Per step the game's compact record (NIp floats, prefetched one step ahead in registers) is
//    scattered into small dense staging blocks (A -> the F buffer, l, R, r); Q_i is added to Z_i
//    straight from the item list, g_k comes precomputed from K_lq.
//  * Shared memory per game: Z_i (N x n x (n+4)), F, P, B'Z, S, a few vectors and the record's
//    values: 8.9 KB for ThreePlayerIntersection.  Padded state dimensions (n < NXP) and control
//    rows (M < MUP) hold zeros / identity and stay so.
//  * S X = Y as in the half-warp kernel: S redundantly in registers, lane c owns column c of P,
//    no pivoting when the Gershgorin step left S column diagonally dominant, partial pivoting in a
//    cold out-of-line path otherwise.
//  * The control-cost terms  P_j' R_ij P_j  and  P_j' (R_ij alpha_j - r_ij)  (:209-212) use the stacked
//    form  P' Omega_i P,  P' (Omega_i alpha - rho_i)  with Omega_i = blockdiag_j R_ij: the first
//    rides on the tensor cores as one more k-step of the second product, the second is an
//    unrolled M-term chain; blocks a player has no cost on are exact zeros.
//  * ExpectedDecrease through the adjoint recursion (see ilqg_backward.cuh).
#pragma once
#include "ilqg_backward.cuh"

namespace ilqg {

constexpr int KTC_WARPS = 4;

// static index tables of one descriptor (device pointers; copied to shared memory per block)
struct TcTables {
  const unsigned* words;   // [scat | bnz | brow_rows | brow | qadd], offsets below
  int total_words;
  int scat, nscat;         // item | smem offset (relative to the game's base) << 16
  int bnz, nbnz;           // q | c << 8 | item << 16, sorted by (c, q)
  int brow_rows, nbrow_rows;  // q | start << 8 | count << 20: rows of B that hold non-zeros
  int brow;                // c | player << 8 | item << 16, grouped by row, sorted by (player, c)
  int qadd, nqadd;         // item (or kItemReg + i) | offset into Z << 16
  int bnz_start[ILQG_MAX_UDIM + 1];
  int NI, NIp;
  // run-time part of the per-game layout (floats): the arrays sized by the number of players
  int off_zeta, off_l, off_Om, off_rho, off_om, off_edr, off_vals, off_Z, per_game;
};

// fixed part of the per-game shared-memory layout (floats); the arrays sized by the number of players
// follow at TcTables::off_*:
//   zeta [N][NXP], l [N][NXP]
//   Om   [N][MUP][MUP]  Omega_i: player i's control Hessians R_ij as blocks of one stacked M x M matrix
//   rho  [N][8]         the matching stacked gradients r_ij;  om, edr [N][8]: per-step control-cost vectors
//   vals [2][NIp]       this step's compact record and the next one in flight;  Z [N][NXP][LD]
// Shared by the kernel and by the host code that builds the tables.
struct TcFixed {
  int LD, F, P, BZt, tv, ya, beta, pv, pn, fixed;
};
__host__ __device__ constexpr TcFixed tc_fixed(int NXP, int MUP) {
  TcFixed L{};
  L.LD = NXP + 4;
  L.F = 0;                          // [NXP][LD]  A, then F = A - B P
  L.P = L.F + NXP * L.LD;           // [MUP][LD]  Y, then the solution P
  L.BZt = L.P + MUP * L.LD;         // [NXP][8]   (B_i' Z_i)' ...
  L.tv = L.BZt;                     // [4][NXP]   ... later zeta_i + Z_i beta (B'Z is dead by then)
  L.ya = L.BZt + (NXP * 8 > 4 * NXP ? NXP * 8 : 4 * NXP);  // [8] y_alpha, then alpha
  L.beta = L.ya + 8;                // [NXP]
  L.pv = L.beta + NXP;              // [NXP]
  L.pn = L.pv + NXP;                // [NXP]
  L.fixed = L.pn + NXP;
  return L;
}

// ---- record staging: one bulk async copy (TMA, 1-D) per game and step, completion on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(mbar),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ unsigned to_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// x = hi + lo with hi on the TF32 grid (round to nearest by integer arithmetic) and lo = x - hi exact in
// fp32.  cvt.rna.tf32.f32 issues at a quarter of the ALU rate on sm_100 (tools/mma_rate_bench.cu:
// 11.7 vs 5.0 cycles per value and sub-partition) and the sweep splits ~56 values per game and step;
// the tensor core ignores the low 13 mantissa bits of its operands (same bench: raw fp32 operands give
// the results of hand-truncated ones, 128 / 128), so lo needs no conversion of its own.
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                         unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// the pivoting solve, kept out of line: it runs only when the Gershgorin step could not make S
// column diagonally dominant (a diagonal below -1e-3), and inlined it costs the hot path registers
template <int MU>
__device__ __noinline__ void lu_solve_pivot(float* Sm /*[MU][MU]*/, float* y /*[MU]*/, float* yal /*[MU]*/) {
  for (int k = 0; k < MU; k++) {
    int piv = k;
    float best = fabsf(Sm[k * MU + k]);
    for (int r = k + 1; r < MU; r++) {
      const float v = fabsf(Sm[r * MU + k]);
      if (v > best) { best = v; piv = r; }
    }
    if (piv != k) {
      for (int c = 0; c < MU; c++) { const float t = Sm[k * MU + c]; Sm[k * MU + c] = Sm[piv * MU + c]; Sm[piv * MU + c] = t; }
      { const float t = y[k]; y[k] = y[piv]; y[piv] = t; }
      { const float t = yal[k]; yal[k] = yal[piv]; yal[piv] = t; }
    }
    const float inv = 1.0f / Sm[k * MU + k];
    Sm[k * MU + k] = inv;
    for (int r = k + 1; r < MU; r++) {
      const float f = Sm[r * MU + k] * inv;
      for (int c = k + 1; c < MU; c++) Sm[r * MU + c] = fmaf(-f, Sm[k * MU + c], Sm[r * MU + c]);
      y[r] = fmaf(-f, y[k], y[r]);
      yal[r] = fmaf(-f, yal[k], yal[r]);
    }
  }
  for (int r = MU - 1; r >= 0; r--) {
    float acc = y[r], acc2 = yal[r];
    for (int c = r + 1; c < MU; c++) {
      acc = fmaf(-Sm[r * MU + c], y[c], acc);
      acc2 = fmaf(-Sm[r * MU + c], yal[c], acc2);
    }
    y[r] = acc * Sm[r * MU + r];
    yal[r] = acc2 * Sm[r * MU + r];
  }
}

template <int NXP, int MUP, int SC, int QA, int MINB>
__global__ void __launch_bounds__(KTC_WARPS * 32, MINB)
k_lq_backward_tc(const __grid_constant__ DevDesc d, const DevParams p, Slab s, const __grid_constant__ TcTables tb,
                 int only_running, Sel sel, const float* x0arg) {
  constexpr TcFixed L = tc_fixed(NXP, MUP);
  constexpr int LD = L.LD;
  constexpr int W = NXP / 4;              // columns per lane in the (control row, column group) mapping
  constexpr int KS = NXP / 8;             // k-steps = n-tiles = 8-column groups
  constexpr int MT = (NXP + 15) / 16;     // 16-row tiles
  constexpr int MM = r4(MUP * MUP);
  static_assert(NXP % 8 == 0 && NXP <= 32 && MUP <= 8 && MUP % 2 == 0, "shape outside the tensor-core kernel's envelope");
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n = d.n, M = d.M, NP = d.N, T = d.T;

  // ---- block prologue: the index tables ----
  unsigned* tab = reinterpret_cast<unsigned*>(smem);
  for (int e = threadIdx.x; e < tb.total_words; e += blockDim.x) tab[e] = __ldg(tb.words + e);
  // one mbarrier per warp (after the tables, 8-byte aligned): completion of the record copy in flight
  const unsigned mbar = smem_u32(smem + r4(tb.total_words)) + 8 * warp;
  if (lane == 0) mbar_init(mbar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const unsigned* t_scat = tab + tb.scat;
  const unsigned* t_bnz = tab + tb.bnz;
  const unsigned* t_brow_rows = tab + tb.brow_rows;
  const unsigned* t_brow = tab + tb.brow;
  const unsigned* t_qadd = tab + tb.qadd;

  bool active = false;
  const int b = sel_instance(s, sel, blockIdx.x * KTC_WARPS + warp, only_running, &active);
  if (!active) return;

  float* sm = smem + r4(tb.total_words) + 2 * KTC_WARPS + (size_t)warp * tb.per_game;
  float* Fs = sm + L.F;
  float* Pm = sm + L.P;
  float* BZt = sm + L.BZt;
  float* tv = sm + L.tv;
  float* ya = sm + L.ya;
  float* zeta = sm + tb.off_zeta;
  float* beta = sm + L.beta;
  float* pv = sm + L.pv;
  float* pn = sm + L.pn;
  float* lst = sm + tb.off_l;
  float* Om = sm + tb.off_Om;
  float* rho = sm + tb.off_rho;
  float* om = sm + tb.off_om;
  float* edr = sm + tb.off_edr;
  float* vals = sm + tb.off_vals;          // the record of this step; the next one lands NIp floats further
  float* vals_next = vals + tb.NIp;
  float* Zs = sm + tb.off_Z;     // [NP][NXP][LD]
  const int NI = tb.NI;
  const unsigned rec_bytes = 4u * tb.NIp;
  unsigned phase = 0;

  // everything starts at zero (padding stays zero for the whole sweep); then the template values
  for (int e = lane; e < tb.per_game; e += 32) sm[e] = 0.f;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the record buffers are written by the async proxy next
  if (lane < n) Fs[lane * LD + lane] = 1.f;
  if (lane < d.num_pairs) {
    const int pi = d.pair_i[lane], pj = d.pair_j[lane];
    const int mj = d.udim[pj], co = d.uoff[pj];
    for (int a = 0; a < mj; a++) Om[pi * MM + (co + a) * MUP + co + a] = d.control_reg[pi];
  }
  if (lane < NP) s.te_quad[(size_t)b * NP + lane] = s.te_new[(size_t)b * NP + lane];

  const int cand = 1 - s.st_cur[b];
  float* outP = s.st_P[cand] + (size_t)b * T * M * n;
  float* outa = s.st_a[cand] + (size_t)b * T * M;
  const float* crec = s.crec + (size_t)b * T * tb.NIp;
  const int Mn = M * n;

  // owner player of the stacked control row this lane works on in the (row, column group) phases
  const int crow = lane >> 2;
  int own = 0;
  for (int i = 1; i < NP; i++)
    if (crow >= d.uoff[i]) own = i;
  int ownS = 0;  // owner of stacked control row `lane` (the S column this lane builds)
  for (int i = 1; i < NP; i++)
    if (lane >= d.uoff[i]) ownS = i;
  // (player, stacked control row) this lane works on in the control-cost phase
  const int wi = lane >> 3, wc = lane & 7;
  const int wi_uo = wi < NP ? d.uoff[wi] : 0, wi_m = wi < NP ? d.udim[wi] : 0;

  // record k -> the buffer that is not in use (one elected lane; everybody waits on the mbarrier later)
  auto prefetch = [&](int k) {
    if (lane == 0) {
      mbar_expect_tx(mbar, rec_bytes);
      bulk_g2s(smem_u32(vals_next), crec + (size_t)k * tb.NIp, rec_bytes, mbar);
    }
  };
  // wait for the record in flight and make it the current one; A = I + items -> Fs (only the rows
  // F = A - B P dirtied need resetting: every other word of Fs is either template or an item's);
  // l, R, r items -> their staging blocks
  auto expand = [&]() {
    mbar_wait(mbar, phase);
    phase ^= 1;
    { float* tmp = vals; vals = vals_next; vals_next = tmp; }
    for (int ri = crow; ri < tb.nbrow_rows; ri += 8) {
      const int q = t_brow_rows[ri] & 0xff;
      float z[W];
#pragma unroll
      for (int j = 0; j < W; j++) z[j] = (t * W + j == q) ? 1.f : 0.f;
      stvec<W>(Fs + q * LD + t * W, z);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SC; j++) {  // (tables are padded to 32 SC entries with writes to a scratch word)
      const unsigned u = t_scat[lane + 32 * j];
      sm[u >> 16] = vals[u & 0xffffu];
    }
    __syncwarp();
  };
  // Z_i += Q_i (items and the constant state regulariser on the diagonal)
  auto add_Q = [&]() {
#pragma unroll
    for (int j = 0; j < QA; j++) {  // (padded likewise; the regulariser constants ride in the record's tail)
      const unsigned u = t_qadd[lane + 32 * j];
      Zs[u >> 16] += vals[u & 0xffffu];
    }
    __syncwarp();
  };

  // ---- terminal condition (:102-105): Z_i = Q_i[T-1], zeta_i = l_i[T-1]; p_{T-1} = g_{T-1} ----
  prefetch(T - 1);
  expand();
  if (T >= 2) prefetch(T - 2);
  add_Q();
  for (int e = lane; e < NP * NXP; e += 32) zeta[e] = lst[e];
  if (lane < n) pv[lane] = vals[NI + lane];
  for (int e = lane; e < M * n; e += 32) outP[(size_t)(T - 1) * M * n + e] = 0.f;
  for (int e = lane; e < M; e += 32) outa[(size_t)(T - 1) * M + e] = 0.f;
  __syncwarp();

  float expected_decrease = 0.f;  // accumulated redundantly by every lane

  for (int kk = T - 2; kk >= 0; kk--) {
    expand();
    if (kk > 0) prefetch(kk - 1);

    // ---- BZ = B_i' Z_i (:128) and y_alpha = B_i' zeta_i + r_ii (:154-157), B through its non-zeros ----
    if (crow < M) {
      float acc[W];
#pragma unroll
      for (int j = 0; j < W; j++) acc[j] = 0.f;
      float ay = 0.f;
      const float* Zo = Zs + own * NXP * LD + t * W;
      const float* zo = zeta + own * NXP;
      for (int e = tb.bnz_start[crow]; e < tb.bnz_start[crow + 1]; e++) {
        const unsigned u = t_bnz[e];
        const int q = u & 0xff;
        const float bv = vals[u >> 16];
        float zr[W];
        ldvec<W>(Zo + q * LD, zr);
#pragma unroll
        for (int j = 0; j < W; j++) acc[j] = fmaf(bv, zr[j], acc[j]);
        ay = fmaf(bv, zo[q], ay);
      }
#pragma unroll
      for (int j = 0; j < W; j++) BZt[(t * W + j) * 8 + crow] = acc[j];
      if (t == 0) ya[crow] = ay + rho[own * 8 + crow];
    }
    __syncwarp();
    // ---- Y = BZ A (:152-153) into Pm; S = BZ B (+ R_ii) (:131-149) ----
    if (crow < M) {
      float acc[1][W];
      mm_tn<NXP, 1, W>(BZt + crow, 8, Fs + t * W, LD, acc);
      stvec<W>(Pm + crow * LD + t * W, acc[0]);
    }
    __syncwarp();

    // ---- Gershgorin (:163-176) + S X = Y (:180), column-distributed: a lane owns one column of the
    //      augmented matrix [S | Y | y_alpha] (MUP rows in registers); pivots and multipliers travel by
    //      shuffle.  Same operations, in the same order, as the per-lane LU of the half-warp kernel. ----
    {
      constexpr bool WIDE = MUP + NXP + 1 <= 32;  // else a lane owns a Y column AND an S / alpha column
      constexpr unsigned FULL = 0xffffffffu;
      // S column c = lane (< MUP): S[r][c] = sum_q BZ[r][q] B[q][c] (+ R_ii) (:131-149), B through its non-zeros
      float sv[MUP];
#pragma unroll
      for (int r = 0; r < MUP; r++) sv[r] = 0.f;
      if (lane < M) {
        for (int e = tb.bnz_start[lane]; e < tb.bnz_start[lane + 1]; e++) {
          const unsigned u = t_bnz[e];
          const float bv = vals[u >> 16];
          float bz[8];
          ldvec<8>(BZt + (u & 0xff) * 8, bz);
#pragma unroll
          for (int r = 0; r < MUP; r++) sv[r] = fmaf(bz[r], bv, sv[r]);
        }
        const float* Oc = Om + ownS * MM + lane;  // Omega_own is block diagonal: zero outside the owner's rows
#pragma unroll
        for (int r = 0; r < MUP; r++) sv[r] = sv[r] + Oc[r * MUP];
      } else if (lane < MUP) {
#pragma unroll
        for (int r = 0; r < MUP; r++) sv[r] = lane == r ? 1.f : 0.f;  // padded control rows: identity
      }
      bool dominant = p.adaptive_regularization != 0;
      if (p.adaptive_regularization) {
        float col1 = 0.f, diag = sv[0];
#pragma unroll
        for (int r = 0; r < MUP; r++) {
          col1 += fabsf(sv[r]);
          if (r > 0) diag = lane == r ? sv[r] : diag;
        }
        const float radius = col1 - fabsf(diag);
        const float eval_lo = diag - radius;
        constexpr float min_eval = 1e-3;
        const float nd = eval_lo < min_eval ? diag + (radius + min_eval) : diag;
#pragma unroll
        for (int r = 0; r < MUP; r++) sv[r] = lane == r ? nd : sv[r];
        dominant = __all_sync(FULL, lane >= MUP || nd > radius);
      }
      // v1: the lane's S column (lane < MUP), else -- WIDE -- its Y column (lanes MUP .. MUP + NXP - 1) or y_alpha
      // (lane MUP + NXP); not WIDE: y_alpha in lane MUP, and the Y column of lane < NXP in v2
      float v1[MUP], v2[WIDE ? 1 : MUP];
      {
        const bool ycl = WIDE && lane >= MUP && lane < MUP + NXP;
        const float* colp = ycl ? Pm + (lane - MUP) : ya;  // (S lanes read y_alpha too and discard it)
        const int cs = ycl ? LD : 1;
#pragma unroll
        for (int r = 0; r < MUP; r++) {
          const float ld = colp[r * cs];
          v1[r] = lane < MUP ? sv[r] : ld;
          if (!WIDE) v2[r] = lane < NXP ? Pm[r * LD + lane] : 0.f;
        }
      }
      float invk[MUP];
#pragma unroll
      for (int k = 0; k < MUP; k++) {
        if (!dominant) {  // partial pivoting (warp-uniform branch): lane k picks the row, every column swaps
          int piv = k;
          float best = fabsf(v1[k]);
#pragma unroll
          for (int r = k + 1; r < MUP; r++) {
            const float v = fabsf(v1[r]);
            if (v > best) { best = v; piv = r; }
          }
          piv = __shfl_sync(FULL, piv, k);
#pragma unroll
          for (int r = k + 1; r < MUP; r++)
            if (piv == r) {
              { const float tmp = v1[k]; v1[k] = v1[r]; v1[r] = tmp; }
              if (!WIDE) { const float tmp = v2[k]; v2[k] = v2[r]; v2[r] = tmp; }
            }
        }
        const float inv = __shfl_sync(FULL, div_rn(1.0f, v1[k]), k);
        invk[k] = inv;
#pragma unroll
        for (int r = k + 1; r < MUP; r++) {
          const float f = __shfl_sync(FULL, v1[r] * inv, k);
          v1[r] = fmaf(-f, v1[k], v1[r]);
          if (!WIDE) v2[r] = fmaf(-f, v2[k], v2[r]);
        }
      }
#pragma unroll
      for (int r = MUP - 1; r >= 0; r--) {
        float acc1 = v1[r], acc2 = WIDE ? 0.f : v2[r];
#pragma unroll
        for (int c = r + 1; c < MUP; c++) {
          const float u = __shfl_sync(FULL, v1[r], c);  // U[r][c]: entry r of S column c
          acc1 = fmaf(-u, v1[c], acc1);
          if (!WIDE) acc2 = fmaf(-u, v2[c], acc2);
        }
        // S columns keep their entries (rows above still read them); right-hand sides take the solution
        if (lane >= MUP) v1[r] = acc1 * invk[r];
        if (!WIDE) v2[r] = acc2 * invk[r];
      }
      __syncwarp();  // every lane has read Pm (= Y) and ya
      const int ycol = WIDE ? lane - MUP : lane;          // the state column this lane solved for
      const int alane = WIDE ? MUP + NXP : MUP;           // the lane that solved for alpha
      if (ycol >= 0 && ycol < n) {
        float* op = outP + kk * Mn + ycol;
        float* pp = Pm + ycol;
#pragma unroll
        for (int r = 0; r < MUP; r++)
          if (r < M) {
            const float x = WIDE ? v1[r] : v2[r];
            pp[r * LD] = x;
            *op = x;
            op += n;
          }
      }
      if (lane == alane) {
        float* oa = outa + kk * M;
#pragma unroll
        for (int r = 0; r < MUP; r++)
          if (r < M) {
            ya[r] = v1[r];
            oa[r] = v1[r];
          }
      }
    }
    __syncwarp();
    // ---- control-cost vectors: omega_i = Omega_i alpha - rho_i (:209-211) and the rows
    //      (alpha_i' R_ii) of ExpectedDecrease (src/ilq_solver.cpp:384-386); lane = (player, control row) ----
    if (wi < NP && wc < MUP) {
      float al[8];
      ldvec<8>(ya, al);
      const float* Oi = Om + wi * MM;
      float o = 0.f, row = 0.f;
#pragma unroll
      for (int a = 0; a < MUP; a++) {
        o = fmaf(Oi[wc * MUP + a], al[a], o);
        if (a >= wi_uo && a < wi_uo + wi_m) row = fmaf(al[a], Oi[a * MUP + wc], row);
      }
      om[lane] = o - rho[lane];
      edr[lane] = row;
    }
    // p_k = A_k' p_{k+1} + g_k; Fs still holds A here
    if (lane < NXP) {
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int q = 0; q < NXP; q += 2) {
        acc0 = fmaf(Fs[q * LD + lane], pv[q], acc0);
        acc1 = fmaf(Fs[(q + 1) * LD + lane], pv[q + 1], acc1);
      }
      pn[lane] = (acc0 + acc1) + (lane < n ? vals[NI + lane] : 0.f);
    }
    __syncwarp();
    {
      // (alpha_i' R_ii) r_ii: one lane per player runs the reference's chain over its own controls
      float t1 = 0.f;
      {
        const int pl = lane & 3;  // (every lane computes some player's sum; lanes 0 .. NP - 1 are read below)
        const int uo = pl < NP ? d.uoff[pl] : 0, mi = pl < NP ? d.udim[pl] : 0;
        float er[8], rr[8];
        ldvec<8>(edr + (pl < NP ? pl : 0) * 8, er);
        ldvec<8>(rho + (pl < NP ? pl : 0) * 8, rr);
#pragma unroll
        for (int c = 0; c < MUP; c++)
          if (c >= uo && c < uo + mi) t1 = fmaf(er[c], rr[c], t1);
      }
#pragma unroll
      for (int i = 0; i < ILQG_MAX_PLAYERS; i++) {
        const float ti = __shfl_sync(0xffffffffu, t1, i);
        if (i < NP) expected_decrease -= ti;
      }
    }
    // ---- F = A - sum_i B_i P_i ; beta = - sum_i B_i alpha_i (:189-194): only the rows B touches ----
    for (int ri = crow; ri < tb.nbrow_rows; ri += 8) {
      const unsigned u = t_brow_rows[ri];
      const int q = u & 0xff, st = (u >> 8) & 0xfff, cnt = u >> 20;
      float f[W], acc[W];
      ldvec<W>(Fs + q * LD + t * W, f);
#pragma unroll
      for (int j = 0; j < W; j++) acc[j] = 0.f;
      float bacc = 0.f, bsum = 0.f;
      int cur = (t_brow[st] >> 8) & 0xff;
      for (int e = st; e < st + cnt; e++) {
        const unsigned v = t_brow[e];
        const int c = v & 0xff, pl = (v >> 8) & 0xff;
        if (pl != cur) {
#pragma unroll
          for (int j = 0; j < W; j++) { f[j] -= acc[j]; acc[j] = 0.f; }
          bsum -= bacc;
          bacc = 0.f;
          cur = pl;
        }
        const float bv = vals[v >> 16];
        float pr[W];
        ldvec<W>(Pm + c * LD + t * W, pr);
#pragma unroll
        for (int j = 0; j < W; j++) acc[j] = fmaf(bv, pr[j], acc[j]);
        bacc = fmaf(bv, ya[c], bacc);
      }
#pragma unroll
      for (int j = 0; j < W; j++) f[j] -= acc[j];
      bsum -= bacc;
      stvec<W>(Fs + q * LD + t * W, f);
      if (t == 0) beta[q] = bsum;
    }
    __syncwarp();
    // state part of ExpectedDecrease via the adjoint: beta_k . p_{k+1}
    {
      float part = lane < NXP ? beta[lane] * pv[lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      expected_decrease -= part;
    }
    // ---- tv_i = zeta_i + Z_i beta (:199), all players at once: one row per lane ----
    {
      float bt[NXP];
      ldvec<NXP>(beta, bt);
      for (int r = lane; r < NP * NXP; r += 32) {
        const float* zr = Zs + r * LD;  // row a of player i: (i NXP + a) LD
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int c = 0; c < NXP; c += 4) {
          const float4 z4 = *reinterpret_cast<const float4*>(zr + c);
          a0 = fmaf(z4.x, bt[c], a0);
          a1 = fmaf(z4.y, bt[c + 1], a1);
          a2 = fmaf(z4.z, bt[c + 2], a2);
          a3 = fmaf(z4.w, bt[c + 3], a3);
        }
        tv[r] = zeta[r] + ((a0 + a1) + (a2 + a3));
      }
    }
    __syncwarp();
    // ---- zeta_i = F' tv_i + l_i + P' omega_i (:197-199, 209-211) ----
    for (int r = lane; r < NP * NXP; r += 32) {
      const int i = r / NXP, a = r - i * NXP;
      const float* tvi = tv + i * NXP;
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int c = 0; c < NXP; c += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(tvi + c);
        acc0 = fmaf(Fs[c * LD + a], t4.x, acc0);
        acc1 = fmaf(Fs[(c + 1) * LD + a], t4.y, acc1);
        acc0 = fmaf(Fs[(c + 2) * LD + a], t4.z, acc0);
        acc1 = fmaf(Fs[(c + 3) * LD + a], t4.w, acc1);
      }
      float tt = 0.f;
      float omi[8];
      ldvec<8>(om + i * 8, omi);
#pragma unroll
      for (int c = 0; c < MUP; c++) tt = fmaf(Pm[c * LD + a], omi[c], tt);
      zeta[r] = ((acc0 + acc1) + lst[r]) + tt;
    }

    // ---- Z_i = F' Z_i F + P' Omega_i P (+ Q_i below) on the tensor cores ----
    // F fragments: element F[8 ks + 2 t + e][8 j + g], hi / lo
    unsigned fh[KS][2][KS], fl[KS][2][KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int j = 0; j < KS; j++) split_tf32(Fs[(8 * ks + 2 * t + e) * LD + 8 * j + g], fh[ks][e][j], fl[ks][e][j]);
    // P fragments: as A operand of  XP_i = P' Omega_i  (rows 8 j + g, slots t, t + 4 <-> control rows t, t + 4)
    // and as B operand of  XP_i P  (slots t, t + 4 <-> control rows 2 t, 2 t + 1): elements P[c][8 j + g]
    unsigned pah[2][KS], pal[2][KS], pbh[2][KS], pbl[2][KS];
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int j = 0; j < KS; j++) {
        const int ca = t + 4 * e, cb = 2 * t + e;
        split_tf32(ca < MUP ? Pm[ca * LD + 8 * j + g] : 0.f, pah[e][j], pal[e][j]);
        split_tf32(cb < MUP ? Pm[cb * LD + 8 * j + g] : 0.f, pbh[e][j], pbl[e][j]);
      }

#pragma unroll 1
    for (int i = 0; i < NP; i++) {
      float* Zi = Zs + i * NXP * LD;
      float x[MT][KS][4];
      {
        unsigned zh[KS][2][KS], zl[KS][2][KS];
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
          for (int e = 0; e < 2; e++)
#pragma unroll
            for (int j = 0; j < KS; j++) split_tf32(Zi[(8 * ks + 2 * t + e) * LD + 8 * j + g], zh[ks][e][j], zl[ks][e][j]);
        // X = F' Z_i: A = F' (rows 16 mt + g, g + 8), B = Z_i
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
          for (int nt = 0; nt < KS; nt++) {
#pragma unroll
            for (int q = 0; q < 4; q++) x[mt][nt][q] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
              const bool up = 2 * mt + 1 < KS;  // rows 16 mt + 8 + g exist
              const int ju = up ? 2 * mt + 1 : 0;
              const unsigned ah0 = fh[ks][0][2 * mt], ah2 = fh[ks][1][2 * mt];
              const unsigned al0 = fl[ks][0][2 * mt], al2 = fl[ks][1][2 * mt];
              const unsigned ah1 = up ? fh[ks][0][ju] : 0u, ah3 = up ? fh[ks][1][ju] : 0u;
              const unsigned al1 = up ? fl[ks][0][ju] : 0u, al3 = up ? fl[ks][1][ju] : 0u;
              mma_tf32(x[mt][nt], al0, al1, al2, al3, zh[ks][0][nt], zh[ks][1][nt]);  // small terms first
              mma_tf32(x[mt][nt], ah0, ah1, ah2, ah3, zl[ks][0][nt], zl[ks][1][nt]);
              mma_tf32(x[mt][nt], ah0, ah1, ah2, ah3, zh[ks][0][nt], zh[ks][1][nt]);
            }
          }
      }
      // XP = P' Omega_i (n x M): one k-step over the stacked control rows
      float xp[MT][4];
      {
        const float* Oi = Om + i * MM;
        unsigned oh[2], ol[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int k = t + 4 * e;
          split_tf32((k < MUP && g < MUP) ? Oi[k * MUP + g] : 0.f, oh[e], ol[e]);
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
          const bool up = 2 * mt + 1 < KS;
          const int ju = up ? 2 * mt + 1 : 0;
#pragma unroll
          for (int q = 0; q < 4; q++) xp[mt][q] = 0.f;
          const unsigned ah0 = pah[0][2 * mt], ah2 = pah[1][2 * mt], al0 = pal[0][2 * mt], al2 = pal[1][2 * mt];
          const unsigned ah1 = up ? pah[0][ju] : 0u, ah3 = up ? pah[1][ju] : 0u;
          const unsigned al1 = up ? pal[0][ju] : 0u, al3 = up ? pal[1][ju] : 0u;
          mma_tf32(xp[mt], al0, al1, al2, al3, oh[0], oh[1]);
          mma_tf32(xp[mt], ah0, ah1, ah2, ah3, ol[0], ol[1]);
          mma_tf32(xp[mt], ah0, ah1, ah2, ah3, oh[0], oh[1]);
        }
      }
      // Z_i' = X F + XP P: the accumulator fragment of X (rows g, g + 8; columns 8 ks + 2 t, + 1) is the A
      // fragment of k-step ks under the slot map (t -> 8 ks + 2 t, t + 4 -> 8 ks + 2 t + 1)
      float z[MT][KS][4];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        unsigned xh[KS + 1][4], xl[KS + 1][4];
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
          for (int q = 0; q < 4; q++) split_tf32(x[mt][ks][q], xh[ks][q], xl[ks][q]);
#pragma unroll
        for (int q = 0; q < 4; q++) split_tf32(xp[mt][q], xh[KS][q], xl[KS][q]);
#pragma unroll
        for (int nt = 0; nt < KS; nt++) {
#pragma unroll
          for (int q = 0; q < 4; q++) z[mt][nt][q] = 0.f;
          mma_tf32(z[mt][nt], xl[KS][0], xl[KS][2], xl[KS][1], xl[KS][3], pbh[0][nt], pbh[1][nt]);
          mma_tf32(z[mt][nt], xh[KS][0], xh[KS][2], xh[KS][1], xh[KS][3], pbl[0][nt], pbl[1][nt]);
          mma_tf32(z[mt][nt], xh[KS][0], xh[KS][2], xh[KS][1], xh[KS][3], pbh[0][nt], pbh[1][nt]);
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            mma_tf32(z[mt][nt], xl[ks][0], xl[ks][2], xl[ks][1], xl[ks][3], fh[ks][0][nt], fh[ks][1][nt]);
            mma_tf32(z[mt][nt], xh[ks][0], xh[ks][2], xh[ks][1], xh[ks][3], fl[ks][0][nt], fl[ks][1][nt]);
            mma_tf32(z[mt][nt], xh[ks][0], xh[ks][2], xh[ks][1], xh[ks][3], fh[ks][0][nt], fh[ks][1][nt]);
          }
        }
      }
      __syncwarp();  // every lane has its fragments of Z_i
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < KS; nt++) {
          *reinterpret_cast<float2*>(Zi + (16 * mt + g) * LD + 8 * nt + 2 * t) = make_float2(z[mt][nt][0], z[mt][nt][1]);
          if (16 * mt + 8 + g < NXP)
            *reinterpret_cast<float2*>(Zi + (16 * mt + 8 + g) * LD + 8 * nt + 2 * t) = make_float2(z[mt][nt][2], z[mt][nt][3]);
        }
    }
    __syncwarp();
    add_Q();
    if (lane < NXP) pv[lane] = pn[lane];
    __syncwarp();
  }
  // a stand-alone solve with delta_x_0 = x0 != 0 (LQFeedbackSolver::Solve's x0 argument): the adjoint sum
  // lacks  sum_{k >= 1} of the part of delta_x_k driven by x0  =  x0 . (A_0' p_1)  =  x0 . (p_0 - g_0)
  if (x0arg) {
    float part = lane < n ? x0arg[(size_t)b * n + lane] * (pv[lane] - vals[NI + lane]) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    expected_decrease -= part;
  }
  if (lane == 0) s.expected_decrease[b] = expected_decrease;
}

}  // namespace ilqg
