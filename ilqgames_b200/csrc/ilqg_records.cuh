// ilqg_records.cuh -- K_lq: fused ComputeLinearization + ComputeCostQuadraticization
// (src/ilq_solver.cpp:437-455, 471-490) producing the dense LQ records.
//
// Every cost / constraint / subsystem kind emits a FIXED sequence of `+= v` updates for given
// descriptor indices (ilqg_device.cuh), so WHERE a record is touched is a static property of
// the problem; only the values depend on (x, u, lambda, mu).  At handle creation
// k_record_pattern records the offsets once and the host groups them into a gather table:
//   entry e of role r  ->  record offset;  gather item g = (offset, role, entries to sum in order)
// A block owns 32 consecutive (instance, timestep) records:
//   phase 1  one ROLE per warp (player i's costs; the last warp all subsystem Jacobians), one
//            record per lane -- the scalar cost code runs at full SIMT width -- writing only the
//            update VALUES into a lane-minor shared array;
//   phase 2  each warp keeps a staging copy of the record template (A = I, Q_i = state_reg I,
//            R = control_reg I, zeros elsewhere); per record, lane g adds its gather item's values
//            (in the reference's accumulation order) onto the template value, the warp streams the
//            record to HBM with 128-bit stores and restores the touched words.
#pragma once
#include "ilqg_kernels.cuh"

namespace ilqg {

struct GatherItem {
  int off;     // offset in the record
  int role;    // which role's value list
  int start;   // first index into gather_idx (used when count > 4)
  int count;   // entries to add, in order
  float base;  // template value at `off` (0, 1 on A's diagonal, the regularisers on Q / R diagonals)
  unsigned short e[4];  // the first four entry indices inline
  int pad;
};

struct RecordPattern {
  const GatherItem* items;          // [num_items]
  const unsigned short* idx;        // entry indices
  const float* tmpl;                // [rec] record template
  int num_items, num_idx;
  int E;                            // value-list capacity per role (multiple of 4)
};

// sink that stores update values (phase 1) ...
// value of update e for record r at val[e * 33 + r]: the odd stride keeps both access patterns
// conflict-free (phase 1: 32 records of one update; phase 2: 32 updates of one record)
constexpr int kValStride = 33;
struct ValueSink {
  float* val;  // [E][33]
  int lane, cnt, cap;
  __device__ __forceinline__ void push(float v) {
    if (cnt < cap) val[cnt * kValStride + lane] = v;
    cnt++;
  }
  __device__ __forceinline__ void H(int, int, float v) { push(v); }
  __device__ __forceinline__ void G(int, float v) { push(v); }
};
struct LinValueSink {
  ValueSink* vs;
  __device__ __forceinline__ void addA(int, int, float v) { vs->push(v); }
  __device__ __forceinline__ void addB(int, int, float v) { vs->push(v); }
};
// ... and sink that stores update offsets (pattern discovery)
struct OffsetSink {
  int* off;
  int cnt, cap;
  int base_H, ld, base_G;
  __device__ __forceinline__ void push(int o) {
    if (cnt < cap) off[cnt] = o;
    cnt++;
  }
  __device__ __forceinline__ void H(int r, int c, float) { push(base_H + r * ld + c); }
  __device__ __forceinline__ void G(int i, float) { push(base_G + i); }
};
struct LinOffsetSink {
  OffsetSink* os;
  int offA, offB, n, M;
  __device__ __forceinline__ void addA(int r, int c, float) { os->push(offA + r * n + c); }
  __device__ __forceinline__ void addB(int r, int c, float) { os->push(offB + r * M + c); }
};

// The walk over one role's records, shared by the value pass and the pattern pass so both see
// the same sequence.  `full` = PlayerCost::Quadraticize, else QuadraticizeControlCosts
// (src/ilq_solver.cpp:483-487); in the latter case the skipped records still emit (zeros are
// pushed by the caller through `skip`).
template <int XS, bool WIDE = true, class PlayerFn, class LinFn, class ChooseFn>
__device__ __forceinline__ void walk_role(const DevDesc& d, int role, PlayerFn&& player, LinFn&& lin, ChooseFn&& choose) {
  if (role < d.N) {
    for (int c = d.cost_begin[role]; c < d.cost_begin[role + 1];) {
      const DevCost& cd = d.cost[c];
      if (WIDE && cd.group > 0) {
        // an ExtremeValueCost: every member emits its updates, only the extreme one non-zero
        const int winner = choose(c);
        for (int m = c; m < cd.group_end; m++) player(d.cost[m], m == winner);
        c = cd.group_end;
      } else {
        player(cd, true);
        c++;
      }
    }
  } else {
    for (int sidx = 0; sidx < d.num_subsystems; sidx++) lin(d.sub[sidx]);
  }
}

// number of updates a record kind emits (HESS = true) -- only used to pad skipped records
__device__ __forceinline__ int record_updates(const DevCost& cd, int dim) {
  switch (cd.kind) {
    case ILQG_COST_QUADRATIC: return cd.d0 >= 0 ? 2 : 2 * dim;
    case ILQG_COST_PROXIMITY:
    case ILQG_CONSTRAINT_PROXIMITY:
    case ILQG_COST_SIGNED_DISTANCE: return 20;
    case ILQG_COST_QUADRATIC_DIFFERENCE: return 6 * cd.flag;
    case ILQG_COST_SEMIQUADRATIC:
    case ILQG_CONSTRAINT_SINGLE_DIMENSION: return 2;
    default: return 6;
  }
}

// one thread per role: record the offsets of every update at a dummy input
__global__ void k_record_pattern(const __grid_constant__ DevDesc d, int* offsets /*[NR][E]*/, int* counts, int E) {
  const int role = threadIdx.x;
  if (role > d.N) return;
  float zeros[ILQG_MAX_XDIM + ILQG_MAX_UDIM];
  for (int a = 0; a < ILQG_MAX_XDIM + ILQG_MAX_UDIM; a++) zeros[a] = 0.5f + 0.37f * a;
  OffsetSink sink;
  sink.off = offsets + (size_t)role * E;
  sink.cnt = 0;
  sink.cap = E;
  const float* x = zeros;
  const float* u = zeros + d.n;
  walk_role<1>(
      d, role,
      [&](const DevCost& cd, bool) {
        if (cd.arg < 0) {
          sink.base_H = d.offQ + role * d.n * d.n;
          sink.ld = d.n;
          sink.base_G = d.offl + role * d.n;
          quadraticize_record_sink<true, 1, false>(d, cd, x, d.n, 0.f, 10.f, sink);
        } else {
          const int mj = d.udim[cd.arg];
          sink.base_H = d.offR + d.pair_Roff[cd.pair];
          sink.ld = mj;
          sink.base_G = d.offr + d.pair_roff[cd.pair];
          quadraticize_record_sink<true, 1, false>(d, cd, u + d.uoff[cd.arg], mj, 0.f, 10.f, sink);
        }
      },
      [&](const DevSubsystem& sub) {
        LinOffsetSink lin{&sink, d.offA, d.offB, d.n, d.M};
        subsystem_linearize_sink<1>(d, sub, x, u, lin);
      },
      [](int c) { return c; });
  counts[role] = sink.cnt;
}

// shared memory (floats unless noted): xu[n+M][32] | vals[NR][E][32] | rec[NR][rec] |
//                                      items[num_items] (GatherItem) | idx[num_idx] (u16)
__host__ __device__ inline size_t klq_smem_bytes(int n, int M, int N, int E, int rec, int num_items, int num_idx) {
  const int NR = N + 1;
  size_t b = sizeof(float) * ((size_t)(n + M) * 32 + (size_t)NR * E * kValStride + (size_t)NR * rec);
  b += 5 * sizeof(int) * (size_t)num_items;
  b += sizeof(unsigned short) * (size_t)((num_idx + 7) & ~7);
  return b;
}

__global__ void __launch_bounds__(160)
k_linearize_quadraticize_v3(const __grid_constant__ DevDesc d, Slab s, RecordPattern pat, int only_running, Sel sel) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = d.n, M = d.M, N = d.N, T = d.T, NR = N + 1, E = pat.E;
  float* xu = smem;                                        // [n+M][32]
  float* vals = xu + (n + M) * 32;                         // [NR][E][33]
  float* recs = vals + (size_t)NR * E * kValStride;        // [NR][rec]  (E % 4 == 0 keeps 16 B alignment)
  // gather table as struct-of-arrays (one conflict-free word per lane and field)
  const int NI = pat.num_items;
  int* g_off = reinterpret_cast<int*>(recs + (size_t)NR * d.rec);   // [NI]
  int* g_meta = g_off + NI;                                          // role | count << 8 | start << 16
  float* g_base = reinterpret_cast<float*>(g_meta + NI);             // [NI]
  unsigned* g_e01 = reinterpret_cast<unsigned*>(g_base + NI);        // entries 0,1
  unsigned* g_e23 = g_e01 + NI;                                      // entries 2,3
  unsigned short* gidx = reinterpret_cast<unsigned short*>(g_e23 + NI);

  const long long first = (long long)blockIdx.x * 32;
  const long long total = (long long)s.B * T;
  const long long w = first + lane;
  bool live = false;
  const int b_sel = w < total ? sel_instance(s, sel, (int)(w / T), only_running, &live) : -1;
  const bool in_range = b_sel >= 0;
  const int b = in_range ? b_sel : 0, k = in_range ? (int)(w % T) : 0;
  if (!__syncthreads_or(live)) return;

  // ---- block prologue: x, u of the 32 records; gather table; per-warp template copy ----
  {
    const int cur = in_range ? s.op_cur[b] : 0;
    const float* xs = s.op_xs[cur] + ((size_t)b * T + k) * n;
    const float* us = s.op_us[cur] + ((size_t)b * T + k) * M;
    for (int e = warp; e < n + M; e += NR) xu[e * 32 + lane] = in_range ? (e < n ? xs[e] : us[e - n]) : 0.f;
    for (int e = threadIdx.x; e < NI; e += blockDim.x) {
      const GatherItem it = pat.items[e];
      g_off[e] = it.off;
      g_meta[e] = it.role | (it.count << 8) | (it.start << 16);
      g_base[e] = it.base;
      g_e01[e] = (unsigned)it.e[0] | ((unsigned)it.e[1] << 16);
      g_e23[e] = (unsigned)it.e[2] | ((unsigned)it.e[3] << 16);
    }
    for (int e = threadIdx.x; e < pat.num_idx; e += blockDim.x) gidx[e] = pat.idx[e];
    float* rec = recs + (size_t)warp * d.rec;
    for (int e = lane; e < d.rec / 4; e += 32)
      reinterpret_cast<float4*>(rec)[e] = __ldg(reinterpret_cast<const float4*>(pat.tmpl) + e);
  }
  __syncthreads();

  // ---- phase 1: role = warp, record = lane: update values only ----
  {
    ValueSink sink;
    sink.val = vals + (size_t)warp * E * kValStride;
    sink.lane = lane;
    sink.cnt = 0;
    sink.cap = E;
    const float* x = xu + lane;  // element idx at x[idx * 32]
    const float* u = xu + n * 32 + lane;
    const float mu = live ? s.mu[b] : 0.f;
    const bool full = warp < N && (d.cost_structure[warp] == ILQG_COST_SUM ||
                                   (live && s.te_quad[(size_t)b * N + warp] == k));
    walk_role<32>(
        d, warp,
        [&](const DevCost& cd, bool chosen) {
          const bool is_con = cd.slot >= 0;
          const int dim = cd.arg < 0 ? n : d.udim[cd.arg];
          if (!full && (cd.arg < 0 || is_con)) {  // QuadraticizeControlCosts: record not visited
            const int cnt = record_updates(cd, dim);
            for (int e = 0; e < cnt; e++) sink.push(0.f);
            return;
          }
          const float lambda =
              (is_con && live) ? s.lambdas[((size_t)b * d.num_constraints + cd.slot) * T + s.lambda_index[k]] : 0.f;
          const float* in = cd.arg < 0 ? x : u + d.uoff[cd.arg] * 32;
          // FinalTimeCost: nothing before its threshold; ExtremeValueCost: the extreme member only
          quadraticize_record_sink<true, 32, false>(d, cd, in, dim, lambda, mu, sink, nullptr,
                                                    chosen && k >= cd.first_step);
        },
        [&](const DevSubsystem& sub) {
          LinValueSink lin{&sink};
          subsystem_linearize_sink<32>(d, sub, x, u, lin);
        },
        [&](int c) { return extreme_member<32>(d, c, x, u); });
  }
  __syncthreads();
  // xu is dead: reuse its first 32 words as the per-record "assemble me" flags
  int* flags = reinterpret_cast<int*>(xu);
  if (warp == 0) flags[lane] = live ? b + 1 : 0;  // 0 = skip, else instance id + 1
  __syncthreads();

  // ---- phase 2: warp w assembles records w, w + NR, ... on top of its template copy ----
  float* rec = recs + (size_t)warp * d.rec;
  for (int r = warp; r < 32; r += NR) {
    const long long wr = first + r;
    if (wr >= total) break;
    if (!flags[r]) continue;
    // every touched word is rewritten from its template value for each record, so the staging
    // copy never needs restoring (the touched set is the same for all records)
    for (int g = lane; g < NI; g += 32) {
      const int meta = g_meta[g];
      const int role = meta & 0xff, count = (meta >> 8) & 0xff, start = meta >> 16;
      const float* v = vals + (size_t)role * E * kValStride + r;
      float acc = g_base[g];
      if (count <= 4) {
        const unsigned e01 = g_e01[g], e23 = g_e23[g];
        if (count > 0) acc += v[(e01 & 0xffff) * kValStride];
        if (count > 1) acc += v[(e01 >> 16) * kValStride];
        if (count > 2) acc += v[(e23 & 0xffff) * kValStride];
        if (count > 3) acc += v[(e23 >> 16) * kValStride];
      } else {
        for (int t = 0; t < count; t++) acc += v[gidx[start + t] * kValStride];
      }
      rec[g_off[g]] = acc;
    }
    __syncwarp();
    // (the instance need not be slot wr / T: SEL_LIST maps slots through the queue list)
    float4* dst = reinterpret_cast<float4*>(s.rec + ((size_t)(flags[r] - 1) * T + (size_t)(wr % T)) * d.rec);
    const float4* src = reinterpret_cast<const float4*>(rec);
    for (int e = lane; e < d.rec / 4; e += 32) dst[e] = src[e];
    __syncwarp();
  }
}


// ===========================================================================
// Compact LQ records (round 2).
//
// A dense record is 1188 floats for ThreePlayerIntersection, but only the ~150 words the gather
// table touches ever differ from the static template: writing the dense record and reading it
// back was 10 x the information content.  K_lq v4 therefore stores, per
// (instance, timestep), only
//     [ item values (NI, in gather-table order) | g_k = sum_i Q_i l_i (n) | state_reg (N) | pad -> multiple of 32 ]
// and the backward sweep (ilqg_backward_tc.cuh) rebuilds what it needs in shared memory from the
// template plus static index tables derived from the same pattern at ilqg_create.  g_k is the
// per-timestep vector ExpectedDecrease's adjoint recursion needs (src/ilq_solver.cpp:387-392);
// computing it here keeps Q out of the sweep's shared memory altogether.
// Dense records (downloads of ILQG_LIN_* / ILQG_QUAD_*, the open-loop solver, stand-alone LQ
// solves on uploaded matrices) are expanded on demand by k_expand_records.
// ===========================================================================
struct CompactPattern {
  int NI;                      // items per record
  int NIp;                     // floats per compact record
  const uint2* gk;             // g_k CSR entries: x = q item | l item << 16, y = player
  const int* gk_start;         // [n + 1]
};

constexpr unsigned kItemReg = 0xFFF0u;  // item code 0xFFF0 + i: the constant state_reg[i] (template diagonal of Q_i)

// shared memory: xu[n+M][32] | vals[NR][E][33] | itemv[NIp][33] | gather table (byte offsets x4 | base, meta) | idx (u16)
__host__ __device__ inline size_t klq4_smem_bytes(int n, int M, int N, int E, int NIp, int num_items, int num_idx) {
  const int NR = N + 1;
  size_t b = sizeof(float) * ((size_t)(n + M) * 32 + (size_t)NR * E * kValStride + (size_t)NIp * kValStride);
  b += 6 * sizeof(int) * (size_t)((num_items + 1) & ~1);
  b += sizeof(unsigned short) * (size_t)((num_idx + 7) & ~7);
  return b;
}

// WIDE = false: the lean instance for descriptors made of round 1's kinds only (no gates, no groups)
template <bool WIDE>
__global__ void __launch_bounds__(160)
k_linearize_quadraticize_v4(const __grid_constant__ DevDesc d, Slab s, RecordPattern pat, CompactPattern cp,
                            int only_running, Sel sel, int linesearch) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = d.n, M = d.M, N = d.N, T = d.T, NR = N + 1, E = pat.E;
  float* xu = smem;                                        // [n+M][32]
  float* vals = xu + (n + M) * 32;                         // [NR][E][33]
  float* itembuf = vals + (size_t)NR * E * kValStride;     // [NIp][33]: the block's 32 compact records, record-minor
  const int NI = pat.num_items;
  // gather table, one int4 + one int2 per item: the byte offsets of its first four values inside `vals`
  // ((role E + entry) 33 floats; the lane's own 4 bytes are added once), then (template value, role | count << 8 |
  // start << 16) -- two vector loads per item instead of four scalar ones and the index arithmetic
  const int NI2 = (NI + 1) & ~1;
  int4* g_off4 = reinterpret_cast<int4*>(itembuf + (size_t)cp.NIp * kValStride);  // [NI2]
  int2* g_bm = reinterpret_cast<int2*>(g_off4 + NI2);                              // [NI2]
  unsigned short* gidx = reinterpret_cast<unsigned short*>(g_bm + NI2);

  const long long first = (long long)blockIdx.x * 32;
  const long long total = (long long)s.B * T;
  const long long w = first + lane;
  bool live = false;
  const int b_sel = w < total ? sel_instance(s, sel, (int)(w / T), only_running, &live) : -1;
  const bool in_range = b_sel >= 0;
  const int b = in_range ? b_sel : 0, k = in_range ? (int)(w % T) : 0;
  if (!__syncthreads_or(live)) return;

  {
    const int cur = in_range ? s.op_cur[b] : 0;
    const float* xs = s.op_xs[cur] + ((size_t)b * T + k) * n;
    const float* us = s.op_us[cur] + ((size_t)b * T + k) * M;
    for (int e = warp; e < n + M; e += NR) xu[e * 32 + lane] = in_range ? (e < n ? xs[e] : us[e - n]) : 0.f;
    for (int e = threadIdx.x; e < NI; e += blockDim.x) {
      const GatherItem it = pat.items[e];
      const int rb = it.role * E;
      g_off4[e] = make_int4((rb + it.e[0]) * kValStride * 4, (rb + it.e[1]) * kValStride * 4,
                            (rb + it.e[2]) * kValStride * 4, (rb + it.e[3]) * kValStride * 4);
      g_bm[e] = make_int2(__float_as_int(it.base), it.role | (it.count << 8) | (it.start << 16));
    }
    for (int e = threadIdx.x; e < pat.num_idx; e += blockDim.x) gidx[e] = pat.idx[e];
  }
  __syncthreads();

  // ---- phase 1: role = warp, record = lane: update values only (as v3) ----
  {
    ValueSink sink;
    sink.val = vals + (size_t)warp * E * kValStride;
    sink.lane = lane;
    sink.cnt = 0;
    sink.cap = E;
    const float* x = xu + lane;
    const float* u = xu + n * 32 + lane;
    const float mu = live ? s.mu[b] : 0.f;
    const bool full = warp < N && (d.cost_structure[warp] == ILQG_COST_SUM ||
                                   (live && s.te_quad[(size_t)b * N + warp] == k));
    walk_role<32, WIDE>(
        d, warp,
        [&](const DevCost& cd, bool chosen) {
          const bool is_con = cd.slot >= 0;
          const int dim = cd.arg < 0 ? n : d.udim[cd.arg];
          if (!full && (cd.arg < 0 || is_con)) {
            const int cnt = record_updates(cd, dim);
            for (int e = 0; e < cnt; e++) sink.push(0.f);
            return;
          }
          const float lambda =
              (is_con && live) ? s.lambdas[((size_t)b * d.num_constraints + cd.slot) * T + s.lambda_index[k]] : 0.f;
          const float* in = cd.arg < 0 ? x : u + d.uoff[cd.arg] * 32;
          // FinalTimeCost: nothing before its threshold; ExtremeValueCost: the extreme member only
          quadraticize_record_sink<true, 32, false, ValueSink, WIDE>(d, cd, in, dim, lambda, mu, sink, nullptr,
                                                                     !WIDE || (chosen && k >= cd.first_step));
        },
        [&](const DevSubsystem& sub) {
          LinValueSink lin{&sink};
          subsystem_linearize_sink<32, LinValueSink, WIDE>(d, sub, x, u, lin);
        },
        [&](int c) { return extreme_member<32>(d, c, x, u); });
  }
  __syncthreads();
  int* flags = reinterpret_cast<int*>(xu);
  // Without a linesearch the reference quadraticizes ONCE, before the first iteration, and only
  // re-linearizes afterwards (ModifyLQStrategies returns before MeritFunction would refresh the
  // quadraticization, src/ilq_solver.cpp:114,322): such records keep their cost items.
  int* keepq = flags + 32;
  if (warp == 0) {
    flags[lane] = live ? b + 1 : 0;
    keepq[lane] = (live && !linesearch && s.iters[b] > 0) ? 1 : 0;
  }
  __syncthreads();

  // ---- phase 2: item values + g_k of the block's 32 records, lane = record ----
  // (Until round 2's last session a warp took one record at a time and its lanes the items: every
  // lane walked its own gather-table entry, and the g_k rows kept 16 of 32 lanes busy -- 74 % of the
  // kernel's instructions.  With the record in the lane, the table walk is uniform across the warp:
  // table words are broadcast reads, the values sit lane-minor in `vals`, nothing diverges.  Same
  // additions in the same order.)
  float* itv = itembuf;  // [g][33]
  const bool keep_quad = keepq[lane] != 0;
  const bool any_keep = __any_sync(0xffffffffu, keep_quad);
  // where this lane's record goes (the instance need not be slot w / T: SEL_LIST maps slots through the queue list)
  float* own = s.crec + ((size_t)b * T + (size_t)k) * cp.NIp;
  const char* vl = reinterpret_cast<const char*>(vals + lane);
  for (int g = warp; g < NI; g += NR) {
    const int4 o = g_off4[g];
    const int2 bm = g_bm[g];
    const int role = bm.y & 0xff, count = (bm.y >> 8) & 0xff;
    float acc = __int_as_float(bm.x);
    if (count <= 4) {
      if (count > 0) acc += *reinterpret_cast<const float*>(vl + o.x);
      if (count > 1) acc += *reinterpret_cast<const float*>(vl + o.y);
      if (count > 2) acc += *reinterpret_cast<const float*>(vl + o.z);
      if (count > 3) acc += *reinterpret_cast<const float*>(vl + o.w);
    } else {
      const float* v = vals + (size_t)role * E * kValStride + lane;
      const int start = bm.y >> 16;
      for (int t = 0; t < count; t++) acc += v[gidx[start + t] * kValStride];
    }
    if (any_keep && role < N && keep_quad) acc = own[g];  // the quadraticization of the solve's first iteration stays
    itv[g * kValStride + lane] = acc;
  }
  __syncthreads();
  // g_k[a] = sum_i (sum_c Q_i[a][c] l_i[c]): per-player fma chains over the non-zero terms, in
  // ascending c (what the dense sweep computed, zeros skipped); warp w takes rows w, w + NR, ...
  for (int a = warp; a < n; a += NR) {
    float gk = 0.f, gi = 0.f;
    int cur = -1;
    const int e1 = __ldg(cp.gk_start + a + 1);
    for (int e = __ldg(cp.gk_start + a); e < e1; e++) {
      const uint2 u = __ldg(cp.gk + e);
      if ((int)u.y != cur) { gk += gi; gi = 0.f; cur = (int)u.y; }
      const unsigned qi = u.x & 0xffffu;
      const float qv = qi >= kItemReg ? d.state_reg[qi - kItemReg] : itv[qi * kValStride + lane];
      gi = fmaf(qv, itv[(u.x >> 16) * kValStride + lane], gi);
    }
    gk += gi;
    itv[(NI + a) * kValStride + lane] = gk;
  }
  __syncthreads();
  // ---- phase 3: warp w writes records w, w + NR, ... (coalesced; the tile is read along its columns,
  //      stride 33: conflict-free) ----
  for (int r = warp; r < 32; r += NR) {
    // record r's address is lane r's `own` (no second division by T)
    const unsigned long long p64 = reinterpret_cast<unsigned long long>(own);
    float* dst = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, p64, r));
    if (first + r >= total) break;
    if (!flags[r]) continue;
    for (int g = lane; g < NI + n; g += 32) dst[g] = itv[g * kValStride + r];
    if (lane < N) dst[NI + n + lane] = d.state_reg[lane];  // constants the sweep adds to the diagonal of Q_i
  }
}

// dense records from compact ones: template + items (one warp per record)
__global__ void __launch_bounds__(128)
k_expand_records(const __grid_constant__ DevDesc d, Slab s, RecordPattern pat, CompactPattern cp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 4 + warp;
  if (w >= (long long)s.B * d.T) return;
  float* dst = s.rec + (size_t)w * d.rec;
  const float* src = s.crec + (size_t)w * cp.NIp;
  for (int e = lane; e < d.rec / 4; e += 32)
    reinterpret_cast<float4*>(dst)[e] = __ldg(reinterpret_cast<const float4*>(pat.tmpl) + e);
  __syncwarp();
  for (int g = lane; g < cp.NI; g += 32) dst[pat.items[g].off] = src[g];
}

}  // namespace ilqg
