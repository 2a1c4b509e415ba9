// mma_chain_bench.cu -- microbenchmark for DESIGN.md section 10, item 1 (not part of the product).
//
// K_bwd's cost is a dependent chain per game: Z <- F' Z F (+ small terms), 99 times, 16 x 16
// operands.  This measures that chain alone, one warp per game, operands in shared memory, for
//   ffma     each lane owns a 2 x 4 tile of the result (what the CUDA-core kernels do)
//   mma3     mma.sync.m16n8k8 TF32 with the three-product split (hi.hi + hi.lo + lo.hi), which
//            profiles/r01_tf32_study.md shows is the cheapest variant inside the parity tolerance
//   mma3i    mma3 with one accumulator per split product (dependent chain of 2 instead of 6 MMAs)
//   mma1     the same with plain TF32 (one product; an upper bound on what the split costs)
// at several warps per SM, and checks warp 0's result against a double-precision chain on the host.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_chain_bench tools/mma_chain_bench.cu
//   /tmp/mma_chain_bench [steps=99] [repeats=20]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int N = 16;    // matrix dimension
constexpr int LD = 20;   // padded leading dimension: fragment loads hit 32 distinct banks
constexpr int WARPS = 4; // warps (games) per block

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
      std::exit(1);                                                                        \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ unsigned tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C (16 x 16, row-major, ld LD) = op(A) * B with A, B row-major in shared memory; TA: use A'.
// INDEP: the three split products go to three accumulators that are added at the end, which cuts the
// dependent MMA chain per output tile from 6 to 2 instructions.
template <bool TA, int SPLIT, bool INDEP = false>
__device__ __forceinline__ void mm_mma(const float* A, const float* B, float* C, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float acc[2][4] = {};
  float acc_a[2][4] = {}, acc_b[2][4] = {};
#pragma unroll
  for (int ks = 0; ks < 2; ks++) {
    float af[4];
    const int k0 = 8 * ks + t;
    af[0] = TA ? A[k0 * LD + g] : A[g * LD + k0];
    af[1] = TA ? A[k0 * LD + g + 8] : A[(g + 8) * LD + k0];
    af[2] = TA ? A[(k0 + 4) * LD + g] : A[g * LD + k0 + 4];
    af[3] = TA ? A[(k0 + 4) * LD + g + 8] : A[(g + 8) * LD + k0 + 4];
    unsigned ah[4], al[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      ah[q] = tf32(af[q]);
      if (SPLIT == 3) al[q] = tf32(af[q] - __uint_as_float(ah[q]));
    }
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      float bf[2] = {B[k0 * LD + g + 8 * nt], B[(k0 + 4) * LD + g + 8 * nt]};
      unsigned bh[2] = {tf32(bf[0]), tf32(bf[1])};
      if (SPLIT == 3) {
        unsigned bl[2] = {tf32(bf[0] - __uint_as_float(bh[0])), tf32(bf[1] - __uint_as_float(bh[1]))};
        mma_tf32(INDEP ? acc_a[nt] : acc[nt], al, bh);  // small terms first
        mma_tf32(INDEP ? acc_b[nt] : acc[nt], ah, bl);
      }
      mma_tf32(acc[nt], ah, bh);
    }
  }
  if (INDEP) {
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[nt][q] += acc_a[nt][q] + acc_b[nt][q];
  }
#pragma unroll
  for (int nt = 0; nt < 2; nt++) {
    *reinterpret_cast<float2*>(&C[g * LD + 2 * t + 8 * nt]) = make_float2(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<float2*>(&C[(g + 8) * LD + 2 * t + 8 * nt]) = make_float2(acc[nt][2], acc[nt][3]);
  }
}

template <bool TA>
__device__ __forceinline__ void mm_ffma(const float* A, const float* B, float* C, int lane) {
  const int r0 = (lane >> 2) * 2, c0 = (lane & 3) * 4;
  float acc[2][4] = {};
#pragma unroll 4
  for (int k = 0; k < N; k++) {
    const float a0 = TA ? A[k * LD + r0] : A[r0 * LD + k];
    const float a1 = TA ? A[k * LD + r0 + 1] : A[(r0 + 1) * LD + k];
    const float4 b = *reinterpret_cast<const float4*>(&B[k * LD + c0]);
    acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
    acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
    acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
    acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
  }
  *reinterpret_cast<float4*>(&C[r0 * LD + c0]) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
  *reinterpret_cast<float4*>(&C[(r0 + 1) * LD + c0]) = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
}

// MODE 0: ffma, 1: mma TF32 x 1, 3: mma TF32 x 3, 4: mma TF32 x 3 with independent accumulators
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32) k_chain(const float* __restrict__ F0, const float* __restrict__ Z0,
                                                     float* __restrict__ out, int games, int steps) {
  __shared__ __align__(16) float sm[WARPS][3][N * LD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int game = blockIdx.x * WARPS + warp;
  if (game >= games) return;
  float *F = sm[warp][0], *Z = sm[warp][1], *W = sm[warp][2];
  for (int e = lane; e < N * N; e += 32) {
    F[(e / N) * LD + e % N] = F0[e];
    Z[(e / N) * LD + e % N] = Z0[e] * (1.0f + 1e-3f * (game % 7));  // per-game data, same flow
  }
  __syncwarp();
  for (int s = 0; s < steps; s++) {
    constexpr int SPLIT = MODE == 1 ? 1 : 3;
    if (MODE == 0) mm_ffma<false>(Z, F, W, lane); else mm_mma<false, SPLIT, MODE == 4>(Z, F, W, lane);
    __syncwarp();
    if (MODE == 0) mm_ffma<true>(F, W, Z, lane); else mm_mma<true, SPLIT, MODE == 4>(F, W, Z, lane);
    __syncwarp();
  }
  for (int e = lane; e < N * N; e += 32) out[(size_t)game * N * N + e] = Z[(e / N) * LD + e % N];
}

template <int MODE>
float run(const float* dF, const float* dZ, float* dout, int games, int steps, int repeats) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int blocks = (games + WARPS - 1) / WARPS;
  k_chain<MODE><<<blocks, WARPS * 32>>>(dF, dZ, dout, games, steps);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int r = 0; r < repeats; r++) k_chain<MODE><<<blocks, WARPS * 32>>>(dF, dZ, dout, games, steps);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / repeats;
}

int main(int argc, char** argv) {
  const int steps = argc > 1 ? std::atoi(argv[1]) : 99;
  const int repeats = argc > 2 ? std::atoi(argv[2]) : 20;
  // F = 0.995 * (product of Givens rotations): the chain neither grows nor dies; Z0 symmetric
  std::vector<double> F(N * N, 0.0), Z(N * N);
  for (int i = 0; i < N; i++) F[i * N + i] = 1.0;
  for (int p = 0; p < N - 1; p++) {
    const double th = 0.3 + 0.1 * p, c = std::cos(th), s = std::sin(th);
    for (int k = 0; k < N; k++) {
      const double x = F[k * N + p], y = F[k * N + p + 1];
      F[k * N + p] = c * x - s * y;
      F[k * N + p + 1] = s * x + c * y;
    }
  }
  for (auto& v : F) v *= 0.995;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) Z[i * N + j] = (i == j ? 2.0 : 0.0) + 0.5 * std::cos(0.37 * (i + 1) * (j + 1)) + 0.5 * std::cos(0.37 * (j + 1) * (i + 1));
  std::vector<float> Ff(F.begin(), F.end()), Zf(Z.begin(), Z.end());
  // host chain in double from the fp32-rounded inputs (game 0)
  std::vector<double> Fd(Ff.begin(), Ff.end()), Zd(Zf.begin(), Zf.end()), Wd(N * N);
  for (int s = 0; s < steps; s++) {
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { double a = 0; for (int k = 0; k < N; k++) a += Zd[i * N + k] * Fd[k * N + j]; Wd[i * N + j] = a; }
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { double a = 0; for (int k = 0; k < N; k++) a += Fd[k * N + i] * Wd[k * N + j]; Zd[i * N + j] = a; }
  }
  double scale = 0;
  for (double v : Zd) scale = std::fmax(scale, std::fabs(v));

  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float *dF, *dZ, *dout;
  const int max_games = sms * 64;
  CK(cudaMalloc(&dF, sizeof(float) * N * N));
  CK(cudaMalloc(&dZ, sizeof(float) * N * N));
  CK(cudaMalloc(&dout, sizeof(float) * N * N * max_games));
  CK(cudaMemcpy(dF, Ff.data(), sizeof(float) * N * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dZ, Zf.data(), sizeof(float) * N * N, cudaMemcpyHostToDevice));

  std::printf("%d SMs, chain of %d steps x 2 products (16x16x16), one warp per game\n", sms, steps);
  std::printf("%-6s %10s %12s %16s %14s %12s\n", "mode", "warps/SM", "ms/launch", "ns/step/warp", "Gproducts/s", "rel err g0");
  std::vector<float> got(N * N);
  for (int wps : {4, 8, 16, 32, 64}) {
    const int games = sms * wps;
    for (int mode : {0, 3, 4, 1}) {
      const float ms = mode == 0 ? run<0>(dF, dZ, dout, games, steps, repeats)
                       : mode == 3 ? run<3>(dF, dZ, dout, games, steps, repeats)
                       : mode == 4 ? run<4>(dF, dZ, dout, games, steps, repeats)
                                   : run<1>(dF, dZ, dout, games, steps, repeats);
      CK(cudaMemcpy(got.data(), dout, sizeof(float) * N * N, cudaMemcpyDeviceToHost));
      double err = 0;
      for (int e = 0; e < N * N; e++) err = std::fmax(err, std::fabs(got[e] - Zd[e]));
      std::printf("%-6s %10d %12.4f %16.1f %14.2f %12.2e\n", mode == 0 ? "ffma" : mode == 3 ? "mma3" : mode == 4 ? "mma3i" : "mma1", wps, ms,
                  1e6 * ms / steps, 2.0 * steps * games / (ms * 1e6), err / scale);
    }
  }
  return 0;
}
