// mma_rate_bench.cu -- issue rate / latency of the legacy tensor path on sm_100a (not part of the product).
//   1. mma.sync.m16n8k8 tf32: cycles per instruction per SM sub-partition, with ACC independent
//      accumulators per warp (ACC = 1: dependent chain = latency) at 1..16 warps per sub-partition
//   2. cvt.rna.tf32.f32 and LOP3+FADD split throughput
//   3. does the tensor core ignore the low 13 mantissa bits of a raw fp32 operand (truncate)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/mma_rate_bench.cu && /tmp/mma_rate
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                         unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int ACC>
__global__ void k_mma(float* out, int iters, long long* cycles) {
  float c[ACC][4];
  for (int a = 0; a < ACC; a++)
    for (int q = 0; q < 4; q++) c[a][q] = threadIdx.x * 1e-3f + a;
  const unsigned a0 = __float_as_uint(1.0f + threadIdx.x * 1e-3f), b0 = __float_as_uint(0.5f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int a = 0; a < ACC; a++) mma_tf32(c[a], a0, a0, a0, a0, b0, b0);
  }
  const long long t1 = clock64();
  float s = 0;
  for (int a = 0; a < ACC; a++)
    for (int q = 0; q < 4; q++) s += c[a][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>  // 0: cvt.rna x2 + sub, 1: lop + sub
__global__ void k_split(float* out, int iters, long long* cycles) {
  float x[8];
  for (int q = 0; q < 8; q++) x[q] = 1.0f + threadIdx.x * 1e-3f + q;
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      unsigned hi, lo;
      if (MODE == 0) {
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x[q]));
        const float l = x[q] - __uint_as_float(hi);
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(l));
      } else {
        hi = __float_as_uint(x[q]) & 0xffffe000u;
        lo = __float_as_uint(x[q] - __uint_as_float(hi));
      }
      acc ^= hi + lo;
      x[q] += 1e-3f;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void k_trunc(const float* a, const float* b, float* out) {
  // one 16x8x8 product with raw fp32 operands (c0) and with operands truncated to tf32 by hand (c1)
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
  const float av[4] = {a[g * 8 + t], a[(g + 8) * 8 + t], a[g * 8 + t + 4], a[(g + 8) * 8 + t + 4]};
  const float bv[2] = {b[t * 8 + g], b[(t + 4) * 8 + g]};
  unsigned ar[4], br[2], at[4], bt[2], an[4], bn[2];
  for (int q = 0; q < 4; q++) {
    ar[q] = __float_as_uint(av[q]);
    at[q] = ar[q] & 0xffffe000u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(an[q]) : "f"(av[q]));
  }
  for (int q = 0; q < 2; q++) {
    br[q] = __float_as_uint(bv[q]);
    bt[q] = br[q] & 0xffffe000u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bn[q]) : "f"(bv[q]));
  }
  mma_tf32(c0, ar[0], ar[1], ar[2], ar[3], br[0], br[1]);
  mma_tf32(c1, at[0], at[1], at[2], at[3], bt[0], bt[1]);
  mma_tf32(c2, an[0], an[1], an[2], an[3], bn[0], bn[1]);
  for (int q = 0; q < 4; q++) {
    out[lane * 12 + q] = c0[q];
    out[lane * 12 + 4 + q] = c1[q];
    out[lane * 12 + 8 + q] = c2[q];
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4 << 20);
  cudaMalloc(&cyc, 8);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000;
  std::printf("mma.sync.m16n8k8.tf32: cycles per MMA per warp (dependent issue) and per sub-partition (throughput)\n");
  std::printf("%6s %10s %14s %18s\n", "ACC", "warps/SMSP", "cyc/MMA/warp", "cyc/MMA/SMSP");
  for (int wps : {1, 2, 4, 8, 16}) {
    const int threads = std::min(1024, wps * 4 * 32), blocks_per_sm = (wps * 4 * 32 + threads - 1) / threads;
    long long c;
#define RUN(ACC)                                                                    \
  k_mma<ACC><<<sms * blocks_per_sm, threads>>>(out, iters, cyc);                     \
  cudaDeviceSynchronize();                                                          \
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);                                   \
  std::printf("%6d %10d %14.2f %18.2f\n", ACC, wps, (double)c / (iters * ACC), (double)c / (iters * ACC * wps));
    RUN(1) RUN(2) RUN(4) RUN(8)
  }
  std::printf("split of 8 values: cycles per value per warp / per SMSP\n");
  for (int wps : {1, 4, 8}) {
    long long c;
    k_split<0><<<sms, wps * 4 * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    std::printf("  cvt.rna x2 + fsub : warps/SMSP %d  %.2f cyc/value/warp  %.2f /SMSP\n", wps, (double)c / (iters * 8), (double)c / (iters * 8 * wps));
    k_split<1><<<sms, wps * 4 * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    std::printf("  lop3 + fsub       : warps/SMSP %d  %.2f cyc/value/warp  %.2f /SMSP\n", wps, (double)c / (iters * 8), (double)c / (iters * 8 * wps));
  }
  // truncation semantics
  std::vector<float> a(128), b(64);
  srand(1);
  for (auto& v : a) v = (float)rand() / RAND_MAX + 1e-4f * rand() / RAND_MAX;
  for (auto& v : b) v = (float)rand() / RAND_MAX - 0.5f;
  float *da, *db;
  cudaMalloc(&da, 512);
  cudaMalloc(&db, 256);
  cudaMemcpy(da, a.data(), 512, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), 256, cudaMemcpyHostToDevice);
  k_trunc<<<1, 32>>>(da, db, out);
  std::vector<float> o(32 * 12);
  cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  int same_trunc = 0, same_rna = 0;
  for (int l = 0; l < 32; l++)
    for (int q = 0; q < 4; q++) {
      same_trunc += o[l * 12 + q] == o[l * 12 + 4 + q];
      same_rna += o[l * 12 + q] == o[l * 12 + 8 + q];
    }
  std::printf("raw fp32 operands: %d / 128 outputs equal hand-truncated operands, %d / 128 equal cvt.rna operands\n", same_trunc, same_rna);
  // accumulate rounding: c + a*b with c large: does the tensor core round to nearest or truncate?
  return 0;
}
