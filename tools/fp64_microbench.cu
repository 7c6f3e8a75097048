// FP64 pipe microbenchmark for B200 (sm_100a).
//
// Answers the design questions DESIGN.md §"FP64 pipe model" depends on:
//   1. what DMMA.8x8x4 (mma.sync.m8n8k4.f64 — the only FP64 tensor shape sm_100a has in SASS)
//      sustains per SM per clock,
//   2. what plain DFMA sustains,
//   3. whether DMMA and DFMA share an execution pipe (mixed streams add up or not),
//   4. what exp() in double costs next to them.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
// Run:   ./fp64_microbench            (prints one line per experiment)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NACC independent DMMA accumulator chains per warp.
template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same warp issues NMMA DMMAs and NFMA DFMAs per iteration (independent chains).
template <int NMMA, int NFMA>
__global__ void k_mixed(double* out, int iters, double a0, double b0) {
  double c[NMMA > 0 ? NMMA : 1][2];
  double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < NMMA; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
  for (int i = 0; i < NFMA; ++i) f[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NMMA > NFMA ? NMMA : NFMA); ++i) {
      if (i < NMMA) dmma884(c[i][0], c[i][1], a, b);
      if (i < NFMA) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NMMA; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NFMA; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Warp-specialised mix: even warps run DMMA, odd warps run DFMA.
__global__ void k_split(double* out, int iters, double a0, double b0, unsigned long long* clk) {
  const int warp = threadIdx.x >> 5;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  double s = 0;
  long long t0 = clock64();
  if ((warp & 1) == 0) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  } else {
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) clk[warp] = (unsigned long long)(t1 - t0);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters, double a0) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = -a0 - i * 0.01 - threadIdx.x * 1e-3;
  double s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { s += exp(x[i]); x[i] += 1e-6; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch();  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s sms %d max_clock_mhz %.0f\n", p.name, sms, clk_khz / 1000.0);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 4 * 1024));
  unsigned long long* clk; CK(cudaMalloc(&clk, sizeof(unsigned long long) * 64));
  const int iters = 20000;

  for (int warps = 4; warps <= 16; warps *= 2) {
    const int threads = warps * 32;
    float ms = time_ms([&] { k_dmma<8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fma = (double)sms * warps * iters * 8 * 256.0;
    printf("dmma884  warps/SM %2d acc 8 : %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM at max clock)\n", warps, ms,
           2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    float ms = time_ms([&] { k_dmma<2><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    double fma = (double)sms * 8 * iters * 2 * 256.0;
    printf("dmma884  warps/SM  8 acc 2 : %8.3f ms  %7.2f TFLOP/s\n", ms, 2 * fma / ms * 1e-9);
    ms = time_ms([&] { k_dmma<16><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    fma = (double)sms * 8 * iters * 16 * 256.0;
    printf("dmma884  warps/SM  8 acc16 : %8.3f ms  %7.2f TFLOP/s\n", ms, 2 * fma / ms * 1e-9);
  }
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int threads = warps * 32;
    float ms = time_ms([&] { k_dfma<8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fma = (double)sms * warps * iters * 8 * 32.0;
    printf("dfma     warps/SM %2d acc 8 : %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM at max clock)\n", warps, ms,
           2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (clk_khz * 1e3));
  }
  {
    // mixed in one warp: 8 DMMA (2048 FMA-lanes... 8*256) + N DFMA (N*32)
    float ms0 = time_ms([&] { k_mixed<8, 0><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    float ms8 = time_ms([&] { k_mixed<8, 8><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    float ms16 = time_ms([&] { k_mixed<8, 16><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    float ms32 = time_ms([&] { k_mixed<8, 32><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    float msf = time_ms([&] { k_mixed<0, 32><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    printf("mixed same-warp (8 warps/SM): 8dmma %.3f ms | +8dfma %.3f | +16dfma %.3f | +32dfma %.3f | 32dfma alone %.3f\n",
           ms0, ms8, ms16, ms32, msf);
  }
  {
    float ms = time_ms([&] { k_split<<<sms, 512>>>(out, iters, 1.0000001, 1e-9, clk); });
    unsigned long long h[16]; CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    printf("split warps (16 warps/SM, even=dmma x8, odd=dfma x8): %.3f ms; clocks dmma-warp %llu dfma-warp %llu\n", ms, h[0], h[1]);
    float msa = time_ms([&] { k_dmma<8><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    float msb = time_ms([&] { k_dfma<8><<<sms, 256>>>(out, iters, 1.0000001, 1e-9); });
    printf("  alone: 8 dmma warps %.3f ms, 8 dfma warps %.3f ms (sum %.3f)\n", msa, msb, msa + msb);
  }
  {
    float ms = time_ms([&] { k_exp<<<sms, 256>>>(out, 5000, 1.0); });
    double n = (double)sms * 256 * 5000 * 4;
    printf("exp(double): %.3f ms  %.2f Gexp/s  (%.2f clk/SM per warp-exp at max clock)\n", ms, n / ms * 1e-6,
           (ms * 1e-3) * (clk_khz * 1e3) / (n / 32 / sms));
  }
  return 0;
}
