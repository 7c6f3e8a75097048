// Microbenchmark for the accumulator parking scheme of pls_gen_gemm.cuh (NS = 2): 8 warps per SM each run blocks of 256
// DMMA.8x8x4 on 64 fp64 accumulators; between blocks the accumulators are swapped with a second set parked in tensor memory
// (tcgen05.st x128 + tcgen05.ld x128).  Prints the FP64 pipe rate with and without the swap, and the cost of a swap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_swap_microbench tools/tmem_swap_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../projected_langevin_sampling_b200/csrc/pls_tmem.cuh"

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>  // 0: no swap, 1: st + wait::st + ld, 2: st + ld (+ wait::st deferred to before the next st)
__global__ void __launch_bounds__(256, 1) bench(double* out, int blocks, double a, double b) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) pls::tmem_alloc_512(&base_s);
  pls::tmem_fence_before_sync();
  __syncthreads();
  pls::tmem_fence_after_sync();
  const uint32_t base = base_s + ((32u * (warp & 3)) << 16) + (warp >> 2) * 256;
  double acc[32][2];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i >> 1][i & 1] = 0.0;
  if (MODE != 0) {
    pls::tmem_store64(base, acc);
    pls::tmem_store64(base + 128, acc);
    pls::tmem_wait_st();
  }
#pragma unroll 1
  for (int blk = 0; blk < blocks; ++blk) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int n = 0; n < 32; ++n) dmma(acc[n][0], acc[n][1], a, b);
    if (MODE == 1) {
      pls::tmem_store64(base + (blk & 1) * 128, acc);
      pls::tmem_wait_st();
      pls::tmem_load64(base + ((blk & 1) ^ 1) * 128, acc);
    } else if (MODE == 2) {
      pls::tmem_wait_st();
      pls::tmem_store64(base + (blk & 1) * 128, acc);
      pls::tmem_load64(base + ((blk & 1) ^ 1) * 128, acc);
    }
  }
  if (MODE == 2) pls::tmem_wait_st();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += acc[i >> 1][i & 1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
  pls::tmem_fence_before_sync();
  __syncthreads();
  if (warp == 0) pls::tmem_dealloc_512(base_s);
}

template <int MODE>
float run(double* out, int sms, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  bench<MODE><<<sms, 256>>>(out, blocks, 1e-3, 1e-3);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<sms, 256>>>(out, blocks, 1e-3, 1e-3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, blocks = 20000;
  double* out;
  cudaMalloc(&out, sms * 256 * sizeof(double));
  const double flops = 2.0 * 8 * 8 * 4 * 256.0 * blocks * 8 * sms;  // per DMMA 2*8*8*4, 256 per block, 8 warps
  const float t0 = run<0>(out, sms, blocks), t1 = run<1>(out, sms, blocks), t2 = run<2>(out, sms, blocks);
  const double clk = prop.clockRate * 1e3;  // Hz (nominal)
  printf("sms=%d blocks=%d\n", sms, blocks);
  printf("no swap                    : %8.3f ms  %6.2f TFLOP/s\n", t0, flops / t0 * 1e-9);
  printf("swap st,wait::st,ld        : %8.3f ms  %6.2f TFLOP/s  (+%.0f clk per swap per SM-block of 8 warps)\n", t1, flops / t1 * 1e-9,
         (t1 - t0) * 1e-3 * clk / blocks);
  printf("swap st,ld (wait::st late) : %8.3f ms  %6.2f TFLOP/s  (+%.0f clk per swap per SM-block of 8 warps)\n", t2, flops / t2 * 1e-9,
         (t2 - t0) * 1e-3 * clk / blocks);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
