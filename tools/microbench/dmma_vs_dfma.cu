// Is the FP64 tensor path (mma.sync m8n8k4 f64, "DMMA") a separate pipe from the FP64 FMA pipe on B200, and what
// are the two rates?  Three kernels with the same grid: DFMA only, DMMA only, and both (even warps DFMA, odd DMMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_vs_dfma dmma_vs_dfma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0 = DFMA, 1 = DMMA, 2 = even warps DFMA / odd warps DMMA
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, double y) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = MODE == 1 || (MODE == 2 && (warp & 1));
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  if (!do_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], x, y);   // 16 independent chains
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(acc[2 * i], acc[2 * i + 1], x, y);  // 8 independent accumulators
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(double* out, int grid, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 4, iters = 20000;
  double* out; cudaMalloc(&out, (size_t)grid * 256 * 8);
  const double warps = (double)grid * 8;
  float t0 = run<0>(out, grid, iters), t1 = run<1>(out, grid, iters), t2 = run<2>(out, grid, iters);
  const double fma_dfma = warps * iters * 16.0 * 32.0, fma_dmma = warps * iters * 8.0 * 256.0;
  printf("SMs %d\n", sms);
  printf("DFMA only : %8.3f ms  %7.2f TFLOP/s\n", t0, 2 * fma_dfma / t0 / 1e9);
  printf("DMMA only : %8.3f ms  %7.2f TFLOP/s\n", t1, 2 * fma_dmma / t1 / 1e9);
  printf("half/half : %8.3f ms  (separate pipes -> ~max(%0.3f, %0.3f) = %0.3f; shared -> ~%0.3f)\n", t2, t0 / 2, t1 / 2,
         (t0 > t1 ? t0 : t1) / 2, (t0 + t1) / 2);
  return 0;
}
