// Micro-benchmark: MUFU.EX2 throughput per SM (ops/clk) with W warps per SM, optionally mixed with FFMA+FADD+F2FP.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MIX>
__global__ void k(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  float s = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float x = a[i];
      if (MIX) x = fmaf(x, 1.4426950f, -0.25f);
      x = ex2(x);
      if (MIX) s += x;
      a[i] = x * 0.5f - 0.3f;
    }
  }
  long long t1 = clock64();
  float r = s;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  for (int mix = 0; mix < 2; ++mix)
    for (int warps : {4, 8, 16}) {
      int iters = 2000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mix) k<1><<<148, warps * 32>>>(d, iters, 0.1f); else k<0><<<148, warps * 32>>>(d, iters, 0.1f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops = (double)warps * 32 * 16 * iters;
      printf("mix=%d warps/SM=%2d: %.2f MUFU/clk/SM (cycles %.0f, %.3f ms)\n", mix, warps, ops / cyc, cyc, ms);
    }
  return 0;
}
