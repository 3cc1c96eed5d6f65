// Micro-benchmark: how far behind a MUFU.EX2 its consumer must sit for ONE warp per scheduler to keep the exp pipe
// busy.  Body = K independent ex2 followed by K dependent FADD/F2FP consumers (distance ~K MUFUs); cycles per ex2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int K, int VARY>
__global__ void k(float* out, int iters, float seed) {
  float a[K];
#pragma unroll
  for (int i = 0; i < K; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  float s0 = 0.f, s1 = 0.f;
  uint32_t acc = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    float e[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { e[i] = ex2(fmaf(a[i], 1.4426950f, -0.25f)); if (VARY) a[i] += 1e-4f; }
#pragma unroll
    for (int i = 0; i < K; i += 2) {
      s0 += e[i]; s1 += e[i + 1];
      __nv_bfloat162 v = __floats2bfloat162_rn(e[i], e[i + 1]);
      acc ^= *reinterpret_cast<uint32_t*>(&v);
    }
  }
  long long t1 = clock64();
  float r = s0 + s1 + __uint_as_float(acc);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
// Consumers work on the results of the PREVIOUS loop iteration (carried in registers): distance >= K MUFUs, whatever
// the scheduler does inside one iteration.
template <int K>
__global__ void kc(float* out, int iters, float seed) {
  float a[K], e[K];
#pragma unroll
  for (int i = 0; i < K; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; e[i] = 0.f; }
  float s0 = 0.f, s1 = 0.f;
  uint32_t acc = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    float n[K];
#pragma unroll
    for (int i = 0; i < K; i += 2) {
      n[i] = ex2(fmaf(a[i], 1.4426950f, -0.25f));
      n[i + 1] = ex2(fmaf(a[i + 1], 1.4426950f, -0.25f));
      s0 += e[i]; s1 += e[i + 1];
      __nv_bfloat162 v = __floats2bfloat162_rn(e[i], e[i + 1]);
      acc ^= *reinterpret_cast<uint32_t*>(&v);
    }
#pragma unroll
    for (int i = 0; i < K; ++i) e[i] = n[i];
  }
  long long t1 = clock64();
  float r = s0 + s1 + __uint_as_float(acc);
#pragma unroll
  for (int i = 0; i < K; ++i) r += e[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int K> void runc(float* d, int warps) {
  int iters = 4096 / K * 8;
  for (int rep = 0; rep < 2; ++rep) { kc<K><<<148, warps * 32>>>(d, iters, 0.1f); cudaDeviceSynchronize(); }
  float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  printf("carried K=%2d warps/SM=%2d: %.2f cycles per ex2 per scheduler-warp-slot (%.2f MUFU/clk/SM)\n", K, warps,
         cyc / ((double)iters * K) / (warps / 4.0), (double)warps * 32 * K * iters / cyc);
}
template <int K, int VARY = 0> void run(float* d, int warps) {
  int iters = 4096 / K * 64;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k<K, VARY><<<148, warps * 32>>>(d, iters, 0.1f); cudaEventRecord(e1); cudaDeviceSynchronize(); }
  float cyc, ms; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost); cudaEventElapsedTime(&ms, e0, e1);
  printf("K=%2d vary=%d warps/SM=%2d: %.2f cycles per warp-ex2 per scheduler (%.2f lanes/clk/SM by clock64; %.2f lanes/ns/SM by events, %.3f ms)\n", K, VARY, warps,
         cyc / ((double)iters * K) / (warps / 4.0), (double)warps * 32 * K * iters / cyc, (double)warps * 32 * K * iters / (ms * 1e6), ms);
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  for (int warps : {4, 8}) { run<2>(d, warps); run<4>(d, warps); run<8>(d, warps); run<16>(d, warps); run<32>(d, warps); run<64>(d, warps); run<32, 1>(d, warps); run<16, 1>(d, warps);
    runc<4>(d, warps); runc<8>(d, warps); runc<16>(d, warps); runc<32>(d, warps); }
  return 0;
}
