// Micro-benchmark: issue throughput (lanes/clk/SM) of the instructions in the softmax sweep, alone and mixed, to find
// which of them share the 16-lane/clk XU pipe with MUFU.EX2.  16 loop-carried independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
enum { OP_MUFU, OP_PACK, OP_FFMA, OP_FADD, OP_MAX3, OP_MUFU_PACK, OP_MUFU_FFMA_FADD, OP_SWEEP, OP_FMUL2 };
template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  float s = 0.f; uint32_t acc = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == OP_MUFU) a[i] = ex2(a[i]);
      if (OP == OP_PACK) a[i] = __uint_as_float(pack(a[i], a[(i + 1) & 15]));
      if (OP == OP_FFMA) a[i] = fmaf(a[i], 1.0001f, -0.25f);
      if (OP == OP_FADD) a[i] = a[i] + 0.25f;
      if (OP == OP_MAX3) a[i] = max3(a[i], a[(i + 1) & 15], a[(i + 5) & 15]);
      if (OP == OP_MUFU_PACK) { a[i] = ex2(a[i]); if (i & 1) acc ^= pack(a[i], a[i - 1]); }
      if (OP == OP_MUFU_FFMA_FADD) { a[i] = ex2(fmaf(a[i], 1.0001f, -0.25f)); s += a[i]; }
      if (OP == OP_SWEEP) { a[i] = ex2(fmaf(a[i], 1.0001f, -0.25f)); s += a[i]; if (i & 1) acc ^= pack(a[i], a[i - 1]); }
      if (OP == OP_FMUL2) { asm volatile("{ .reg .b64 t, u, w; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %2}; mov.b64 w, {%3, %3}; fma.rn.f32x2 t, t, u, w; mov.b64 {%0, %1}, t; }" : "+f"(a[i]), "+f"(a[(i + 1) & 15]) : "f"(1.0001f), "f"(-0.25f)); ++i; }
    }
  }
  long long t1 = clock64();
  float r = s + __uint_as_float(acc);
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int OP> void run(const char* name, float* d) {
  for (int warps : {4, 8, 16}) {
    int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) { k<OP><<<148, warps * 32>>>(d, iters, 0.1f); cudaDeviceSynchronize(); }
    float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    printf("%-16s warps/SM=%2d: %7.2f elements/clk/SM  (%.2f cycles per 16-element group per warp)\n", name, warps,
           (double)warps * 32 * 16 * iters / cyc, cyc / iters);
  }
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  run<OP_MUFU>("mufu", d); run<OP_PACK>("f2fp.pack", d); run<OP_FFMA>("ffma", d); run<OP_FADD>("fadd", d);
  run<OP_MAX3>("fmnmx3", d); run<OP_MUFU_PACK>("mufu+pack/2", d); run<OP_MUFU_FFMA_FADD>("ffma+mufu+fadd", d);
  run<OP_SWEEP>("sweep(all)", d); run<OP_FMUL2>("ffma2", d);
  return 0;
}
