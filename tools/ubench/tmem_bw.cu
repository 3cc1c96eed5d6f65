// Micro-benchmark: TMEM read/write bandwidth seen by tcgen05.ld / tcgen05.st (32x32b.x32) with W warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../taiwan-tongues-asr-ce_b200/csrc/ptx_sm100.cuh"
using namespace ttasr;
template <int MODE>  // 0: ld only, 1: st only, 2: ld + dependent max over the values
__global__ void k(float* out, int iters) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&tmem_ptr), 512); tmem_relinquish<1>(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  tmem_st_32x32(base, v); tmem_st_32x32(base + 32, v); tmem_st_32x32(base + 64, v); tmem_st_32x32(base + 96, v);
  tmem_wait_st();
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld_32x32(base, a); tmem_ld_32x32(base + 32, b); tmem_ld_32x32(base + 64, c); tmem_ld_32x32(base + 96, d);
      tmem_wait_ld();
      if (MODE == 2) {
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(fmaxf(m, __uint_as_float(a[i])), fmaxf(__uint_as_float(b[i]), fmaxf(__uint_as_float(c[i]), __uint_as_float(d[i]))));
        acc += m;
      } else {
        acc += __uint_as_float(a[0] ^ b[1] ^ c[2] ^ d[3]);
      }
    } else {
      v[0] = it;
      tmem_st_32x32(base, v); tmem_st_32x32(base + 32, v); tmem_st_32x32(base + 64, v); tmem_st_32x32(base + 96, v);
      tmem_wait_st();
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
  if (warp == 0) tmem_dealloc<1>(tmem_ptr, 512);
}
template <int MODE> void run(const char* name, float* d) {
  for (int warps : {4, 8, 16}) {
    int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) { k<MODE><<<148, warps * 32>>>(d, iters); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return; } }
    float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    double bytes = (double)warps * 32 * 128 * 4 * iters;
    printf("%-10s warps/SM=%2d: %.1f B/clk/SM  (%.0f cycles per 4 x (32x32b.x32) per warp)\n", name, warps, bytes / cyc, cyc / iters);
  }
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  run<0>("ld", d); run<2>("ld+max", d); run<1>("st", d);
  return 0;
}
