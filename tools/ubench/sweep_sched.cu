// Micro-benchmark of the softmax exp sweep of ONE warp per scheduler (what the token-exclusive sweep looks like):
// 128 scores per thread from shared memory -> p = 2^(s*log2e - m) -> row sum + bf16 pack -> shared memory.
//   V0: straight-line code, ptxas schedules (consumers end up right behind their MUFU.EX2)
//   V1: staged pipeline, stages separated by opaque branches: block c holds the exps of chunk c+1 and the sum/pack of
//       chunk c, so consumers can never be scheduled next to their producers
//   V2: V1 with packed fma.rn.f32x2 / add.f32x2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ void expc(float (&v)[32], float m) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = ex2(fmaf(v[i], kLog2e, -m));
}
__device__ __forceinline__ void expc2(float (&v)[32], float m) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    asm("{ .reg .b64 t, u, w; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %2}; mov.b64 w, {%3, %3}; fma.rn.f32x2 t, t, u, w; mov.b64 {%0, %1}, t; }"
        : "+f"(v[i]), "+f"(v[i + 1]) : "f"(kLog2e), "f"(-m));
    v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]);
  }
}
__device__ __forceinline__ float sumpack(const float (&v)[32], uint32_t (&pk)[16]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s0 += v[2 * i]; s1 += v[2 * i + 1]; pk[i] = pack(v[2 * i], v[2 * i + 1]); }
  return s0 + s1;
}
__device__ __forceinline__ float sumpack2(const float (&v)[32], uint32_t (&pk)[16]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    asm("{ .reg .b64 t, u; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %3}; add.rn.f32x2 t, t, u; mov.b64 {%0, %1}, t; }" : "+f"(s0), "+f"(s1) : "f"(v[2 * i]), "f"(v[2 * i + 1]));
    pk[i] = pack(v[2 * i], v[2 * i + 1]);
  }
  return s0 + s1;
}
__device__ __forceinline__ void ldc(float (&v)[32], const float* src) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) { float4 t = *reinterpret_cast<const float4*>(src + i); v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w; }
}
__device__ __forceinline__ void stc(uint32_t* dst, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(dst + i) = make_uint4(pk[i], pk[i + 1], pk[i + 2], pk[i + 3]);
}
template <int V>
__global__ void k(float* out, int iters, int opq0, int opq1, int opq2, int opq3) {
  extern __shared__ float sm[];
  float* mine = sm + threadIdx.x * 132;   // 128 scores + pad
  for (int i = 0; i < 128; ++i) mine[i] = -0.01f * ((i * 37 + threadIdx.x) & 63);
  uint32_t* pdst = reinterpret_cast<uint32_t*>(mine);
  float l = 0.f, m = 0.25f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    float v0[32], v1[32], v2[32], v3[32];
    uint32_t pk[16];
    ldc(v0, mine); ldc(v1, mine + 32); ldc(v2, mine + 64); ldc(v3, mine + 96);
    m += 1e-3f;
    float ls = 0.f;
    if (V == 0) {
      expc(v0, m); ls += sumpack(v0, pk); stc(pdst, pk);
      expc(v1, m); ls += sumpack(v1, pk); stc(pdst + 16, pk);
      expc(v2, m); ls += sumpack(v2, pk); stc(pdst + 32, pk);
      expc(v3, m); ls += sumpack(v3, pk); stc(pdst + 48, pk);
    } else if (V == 1) {
      expc(v0, m);
      if (opq0) { expc(v1, m); ls += sumpack(v0, pk); stc(pdst, pk); }
      if (opq1) { expc(v2, m); ls += sumpack(v1, pk); stc(pdst + 16, pk); }
      if (opq2) { expc(v3, m); ls += sumpack(v2, pk); stc(pdst + 32, pk); }
      if (opq3) { ls += sumpack(v3, pk); stc(pdst + 48, pk); }
    } else {
      expc2(v0, m);
      if (opq0) { expc2(v1, m); ls += sumpack2(v0, pk); stc(pdst, pk); }
      if (opq1) { expc2(v2, m); ls += sumpack2(v1, pk); stc(pdst + 16, pk); }
      if (opq2) { expc2(v3, m); ls += sumpack2(v2, pk); stc(pdst + 32, pk); }
      if (opq3) { ls += sumpack2(v3, pk); stc(pdst + 48, pk); }
    }
    l += ls;
    // restore the scores for the next round (cheap, off the measured critical path would be nicer; same for all V)
#pragma unroll
    for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(mine + i) = make_float4(-0.1f * (it & 7), -0.3f, -1.f, -2.f);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int V> void run(const char* name, float* d) {
  for (int warps : {4, 8}) {
    int iters = 2000;
    size_t smem = (size_t)warps * 32 * 132 * 4;
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) { k<V><<<148, warps * 32, smem>>>(d, iters, 1, 1, 1, 1); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return; } }
    float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM=%d: %.0f cycles per 128-score row sweep per warp (%.2f cycles per MUFU per scheduler)\n", name, warps, cyc / iters, cyc / iters / 128 / (warps / 4));
  }
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  run<0>("V0 straight-line", d); run<1>("V1 staged (opaque branches)", d); run<2>("V2 staged + f32x2", d);
  return 0;
}
