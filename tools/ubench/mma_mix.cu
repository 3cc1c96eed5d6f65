// Micro-benchmark: cycles per attention "pass" of tcgen05.mma issued by one thread from a fully unrolled, static
// instruction sequence (no index arithmetic between MMAs): S0 PV1 S1 PV0 = 8 x (128x128x16) + 16 x (128x64x16).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../taiwan-tongues-asr-ce_b200/csrc/ptx_sm100.cuh"
using namespace ttasr;
enum { PASS_TS_MN = 0, PASS_SS_MN, S_ONLY, PV_ONLY, PASS_TS_K, PASS_PVN128, PASS_GROUPED, PASS_TS_MN_COMMITS };
template <int MODE>
__global__ void k(float* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_ptr;
  __shared__ unsigned long long bar, bar2[4];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&tmem_ptr), 512); tmem_relinquish<1>(); }
  if (threadIdx.x == 32) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar2[i]), 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 32) {
    const uint64_t q0 = umma_desc_sw128(smem0, 16, 1024), q1 = umma_desc_sw128(smem0 + 16384, 16, 1024);
    const uint64_t kd = umma_desc_sw128(smem0 + 32768, 16, 1024);
    const uint64_t vmn = umma_desc_sw128(smem0 + 49152, 16384, 1024), vk = umma_desc_sw128(smem0 + 49152, 16, 1024);
    const uint64_t p0s = umma_desc_sw128(smem0 + 65536, 16, 1024), p1s = umma_desc_sw128(smem0 + 65536 + 32768, 16, 1024);
    constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0), idPV = umma_idesc_bf16(128, 64, 0, 1), idPVk = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idPV128 = umma_idesc_bf16(128, 128, 0, 1);
    uint32_t phase = 0;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      auto S = [&](int t) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss<1>(tm + t * 128, (t ? q1 : q0) + 2 * kk, kd + 2 * kk, idS, kk != 0);
      };
      auto PV = [&](int t) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          if (MODE == PASS_SS_MN) umma_ss<1>(tm + 384 + t * 64, (t ? p1s : p0s) + 2 * (kk & 3) + (kk >> 2) * 1024, vmn + 128 * kk, idPV, 1);
          else if (MODE == PASS_TS_K) umma_ts(tm + 384 + t * 64, tm + 256 + t * 64 + kk * 8, vk + 2 * (kk & 3) + (kk >> 2) * 512, idPVk, 1);
          else if (MODE == PASS_PVN128) umma_ts(tm + 256, tm + 256 + t * 64 + kk * 8, vmn + 128 * kk, idPV128, 1);
          else umma_ts(tm + 384 + t * 64, tm + 256 + t * 64 + kk * 8, vmn + 128 * kk, idPV, 1);
        }
      };
      if (MODE == S_ONLY) { S(0); S(1); S(0); S(1); }
      else if (MODE == PV_ONLY) { PV(1); PV(0); }
      else if (MODE == PASS_GROUPED) { S(0); S(1); PV(1); PV(0); }
      else if (MODE == PASS_TS_MN_COMMITS) {
        S(0); umma_commit(smem_u32(&bar2[0])); PV(1); umma_commit(smem_u32(&bar2[1])); S(1); umma_commit(smem_u32(&bar2[2])); PV(0); umma_commit(smem_u32(&bar2[3]));
      } else { S(0); PV(1); S(1); PV(0); }
      if ((it & 3) == 3 || MODE == PASS_TS_MN_COMMITS) {  // let a few passes queue up
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
      }
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tm, 512);
}
template <int MODE>
__global__ void ku(float* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_ptr;
  __shared__ unsigned long long bar;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) { tmem_alloc<1>(smem_u32(&tmem_ptr), 512); tmem_relinquish<1>(); }
  if (threadIdx.x == 32) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint64_t q0 = umma_desc_sw128(smem0, 16, 1024), q1 = umma_desc_sw128(smem0 + 16384, 16, 1024);
    const uint64_t kd = umma_desc_sw128(smem0 + 32768, 16, 1024);
    const uint64_t vmn = umma_desc_sw128(smem0 + 49152, 16384, 1024);
    constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0), idPV = umma_idesc_bf16(128, 64, 0, 1);
    uint32_t phase = 0;
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      auto S = [&](int t) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss<1>(tm + t * 128, (t ? q1 : q0) + 2 * kk, kd + 2 * kk, idS, kk != 0);
        }
        __syncwarp();
      };
      auto PV = [&](int t) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ts(tm + 384 + t * 64, tm + 256 + t * 64 + kk * 8, vmn + 128 * kk, idPV, 1);
        }
        __syncwarp();
      };
      if (MODE == PV_ONLY) { PV(1); PV(0); } else { S(0); PV(1); S(1); PV(0); }
      if ((it & 3) == 3) {
        if (elect_one()) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
      }
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tm, 512);
}
template <int MODE> void runu(const char* name, float* d) {
  cudaFuncSetAttribute(ku<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  int iters = 400;
  for (int rep = 0; rep < 2; ++rep) { ku<MODE><<<148, 128, 160 * 1024>>>(d, iters); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return; } }
  float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  printf("%-34s: %.0f cycles per pass\n", name, cyc / iters);
}
template <int MODE> void run(const char* name, float* d) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  int iters = 400;
  for (int rep = 0; rep < 2; ++rep) { k<MODE><<<148, 128, 160 * 1024>>>(d, iters); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return; } }
  float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
  printf("%-34s: %.0f cycles per pass\n", name, cyc / iters);
}
int main() {
  float* d; cudaMalloc(&d, 1 << 20);
  run<S_ONLY>("S0 S1 S0 S1 (16 x 128x128x16)", d);
  run<PV_ONLY>("PV1 PV0 (16 x TS 128x64 MN-B)", d);
  run<PASS_TS_MN>("S0 PV1 S1 PV0 (TS, MN-major V)", d);
  run<PASS_GROUPED>("S0 S1 PV1 PV0 (TS, MN-major V)", d);
  run<PASS_SS_MN>("S0 PV1 S1 PV0 (P in smem)", d);
  run<PASS_TS_K>("S0 PV1 S1 PV0 (TS, K-major V)", d);
  run<PASS_PVN128>("S0 PV1 S1 PV0 (PV N=128 junk)", d);
  run<PASS_TS_MN_COMMITS>("S0 PV1 S1 PV0 + commit each, wait", d);
  runu<PASS_TS_MN>("uniform-issue S0 PV1 S1 PV0", d);
  runu<PV_ONLY>("uniform-issue PV1 PV0", d);
  return 0;
}
