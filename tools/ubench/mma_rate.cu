// Micro-benchmark: sustained cycles per tcgen05.mma for the shapes the attention kernel issues (one CTA per SM,
// one issuing thread, R back-to-back MMAs per commit).  Operand contents are irrelevant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../taiwan-tongues-asr-ce_b200/csrc/ptx_sm100.cuh"
using namespace ttasr;
enum { SS_128x128 = 0, SS_128x256, SS_128x64, TS_128x64_MN, TS_128x128_MN, TS_128x64_K, SS_128x64_MN, MIX_ATTN, SS_ALT_ACC, TS_ALT_ACC, SS_OVERWRITE, MIX_SS_TSK, MIX_ATTN_2T, SS_ALT_ACC1, TS_K_N128, MIX_ALLSS, MIX_ALLTS, MIX_ALLSS_2T, T1, T2, T3, T4, T5, T6 };
template <int MODE>
__global__ void k(float* out, int iters, int burst) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_ptr;
  __shared__ unsigned long long bar;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&tmem_ptr), 512); tmem_relinquish<1>(); }
  if (threadIdx.x == 32) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 32) {
    const uint32_t a_s = smem0, b_s = smem0 + 32768;   // A tile 128x64 (16 KB), B tile up to 256x64 (32 KB)
    uint32_t phase = 0;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int r = 0; r < burst; ++r) {
        const int kk = r & 3;
        if (MODE == SS_128x128) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * kk, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 128, 0, 0), 1);
        if (MODE == SS_128x256) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * kk, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 256, 0, 0), 1);
        if (MODE == SS_128x64) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * kk, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 64, 0, 0), 1);
        if (MODE == TS_128x64_MN) umma_ts(tm + 384, tm + 256 + (r & 7) * 8, umma_desc_sw128(b_s, 16384, 1024) + 128 * (r & 7), umma_idesc_bf16(128, 64, 0, 1), 1);
        if (MODE == TS_128x128_MN) umma_ts(tm, tm + 256 + (r & 7) * 8, umma_desc_sw128(b_s, 16384, 1024) + 128 * (r & 7), umma_idesc_bf16(128, 128, 0, 1), 1);
        if (MODE == TS_128x64_K) umma_ts(tm + 384, tm + 256 + (r & 3) * 8, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 64, 0, 0), 1);
        if (MODE == SS_128x64_MN) umma_ss<1>(tm + 384, umma_desc_sw128(a_s, 16, 1024) + 2 * kk, umma_desc_sw128(b_s, 16384, 1024) + 128 * (r & 7), umma_idesc_bf16(128, 64, 0, 1), 1);
        if (MODE == SS_ALT_ACC)  // 4 MMAs into accumulator 0, 4 into accumulator 1, ...
          umma_ss<1>(tm + ((r >> 2) & 1) * 128, umma_desc_sw128(a_s, 16, 1024) + 2 * kk, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 128, 0, 0), 1);
        if (MODE == SS_ALT_ACC1)  // alternate accumulators on every MMA
          umma_ss<1>(tm + (r & 1) * 128, umma_desc_sw128(a_s, 16, 1024) + 2 * kk, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 128, 0, 0), 1);
        if (MODE == TS_ALT_ACC)  // 8 PV MMAs into O0, 8 into O1
          umma_ts(tm + 384 + ((r >> 3) & 1) * 64, tm + 256 + ((r >> 3) & 1) * 64 + (r & 7) * 8, umma_desc_sw128(b_s, 16384, 1024) + 128 * (r & 7), umma_idesc_bf16(128, 64, 0, 1), 1);
        if (MODE == SS_OVERWRITE)  // first MMA of each group of 4 overwrites
          umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * kk, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 128, 0, 0), kk != 0);
        if (MODE == TS_K_N128)
          umma_ts(tm, tm + 256 + (r & 3) * 8, umma_desc_sw128(b_s, 16, 1024) + 2 * kk, umma_idesc_bf16(128, 128, 0, 0), 1);
        if (MODE == MIX_SS_TSK) {  // 4 S + 8 PV with K-major V
          const int q = r % 12;
          if (q < 4) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * q, umma_desc_sw128(b_s, 16, 1024) + 2 * q, umma_idesc_bf16(128, 128, 0, 0), 1);
          else umma_ts(tm + 384, tm + 256 + ((q - 4) & 3) * 8, umma_desc_sw128(b_s + 16384, 16, 1024) + 2 * ((q - 4) & 3), umma_idesc_bf16(128, 64, 0, 0), 1);
        }
        if (MODE == MIX_ATTN_2T) {  // the kernel's order: S0 PV1 S1 PV0 (two tiles, separate accumulators)
          const int q = r % 24;
          if (q < 4) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * q, umma_desc_sw128(b_s, 16, 1024) + 2 * q, umma_idesc_bf16(128, 128, 0, 0), q != 0);
          else if (q < 12) umma_ts(tm + 448, tm + 320 + (q - 4) * 8, umma_desc_sw128(b_s + 16384, 16384, 1024) + 128 * (q - 4), umma_idesc_bf16(128, 64, 0, 1), 1);
          else if (q < 16) umma_ss<1>(tm + 128, umma_desc_sw128(a_s + 16384, 16, 1024) + 2 * (q - 12), umma_desc_sw128(b_s, 16, 1024) + 2 * (q - 12), umma_idesc_bf16(128, 128, 0, 0), q != 12);
          else umma_ts(tm + 384, tm + 256 + (q - 16) * 8, umma_desc_sw128(b_s + 16384, 16384, 1024) + 128 * (q - 16), umma_idesc_bf16(128, 64, 0, 1), 1);
        }
        if (MODE == MIX_ALLSS) {  // 4 S + 8 PV, P read from shared memory (SS form, MN-major V)
          const int q = r % 12;
          if (q < 4) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * q, umma_desc_sw128(b_s, 16, 1024) + 2 * q, umma_idesc_bf16(128, 128, 0, 0), q != 0);
          else umma_ss<1>(tm + 384, umma_desc_sw128(a_s + 16384 + ((q - 4) >> 2) * 16384, 16, 1024) + 2 * ((q - 4) & 3), umma_desc_sw128(b_s + 16384, 16384, 1024) + 128 * (q - 4), umma_idesc_bf16(128, 64, 0, 1), 1);
        }
        if (MODE == MIX_ALLSS_2T) {  // S0 PV1 S1 PV0 with P in shared memory
          const int q = r % 24;
          if (q < 4) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * q, umma_desc_sw128(b_s, 16, 1024) + 2 * q, umma_idesc_bf16(128, 128, 0, 0), q != 0);
          else if (q < 12) umma_ss<1>(tm + 448, umma_desc_sw128(a_s + 16384 + ((q - 4) >> 2) * 16384, 16, 1024) + 2 * ((q - 4) & 3), umma_desc_sw128(b_s + 16384, 16384, 1024) + 128 * (q - 4), umma_idesc_bf16(128, 64, 0, 1), 1);
          else if (q < 16) umma_ss<1>(tm + 128, umma_desc_sw128(a_s + 16384, 16, 1024) + 2 * (q - 12), umma_desc_sw128(b_s, 16, 1024) + 2 * (q - 12), umma_idesc_bf16(128, 128, 0, 0), q != 12);
          else umma_ss<1>(tm + 384, umma_desc_sw128(a_s + 16384 + ((q - 16) >> 2) * 16384, 16, 1024) + 2 * ((q - 16) & 3), umma_desc_sw128(b_s + 16384, 16384, 1024) + 128 * (q - 16), umma_idesc_bf16(128, 64, 0, 1), 1);
        }
        if (MODE == MIX_ALLTS) {  // 4 S (Q in TMEM) + 8 PV (P in TMEM)
          const int q = r % 12;
          if (q < 4) umma_ts(tm, tm + 448 + q * 8, umma_desc_sw128(b_s, 16, 1024) + 2 * q, umma_idesc_bf16(128, 128, 0, 0), q != 0);
          else umma_ts(tm + 384, tm + 256 + (q - 4) * 8, umma_desc_sw128(b_s + 16384, 16384, 1024) + 128 * (q - 4), umma_idesc_bf16(128, 64, 0, 1), 1);
        }
        if (MODE >= T1 && MODE <= T6) {
          const int q = r % 12;
          const uint64_t a0 = umma_desc_sw128(a_s, 16, 1024), a1 = umma_desc_sw128(a_s + 16384, 16, 1024);
          const uint64_t bk = umma_desc_sw128(b_s, 16, 1024), bmn = umma_desc_sw128(b_s + 16384, 16384, 1024);
          if (q < 4) umma_ss<1>(tm, a0 + 2 * q, bk + 2 * q, umma_idesc_bf16(128, 128, 0, 0), 1);
          else if (MODE == T1) umma_ss<1>(tm + 384, a1 + 2 * (q & 3), bk + 2 * (q & 3), umma_idesc_bf16(128, 64, 0, 0), 1);      // N differs
          else if (MODE == T2) umma_ss<1>(tm + 256, a1 + 2 * (q & 3), bmn + 128 * (q - 4), umma_idesc_bf16(128, 128, 0, 1), 1);   // B major differs
          else if (MODE == T3) umma_ss<1>(tm + 256, a1 + 2 * (q & 3), bk + 2 * (q & 3), umma_idesc_bf16(128, 128, 0, 0), 1);      // same idesc
          else if (MODE == T4) umma_ss<1>(tm + 256, a1 + 2 * (q & 3), bk + 2 * (q & 3), umma_idesc_bf16(128, 256, 0, 0), 1);      // N=256
          else if (MODE == T5) umma_ts(tm + 256, tm + 448 + (q & 3) * 8, bk + 2 * (q & 3), umma_idesc_bf16(128, 128, 0, 0), 1);   // same idesc, TS
          else if (MODE == T6) umma_ss<1>(tm + 384, a1 + 2 * (q & 3), bk + 2 * (q & 3), umma_idesc_bf16(64, 128, 0, 0), 1);       // M=64
        }
        if (MODE == MIX_ATTN) {  // one attention tile: 4 S MMAs + 8 PV MMAs per 12
          const int q = r % 12;
          if (q < 4) umma_ss<1>(tm, umma_desc_sw128(a_s, 16, 1024) + 2 * q, umma_desc_sw128(b_s, 16, 1024) + 2 * q, umma_idesc_bf16(128, 128, 0, 0), 1);
          else umma_ts(tm + 384, tm + 256 + (q - 4) * 8, umma_desc_sw128(b_s + 16384, 16384, 1024) + 128 * (q - 4), umma_idesc_bf16(128, 64, 0, 1), 1);
        }
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tm, 512);
}
template <int MODE> void run(const char* name, float* d, double macs_per_mma) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int burst : {24, 48}) {
    int iters = 400;
    for (int rep = 0; rep < 2; ++rep) { k<MODE><<<148, 128, 100 * 1024>>>(d, iters, burst); cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return; } }
    float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    printf("%-16s burst=%2d: %.1f cycles per MMA (%.1f per burst incl. commit+wait); %.0f MAC/clk/SM\n", name, burst, cyc / iters / burst, cyc / iters, macs_per_mma * burst * iters / cyc);
  }
}
int main() {
  float* d; cudaMalloc(&d, 1 << 20);
  run<SS_128x128>("SS 128x128x16", d, 128. * 128 * 16);
  run<SS_128x256>("SS 128x256x16", d, 128. * 256 * 16);
  run<SS_128x64>("SS 128x64x16", d, 128. * 64 * 16);
  run<TS_128x64_MN>("TS 128x64 MN-B", d, 128. * 64 * 16);
  run<TS_128x128_MN>("TS 128x128 MN-B", d, 128. * 128 * 16);
  run<TS_128x64_K>("TS 128x64 K-B", d, 128. * 64 * 16);
  run<SS_128x64_MN>("SS 128x64 MN-B", d, 128. * 64 * 16);
  run<MIX_ATTN>("attn mix 4S+8PV", d, 128. * 64 * 16 * 4 / 3);
  run<SS_ALT_ACC>("SS alt acc /4", d, 128. * 128 * 16);
  run<SS_ALT_ACC1>("SS alt acc /1", d, 128. * 128 * 16);
  run<TS_ALT_ACC>("TS-MN alt acc /8", d, 128. * 64 * 16);
  run<SS_OVERWRITE>("SS overwrite/4", d, 128. * 128 * 16);
  run<TS_K_N128>("TS 128x128 K-B", d, 128. * 128 * 16);
  run<MIX_SS_TSK>("mix 4S+8PV(K-B)", d, 128. * 64 * 16 * 4 / 3);
  run<MIX_ATTN_2T>("mix S0 PV1 S1 PV0", d, 128. * 64 * 16 * 4 / 3);
  run<MIX_ALLSS>("all-SS 4S+8PV", d, 128. * 64 * 16 * 4 / 3);
  run<MIX_ALLSS_2T>("all-SS S0PV1S1PV0", d, 128. * 64 * 16 * 4 / 3);
  run<MIX_ALLTS>("all-TS 4S+8PV", d, 128. * 64 * 16 * 4 / 3);
  run<T1>("4 N128 + 8 N64 (K-B)", d, 128. * 64 * 16 * 4 / 3);
  run<T2>("4 K-B + 8 MN-B (N128)", d, 128. * 128 * 16);
  run<T3>("4 + 8 same idesc", d, 128. * 128 * 16);
  run<T4>("4 N128 + 8 N256", d, 128. * 128 * 16 * 5 / 3);
  run<T5>("4 SS + 8 TS same idesc", d, 128. * 128 * 16);
  run<T6>("4 M128 + 8 M64", d, 128. * 128 * 16 * 2 / 3);
  return 0;
}
