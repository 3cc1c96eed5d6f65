// Internal interface of the tcgen05 GEMM / implicit-GEMM-conv kernel family (gemm_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttasr {

enum GemmMode {
  kGemmPlain = 0,  // A[M, K] row-major
  kGemmConv1 = 1,  // Conv1d k=3 s=1 p=1 over a time-major [B, T, C] input
  kGemmConv2 = 2,  // Conv1d k=3 s=2 p=1 over a time-major [B, 2*T, C] input (T = output rows per chunk)
};

struct GemmCall {
  int mode = kGemmPlain;
  // A operand (bf16). plain: a[M, lda]; conv: time-major activations [B, T_in, lda]
  const void* a = nullptr;
  int64_t lda = 0;          // row stride in elements
  int a_inner = 0;          // valid channels per row (K for plain; C_in for conv) — reads beyond are zero-filled
  int rows = 0;             // OUTPUT rows per batch item (M for plain)
  int nbatch = 1;
  // B operand: packed weights [N, kp] bf16 row-major, kp = k_blocks * 64 (conv: tap-major, channels padded per tap)
  const void* w = nullptr;
  int n = 0;
  int k_blocks = 0;
  int kb_per_tap = 0;       // k-blocks per filter tap (== k_blocks for plain)
  // epilogue: out = act(acc + bias) (+ addend)
  const float* bias = nullptr;      // [n] fp32 (required; pass zeros for "no bias")
  const float* addend = nullptr;    // fp32 [nbatch*rows, n] or, if addend_bcast, [rows, n] shared by every batch item
  int addend_bcast = 0;
  int act = 0;                      // 0 identity, 1 GELU(erf)
  void* out = nullptr;              // [nbatch*rows, n]
  int out_f32 = 0;
  int cta_group = 0;                // 0 = default, 1, 2
  // ---- split residual stream + LayerNorm fusion (the default encoder path).
  // The residual stream x is kept as TWO bf16 arrays, x = hi + lo (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits
  // at the HBM cost of one fp32 array).  A residual GEMM with `split` set computes
  //     x' = act(acc + bias) + (addend_hi + addend_lo)      (fp32 in registers)
  // and writes out = hi(x'), out_lo = lo(x') (out_lo / addend_lo may be null: plain bf16 stream).  `hi` IS the bf16 A
  // operand of the GEMM that consumes LN(x'), so no LayerNorm kernel and no extra copy exist: the residual GEMM also
  // emits, per row, n / 64 partial statistics (mean_i, M2_i) of its hi values, one per 64 columns (independent of the tile shape)
  // (ln_stats_out; *ln_parts_out receives the count), and the consuming GEMM runs on hi with gamma folded into W and
  // applies  out = rstd * (acc - mean * c1[n]) + bias[n]  in its epilogue (bf16 out), c1[n] = sum_k W'[n][k],
  // bias[n] = b[n] + sum_k beta[k] W[n][k]; mean / rstd are combined from the partials (ln_stats_in, ln_parts_in per
  // row) with Chan's parallel-variance formula in a fixed order.
  int split = 0;
  void* out_lo = nullptr;
  const void* addend_hi = nullptr;  // bf16 [nbatch*rows, n] (or [rows, n] with addend_bcast)
  const void* addend_lo = nullptr;
  void* ln_stats_out = nullptr;
  int* ln_parts_out = nullptr;
  const void* ln_stats_in = nullptr;
  int ln_parts_in = 0;
  const float* ln_c1 = nullptr;
  float ln_eps = 1e-5f;
};

// TMA descriptor helper (driver entry point resolved at run time; libcuda is not linked)
CUresult encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* ptr, const uint64_t* dims,
                     const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box, CUtensorMapSwizzle swizzle);

// Returns cudaSuccess or the launch error; *why (optional) gets a static string on argument errors.
cudaError_t gemm_launch(const GemmCall& c, int num_sms, cudaStream_t stream, const char** why);

}  // namespace ttasr
