#pragma once
#include <cuda_runtime.h>
namespace ttasr {
// y[rows, d] = LayerNorm(x[rows, d]) * g + b, eps 1e-5; x fp32; y bf16 (out_f32 = 0) or fp32. d % 128 == 0, d <= 1280.
cudaError_t layernorm_launch(const float* x, const float* g, const float* b, void* y, long long rows, int d,
                             int out_f32, cudaStream_t stream);
// the same over the split residual stream x = hi + lo (two bf16 arrays [rows, d]; x_lo may be null)
cudaError_t layernorm_split_launch(const void* x_hi, const void* x_lo, const float* g, const float* b, void* y,
                                   long long rows, int d, int out_f32, cudaStream_t stream);
}  // namespace ttasr
