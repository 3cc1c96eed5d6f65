#pragma once
#include <cuda_runtime.h>
namespace ttasr {
// qkv: [batch, n_ctx, 3*d] bf16 (q | k | v; head h occupies columns [64h, 64h+64) of each third; q pre-scaled).
// out: [batch, n_ctx, d] bf16.  d = 64 * n_heads.
cudaError_t attention_launch(const void* qkv, void* out, int batch, int n_ctx, int n_heads, int num_sms,
                             cudaStream_t stream, const char** why);
// the four-warpgroup kernel (attention4_sm100.cu); same contract
cudaError_t attention4_launch(const void* qkv, void* out, int batch, int n_ctx, int n_heads, int num_sms,
                              cudaStream_t stream, const char** why);
int attention_variant();  // 2 or 4 (TTASR_ATTN_KERNEL=2wg|4wg, else the build's default)
}  // namespace ttasr
