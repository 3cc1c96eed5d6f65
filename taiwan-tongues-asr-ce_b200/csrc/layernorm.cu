// Row LayerNorm (eps 1e-5, torch semantics: biased variance) over an fp32 residual stream, emitting the bf16
// A-operand of the next tcgen05 GEMM (or fp32/bf16 final hidden states).
// Reference: nn.LayerNorm uses at modeling_whisper.py:372,378,393,403,574,643.
// One warp per row, the row lives in registers (<= 1280 channels = 10 float4 per lane), two-pass statistics,
// 128-bit loads / 64- or 128-bit stores; pure HBM streaming (read 4 B, write 2 B per element).
#include "layernorm.h"
#include "ptx_sm100.cuh"

namespace ttasr {
namespace {

constexpr int kMaxVec = 10;  // d <= 1280

__device__ __forceinline__ float4 bf16x4_to_f32(uint2 w) {
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                     __uint_as_float(w.y & 0xffff0000u));
}

// SPLIT_IN: the row is the split residual stream x = hi + lo (two bf16 arrays, `x` = hi, `x_lo` = lo or null)
template <bool OUT_F32, bool SPLIT_IN>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x, const void* __restrict__ x_lo,
                                                        const float* __restrict__ g, const float* __restrict__ bta,
                                                        void* __restrict__ y, long long rows, int d) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();
  if (row >= rows) return;
  const int nvec = d >> 7;  // float4 per lane
  float4 v[kMaxVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      if constexpr (SPLIT_IN) {
        const uint2* hr = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x) + row * d);
        v[i] = bf16x4_to_f32(__ldg(hr + i * 32 + lane));
        if (x_lo) {
          const uint2* lr = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x_lo) + row * d);
          const float4 l = bf16x4_to_f32(__ldg(lr + i * 32 + lane));
          v[i].x += l.x; v[i].y += l.y; v[i].z += l.z; v[i].w += l.w;
        }
      } else {
        v[i] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(x) + row * d) + i * 32 + lane);
      }
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(d);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + e * e);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(d) + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(bta);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nvec) {
      const float4 gg = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gg.x + bb.x;
      o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
      o.z = (v[i].z - mean) * rstd * gg.z + bb.z;
      o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
      if constexpr (OUT_F32) {
        reinterpret_cast<float4*>(static_cast<float*>(y) + row * d)[i * 32 + lane] = o;
      } else {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + row * d)[i * 32 + lane] = pk;
      }
    }
  }
}

template <bool OUT_F32, bool SPLIT_IN>
cudaError_t launch_ln(unsigned grid, cudaStream_t stream, const void* x, const void* x_lo, const float* g, const float* b,
                      void* y, long long rows, int d) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_launch_attr(&attr[0]);
  return cudaLaunchKernelEx(&cfg, layernorm_kernel<OUT_F32, SPLIT_IN>, x, x_lo, g, b, y, rows, d);
}

}  // namespace

cudaError_t layernorm_launch(const float* x, const float* g, const float* b, void* y, long long rows, int d,
                             int out_f32, cudaStream_t stream) {
  if (d % 128 != 0 || d <= 0 || d > 128 * kMaxVec) return cudaErrorInvalidValue;
  if (rows <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  return out_f32 ? launch_ln<true, false>(grid, stream, x, nullptr, g, b, y, rows, d)
                 : launch_ln<false, false>(grid, stream, x, nullptr, g, b, y, rows, d);
}

cudaError_t layernorm_split_launch(const void* x_hi, const void* x_lo, const float* g, const float* b, void* y,
                                   long long rows, int d, int out_f32, cudaStream_t stream) {
  if (d % 128 != 0 || d <= 0 || d > 128 * kMaxVec) return cudaErrorInvalidValue;
  if (rows <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  return out_f32 ? launch_ln<true, true>(grid, stream, x_hi, x_lo, g, b, y, rows, d)
                 : launch_ln<false, true>(grid, stream, x_hi, x_lo, g, b, y, rows, d);
}

}  // namespace ttasr
