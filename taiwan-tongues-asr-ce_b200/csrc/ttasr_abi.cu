// C ABI (include/ttasr_abi.h) over the sm_100a kernels: handle management, weight packing, the encoder's launch
// sequence.  No torch types, no exceptions across the boundary, no CPU fallback.
#include "../../include/ttasr_abi.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "attention_sm100.h"
#include "fft400.cuh"
#include "frontend_logmel.h"
#include "gemm_sm100.h"
#include "ingest_resample.h"
#include "layernorm.h"
#include "ptx_sm100.cuh"

using namespace ttasr;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return fail(TTASR_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

int check_arch(int device, int* num_sms) {
  int major = 0, sms = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess)
    return fail(TTASR_E_CUDA, "no usable CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
  if (major != 10)
    return fail(TTASR_E_ARCH, "device %d has compute capability %d.x; this library only runs on sm_100 (B200)", device, major);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (num_sms) *num_sms = sms;
  return TTASR_OK;
}

// ------------------------------------------------------------------------------------------- small device helpers
__global__ void scale_copy_bf16_kernel(__nv_bfloat16* dst, const __nv_bfloat16* src, long long n, float scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(__bfloat162float(src[i]) * scale);
}
__global__ void scale_copy_f32_kernel(float* dst, const float* src, long long n, float scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src ? src[i] * scale : 0.f;
}
// LayerNorm folded into the consuming projection:  LN(x) W^T + b = rstd * (x W'^T - mean * c1) + c2  with
// W'[n][k] = bf16(scale * W[n][k] * gamma[k]),  c1[n] = sum_k W'[n][k] (of the ROUNDED weights, as the MMA sees them),
// c2[n] = scale * b[n] + sum_k beta[k] * scale * W[n][k].  One warp per output row n.
__global__ void fold_ln_kernel(__nv_bfloat16* dst, const __nv_bfloat16* src, int n_rows, int k, float scale,
                               const float* gamma, const float* beta, const float* bias, float* c1, float* c2) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= n_rows) return;
  float s1 = 0.f, s2 = 0.f;
  for (int j = lane; j < k; j += 32) {
    const float w = __bfloat162float(src[static_cast<size_t>(n) * k + j]) * scale;
    const __nv_bfloat16 wf = __float2bfloat16(w * gamma[j]);
    dst[static_cast<size_t>(n) * k + j] = wf;
    s1 += __bfloat162float(wf);
    s2 = fmaf(beta[j], w, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    c1[n] = s1;
    c2[n] = s2 + (bias ? bias[n] * scale : 0.f);
  }
}
// fp32 -> split pair: hi = bf16(x), lo = bf16(x - hi)
__global__ void split_f32_kernel(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* src, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  const __nv_bfloat16 h = __float2bfloat16(v);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16(v - __bfloat162float(h));
}
// conv weight [n_out, c_in, 3] -> tap-major [n_out, 3 * c_pad] (zero padded channels)
__global__ void pack_conv_kernel(__nv_bfloat16* dst, const __nv_bfloat16* src, int n_out, int c_in, int c_pad) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_out) * 3 * c_pad;
  if (i >= total) return;
  const int c = static_cast<int>(i % c_pad);
  const int tap = static_cast<int>((i / c_pad) % 3);
  const int n = static_cast<int>(i / (3LL * c_pad));
  dst[i] = c < c_in ? src[(static_cast<long long>(n) * c_in + c) * 3 + tap] : __float2bfloat16(0.f);
}
// [B, C, T] fp32 -> [B, T, ld] bf16 (channels zero padded to ld)
__global__ void feats_to_tmajor_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int T, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 256 threads: 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    tile[r][tx] = (c < C && t < T) ? in[(static_cast<long long>(b) * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    if (t < T && c < ld) out[(static_cast<long long>(b) * T + t) * ld + c] = __float2bfloat16(tile[tx][r]);
  }
}

inline unsigned blocks_for(long long n, int bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

}  // namespace

// =================================================================================================== front end
struct ttasr_frontend {
  int device = 0;
  int num_sms = 0;
  int n_mels = 0;
  int n_samples = 0;
  FrontTables tables{};
  MelProgram mel{};               // host copy: passed to the frames kernel as a parameter (constant bank)
  int mel_baked = 0;              // n_mels when the bank equals one compiled into the kernel (mel_baked.inc), else 0
  void* table_mem = nullptr;
  unsigned* chunk_max = nullptr;  // per-chunk running maxima (order-encoded), capacity max_batch
  float* tile_min = nullptr;      // per-tile minima, capacity max_batch * tiles per chunk
  int64_t max_batch = 0;
};

extern "C" {

int ttasr_abi_version(void) { return TTASR_ABI_VERSION; }
const char* ttasr_last_error(void) { return g_err; }

int ttasr_device_check(int device) { return check_arch(device, nullptr); }

int ttasr_frontend_create(int n_mels, int n_fft, int hop, int n_samples, const float* mel_filters, const float* window,
                          ttasr_frontend_t** out) {
  if (!out || !mel_filters || !window) return fail(TTASR_E_ARG, "frontend_create: null argument");
  *out = nullptr;
  if (n_fft != kNfft || hop != kHop) return fail(TTASR_E_ARG, "frontend_create: only n_fft=400, hop=160 are implemented (got %d, %d)", n_fft, hop);
  if (n_mels <= 0 || n_mels > kMaxMels) return fail(TTASR_E_SHAPE, "frontend_create: n_mels must be in [1, %d] (got %d)", kMaxMels, n_mels);
  if (n_samples <= 0 || n_samples % hop != 0) return fail(TTASR_E_SHAPE, "frontend_create: n_samples must be a positive multiple of %d", hop);
  int device = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, &sms);
  if (rc != TTASR_OK) return rc;

  // ---- host tables
  std::vector<float2> tw(kNfft);
  for (int k1 = 0; k1 < kRadix; ++k1)
    for (int n2 = 0; n2 < kRadix; ++n2) {
      const double ang = -2.0 * M_PI * static_cast<double>(n2 * k1) / kNfft;
      tw[k1 * kRadix + n2] = make_float2(static_cast<float>(cos(ang)), static_cast<float>(sin(ang)));
    }
  // ---- mel projection as a streaming program per warp (frontend_logmel.cu, mel pass): a frequency bin of a
  // triangular bank feeds at most two ADJACENT filters, so a warp that owns filters [m0, m1) walks its bins once in
  // ascending order with two accumulators (filter m_cur and m_cur + 1) and "emits" a filter when its last bin is
  // behind it.  op = {bin, weight for m_cur, weight for m_cur + 1, filters to emit after this bin}.
  std::vector<int> first(n_mels, -1), last(n_mels, -1);
  for (int m = 0; m < n_mels; ++m)
    for (int k = 0; k < kNfreq; ++k)
      if (mel_filters[k * n_mels + m] != 0.f) {
        if (first[m] < 0) first[m] = k;
        last[m] = k;
      }
  // contiguous filter ranges of roughly equal cost (bins walked + ~2 bin-equivalents per emitted filter)
  std::vector<int> warp_m0(kMelWarps + 1, n_mels);
  {
    std::vector<double> cost(n_mels);
    double total = 0;
    for (int m = 0; m < n_mels; ++m) {
      const int prev_last = (m > 0 && last[m - 1] >= 0) ? last[m - 1] : -1;
      cost[m] = 2.0 + (last[m] >= 0 ? std::max(0, last[m] - std::max(prev_last, first[m] - 1)) : 0);
      total += cost[m];
    }
    double acc = 0;
    int w = 0;
    warp_m0[0] = 0;
    for (int m = 0; m < n_mels; ++m) {
      if (w + 1 < kMelWarps && acc >= total * (w + 1) / kMelWarps) warp_m0[++w] = m;
      acc += cost[m];
    }
    for (int i = w + 1; i <= kMelWarps; ++i) warp_m0[i] = n_mels;
  }
  std::vector<int4> ops;
  std::vector<int> op_off(kMelWarps + 1, 0);
  auto weight = [&](int k, int m) { return mel_filters[k * n_mels + m]; };
  for (int w = 0; w < kMelWarps; ++w) {
    op_off[w] = static_cast<int>(ops.size());
    const int m0 = warp_m0[w], m1 = warp_m0[w + 1];
    if (m0 >= m1) continue;
    int k_begin = kNfreq, k_end = -1;
    for (int m = m0; m < m1; ++m)
      if (first[m] >= 0) {
        k_begin = std::min(k_begin, first[m]);
        k_end = std::max(k_end, last[m]);
      }
    int m_cur = m0;
    // n more filters are complete: the flag goes on the last op; further completions (empty filters, narrow
    // triangles) get weightless ops of their own, so the kernel only ever tests "emit one filter after this op"
    auto emit_into_last = [&](int n) {
      for (; n > 0; --n, ++m_cur) {
        if (static_cast<int>(ops.size()) == op_off[w] || ops.back().w != 0) ops.push_back(make_int4(0, 0, 0, 0));
        ops.back().w = 1;
      }
    };
    for (int k = k_begin; k <= k_end && m_cur < m1; ++k) {
      int done = 0;  // filters (from m_cur on) whose last non-zero bin lies before k
      while (m_cur + done < m1 && last[m_cur + done] < k) ++done;
      emit_into_last(done);
      if (m_cur >= m1) break;
      for (int m = m0; m < m1; ++m)
        if (weight(k, m) != 0.f && (m < m_cur || m > m_cur + 1)) {
          return fail(TTASR_E_SHAPE, "frontend_create: mel filter bank is not a bank of overlapping triangles "
                                     "(frequency bin %d feeds filter %d besides %d and %d)", k, m, m_cur, m_cur + 1);
        }
      const float wa = weight(k, m_cur);
      const float wb = (m_cur + 1 < m1) ? weight(k, m_cur + 1) : 0.f;
      int4 op;
      op.x = k * kPowerPitchBytes;  // byte offset of the bin's row in the power-spectrum buffer
      memcpy(&op.y, &wa, 4);
      memcpy(&op.z, &wb, 4);
      op.w = 0;
      ops.push_back(op);
    }
    emit_into_last(m1 - m_cur);
  }
  op_off[kMelWarps] = static_cast<int>(ops.size());
  if (static_cast<int>(ops.size()) > kMaxMelOps)
    return fail(TTASR_E_SHAPE, "frontend_create: mel program too long (%d > %d ops)", static_cast<int>(ops.size()), kMaxMelOps);
  ops.resize(kMaxMelOps, make_int4(0, 0, 0, 0));

  ttasr_frontend* h = new ttasr_frontend();
  h->device = device;
  h->num_sms = sms;
  h->n_mels = n_mels;
  h->n_samples = n_samples;
  memcpy(h->mel.ops, ops.data(), sizeof(int4) * kMaxMelOps);
  memcpy(h->mel.op_off, op_off.data(), sizeof(int) * (kMelWarps + 1));
  memcpy(h->mel.m0, warp_m0.data(), sizeof(int) * (kMelWarps + 1));
  {  // the two Whisper banks have their projection compiled into the kernel; any other table runs the program above
    unsigned long long hash = 0xcbf29ce484222325ull;  // FNV-1a 64 over the bytes of the fp32 table
    const unsigned char* bytes_in = reinterpret_cast<const unsigned char*>(mel_filters);
    for (size_t i = 0; i < sizeof(float) * kNfreq * static_cast<size_t>(n_mels); ++i) hash = (hash ^ bytes_in[i]) * 0x100000001b3ull;
    const char* force = getenv("TTASR_FRONTEND_MEL");
    const bool generic = force && !strcmp(force, "generic");
    if (!generic && frontend_baked_hash(n_mels) != 0ull && hash == frontend_baked_hash(n_mels)) h->mel_baked = n_mels;
  }
  const size_t bytes = sizeof(float2) * kNfft + sizeof(float) * kNfft;
  cudaError_t e = cudaMalloc(&h->table_mem, bytes);
  if (e != cudaSuccess) { delete h; return fail(TTASR_E_NOMEM, "frontend_create: cudaMalloc tables: %s", cudaGetErrorString(e)); }
  char* pdev = static_cast<char*>(h->table_mem);
  auto put = [&](const void* src, size_t n) -> const void* {
    void* dst = pdev;
    if (e == cudaSuccess) e = cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
    pdev += n;
    return dst;
  };
  h->tables.twiddle = static_cast<const float2*>(put(tw.data(), sizeof(float2) * kNfft));
  h->tables.window = static_cast<const float*>(put(window, sizeof(float) * kNfft));
  h->max_batch = 1 << 14;
  if (e == cudaSuccess) e = cudaMalloc(&h->chunk_max, sizeof(unsigned) * h->max_batch);
  if (e == cudaSuccess) e = cudaMalloc(&h->tile_min, sizeof(float) * h->max_batch * frontend_tiles(n_samples));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(h->table_mem);
    cudaFree(h->chunk_max);
    cudaFree(h->tile_min);
    delete h;
    return fail(TTASR_E_CUDA, "frontend_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return TTASR_OK;
}

int ttasr_frontend_max_batch(const ttasr_frontend_t* h, int64_t* out) {
  if (!h || !out) return fail(TTASR_E_ARG, "frontend_max_batch: null argument");
  *out = h->max_batch;
  return TTASR_OK;
}

int ttasr_frontend_mel_mode(const ttasr_frontend_t* h, int* out) {
  if (!h || !out) return fail(TTASR_E_ARG, "frontend_mel_mode: null argument");
  *out = h->mel_baked;
  return TTASR_OK;
}

int ttasr_frontend_run(const ttasr_frontend_t* h, const void* pcm_dev, int pcm_dtype, int64_t batch, int64_t row_stride,
                       const int32_t* n_valid_dev, float* feats_dev, void* tmajor_dev, int tmajor_ld, void* stream) {
  return ttasr_frontend_run_ex(h, pcm_dev, pcm_dtype, batch, row_stride, n_valid_dev, feats_dev, tmajor_dev, tmajor_ld,
                               8.0f, stream);
}

int ttasr_frontend_run_ex(const ttasr_frontend_t* h, const void* pcm_dev, int pcm_dtype, int64_t batch, int64_t row_stride,
                          const int32_t* n_valid_dev, float* feats_dev, void* tmajor_dev, int tmajor_ld,
                          float clamp_decades, void* stream) {
  if (!h) return fail(TTASR_E_ARG, "frontend_run: null handle");
  if (!(clamp_decades > 0.f)) return fail(TTASR_E_ARG, "frontend_run: clamp_decades must be positive (8 = Whisper, +inf = none)");
  if (batch < 0 || batch > h->max_batch) return fail(TTASR_E_SHAPE, "frontend_run: batch %lld outside [0, %lld]", (long long)batch, (long long)h->max_batch);
  if (batch == 0) return TTASR_OK;
  if (!pcm_dev) return fail(TTASR_E_ARG, "frontend_run: null pcm buffer");
  if (!feats_dev && !tmajor_dev) return fail(TTASR_E_ARG, "frontend_run: give feats_dev, tmajor_dev or both");
  if (pcm_dtype != TTASR_PCM_F32 && pcm_dtype != TTASR_PCM_I16) return fail(TTASR_E_ARG, "frontend_run: bad pcm_dtype %d", pcm_dtype);
  if (!n_valid_dev && row_stride < h->n_samples) return fail(TTASR_E_SHAPE, "frontend_run: row_stride %lld < n_samples %d without n_valid", (long long)row_stride, h->n_samples);
  if (tmajor_dev && (tmajor_ld < h->n_mels || (tmajor_ld & 1))) return fail(TTASR_E_SHAPE, "frontend_run: tmajor_ld must be even and >= n_mels");
  cudaError_t e = launch_logmel(pcm_dev, pcm_dtype == TTASR_PCM_I16, row_stride, n_valid_dev, h->n_samples, h->n_mels,
                                static_cast<int>(batch), h->tables, h->mel, feats_dev, h->chunk_max, h->tile_min,
                                static_cast<__nv_bfloat16*>(tmajor_dev), tmajor_ld, h->num_sms,
                                static_cast<cudaStream_t>(stream), clamp_decades, h->mel_baked);
  if (e != cudaSuccess) return fail(TTASR_E_CUDA, "frontend_run: launch failed: %s", cudaGetErrorString(e));
  return TTASR_OK;
}

void ttasr_frontend_destroy(ttasr_frontend_t* h) {
  if (!h) return;
  cudaFree(h->table_mem);
  cudaFree(h->chunk_max);
  cudaFree(h->tile_min);
  delete h;
}

}  // extern "C"

// =================================================================================================== ingest
struct ttasr_ingest {
  IngestPlan plan;
  float* table = nullptr;
};

extern "C" {

int ttasr_ingest_create(int up, int down, const float* taps, int n_taps, ttasr_ingest_t** out) {
  if (!out) return fail(TTASR_E_ARG, "ingest_create: null argument");
  *out = nullptr;
  if (up < 1 || down < 1 || up > 4096 || down > 4096) return fail(TTASR_E_ARG, "ingest_create: up/down must be in [1, 4096] (got %d, %d)", up, down);
  if (std::__gcd(up, down) != 1) return fail(TTASR_E_ARG, "ingest_create: up/down must be reduced by their gcd (got %d/%d)", up, down);
  const bool identity = (up == 1 && down == 1);
  if (!identity && (!taps || n_taps < 1 || (n_taps & 1) == 0 || n_taps > (1 << 20)))
    return fail(TTASR_E_ARG, "ingest_create: the prototype filter must have an odd number of taps (got %d)", n_taps);
  int device = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, nullptr);
  if (rc != TTASR_OK) return rc;
  auto* h = new ttasr_ingest();
  h->plan.up = up;
  h->plan.down = down;
  if (!identity) {
    // scipy.signal.resample_poly: centre the filter on the output grid
    const int half_len = (n_taps - 1) / 2;
    h->plan.n_taps = n_taps;
    h->plan.n_pre_pad = down - half_len % down;
    h->plan.n_pre_remove = (half_len + h->plan.n_pre_pad) / down;
    h->plan.taps_per_phase = (n_taps + up - 1) / up;
    std::vector<float> tab(static_cast<size_t>(up) * h->plan.taps_per_phase, 0.f);
    for (int j = 0; j < n_taps; ++j) tab[static_cast<size_t>(j % up) * h->plan.taps_per_phase + j / up] = taps[j];
    cudaError_t e = cudaMalloc(&h->table, tab.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->table, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(h->table);
      delete h;
      return fail(TTASR_E_CUDA, "ingest_create: %s", cudaGetErrorString(e));
    }
    h->plan.phase_taps = h->table;
  }
  *out = h;
  return TTASR_OK;
}

int ttasr_ingest_out_len(const ttasr_ingest_t* h, int64_t n_in, int64_t* n_out) {
  if (!h || !n_out) return fail(TTASR_E_ARG, "ingest_out_len: null argument");
  if (n_in < 0) return fail(TTASR_E_SHAPE, "ingest_out_len: negative length");
  const int64_t x = n_in * h->plan.up;
  *n_out = x / h->plan.down + (x % h->plan.down != 0);
  return TTASR_OK;
}

int ttasr_ingest_run(const ttasr_ingest_t* h, const void* pcm_dev, int pcm_dtype, int channels, int64_t n_in,
                     float* out_dev, int64_t out_capacity, void* stream) {
  if (!h) return fail(TTASR_E_ARG, "ingest_run: null handle");
  if (pcm_dtype != TTASR_PCM_F32 && pcm_dtype != TTASR_PCM_I16) return fail(TTASR_E_ARG, "ingest_run: bad pcm_dtype %d", pcm_dtype);
  if (channels < 1 || channels > 8) return fail(TTASR_E_SHAPE, "ingest_run: channels must be in [1, 8] (got %d)", channels);
  if (n_in < 0 || out_capacity < 0) return fail(TTASR_E_SHAPE, "ingest_run: negative length");
  int64_t n_out = 0;
  ttasr_ingest_out_len(h, n_in, &n_out);
  if (out_capacity < n_out) return fail(TTASR_E_SHAPE, "ingest_run: out_capacity %lld < %lld resampled samples", (long long)out_capacity, (long long)n_out);
  if (out_capacity == 0) return TTASR_OK;
  if (!out_dev || (n_in > 0 && !pcm_dev)) return fail(TTASR_E_ARG, "ingest_run: null buffer");
  const size_t esz = pcm_dtype == TTASR_PCM_I16 ? 2 : 4;
  if (channels == 2 && (reinterpret_cast<uintptr_t>(pcm_dev) % (2 * esz)) != 0) return fail(TTASR_E_ARG, "ingest_run: stereo input must be aligned to one frame");
  cudaError_t e = launch_ingest(h->plan, pcm_dev, pcm_dtype == TTASR_PCM_I16, channels, n_in, out_dev, n_out, out_capacity,
                                static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(TTASR_E_CUDA, "ingest_run: launch failed: %s", cudaGetErrorString(e));
  return TTASR_OK;
}

int ttasr_ingest_frame_energy(const float* pcm_dev, int64_t n, int win, int hop, float* db_dev, int64_t n_frames,
                              void* stream) {
  if (!pcm_dev || !db_dev) return fail(TTASR_E_ARG, "ingest_frame_energy: null buffer");
  if (n < 0 || win <= 0 || hop <= 0 || n_frames < 0) return fail(TTASR_E_SHAPE, "ingest_frame_energy: bad sizes");
  cudaError_t e = launch_frame_energy(pcm_dev, n, win, hop, db_dev, n_frames, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(TTASR_E_CUDA, "ingest_frame_energy: %s", cudaGetErrorString(e));
  return TTASR_OK;
}

int ttasr_ingest_gather_rows(const float* pcm_dev, const int64_t* start_dev, const int32_t* len_dev, float* out_dev,
                             int n_rows, int row_samples, void* stream) {
  if (!pcm_dev || !start_dev || !len_dev || !out_dev) return fail(TTASR_E_ARG, "ingest_gather_rows: null buffer");
  if (n_rows < 0 || n_rows > 65535 || row_samples <= 0) return fail(TTASR_E_SHAPE, "ingest_gather_rows: bad sizes");
  cudaError_t e = launch_gather_rows(pcm_dev, reinterpret_cast<const long long*>(start_dev), len_dev, out_dev, n_rows,
                                     row_samples, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(TTASR_E_CUDA, "ingest_gather_rows: %s", cudaGetErrorString(e));
  return TTASR_OK;
}

void ttasr_ingest_destroy(ttasr_ingest_t* h) {
  if (!h) return;
  cudaFree(h->table);
  delete h;
}

}  // extern "C"

// =================================================================================================== encoder
struct LayerDev {
  const float *ln1_g, *ln1_b, *bqkv, *bo, *ln2_g, *ln2_b, *b1, *b2;
  const __nv_bfloat16 *wqkv, *wo, *w1, *w2;
  // LayerNorm-fused mode: wqkv / w1 hold the gamma-folded weights, bqkv / b1 the folded biases (c2), and
  const float *c1_qkv = nullptr, *c1_fc1 = nullptr;   // column sums of the folded weights
};

// How the residual stream x is held between the blocks (ttasr_encoder_create_ex / TTASR_RESIDUAL):
//   split (default): two bf16 arrays hi + lo; the per-layer LayerNorms are folded into the QKV / fc1 GEMMs, whose A
//                    operand is `hi` itself; statistics come from the residual GEMMs' epilogues.  No LayerNorm kernel
//                    between the blocks, residual traffic = one fp32 array.
//   f32:             fp32 array + a LayerNorm kernel before each QKV / fc1 GEMM (the reference-precision path)
//   bf16:            `hi` only (what an all-bf16 Hugging Face run does); half the residual traffic, measured in
//                    DESIGN.md, not the default
#ifndef TTASR_RESIDUAL_DEFAULT
#define TTASR_RESIDUAL_DEFAULT TTASR_RESIDUAL_SPLIT
#endif

enum ProfileKind { kProfPrep = 0, kProfConv1, kProfConv2, kProfLayerNorm, kProfQkv, kProfAttention, kProfOutProj,
                   kProfFc1, kProfFc2, kProfKinds };
static_assert(kProfKinds == TTASR_PROFILE_KINDS, "profile kinds out of sync with the header");
const char* const kProfNames[kProfKinds] = {"feats_to_time_major", "conv1_gemm", "conv2_gemm", "layernorm", "qkv_gemm",
                                            "attention", "out_proj_gemm", "fc1_gemm", "fc2_gemm"};

struct ProfileState {
  bool on = false;
  struct Span { int kind; cudaEvent_t a, b; };
  std::vector<Span> pending;
  std::vector<cudaEvent_t> pool;
  double ms[kProfKinds] = {0};
  int64_t launches[kProfKinds] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};

struct ttasr_encoder {
  mutable ProfileState prof;
  int device = 0, num_sms = 0;
  ttasr_encoder_cfg cfg{};
  int c1_pad = 128;        // conv1 input channels padded per tap
  int residual = TTASR_RESIDUAL_SPLIT;
  bool fuse_ln = true;     // residual != f32: per-layer LayerNorms folded into the QKV / fc1 GEMMs
  void* arena = nullptr;   // all packed weights
  const __nv_bfloat16 *conv1_w = nullptr, *conv2_w = nullptr, *pos_hi = nullptr, *pos_lo = nullptr;
  const float *conv1_b = nullptr, *conv2_b = nullptr, *pos = nullptr, *lnp_g = nullptr, *lnp_b = nullptr;
  std::vector<LayerDev> layers;
};

namespace {

struct WsLayout {
  size_t x, h, qkv, ffn, st0, st1, total;
};
WsLayout ws_layout(const ttasr_encoder_cfg& c, int64_t B, bool fuse_ln) {
  auto up = [](size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); };
  const size_t rows = static_cast<size_t>(B) * c.n_ctx;
  WsLayout w;
  size_t o = 0;
  // residual stream: fp32 [rows, d], or the split pair hi | lo (two bf16 [rows, d] halves of the same region)
  w.x = o;   o += up(rows * c.d_model * 2) * 2;
  w.h = o;   o += up(rows * c.d_model * 2);
  // qkv region also hosts the bf16 time-major features of the stem (2*n_ctx x 128 <= n_ctx x 3d)
  w.qkv = o; o += up(std::max(rows * 3 * c.d_model * 2, rows * 2 * 128 * 2));
  // ffn region also hosts conv1's output [B, 2*n_ctx, d] (ffn >= 2d always holds for Whisper; guarded in create)
  w.ffn = o; o += up(std::max(rows * c.ffn_dim * 2, rows * 2 * c.d_model * 2));
  w.st0 = w.st1 = o;
  if (fuse_ln) {  // two sets of per-row LayerNorm partials (d/64 (mean, M2) pairs per row)
    const size_t parts = static_cast<size_t>(c.d_model / 64);
    w.st0 = o; o += up(rows * parts * 8);
    w.st1 = o; o += up(rows * parts * 8);
  }
  w.total = o;
  return w;
}

}  // namespace

extern "C" {

int ttasr_encoder_create(const ttasr_encoder_cfg* cfg, const ttasr_weights* w, ttasr_encoder_t** out) {
  return ttasr_encoder_create_ex(cfg, w, TTASR_RESIDUAL_AUTO, out);
}

int ttasr_encoder_create_ex(const ttasr_encoder_cfg* cfg, const ttasr_weights* w, int residual, ttasr_encoder_t** out) {
  if (!cfg || !w || !out || !w->layers) return fail(TTASR_E_ARG, "encoder_create: null argument");
  *out = nullptr;
  if (residual == TTASR_RESIDUAL_AUTO) {
    residual = TTASR_RESIDUAL_DEFAULT;
    if (const char* env = getenv("TTASR_RESIDUAL")) {
      if (!strcmp(env, "f32")) residual = TTASR_RESIDUAL_F32;
      else if (!strcmp(env, "split")) residual = TTASR_RESIDUAL_SPLIT;
      else if (!strcmp(env, "bf16")) residual = TTASR_RESIDUAL_BF16;
      else return fail(TTASR_E_ARG, "encoder_create: TTASR_RESIDUAL must be f32, split or bf16 (got '%s')", env);
    } else if (const char* env2 = getenv("TTASR_FUSE_LN")) {   // round-1 switch, kept: 0 selects the fp32 stream
      residual = atoi(env2) != 0 ? TTASR_RESIDUAL_SPLIT : TTASR_RESIDUAL_F32;
    }
  }
  if (residual != TTASR_RESIDUAL_F32 && residual != TTASR_RESIDUAL_SPLIT && residual != TTASR_RESIDUAL_BF16)
    return fail(TTASR_E_ARG, "encoder_create: bad residual mode %d", residual);
  const int d = cfg->d_model, f = cfg->ffn_dim, L = cfg->n_layers;
  if (d <= 0 || d % 128 != 0 || d > 1280) return fail(TTASR_E_SHAPE, "encoder_create: d_model must be a multiple of 128, <= 1280 (got %d)", d);
  if (cfg->n_heads <= 0 || d != cfg->n_heads * 64) return fail(TTASR_E_SHAPE, "encoder_create: head_dim must be 64 (d_model %d, heads %d)", d, cfg->n_heads);
  if (f <= 0 || f % 128 != 0) return fail(TTASR_E_SHAPE, "encoder_create: ffn_dim must be a multiple of 128 (got %d)", f);
  if (cfg->n_mels <= 0 || cfg->n_mels > 128 || cfg->n_mels % 8 != 0) return fail(TTASR_E_SHAPE, "encoder_create: n_mels must be a multiple of 8, <= 128 (got %d)", cfg->n_mels);
  if (cfg->n_ctx <= 0 || L <= 0) return fail(TTASR_E_SHAPE, "encoder_create: n_ctx and n_layers must be positive");
  int device = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, &sms);
  if (rc != TTASR_OK) return rc;
  if (!w->conv1_w || !w->conv1_b || !w->conv2_w || !w->conv2_b || !w->pos || !w->ln_post_g || !w->ln_post_b)
    return fail(TTASR_E_ARG, "encoder_create: null stem / final-norm weight");
  for (int i = 0; i < L; ++i) {
    const ttasr_layer_weights& l = w->layers[i];
    if (!l.ln1_g || !l.ln1_b || !l.wq || !l.bq || !l.wk || !l.wv || !l.bv || !l.wo || !l.bo || !l.ln2_g || !l.ln2_b ||
        !l.w1 || !l.b1 || !l.w2 || !l.b2)
      return fail(TTASR_E_ARG, "encoder_create: null weight in layer %d", i);
  }

  ttasr_encoder* h = new ttasr_encoder();
  h->device = device;
  h->num_sms = sms;
  h->cfg = *cfg;
  h->residual = residual;
  h->fuse_ln = residual != TTASR_RESIDUAL_F32;
  const size_t dd = static_cast<size_t>(d) * d;
  size_t bytes = 0;
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  const size_t conv1_bytes = up(static_cast<size_t>(d) * 3 * h->c1_pad * 2), conv2_bytes = up(3 * dd * 2);
  bytes += conv1_bytes + conv2_bytes + 2 * up(d * 4) + up(static_cast<size_t>(cfg->n_ctx) * d * 4) + 2 * up(d * 4) +
           2 * up(static_cast<size_t>(cfg->n_ctx) * d * 2);
  const size_t per_layer = up(3 * dd * 2) + up(dd * 2) + 2 * up(static_cast<size_t>(d) * f * 2) + 4 * up(d * 4) +
                           up(3 * d * 4) + up(d * 4) + up(f * 4) + up(d * 4) + up(3 * d * 4) + up(f * 4);
  bytes += per_layer * L;
  cudaError_t e = cudaMalloc(&h->arena, bytes);
  if (e != cudaSuccess) { delete h; return fail(TTASR_E_NOMEM, "encoder_create: cudaMalloc(%zu) for packed weights: %s", bytes, cudaGetErrorString(e)); }
  char* cur = static_cast<char*>(h->arena);
  auto take = [&](size_t n) { char* p = cur; cur += up(n); return p; };
  auto copy_bf16 = [&](const void* src, size_t n, float scale) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(take(n * 2));
    scale_copy_bf16_kernel<<<blocks_for(n, 256), 256>>>(dst, static_cast<const __nv_bfloat16*>(src), n, scale);
    return dst;
  };
  auto copy_f32 = [&](const float* src, size_t n, float scale) {
    float* dst = reinterpret_cast<float*>(take(n * 4));
    scale_copy_f32_kernel<<<blocks_for(n, 256), 256>>>(dst, src, n, scale);
    return dst;
  };
  {
    __nv_bfloat16* c1 = reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(d) * 3 * h->c1_pad * 2));
    pack_conv_kernel<<<blocks_for(static_cast<long long>(d) * 3 * h->c1_pad, 256), 256>>>(
        c1, static_cast<const __nv_bfloat16*>(w->conv1_w), d, cfg->n_mels, h->c1_pad);
    h->conv1_w = c1;
    __nv_bfloat16* c2 = reinterpret_cast<__nv_bfloat16*>(take(3 * dd * 2));
    pack_conv_kernel<<<blocks_for(3LL * dd, 256), 256>>>(c2, static_cast<const __nv_bfloat16*>(w->conv2_w), d, d, d);
    h->conv2_w = c2;
  }
  h->conv1_b = copy_f32(w->conv1_b, d, 1.f);
  h->conv2_b = copy_f32(w->conv2_b, d, 1.f);
  h->pos = copy_f32(w->pos, static_cast<size_t>(cfg->n_ctx) * d, 1.f);
  if (h->fuse_ln) {  // the positional table as a split pair: it is the "residual" the conv stem's epilogue adds
    const size_t n = static_cast<size_t>(cfg->n_ctx) * d;
    __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(take(n * 2));
    __nv_bfloat16* pl = residual == TTASR_RESIDUAL_SPLIT ? reinterpret_cast<__nv_bfloat16*>(take(n * 2)) : nullptr;
    split_f32_kernel<<<blocks_for(static_cast<long long>(n), 256), 256>>>(ph, pl, w->pos, static_cast<long long>(n));
    h->pos_hi = ph;
    h->pos_lo = pl;
  }
  h->lnp_g = copy_f32(w->ln_post_g, d, 1.f);
  h->lnp_b = copy_f32(w->ln_post_b, d, 1.f);
  h->layers.resize(L);
  const float qscale = 0.125f;  // head_dim^-0.5 with head_dim = 64: exact in bf16
  for (int i = 0; i < L; ++i) {
    const ttasr_layer_weights& l = w->layers[i];
    LayerDev& o = h->layers[i];
    o.ln1_g = copy_f32(l.ln1_g, d, 1.f);
    o.ln1_b = copy_f32(l.ln1_b, d, 1.f);
    __nv_bfloat16* wqkv = reinterpret_cast<__nv_bfloat16*>(take(3 * dd * 2));
    scale_copy_bf16_kernel<<<blocks_for(dd, 256), 256>>>(wqkv, static_cast<const __nv_bfloat16*>(l.wq), dd, qscale);
    scale_copy_bf16_kernel<<<blocks_for(dd, 256), 256>>>(wqkv + dd, static_cast<const __nv_bfloat16*>(l.wk), dd, 1.f);
    scale_copy_bf16_kernel<<<blocks_for(dd, 256), 256>>>(wqkv + 2 * dd, static_cast<const __nv_bfloat16*>(l.wv), dd, 1.f);
    o.wqkv = wqkv;
    float* bqkv = reinterpret_cast<float*>(take(3 * d * 4));
    scale_copy_f32_kernel<<<blocks_for(d, 256), 256>>>(bqkv, l.bq, d, qscale);
    scale_copy_f32_kernel<<<blocks_for(d, 256), 256>>>(bqkv + d, nullptr, d, 0.f);  // k_proj has no bias
    scale_copy_f32_kernel<<<blocks_for(d, 256), 256>>>(bqkv + 2 * d, l.bv, d, 1.f);
    o.bqkv = bqkv;
    o.wo = copy_bf16(l.wo, dd, 1.f);
    o.bo = copy_f32(l.bo, d, 1.f);
    o.ln2_g = copy_f32(l.ln2_g, d, 1.f);
    o.ln2_b = copy_f32(l.ln2_b, d, 1.f);
    o.w1 = copy_bf16(l.w1, static_cast<size_t>(d) * f, 1.f);
    o.b1 = copy_f32(l.b1, f, 1.f);
    o.w2 = copy_bf16(l.w2, static_cast<size_t>(d) * f, 1.f);
    o.b2 = copy_f32(l.b2, d, 1.f);
    if (h->fuse_ln) {
      // fold LN1 into the fused QKV projection and LN2 into fc1, in place of the plain copies made above
      float* c1q = reinterpret_cast<float*>(take(3 * d * 4));
      float* c1f = reinterpret_cast<float*>(take(f * 4));
      auto fold = [&](__nv_bfloat16* dst, const void* src, int n_rows, float scale, const float* g, const float* bta,
                      const float* bias, float* c1, float* c2) {
        fold_ln_kernel<<<blocks_for(static_cast<long long>(n_rows) * 32, 256), 256>>>(
            dst, static_cast<const __nv_bfloat16*>(src), n_rows, d, scale, g, bta, bias, c1, c2);
      };
      fold(wqkv, l.wq, d, qscale, l.ln1_g, l.ln1_b, l.bq, c1q, bqkv);
      fold(wqkv + dd, l.wk, d, 1.f, l.ln1_g, l.ln1_b, nullptr, c1q + d, bqkv + d);
      fold(wqkv + 2 * dd, l.wv, d, 1.f, l.ln1_g, l.ln1_b, l.bv, c1q + 2 * d, bqkv + 2 * d);
      fold(const_cast<__nv_bfloat16*>(o.w1), l.w1, f, 1.f, l.ln2_g, l.ln2_b, l.b1, c1f, const_cast<float*>(o.b1));
      o.c1_qkv = c1q;
      o.c1_fc1 = c1f;
    }
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(h->arena);
    delete h;
    return fail(TTASR_E_CUDA, "encoder_create: packing weights failed: %s", cudaGetErrorString(e));
  }
  if (static_cast<size_t>(cur - static_cast<char*>(h->arena)) > bytes) {
    cudaFree(h->arena);
    delete h;
    return fail(TTASR_E_NOMEM, "encoder_create: internal arena accounting error");
  }
  *out = h;
  return TTASR_OK;
}

int ttasr_encoder_workspace_bytes(const ttasr_encoder_t* h, int64_t batch, size_t* out) {
  if (!h || !out) return fail(TTASR_E_ARG, "encoder_workspace_bytes: null argument");
  if (batch < 0) return fail(TTASR_E_SHAPE, "encoder_workspace_bytes: negative batch");
  *out = ws_layout(h->cfg, batch, h->fuse_ln).total;
  return TTASR_OK;
}

int ttasr_encoder_launch_count(const ttasr_encoder_t* h, int64_t* out) {
  if (!h || !out) return fail(TTASR_E_ARG, "encoder_launch_count: null argument");
  *out = 1 + 2 + (h->fuse_ln ? 5LL : 7LL) * h->cfg.n_layers + 1;
  return TTASR_OK;
}

int ttasr_encoder_forward(const ttasr_encoder_t* h, const void* feats_dev, int feats_layout, int tmajor_ld, int64_t batch,
                          void* workspace_dev, size_t workspace_bytes, void* out_dev, int out_dtype, void* stream_v) {
  if (!h) return fail(TTASR_E_ARG, "encoder_forward: null handle");
  if (batch < 0) return fail(TTASR_E_SHAPE, "encoder_forward: negative batch");
  if (batch == 0) return TTASR_OK;
  if (!feats_dev || !workspace_dev || !out_dev) return fail(TTASR_E_ARG, "encoder_forward: null buffer");
  if (out_dtype != TTASR_OUT_BF16 && out_dtype != TTASR_OUT_F32) return fail(TTASR_E_ARG, "encoder_forward: bad out_dtype %d", out_dtype);
  const ttasr_encoder_cfg& c = h->cfg;
  if (batch * c.n_ctx * 2 > 0x7fffffffLL) return fail(TTASR_E_SHAPE, "encoder_forward: batch %lld too large for one call", (long long)batch);
  const WsLayout ws = ws_layout(c, batch, h->fuse_ln);
  if (workspace_bytes < ws.total) return fail(TTASR_E_NOMEM, "encoder_forward: workspace %zu < required %zu bytes", workspace_bytes, ws.total);
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 1023) return fail(TTASR_E_ARG, "encoder_forward: workspace must be 1024-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* wsb = static_cast<char*>(workspace_dev);
  float* x = reinterpret_cast<float*>(wsb + ws.x);
  __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(wsb + ws.h);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(wsb + ws.qkv);
  __nv_bfloat16* ffn = reinterpret_cast<__nv_bfloat16*>(wsb + ws.ffn);
  const bool fuse = h->fuse_ln;
  const size_t half = (static_cast<size_t>(batch) * c.n_ctx * c.d_model * 2 + 1023) & ~static_cast<size_t>(1023);
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(wsb + ws.x);            // split residual stream: hi ...
  __nv_bfloat16* xl = h->residual == TTASR_RESIDUAL_SPLIT ? reinterpret_cast<__nv_bfloat16*>(wsb + ws.x + half) : nullptr;  // ... lo
  void* st0 = wsb + ws.st0;   // LayerNorm partials of x as the attention block sees it (LN1)
  void* st1 = wsb + ws.st1;   // ... as the MLP block sees it (LN2)
  int parts0 = 0, parts1 = 0;
  const int d = c.d_model, f = c.ffn_dim, T = c.n_ctx, Tin = 2 * c.n_ctx;
  const int B = static_cast<int>(batch);
  const char* why = nullptr;
  cudaError_t e;
  ProfileState& prof = h->prof;
  ProfileState::Span span{};
  auto prof_begin = [&](int kind) {
    nvtxRangePushA(kProfNames[kind]);
    if (!prof.on) return;
    span.kind = kind;
    span.a = prof.get();
    span.b = prof.get();
    cudaEventRecord(span.a, stream);
  };
  auto prof_end = [&]() {
    nvtxRangePop();
    if (!prof.on) return;
    cudaEventRecord(span.b, stream);
    prof.pending.push_back(span);
  };
#define GEMM_TRY(call, name, kind)                                                                                \
  do {                                                                                                            \
    prof_begin(kind);                                                                                             \
    e = gemm_launch(call, h->num_sms, stream, &why);                                                              \
    prof_end();                                                                                                   \
    if (e != cudaSuccess) return fail(TTASR_E_CUDA, "encoder_forward: %s: %s", name, why ? why : cudaGetErrorString(e)); \
  } while (0)

  // ---- stem input as bf16 time-major
  const __nv_bfloat16* ft;
  int ld;
  if (feats_layout == TTASR_FEATS_F32_MEL_MAJOR) {
    ld = c.n_mels;
    __nv_bfloat16* dst = qkv;  // scratch: dead before the first QKV GEMM
    dim3 grid((Tin + 31) / 32, (ld + 31) / 32, B);
    prof_begin(kProfPrep);
    feats_to_tmajor_kernel<<<grid, 256, 0, stream>>>(static_cast<const float*>(feats_dev), dst, c.n_mels, Tin, ld);
    prof_end();
    ft = dst;
  } else if (feats_layout == TTASR_FEATS_BF16_TIME_MAJOR) {
    if (tmajor_ld < c.n_mels || tmajor_ld % 8 != 0) return fail(TTASR_E_SHAPE, "encoder_forward: tmajor_ld must be a multiple of 8 and >= n_mels");
    ld = tmajor_ld;
    ft = static_cast<const __nv_bfloat16*>(feats_dev);
  } else {
    return fail(TTASR_E_ARG, "encoder_forward: bad feats_layout %d", feats_layout);
  }
  __nv_bfloat16* c1 = ffn;  // conv1 output [B, 2T, d], dead before fc1 of layer 0
  {
    GemmCall g;
    g.mode = kGemmConv1;
    g.a = ft; g.lda = ld; g.a_inner = c.n_mels; g.rows = Tin; g.nbatch = B;
    g.w = h->conv1_w; g.n = d; g.kb_per_tap = h->c1_pad / 64; g.k_blocks = 3 * g.kb_per_tap;
    g.bias = h->conv1_b; g.act = 1; g.out = c1; g.out_f32 = 0;
    GEMM_TRY(g, "conv1", kProfConv1);
  }
  {
    GemmCall g;
    g.mode = kGemmConv2;
    g.a = c1; g.lda = d; g.a_inner = d; g.rows = T; g.nbatch = B;
    g.w = h->conv2_w; g.n = d; g.kb_per_tap = d / 64; g.k_blocks = 3 * g.kb_per_tap;
    g.bias = h->conv2_b; g.act = 1; g.addend_bcast = 1;
    if (fuse) {
      g.split = 1; g.addend_hi = h->pos_hi; g.addend_lo = h->pos_lo; g.out = xh; g.out_lo = xl;
      g.ln_stats_out = st0; g.ln_parts_out = &parts0;
    } else {
      g.addend = h->pos; g.out = x; g.out_f32 = 1;
    }
    GEMM_TRY(g, "conv2", kProfConv2);
  }
  const long long M = static_cast<long long>(B) * T;
  // residual GEMM: x += a W^T + b, on whichever representation of x this handle uses
  auto residual_gemm = [&](GemmCall& g, void* stats, int* parts) {
    if (fuse) {
      g.split = 1; g.addend_hi = xh; g.addend_lo = xl; g.out = xh; g.out_lo = xl;
      g.ln_stats_out = stats; g.ln_parts_out = parts;
    } else {
      g.addend = x; g.out = x; g.out_f32 = 1;
    }
  };
  for (int i = 0; i < c.n_layers; ++i) {
    const LayerDev& l = h->layers[i];
    if (!fuse) {
      prof_begin(kProfLayerNorm);
      e = layernorm_launch(x, l.ln1_g, l.ln1_b, hbuf, M, d, 0, stream);
      prof_end();
      if (e != cudaSuccess) return fail(TTASR_E_CUDA, "encoder_forward: layer %d ln1: %s", i, cudaGetErrorString(e));
    }
    {
      GemmCall g;
      g.a = fuse ? xh : hbuf; g.lda = d; g.a_inner = d; g.rows = static_cast<int>(M); g.nbatch = 1;
      g.w = l.wqkv; g.n = 3 * d; g.k_blocks = d / 64; g.kb_per_tap = g.k_blocks;
      g.bias = l.bqkv; g.out = qkv;
      if (fuse) { g.ln_stats_in = st0; g.ln_parts_in = parts0; g.ln_c1 = l.c1_qkv; }
      GEMM_TRY(g, "qkv", kProfQkv);
    }
    prof_begin(kProfAttention);
    e = attention_launch(qkv, hbuf, B, T, c.n_heads, h->num_sms, stream, &why);
    prof_end();
    if (e != cudaSuccess) return fail(TTASR_E_CUDA, "encoder_forward: layer %d attention: %s", i, why ? why : cudaGetErrorString(e));
    {
      GemmCall g;
      g.a = hbuf; g.lda = d; g.a_inner = d; g.rows = static_cast<int>(M); g.nbatch = 1;
      g.w = l.wo; g.n = d; g.k_blocks = d / 64; g.kb_per_tap = g.k_blocks;
      g.bias = l.bo;
      residual_gemm(g, st1, &parts1);
      GEMM_TRY(g, "out_proj", kProfOutProj);
    }
    if (!fuse) {
      prof_begin(kProfLayerNorm);
      e = layernorm_launch(x, l.ln2_g, l.ln2_b, hbuf, M, d, 0, stream);
      prof_end();
      if (e != cudaSuccess) return fail(TTASR_E_CUDA, "encoder_forward: layer %d ln2: %s", i, cudaGetErrorString(e));
    }
    {
      GemmCall g;
      g.a = fuse ? xh : hbuf; g.lda = d; g.a_inner = d; g.rows = static_cast<int>(M); g.nbatch = 1;
      g.w = l.w1; g.n = f; g.k_blocks = d / 64; g.kb_per_tap = g.k_blocks;
      g.bias = l.b1; g.act = 1; g.out = ffn;
      if (fuse) { g.ln_stats_in = st1; g.ln_parts_in = parts1; g.ln_c1 = l.c1_fc1; }
      GEMM_TRY(g, "fc1", kProfFc1);
    }
    {
      GemmCall g;
      g.a = ffn; g.lda = f; g.a_inner = f; g.rows = static_cast<int>(M); g.nbatch = 1;
      g.w = l.w2; g.n = d; g.k_blocks = f / 64; g.kb_per_tap = g.k_blocks;
      g.bias = l.b2;
      residual_gemm(g, (i + 1 < c.n_layers) ? st0 : nullptr, &parts0);  // the last block feeds the final LayerNorm kernel
      GEMM_TRY(g, "fc2", kProfFc2);
    }
  }
  prof_begin(kProfLayerNorm);
  e = fuse ? layernorm_split_launch(xh, xl, h->lnp_g, h->lnp_b, out_dev, M, d, out_dtype == TTASR_OUT_F32, stream)
           : layernorm_launch(x, h->lnp_g, h->lnp_b, out_dev, M, d, out_dtype == TTASR_OUT_F32, stream);
  prof_end();
  if (e != cudaSuccess) return fail(TTASR_E_CUDA, "encoder_forward: final layer norm: %s", cudaGetErrorString(e));
#undef GEMM_TRY
  return TTASR_OK;
}

int ttasr_encoder_profile_enable(ttasr_encoder_t* h, int on) {
  if (!h) return fail(TTASR_E_ARG, "encoder_profile_enable: null handle");
  h->prof.on = on != 0;
  return TTASR_OK;
}

int ttasr_encoder_profile_read(ttasr_encoder_t* h, double* ms_by_kind, int64_t* launches_by_kind, int reset) {
  if (!h || !ms_by_kind || !launches_by_kind) return fail(TTASR_E_ARG, "encoder_profile_read: null argument");
  ProfileState& p = h->prof;
  for (auto& sp : p.pending) {
    cudaError_t e = cudaEventSynchronize(sp.b);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, sp.a, sp.b);
    if (e != cudaSuccess) return fail(TTASR_E_CUDA, "encoder_profile_read: %s", cudaGetErrorString(e));
    p.ms[sp.kind] += ms;
    p.launches[sp.kind] += 1;
    p.pool.push_back(sp.a);
    p.pool.push_back(sp.b);
  }
  p.pending.clear();
  for (int k = 0; k < kProfKinds; ++k) {
    ms_by_kind[k] = p.ms[k];
    launches_by_kind[k] = p.launches[k];
    if (reset) { p.ms[k] = 0; p.launches[k] = 0; }
  }
  return TTASR_OK;
}

const char* ttasr_encoder_profile_kind_name(int kind) {
  return (kind >= 0 && kind < kProfKinds) ? kProfNames[kind] : "";
}

void ttasr_encoder_destroy(ttasr_encoder_t* h) {
  if (!h) return;
  for (auto& sp : h->prof.pending) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : h->prof.pool) cudaEventDestroy(e);
  cudaFree(h->arena);
  delete h;
}

// =================================================================================================== single ops
int ttasr_op_gemm(const void* a_dev, const void* w_dev, const float* bias_dev, const float* addend_dev, void* out_dev,
                  int64_t M, int64_t N, int64_t K, int act, int out_dtype, int cta_group, void* stream) {
  if (!a_dev || !w_dev || !out_dev) return fail(TTASR_E_ARG, "op_gemm: null buffer");
  if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0 || N % 128 != 0 || M > 0x7fffffff)
    return fail(TTASR_E_SHAPE, "op_gemm: need M > 0, N %% 128 == 0, K %% 64 == 0 (got %lld, %lld, %lld)", (long long)M, (long long)N, (long long)K);
  int device = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, &sms);
  if (rc != TTASR_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* zero_bias = nullptr;  // stream-ordered scratch, freed right behind the launch
  if (!bias_dev) {
    CUDA_TRY(cudaMallocAsync(&zero_bias, sizeof(float) * N, st));
    CUDA_TRY(cudaMemsetAsync(zero_bias, 0, sizeof(float) * N, st));
    bias_dev = zero_bias;
  }
  GemmCall g;
  g.a = a_dev; g.lda = K; g.a_inner = static_cast<int>(K); g.rows = static_cast<int>(M); g.nbatch = 1;
  g.w = w_dev; g.n = static_cast<int>(N); g.k_blocks = static_cast<int>(K / 64); g.kb_per_tap = g.k_blocks;
  g.bias = bias_dev; g.addend = addend_dev; g.act = act; g.out = out_dev; g.out_f32 = (out_dtype == TTASR_OUT_F32);
  g.cta_group = cta_group;
  const char* why = nullptr;
  cudaError_t e = gemm_launch(g, sms, st, &why);
  if (zero_bias) cudaFreeAsync(zero_bias, st);
  if (e != cudaSuccess) return fail(why ? TTASR_E_ARG : TTASR_E_CUDA, "op_gemm: %s", why ? why : cudaGetErrorString(e));
  return TTASR_OK;
}

int ttasr_op_gemm_split(const void* a_dev, const void* w_dev, const float* bias_dev, const void* addh_dev,
                        const void* addl_dev, void* outh_dev, void* outl_dev, void* stats_out_dev, int64_t M, int64_t N,
                        int64_t K, int act, int cta_group, void* stream) {
  if (!a_dev || !w_dev || !bias_dev || !addh_dev || !outh_dev) return fail(TTASR_E_ARG, "op_gemm_split: null buffer");
  if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0 || N % 128 != 0 || M > 0x7fffffff)
    return fail(TTASR_E_SHAPE, "op_gemm_split: need M > 0, N %% 128 == 0, K %% 64 == 0 (got %lld, %lld, %lld)", (long long)M, (long long)N, (long long)K);
  int device = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, &sms);
  if (rc != TTASR_OK) return rc;
  GemmCall g;
  g.a = a_dev; g.lda = K; g.a_inner = static_cast<int>(K); g.rows = static_cast<int>(M); g.nbatch = 1;
  g.w = w_dev; g.n = static_cast<int>(N); g.k_blocks = static_cast<int>(K / 64); g.kb_per_tap = g.k_blocks;
  g.bias = bias_dev; g.act = act; g.split = 1; g.addend_hi = addh_dev; g.addend_lo = addl_dev;
  g.out = outh_dev; g.out_lo = outl_dev; g.ln_stats_out = stats_out_dev; g.cta_group = cta_group;
  const char* why = nullptr;
  cudaError_t e = gemm_launch(g, sms, static_cast<cudaStream_t>(stream), &why);
  if (e != cudaSuccess) return fail(why ? TTASR_E_ARG : TTASR_E_CUDA, "op_gemm_split: %s", why ? why : cudaGetErrorString(e));
  return TTASR_OK;
}

int ttasr_op_gemm_lnfold(const void* a_dev, const void* w_dev, const float* c1_dev, const float* c2_dev,
                         const void* stats_in_dev, int parts, void* out_dev, int64_t M, int64_t N, int64_t K, int act,
                         float eps, int cta_group, void* stream) {
  if (!a_dev || !w_dev || !c1_dev || !c2_dev || !stats_in_dev || !out_dev) return fail(TTASR_E_ARG, "op_gemm_lnfold: null buffer");
  if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0 || N % 128 != 0 || M > 0x7fffffff || parts <= 0 || K % parts != 0)
    return fail(TTASR_E_SHAPE, "op_gemm_lnfold: need M > 0, N %% 128 == 0, K %% 64 == 0, parts | K (got %lld, %lld, %lld, %d)", (long long)M, (long long)N, (long long)K, parts);
  int device = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, &sms);
  if (rc != TTASR_OK) return rc;
  GemmCall g;
  g.a = a_dev; g.lda = K; g.a_inner = static_cast<int>(K); g.rows = static_cast<int>(M); g.nbatch = 1;
  g.w = w_dev; g.n = static_cast<int>(N); g.k_blocks = static_cast<int>(K / 64); g.kb_per_tap = g.k_blocks;
  g.bias = c2_dev; g.act = act; g.out = out_dev; g.ln_stats_in = stats_in_dev; g.ln_parts_in = parts; g.ln_c1 = c1_dev;
  g.ln_eps = eps; g.cta_group = cta_group;
  const char* why = nullptr;
  cudaError_t e = gemm_launch(g, sms, static_cast<cudaStream_t>(stream), &why);
  if (e != cudaSuccess) return fail(why ? TTASR_E_ARG : TTASR_E_CUDA, "op_gemm_lnfold: %s", why ? why : cudaGetErrorString(e));
  return TTASR_OK;
}

int ttasr_op_conv_stem(const void* feats_tm_dev, int ld, int n_mels, int64_t B, int T, int d, const void* conv1_w_dev,
                       const float* conv1_b_dev, const void* conv2_w_dev, const float* conv2_b_dev, const float* pos_dev,
                       void* scratch_dev, float* out_dev, void* stream) {
  if (!feats_tm_dev || !conv1_w_dev || !conv1_b_dev || !conv2_w_dev || !conv2_b_dev || !pos_dev || !scratch_dev || !out_dev)
    return fail(TTASR_E_ARG, "op_conv_stem: null buffer");
  if (B <= 0 || T <= 0 || d <= 0 || d % 128 != 0 || n_mels <= 0 || n_mels > 128 || ld < n_mels || ld % 8 != 0)
    return fail(TTASR_E_SHAPE, "op_conv_stem: need d %% 128 == 0, n_mels <= 128 <= ld-compatible, ld %% 8 == 0");
  int device = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, &sms);
  if (rc != TTASR_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int c1_pad = 128;
  __nv_bfloat16 *w1 = nullptr, *w2 = nullptr;
  CUDA_TRY(cudaMallocAsync(&w1, static_cast<size_t>(d) * 3 * c1_pad * 2, st));
  cudaError_t e = cudaMallocAsync(&w2, static_cast<size_t>(d) * 3 * d * 2, st);
  if (e != cudaSuccess) { cudaFreeAsync(w1, st); return fail(TTASR_E_NOMEM, "op_conv_stem: %s", cudaGetErrorString(e)); }
  pack_conv_kernel<<<blocks_for(static_cast<long long>(d) * 3 * c1_pad, 256), 256, 0, st>>>(
      w1, static_cast<const __nv_bfloat16*>(conv1_w_dev), d, n_mels, c1_pad);
  pack_conv_kernel<<<blocks_for(3LL * d * d, 256), 256, 0, st>>>(w2, static_cast<const __nv_bfloat16*>(conv2_w_dev), d, d, d);
  const char* why = nullptr;
  {
    GemmCall g;
    g.mode = kGemmConv1;
    g.a = feats_tm_dev; g.lda = ld; g.a_inner = n_mels; g.rows = 2 * T; g.nbatch = static_cast<int>(B);
    g.w = w1; g.n = d; g.kb_per_tap = c1_pad / 64; g.k_blocks = 3 * g.kb_per_tap;
    g.bias = conv1_b_dev; g.act = 1; g.out = scratch_dev; g.out_f32 = 0;
    e = gemm_launch(g, sms, st, &why);
  }
  if (e == cudaSuccess) {
    GemmCall g;
    g.mode = kGemmConv2;
    g.a = scratch_dev; g.lda = d; g.a_inner = d; g.rows = T; g.nbatch = static_cast<int>(B);
    g.w = w2; g.n = d; g.kb_per_tap = d / 64; g.k_blocks = 3 * g.kb_per_tap;
    g.bias = conv2_b_dev; g.act = 1; g.addend = pos_dev; g.addend_bcast = 1; g.out = out_dev; g.out_f32 = 1;
    e = gemm_launch(g, sms, st, &why);
  }
  cudaFreeAsync(w1, st);
  cudaFreeAsync(w2, st);
  if (e != cudaSuccess) return fail(why ? TTASR_E_ARG : TTASR_E_CUDA, "op_conv_stem: %s", why ? why : cudaGetErrorString(e));
  return TTASR_OK;
}

int ttasr_op_layernorm(const float* x_dev, const float* g_dev, const float* b_dev, void* y_dev, int64_t rows, int d,
                       int out_dtype, void* stream) {
  if (!x_dev || !g_dev || !b_dev || !y_dev) return fail(TTASR_E_ARG, "op_layernorm: null buffer");
  if (d <= 0 || d % 128 != 0 || d > 1280) return fail(TTASR_E_SHAPE, "op_layernorm: d must be a multiple of 128, <= 1280");
  cudaError_t e = layernorm_launch(x_dev, g_dev, b_dev, y_dev, rows, d, out_dtype == TTASR_OUT_F32, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(TTASR_E_CUDA, "op_layernorm: %s", cudaGetErrorString(e));
  return TTASR_OK;
}

int ttasr_op_attention(const void* qkv_dev, void* out_dev, int64_t batch, int n_ctx, int n_heads, void* stream) {
  if (!qkv_dev || !out_dev) return fail(TTASR_E_ARG, "op_attention: null buffer");
  if (batch <= 0 || n_ctx <= 0 || n_heads <= 0) return fail(TTASR_E_SHAPE, "op_attention: empty problem");
  int device = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&device));
  int rc = check_arch(device, &sms);
  if (rc != TTASR_OK) return rc;
  const char* why = nullptr;
  cudaError_t e = attention_launch(qkv_dev, out_dev, static_cast<int>(batch), n_ctx, n_heads, sms, static_cast<cudaStream_t>(stream), &why);
  if (e != cudaSuccess) return fail(why ? TTASR_E_ARG : TTASR_E_CUDA, "op_attention: %s", why ? why : cudaGetErrorString(e));
  return TTASR_OK;
}

}  // extern "C"
