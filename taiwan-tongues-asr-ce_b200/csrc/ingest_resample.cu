// Ingest step in front of the log-mel path: decoded PCM frames -> mono -> 16 kHz, written
// straight into the zero-padded [n_chunks, 480000] fp32 layout the front end consumes.
//
// Reference semantics: librosa.load(path, sr=16000, mono=True) as called at asr_core.py:156 and api/file_asr.py:271
// (librosa is a third-party dependency that is not on disk, requirements.txt:3 "librosa>=0.9.0"): samples to float32
// (int16 / 32768), librosa.to_mono = mean over channels, then librosa.resample.  The resampler implemented here is
// librosa's res_type="polyphase" branch, i.e. scipy.signal.resample_poly(y, up, down) with its default Kaiser(5.0)
// windowed-sinc prototype of 2 * 10 * max(up, down) + 1 taps: y[n] = sum_k hpad[(n + n_pre_remove) * down - k * up] x[k].
// (librosa's default res_type, "soxr_hq", is a different low-pass design; see INTEGRATION.md.)
//
// HBM-bound byte work: one thread per output sample, the <= ceil(n_taps / up) taps of its phase are contiguous in a
// polyphase table (L1-resident, <= 36 KB), the input window is contiguous and shared by neighbouring threads
// through L1; every input byte is fetched from HBM once.
#include "ingest_resample.h"

namespace ttasr {
namespace {

template <typename T>
__device__ __forceinline__ float frame_mono(const T* __restrict__ pcm, long long k, int channels);
template <>
__device__ __forceinline__ float frame_mono<float>(const float* __restrict__ pcm, long long k, int channels) {
  if (channels == 1) return __ldg(pcm + k);
  if (channels == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(pcm) + k);
    return (v.x + v.y) * 0.5f;
  }
  float s = 0.f;
  for (int c = 0; c < channels; ++c) s += __ldg(pcm + k * channels + c);
  return s / static_cast<float>(channels);
}
template <>
__device__ __forceinline__ float frame_mono<int16_t>(const int16_t* __restrict__ pcm, long long k, int channels) {
  constexpr float kScale = 1.0f / 32768.0f;
  if (channels == 1) return static_cast<float>(__ldg(pcm + k)) * kScale;
  if (channels == 2) {
    const short2 v = __ldg(reinterpret_cast<const short2*>(pcm) + k);
    return (static_cast<float>(v.x) * kScale + static_cast<float>(v.y) * kScale) * 0.5f;
  }
  float s = 0.f;
  for (int c = 0; c < channels; ++c) s += static_cast<float>(__ldg(pcm + k * channels + c)) * kScale;
  return s / static_cast<float>(channels);
}

template <typename T>
__global__ void __launch_bounds__(256) ingest_kernel(IngestPlan plan, const T* __restrict__ pcm, int channels,
                                                     long long n_in, float* __restrict__ out, long long n_out,
                                                     long long out_capacity) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; n < out_capacity; n += stride) {
    float acc = 0.f;
    if (n < n_out) {
      if (plan.up == 1 && plan.down == 1) {
        acc = frame_mono<T>(pcm, n, channels);
      } else {
        // index into the unpadded prototype of the tap that meets input frame k = 0
        const long long j0 = (n + plan.n_pre_remove) * plan.down - plan.n_pre_pad;
        long long q = j0 / plan.up;           // floor division (j0 may be negative for the first outputs)
        int p = static_cast<int>(j0 - q * plan.up);
        if (p < 0) { p += plan.up; --q; }
        // term i uses h[p + i * up] and input frame k = q - i
        const float* taps = plan.phase_taps + static_cast<long long>(p) * plan.taps_per_phase;
        int i_lo = 0, i_hi = plan.taps_per_phase;              // [i_lo, i_hi)
        if (q >= n_in) i_lo = static_cast<int>(min(static_cast<long long>(i_hi), q - (n_in - 1)));
        if (q - (i_hi - 1) < 0) i_hi = static_cast<int>(max(static_cast<long long>(i_lo), q + 1));
        for (int i = i_lo; i < i_hi; ++i) acc = fmaf(__ldg(taps + i), frame_mono<T>(pcm, q - i, channels), acc);
      }
    }
    out[n] = acc;
  }
}

// ---- energy track for VAD-aligned chunk cuts: one warp per 10 ms hop, mean square of a 20 ms window, in dB
__global__ void __launch_bounds__(256) frame_energy_kernel(const float* __restrict__ x, long long n, int win, int hop,
                                                           float* __restrict__ db, long long n_frames) {
  const long long f = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= n_frames) return;
  const long long start = f * hop;
  float s = 0.f;
  for (int i = lane; i < win; i += 32) {
    const long long k = start + i;
    const float v = k < n ? __ldg(x + k) : 0.f;
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) db[f] = 10.0f * log10f(s / static_cast<float>(win) + 1e-12f);
}

// ---- ragged rows: out[r, 0 .. len[r]) = x[start[r] .. start[r] + len[r]), zero tail
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ x, const long long* __restrict__ start,
                                                          const int* __restrict__ len, float* __restrict__ out,
                                                          int row_samples) {
  const int r = blockIdx.y;
  const long long s0 = start[r];
  const int n = len[r];
  float* dst = out + static_cast<long long>(r) * row_samples;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_samples; i += gridDim.x * blockDim.x)
    dst[i] = i < n ? __ldg(x + s0 + i) : 0.f;
}

}  // namespace

cudaError_t launch_frame_energy(const float* x, long long n, int win, int hop, float* db, long long n_frames,
                                cudaStream_t stream) {
  if (n_frames <= 0) return cudaSuccess;
  const long long blocks = (n_frames * 32 + 255) / 256;
  frame_energy_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, n, win, hop, db, n_frames);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* x, const long long* start, const int* len, float* out, int n_rows,
                               int row_samples, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  dim3 grid(64, n_rows);
  gather_rows_kernel<<<grid, 256, 0, stream>>>(x, start, len, out, row_samples);
  return cudaGetLastError();
}

cudaError_t launch_ingest(const IngestPlan& plan, const void* pcm, int pcm_is_i16, int channels, long long n_in,
                          float* out, long long n_out, long long out_capacity, cudaStream_t stream) {
  if (out_capacity <= 0) return cudaSuccess;
  const long long blocks = (out_capacity + 255) / 256;
  const int grid = static_cast<int>(blocks < 148LL * 32 ? blocks : 148LL * 32);
  if (pcm_is_i16)
    ingest_kernel<int16_t><<<grid, 256, 0, stream>>>(plan, static_cast<const int16_t*>(pcm), channels, n_in, out, n_out,
                                                     out_capacity);
  else
    ingest_kernel<float><<<grid, 256, 0, stream>>>(plan, static_cast<const float*>(pcm), channels, n_in, out, n_out,
                                                   out_capacity);
  return cudaGetLastError();
}

}  // namespace ttasr
