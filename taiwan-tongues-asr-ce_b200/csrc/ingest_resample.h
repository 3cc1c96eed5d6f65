// Internal interface of the ingest kernel (see ingest_resample.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttasr {

// Polyphase tables of a rational resampler, device resident (owned by the ingest handle).
struct IngestPlan {
  int up = 1, down = 1;
  int n_taps = 0;          // taps of the prototype filter h (already multiplied by `up`)
  int n_pre_pad = 0;       // scipy.signal.resample_poly: zeros prepended to h
  int n_pre_remove = 0;    // scipy.signal.resample_poly: leading outputs of upfirdn dropped
  int taps_per_phase = 0;  // ceil(n_taps / up)
  const float* phase_taps = nullptr;  // [up][taps_per_phase]: phase_taps[p][i] = h[p + i * up] (0 beyond n_taps)
};

// pcm: interleaved frames [n_in][channels], int16 (scaled by 1/32768) or float32.  out[0 .. n_out) = resampled mono
// signal, out[n_out .. out_capacity) = 0.
cudaError_t launch_ingest(const IngestPlan& plan, const void* pcm, int pcm_is_i16, int channels, long long n_in,
                          float* out, long long n_out, long long out_capacity, cudaStream_t stream);

// db[f] = 10 log10(mean(x[f*hop .. f*hop + win)^2) + 1e-12), samples past n read as 0
cudaError_t launch_frame_energy(const float* x, long long n, int win, int hop, float* db, long long n_frames,
                                cudaStream_t stream);
// out[r, :] = x[start[r] .. start[r] + len[r]) followed by zeros up to row_samples
cudaError_t launch_gather_rows(const float* x, const long long* start, const int* len, float* out, int n_rows,
                               int row_samples, cudaStream_t stream);

}  // namespace ttasr
