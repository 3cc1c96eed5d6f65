// Internal interface of the fused log-mel front end (see frontend_logmel.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace ttasr {

constexpr int kMaxMels = 128;
constexpr int kMelWarps = 10;    // warps of the frames kernel (320 threads); each owns a contiguous range of filters
constexpr int kMaxMelOps = 512;  // entries of the streaming mel program (one per frequency bin walked, per warp)
constexpr int kPowerPitchBytes = 33 * 4;  // row pitch of the [bin][frame] power-spectrum buffer the program indexes

// device-resident constant tables owned by the front-end handle
struct FrontTables {
  const float2* twiddle;  // [400]  W400^(n2*k1) = (cos, -sin)(2 pi n2 k1 / 400) at [k1*20 + n2]
  const float* window;    // [400]  periodic Hann
};

// The mel projection as a streaming program (built on the host at ttasr_frontend_create).  It travels as a KERNEL
// PARAMETER: the ops are warp-uniform, so the kernel reads them through the constant bank (LDC / uniform datapath)
// and the load/store unit — the busiest pipe of the frames kernel — only carries the power-spectrum reads.
struct MelProgram {
  int4 ops[kMaxMelOps];        // {bin * kPowerPitchBytes, weight of filter m_cur, of m_cur + 1, 1 = emit a filter}
  int op_off[kMelWarps + 1];   // warp w runs ops [op_off[w], op_off[w + 1])
  int m0[kMelWarps + 1];       // warp w owns filters [m0[w], m0[w + 1])
};

size_t frontend_smem_bytes();

// feats: [B, n_mels, n_samples/160] fp32 (HF layout).  tmajor (optional): [B, T, tmajor_ld] bf16, zero-padded channels.
// n_valid (optional, device): samples >= n_valid[b] are treated as zero and never read.
// clamp_decades: the reference's max(x, x.max() - 8) range in log10 units; +infinity disables the clamp (the features
// are then plain (log10(max(mel, 1e-10)) + 4) / 4 and the caller applies its own, e.g. faster-whisper's per-file maximum).
cudaError_t launch_logmel(const void* pcm, int pcm_is_i16, long long row_stride, const int* n_valid, int n_samples,
                          int n_mels, int batch, const FrontTables& tables, const MelProgram& mel, float* feats, unsigned* chunk_max,
                          float* tile_min, __nv_bfloat16* tmajor, int tmajor_ld, int num_sms, cudaStream_t stream,
                          float clamp_decades = 8.0f, int mel_baked = 0);
// FNV-1a 64 hash of the fp32 [201, n_mels] filter table whose projection is compiled into the kernel (mel_baked.inc:
// the two Whisper banks), 0 if there is none for this n_mels; launch_logmel(mel_baked = n_mels) selects that code.
unsigned long long frontend_baked_hash(int n_mels);
// scratch the caller provides: chunk_max[batch] (order-encoded running maxima), tile_min[batch * frontend_tiles(n_samples)]
int frontend_tiles(int n_samples);

}  // namespace ttasr
