/* ttasr_abi.h — C ABI of the B200-native Whisper log-mel front end + encoder (libttasr_b200.so).
 *
 * This is the drop-in boundary beneath the reference's Python plugin surface.  The reference repository
 * (adi-gov-tw/Taiwan-Tongues-ASR-CE) has no FFI of its own: its hot path runs inside third-party packages, so each
 * entry point cites the Python call it replaces:
 *
 *   ttasr_frontend_*   <- WhisperFeatureExtractor.__call__/_np_extract_fbank_features
 *                         (transformers/models/whisper/feature_extraction_whisper.py:105-133,189-342),
 *                         called by the reference at train_asr.py:607-616, and the equivalent
 *                         faster_whisper.FeatureExtractor behind asr_core.py:159-167, api/file_asr.py:457-465,
 *                         api/stt_streaming/src/asr/faster_whisper_asr.py:170-172.
 *   ttasr_encoder_*    <- WhisperEncoder.forward (transformers/models/whisper/modeling_whisper.py:593-647) reached
 *                         from train_asr.py:697-716,736-740, and faster_whisper.WhisperModel.encode behind the
 *                         three transcribe() call sites above.
 *
 * Conventions
 *   - plain pointers and sizes only; every *_dev pointer is CUDA device memory owned by the caller.
 *   - every function returns a status (0 = ok, negative = TTASR_E_*), never throws; ttasr_last_error() returns a
 *     thread-local message for the last failing call on this thread.
 *   - all device work is ordered on the `stream` argument (a cudaStream_t passed as void*); calls are asynchronous.
 *   - handles are immutable after create, usable from any thread, one per device (the device current at create).
 *     Exception: a front-end handle owns the per-call scratch of its kernels (per-chunk maxima, per-tile minima), so
 *     calls of ttasr_frontend_run on ONE handle must be ordered (same stream, or serialised by the caller); use one
 *     handle per concurrent stream.  Encoder calls take their workspace from the caller and may run concurrently
 *     with distinct workspaces.  ttasr_ingest_* keeps no per-call state.
 *   - create fails with TTASR_E_ARCH on anything but compute capability 10.x: there is no fallback path.
 */
#ifndef TTASR_ABI_H_
#define TTASR_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTASR_ABI_VERSION 1

#if defined(__GNUC__)
#define TTASR_API __attribute__((visibility("default")))
#else
#define TTASR_API
#endif

enum {
  TTASR_OK = 0,
  TTASR_E_ARG = -1,   /* null pointer, bad enum, unsupported constant (n_fft != 400, hop != 160, ...) */
  TTASR_E_SHAPE = -2, /* shape the kernels cannot take (d_model % 128, head_dim != 64, batch < 0, ...) */
  TTASR_E_ARCH = -3,  /* not an sm_100 device */
  TTASR_E_CUDA = -4,  /* a CUDA runtime/driver call failed; message has the CUDA error string */
  TTASR_E_NOMEM = -5  /* workspace too small / allocation failed */
};

enum { TTASR_PCM_F32 = 0, TTASR_PCM_I16 = 1 };
enum { TTASR_FEATS_F32_MEL_MAJOR = 0, /* [B, n_mels, 3000] fp32, the HF `input_features` layout */
       TTASR_FEATS_BF16_TIME_MAJOR = 1 /* [B, 3000, ld] bf16 as written by ttasr_frontend_run */ };
enum { TTASR_OUT_BF16 = 0, TTASR_OUT_F32 = 1 };

typedef struct ttasr_frontend ttasr_frontend_t;
typedef struct ttasr_encoder ttasr_encoder_t;
typedef struct ttasr_ingest ttasr_ingest_t;

TTASR_API int ttasr_abi_version(void);
TTASR_API const char* ttasr_last_error(void);
/* TTASR_OK iff `device` is a compute-capability-10.x GPU. */
TTASR_API int ttasr_device_check(int device);

/* ------------------------------------------------------------------ log-mel front end ------------------------
 * mel_filters: host, [201, n_mels] fp32 row-major (HF `mel_filters`, audio_utils.py:453-544).
 * window:      host, [400] fp32 periodic Hann (audio_utils.py:560-620).
 * n_fft must be 400, hop 160, n_samples a multiple of 160 (480000 for Whisper). */
TTASR_API int ttasr_frontend_create(int n_mels, int n_fft, int hop, int n_samples, const float* mel_filters,
                          const float* window, ttasr_frontend_t** out);
/* pcm_dev:   [batch, row_stride] samples (fp32 in [-1,1], or int16 scaled by 1/32768 in-kernel).
 * n_valid_dev: optional int32[batch]; samples at index >= n_valid[b] are taken as 0.0 and never read
 *            (right-padding to 30 s without materialising it); NULL = every row holds n_samples samples.
 * feats_dev: [batch, n_mels, n_samples/160] fp32 (the HF `input_features`), or NULL when only tmajor_dev is wanted
 *            (the PCM -> hidden-state pipeline: nothing downstream reads the fp32 features).
 * tmajor_dev: optional (required if feats_dev is NULL) [batch, n_samples/160, tmajor_ld] bf16 copy (channels zero-padded to tmajor_ld, which must
 *            be even and >= n_mels) for ttasr_encoder_forward(TTASR_FEATS_BF16_TIME_MAJOR). */
TTASR_API int ttasr_frontend_run(const ttasr_frontend_t* h, const void* pcm_dev, int pcm_dtype, int64_t batch,
                       int64_t row_stride, const int32_t* n_valid_dev, float* feats_dev, void* tmajor_dev,
                       int tmajor_ld, void* stream);
/* The same with the dynamic-range clamp as a parameter: the reference's `max(x, x.max() - 8.0)` per chunk
 * (feature_extraction_whisper.py:129) is clamp_decades = 8; +INFINITY writes the unclamped (log10(max(mel, 1e-10)) + 4) / 4
 * so that a caller with a different maximum (faster-whisper takes it over the whole file) can apply
 * its own. */
TTASR_API int ttasr_frontend_run_ex(const ttasr_frontend_t* h, const void* pcm_dev, int pcm_dtype, int64_t batch,
                          int64_t row_stride, const int32_t* n_valid_dev, float* feats_dev, void* tmajor_dev,
                          int tmajor_ld, float clamp_decades, void* stream);
/* largest batch one ttasr_frontend_run call accepts (sizes the handle's scratch: per-chunk maxima, per-tile minima) */
TTASR_API int ttasr_frontend_max_batch(const ttasr_frontend_t* h, int64_t* out);
/* which mel projection the handle runs: 80 = the straight-line code compiled in for the 80-filter Whisper bank (chosen
 * when mel_filters is bit-identical to the HF table: audio_utils.py:453-544 with 80 slaney triangles, 16 kHz, n_fft 400),
 * 0 = the generic per-bin program built from mel_filters (any bank of overlapping triangles, incl. the 128-filter one;
 * also forced by TTASR_FRONTEND_MEL=generic in the environment at create time).  Both give bit-identical features. */
TTASR_API int ttasr_frontend_mel_mode(const ttasr_frontend_t* h, int* out);
TTASR_API void ttasr_frontend_destroy(ttasr_frontend_t* h);

/* ------------------------------------------------------------------ ingest (step before the path) ---
 * Replaces the numeric part of `librosa.load(path, sr=16000, mono=True)` (reference: asr_core.py:156,
 * api/file_asr.py:271-275): decoded PCM frames -> float32 -> mean over channels -> rational resampling to 16 kHz with
 * librosa's res_type="polyphase" (= scipy.signal.resample_poly), written into a zero-padded buffer that is the
 * [n_chunks, 480000] input of ttasr_frontend_run.  File decoding stays with the caller.
 * up/down: target_sr / orig_sr reduced by their gcd (1/3 for 48 kHz, 160/441 for 44.1 kHz, 1/1 = convert only).
 * taps: host, [n_taps] fp32 = resample_poly's prototype filter firwin(2*10*max(up,down)+1, 1/max(up,down),
 *       window=("kaiser", 5.0)) * up; ignored for 1/1. */
TTASR_API int ttasr_ingest_create(int up, int down, const float* taps, int n_taps, ttasr_ingest_t** out);
/* ceil(n_in * up / down), the length librosa.resample / resample_poly return */
TTASR_API int ttasr_ingest_out_len(const ttasr_ingest_t* h, int64_t n_in, int64_t* n_out);
/* pcm_dev: [n_in, channels] interleaved frames (TTASR_PCM_I16 scaled by 1/32768, or TTASR_PCM_F32), channels 1..8.
 * out_dev: fp32 [out_capacity], out_capacity >= out_len(n_in); samples past the resampled signal are zero-filled. */
TTASR_API int ttasr_ingest_run(const ttasr_ingest_t* h, const void* pcm_dev, int pcm_dtype, int channels, int64_t n_in,
                     float* out_dev, int64_t out_capacity, void* stream);
/* Building blocks of VAD-aligned chunking (long-form files: the reference runs faster-whisper with vad_filter=True,
 * asr_core.py:159-167, api/file_asr.py:457-465, so its 30 s windows never start or end inside a word):
 * frame_energy: db[f] = 10 log10(mean square of pcm[f*hop .. f*hop+win)) over the 16 kHz mono signal — the track the
 *   host scans for pauses; gather_rows: out[r, :] = pcm[start[r] .. start[r]+len[r]) zero-padded to row_samples — the
 *   ragged [n_rows, 480000] + n_valid layout ttasr_frontend_run takes.  start: int64[n_rows], len: int32[n_rows]. */
TTASR_API int ttasr_ingest_frame_energy(const float* pcm_dev, int64_t n, int win, int hop, float* db_dev,
                                        int64_t n_frames, void* stream);
TTASR_API int ttasr_ingest_gather_rows(const float* pcm_dev, const int64_t* start_dev, const int32_t* len_dev,
                                       float* out_dev, int n_rows, int row_samples, void* stream);
TTASR_API void ttasr_ingest_destroy(ttasr_ingest_t* h);

/* ------------------------------------------------------------------ Whisper encoder --------------------------*/
typedef struct {
  int d_model;  /* multiple of 128 */
  int n_layers;
  int n_heads;  /* d_model / n_heads must be 64 */
  int ffn_dim;  /* multiple of 128 */
  int n_mels;   /* 80 or 128 */
  int n_ctx;    /* 1500 */
} ttasr_encoder_cfg;

/* One pre-LN block; HF names in comments (modeling_whisper.py:361-379).  Matrices: device bf16, HF [out, in]
 * row-major.  Vectors: device fp32. */
typedef struct {
  const float* ln1_g; const float* ln1_b;           /* self_attn_layer_norm.{weight,bias} */
  const void* wq; const float* bq;                  /* self_attn.q_proj */
  const void* wk;                                   /* self_attn.k_proj (no bias) */
  const void* wv; const float* bv;                  /* self_attn.v_proj */
  const void* wo; const float* bo;                  /* self_attn.out_proj */
  const float* ln2_g; const float* ln2_b;           /* final_layer_norm */
  const void* w1; const float* b1;                  /* fc1 */
  const void* w2; const float* b2;                  /* fc2 */
} ttasr_layer_weights;

typedef struct {
  const void* conv1_w; const float* conv1_b;        /* conv1.weight [d, n_mels, 3] bf16, conv1.bias fp32 */
  const void* conv2_w; const float* conv2_b;        /* conv2.weight [d, d, 3] bf16 */
  const float* pos;                                 /* embed_positions.weight [n_ctx, d] fp32 */
  const float* ln_post_g; const float* ln_post_b;   /* layer_norm */
  const ttasr_layer_weights* layers;                /* host array, n_layers entries */
} ttasr_weights;

/* How the residual stream x is held between the encoder blocks.
 *   TTASR_RESIDUAL_SPLIT (library default): x = hi + lo, two bf16 arrays (16 mantissa bits, the HBM cost of one fp32
 *     array).  `hi` is itself the bf16 A operand of the QKV / fc1 projections: the per-layer LayerNorms
 *     (modeling_whisper.py:393,403) are folded into those GEMMs (gamma into the weights, mean / rstd applied in the
 *     epilogue from row statistics the residual GEMMs emit), so no LayerNorm kernel runs between the blocks.
 *   TTASR_RESIDUAL_F32: fp32 array and one LayerNorm kernel in front of each QKV / fc1 GEMM (normalises before the
 *     bf16 rounding; the reference-precision path, ~5 % slower).
 *   TTASR_RESIDUAL_BF16: `hi` only — what an all-bf16 Hugging Face run keeps; measured in DESIGN.md, not recommended.
 *   TTASR_RESIDUAL_AUTO: the environment (TTASR_RESIDUAL=f32|split|bf16) or else the library default. */
enum { TTASR_RESIDUAL_AUTO = -1, TTASR_RESIDUAL_F32 = 0, TTASR_RESIDUAL_SPLIT = 1, TTASR_RESIDUAL_BF16 = 2 };

/* Packs the weights into its own device buffers (fused QKV with the d_h^-1/2 query scale folded in, conv filters
 * tap-major), builds TMA descriptors.  The caller may free its weight tensors afterwards. */
TTASR_API int ttasr_encoder_create(const ttasr_encoder_cfg* cfg, const ttasr_weights* w, ttasr_encoder_t** out);
/* ttasr_encoder_create with an explicit residual-stream representation (ttasr_encoder_create = TTASR_RESIDUAL_AUTO) */
TTASR_API int ttasr_encoder_create_ex(const ttasr_encoder_cfg* cfg, const ttasr_weights* w, int residual,
                                      ttasr_encoder_t** out);
TTASR_API int ttasr_encoder_workspace_bytes(const ttasr_encoder_t* h, int64_t batch, size_t* out);
/* feats: layout per `feats_layout` (tmajor_ld only read for the bf16 layout).  workspace_dev: >= workspace_bytes,
 * 1024-byte aligned.  out_dev: [batch, n_ctx, d_model] bf16 or fp32 per `out_dtype` (last_hidden_state). */
TTASR_API int ttasr_encoder_forward(const ttasr_encoder_t* h, const void* feats_dev, int feats_layout, int tmajor_ld,
                          int64_t batch, void* workspace_dev, size_t workspace_bytes, void* out_dev, int out_dtype,
                          void* stream);
/* number of kernel launches one forward of `batch` chunks enqueues (bench.py's gpu_launches accounting) */
TTASR_API int ttasr_encoder_launch_count(const ttasr_encoder_t* h, int64_t* out);
/* Optional per-stage timing (CUDA events recorded around every launch of ttasr_encoder_forward on its stream).
 * Off by default; enabling it makes the handle single-threaded.  ttasr_encoder_profile_read waits for the recorded
 * events and returns accumulated milliseconds and launch counts per stage kind (TTASR_PROFILE_KINDS entries:
 * see ttasr_encoder_profile_kind_name). */
#define TTASR_PROFILE_KINDS 9
TTASR_API int ttasr_encoder_profile_enable(ttasr_encoder_t* h, int on);
TTASR_API int ttasr_encoder_profile_read(ttasr_encoder_t* h, double* ms_by_kind, int64_t* launches_by_kind, int reset);
TTASR_API const char* ttasr_encoder_profile_kind_name(int kind);
TTASR_API void ttasr_encoder_destroy(ttasr_encoder_t* h);

/* ------------------------------------------------------------------ single ops (parity tests / profiling) -----
 * The building blocks of ttasr_encoder_forward, exposed so tests can pin each kernel separately.
 *
 * gemm:  out[M, N] = act(a[M, K] * w[N, K]^T + bias[N]) (+ addend[M, N]),  a, w bf16 row-major; bias fp32 or NULL;
 *        act: 0 = identity, 1 = exact-erf GELU; addend fp32 or NULL (may alias out when out is fp32);
 *        out_dtype TTASR_OUT_*; cta_group 1 or 2 (CTA-pair MMA), 0 = library default. */
TTASR_API int ttasr_op_gemm(const void* a_dev, const void* w_dev, const float* bias_dev, const float* addend_dev, void* out_dev,
                  int64_t M, int64_t N, int64_t K, int act, int out_dtype, int cta_group, void* stream);
/* gemm_split: the residual GEMM on the split stream (TTASR_RESIDUAL_SPLIT):
 *        (outh, outl) = split(act(a[M, K] * w[N, K]^T + bias[N]) + (addh[M, N] + addl[M, N])), all bf16; split(x) =
 *        (bf16(x), bf16(x - bf16(x))); outl and addl may both be NULL (plain bf16 stream); add* may alias out*.
 *        stats_out: optional float2 [M, N / 64]: per row and per 64 columns the (mean, sum of squared deviations) of
 *        the new outh values — what ttasr_op_gemm_lnfold consumes.  N % 128 == 0. */
TTASR_API int ttasr_op_gemm_split(const void* a_dev, const void* w_dev, const float* bias_dev, const void* addh_dev,
                                  const void* addl_dev, void* outh_dev, void* outl_dev, void* stats_out_dev, int64_t M,
                                  int64_t N, int64_t K, int act, int cta_group, void* stream);
/* gemm_lnfold: out[M, N] bf16 = act(LayerNorm(a)[M, K] * W^T + b) computed on the UN-normalised bf16 rows `a`:
 *        w_dev = bf16(W * gamma) [N, K], c1[n] = sum_k w_dev[n, k], c2[n] = b[n] + sum_k beta[k] W[n, k] (fp32), and
 *        out = act(rstd * (a w_dev^T - mean * c1) + c2) with mean / rstd per row combined from stats_in
 *        (float2 [M, parts] partial (mean, M2) over K / parts columns each, as written by ttasr_op_gemm_split). */
TTASR_API int ttasr_op_gemm_lnfold(const void* a_dev, const void* w_dev, const float* c1_dev, const float* c2_dev,
                                   const void* stats_in_dev, int parts, void* out_dev, int64_t M, int64_t N, int64_t K,
                                   int act, float eps, int cta_group, void* stream);
/* conv_stem: the encoder's two-layer convolutional stem on its own (modeling_whisper.py:567-568,619-626):
 *        feats_tm [B, 2T, ld] bf16 time-major (channels >= n_mels are ignored) -> GELU(conv1, k=3, p=1) ->
 *        GELU(conv2, k=3, s=2, p=1) + pos[T, d] -> out [B, T, d] fp32.  Weights in the HF layout: conv1_w
 *        [d, n_mels, 3], conv2_w [d, d, 3] bf16; biases and pos fp32.  scratch_dev: [B, 2T, d] bf16. */
TTASR_API int ttasr_op_conv_stem(const void* feats_tm_dev, int ld, int n_mels, int64_t B, int T, int d,
                                 const void* conv1_w_dev, const float* conv1_b_dev, const void* conv2_w_dev,
                                 const float* conv2_b_dev, const float* pos_dev, void* scratch_dev, float* out_dev,
                                 void* stream);
/* layernorm: y[rows, d] = (x - mean) / sqrt(var + 1e-5) * g + b, x fp32, y per out_dtype */
TTASR_API int ttasr_op_layernorm(const float* x_dev, const float* g_dev, const float* b_dev, void* y_dev, int64_t rows, int d,
                       int out_dtype, void* stream);
/* attention: qkv [batch, n_ctx, 3*d] bf16 (q | k | v, heads of 64 inside each), no mask, scale already folded in q;
 *        out [batch, n_ctx, d] bf16. */
TTASR_API int ttasr_op_attention(const void* qkv_dev, void* out_dev, int64_t batch, int n_ctx, int n_heads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTASR_ABI_H_ */
