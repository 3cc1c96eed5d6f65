// Fused Whisper log-mel front end for sm_100a.
//
//   kernel 1  logmel_frames_kernel : PCM -> framing (centre/reflect) -> Hann -> 400-pt FFT -> |.|^2 -> sparse mel
//                                    -> max(1e-10, .) -> log10 -> y = (x + 4) / 4 written once, as fp32 [B, n_mels, T]
//                                    and (optionally) as the bf16 time-major copy [B, T, ld] the conv stem's implicit
//                                    GEMM reads; per-chunk max (atomic) and per-tile min on the side.
//   kernel 2  logmel_clamp_kernel  : the reference's max(x, chunk_max - 8) is max(y, y_max - 2) after the affine map
//                                    (both monotone), so only tiles whose minimum lies below y_max - 2 are touched
//                                    again, plus the all-padding tiles kernel 1 skipped (one constant fill).  On
//                                    ordinary audio (dynamic range < 8 decades inside a tile) this pass reads two
//                                    small tables and exits: every feature byte is written to HBM exactly once.
//
// Reference semantics: HF WhisperFeatureExtractor._np_extract_fbank_features
// (transformers/models/whisper/feature_extraction_whisper.py:105-133, audio_utils.py:624-832), the extractor the
// reference repo calls at train_asr.py:607-616.
//
// Data movement: a persistent CTA walks tiles of 32 frames; the 5360-sample PCM span of the NEXT tile is brought
// into shared memory by one cp.async.bulk (TMA 1-D bulk copy, mbarrier completion) while the current tile is
// transformed, so every PCM byte is read from HBM once (+4.5 % halo, L2 hits).  Twenty threads own one pair of
// frames: 400 = 20 x 20, each thread a register-resident twiddle-free 20-point PFA DFT (fft400.cuh), one
// shared-memory transpose between the passes (strides chosen bank-conflict-free), then Hermitian split, power,
// the mel projection as a host-built streaming program (a bin feeds <= 2 adjacent triangles: each warp walks the bins
// of its filter range once with two accumulators), lg2, and 128-byte coalesced stores along time.
#include "fft400.cuh"
#include "frontend_logmel.h"
#include "ptx_sm100.cuh"

namespace ttasr {

namespace {

constexpr int kGroups = 16;                    // frame pairs per tile
constexpr int kTileFrames = 2 * kGroups;       // 32
constexpr int kThreads = kRadix * kGroups;     // 320
constexpr int kSpan = kHop * (kTileFrames - 1) + kNfft;  // 5360 samples per tile
constexpr int kTStride = 500;                  // per-group stride of the transpose buffer (== 20 mod 32)
constexpr int kTRow = 25;                      // k1 stride inside a group (== 1 mod 8, >= 20)
constexpr int kPwPitch = 33;                   // power spectra [bin][frame], odd pitch
static_assert(kPwPitch * 4 == kPowerPitchBytes, "mel program offsets");
static_assert(kNfreq * kPwPitch <= kGroups * kTStride, "power spectra must fit in the transpose buffer");
constexpr float kMelFloor = 1e-10f;
constexpr float kLog10Floor = -10.0f;

struct __align__(16) FrontSmem {
  float pcm[kSpan];                // raw samples (float, or int16 packed in the first half)
  float tr[kGroups * kTStride];    // transpose buffer (re); then Z (re); then power of the even frame
  float ti[kGroups * kTStride];    // (im);                 Z (im);       power of the odd frame
  float tw_re[kNfft], tw_im[kNfft];  // W400^(n2*k1) at [k1*20 + n2]: two 4-byte tables (the lanes of the second frame
                                   // pair in a warp re-read the first one's words: a broadcast for 32-bit loads, a
                                   // 2-way conflict per half-warp for one 64-bit load)
  float win[kNfft];
  float wmin[2][kThreads / 32];   // per-warp tile minima, double-buffered by tile parity (see the reduction)
  unsigned long long bar;
};

// column of frame f (0..31) inside a [bin][32 frames] power-spectrum row: a bijection with pw_col(2g + 2) - pw_col(2g) = 20
// (mod 32) except across g = 7 -> 8, and pw_col(2g + 1) = pw_col(2g) + 1
__device__ __forceinline__ int pw_col(int f) {
  const int g = f >> 1;
  return ((20 * g) & 31) + ((g >> 3) << 1) + (f & 1);
}

template <typename T>
__device__ __forceinline__ float load_sample(const float* buf, int i);
template <>
__device__ __forceinline__ float load_sample<float>(const float* buf, int i) { return buf[i]; }
template <>
__device__ __forceinline__ float load_sample<int16_t>(const float* buf, int i) {
  return static_cast<float>(reinterpret_cast<const int16_t*>(buf)[i]) * (1.0f / 32768.0f);
}

// order-preserving float -> uint map, so that a zero-filled (cudaMemsetAsync) word is below every float and
// atomicMax on the encoded value is a float max
__device__ __forceinline__ unsigned enc_ordered(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
constexpr float kYFloor = -1.5f;                         // (log10(1e-10) + 4) / 4
constexpr float kLog2ToY = 0.25f * 0.30102999566398120f; // y = log2(p) * log10(2) / 4 + 1

// (log10(max(acc, 1e-10)) + 4) / 4 through the hardware lg2 (abs err ~1e-7 << the 1e-4 gate; acc > 1e-10 is a normal
// number, so the flush-to-zero form needs no denormal path); the floor is returned exactly
__device__ __forceinline__ float mel_to_y(float acc) {
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(acc));
  return acc > kMelFloor ? fmaf(lg, kLog2ToY, 1.0f) : kYFloor;
}

// Sink of the baked mel code (mel_baked.inc): filters complete in ascending order and are written out in pairs
// (m even, m + 1): two fp32 stores along time (lane = frame), one packed bf16x2 word of the time-major staging row.
// RAW: 0 = no fp32 output, 1 = every lane of the warp stores (the tile lies inside the chunk), 2 = lanes past the last
// frame do not (last tile of a chunk) — a warp-uniform choice, so the common paths carry no per-lane branches.
template <int RAW>
struct MelEmit {
  float* out;          // &raw[b][0][t0 + lane]
  long long nf;        // n_frames
  uint32_t* stg;       // the lane's staging row (bf16 pairs)
  float vmax, vmin;    // over everything this lane emitted (the caller discards them for lanes past the last frame)
  bool live;
  __device__ __forceinline__ float y(float acc) const { return mel_to_y(acc); }
  __device__ __forceinline__ void emit2(int m, float y0, float y1) {
    if (RAW == 1 || (RAW == 2 && live)) {
      out[m * nf] = y0;
      out[(m + 1) * nf] = y1;
    }
    vmax = fmaxf(vmax, fmaxf(y0, y1));
    vmin = fminf(vmin, fminf(y0, y1));
    stg[m >> 1] = pack_bf16x2(y0, y1);
  }
};

#include "mel_baked.inc"

// MELB: 0 = generic mel program (any bank of overlapping triangles); 80 = the baked 80-filter Whisper bank (mel_baked.inc)
template <typename T, int MELB>
__global__ void __launch_bounds__(kThreads, 2)
logmel_frames_kernel(const T* __restrict__ pcm, long long row_stride, const int* __restrict__ n_valid, int n_samples,
                     int n_frames, int n_mels, int batch, FrontTables tables, const __grid_constant__ MelProgram mel,
                     float* __restrict__ raw,
                     unsigned* __restrict__ chunk_max, float* __restrict__ tile_min, __nv_bfloat16* __restrict__ tmajor,
                     int tmajor_ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FrontSmem& s = *reinterpret_cast<FrontSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int grp = tid / kRadix;   // frame pair
  const int sub = tid % kRadix;   // n2 in pass 1, k1 in pass 2
  const int tiles_per_chunk = (n_frames + kTileFrames - 1) / kTileFrames;
  const long long total_tiles = static_cast<long long>(batch) * tiles_per_chunk;

  // ---- one-time tables -> shared memory
  for (int i = tid; i < kNfft; i += kThreads) {
    const float2 w = tables.twiddle[i];
    s.tw_re[i] = w.x;
    s.tw_im[i] = w.y;
    s.win[i] = tables.window[i];
  }
  const uint32_t bar = smem_u32(&s.bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const bool rows_aligned = ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && ((row_stride * sizeof(T)) % 16 == 0);
  float2 w20;   // W20^sub = exp(-i pi sub / 10), the thread's constant of the twiddle symmetry in pass 1
  {
    double sn, cs;   // once per thread, in double so that the constant is correctly rounded like the table entries
    sincospi(-0.1 * static_cast<double>(sub), &sn, &cs);
    w20 = make_float2(static_cast<float>(cs), static_cast<float>(sn));
  }

  // tile classification: 0 = all-zero (skip the transform), 1 = interior (bulk copy), 2 = edge (reflect / zero fill)
  auto classify = [&](int b, int tt, int& valid) -> int {
    valid = n_valid ? min(max(n_valid[b], 0), n_samples) : n_samples;
    valid = static_cast<int>(min(static_cast<long long>(valid), row_stride));  // never read past the row
    const int start = tt * kTileFrames * kHop - kNfft / 2;
    if (start >= valid && valid < n_samples - kNfft) return 0;
    if (rows_aligned && start >= 0 && start + kSpan <= valid) return 1;
    return 2;
  };
  // bring a tile's PCM span into the staging buffer (asynchronously when interior); caller guarantees it is free.
  // (Measured and dropped: staging the span in 320-sample blocks at a pitch of 340 words, which makes the pass-1 sample
  // loads of the frame pairs sharing a warp conflict-free — 7 % fewer shared-memory wavefronts for 17 bulk copies per
  // tile instead of one: 0.556 -> 0.554 ms per 256 chunks with the copies issued by 17 lanes, 0.570 by one.)
  auto stage = [&](int b, int tt) -> int {
    int valid;
    const int kind = classify(b, tt, valid);
    const int start = tt * kTileFrames * kHop - kNfft / 2;
    const T* row = pcm + static_cast<long long>(b) * row_stride;
    if (kind == 1) {
      if (tid == 0) {
        fence_proxy_async_smem();  // earlier generic-proxy accesses to the buffer vs. the async-proxy write
        mbar_arrive_expect_tx(bar, kSpan * sizeof(T));
        bulk_load_1d(smem_u32(&s.pcm[0]), row + start, kSpan * sizeof(T), bar);
      }
    } else if (kind == 2) {
      T* dst = reinterpret_cast<T*>(&s.pcm[0]);
      for (int i = tid; i < kSpan; i += kThreads) {
        int idx = start + i;
        if (idx < 0) idx = -idx;  // reflect (edge sample not repeated)
        if (idx >= n_samples) idx = 2 * n_samples - 2 - idx;
        T v = T(0);
        if (idx >= 0 && idx < valid) v = row[idx];
        dst[i] = v;
      }
    }
    return kind;
  };

  // tiles are walked with stride gridDim.x; (chunk, tile-in-chunk) advance incrementally (no 64-bit division per tile)
  const int step_b = static_cast<int>(gridDim.x) / tiles_per_chunk, step_t = static_cast<int>(gridDim.x) % tiles_per_chunk;
  uint32_t phase = 0;
  int wbuf = 0;
  long long tile = blockIdx.x;
  int b = static_cast<int>(blockIdx.x) / tiles_per_chunk, tt = static_cast<int>(blockIdx.x) % tiles_per_chunk;
  int kind_cur = 0;
  if (tile < total_tiles) kind_cur = stage(b, tt);

  for (; tile < total_tiles; tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    int b_next = b + step_b, tt_next = tt + step_t;
    if (tt_next >= tiles_per_chunk) { tt_next -= tiles_per_chunk; ++b_next; }
    int kind_next = 0;
    const int t0 = tt * kTileFrames;
    float tmax = -INFINITY, tmin = INFINITY;

    if (kind_cur == 0) {
      // every sample the tile touches is zero padding: mel power 0 -> floor.  Nothing is written here; the clamp
      // kernel fills the tile with max(floor, y_max - 2) in one go (tile_min = -inf marks "not written").
      tmax = kYFloor;
      tmin = -INFINITY;
      if (next < total_tiles) kind_next = stage(b_next, tt_next);  // staging buffer is idle on this path
    } else {
      if (kind_cur == 1) {
        mbar_wait(bar, phase);
        phase ^= 1;
      } else {
        __syncthreads();
      }
      // ---------------- pass 1: thread (grp, n2) transforms z[20 n1 + n2], n1 = 0..19
      {
        const float* x = &s.pcm[0];
        const int base = grp * 2 * kHop + sub;
        float yr[20], yi[20];
        {
          c32_t z[20], y[20];   // frame a in the real lane, frame b in the imaginary lane: packed f32x2 butterflies
#pragma unroll
          for (int n1 = 0; n1 < 20; ++n1) {
            const float w = s.win[20 * n1 + sub];
            z[n1] = c_pack(w * load_sample<T>(x, base + 20 * n1), w * load_sample<T>(x, base + kHop + 20 * n1));
          }
          dft20_packed(z, y);
#pragma unroll
          for (int k1 = 0; k1 < 20; ++k1) c_unpack(y[k1], yr[k1], yi[k1]);
        }
        float* tr = &s.tr[grp * kTStride + sub];
        float* ti = &s.ti[grp * kTStride + sub];
        tr[0] = yr[0];
        ti[0] = yi[0];
        // twiddles W400^(n2 k1): k1 = 1..10 come from the table; W400^(n2 (20 - k1)) = W20^n2 * conj(W400^(n2 k1)) is
        // derived with the thread's constant W20^n2 (4 FMA-pipe ops instead of 2 shared-memory loads: the load/store
        // pipe is the busy one in this kernel)
#pragma unroll
        for (int k1 = 1; k1 <= 10; ++k1) {
          const float wx = s.tw_re[k1 * 20 + sub], wy = s.tw_im[k1 * 20 + sub];
          tr[k1 * kTRow] = yr[k1] * wx - yi[k1] * wy;
          ti[k1 * kTRow] = yr[k1] * wy + yi[k1] * wx;
          if (k1 < 10) {
            const int k1m = 20 - k1;
            const float vx = fmaf(w20.x, wx, w20.y * wy), vy = fmaf(w20.y, wx, -(w20.x * wy));
            tr[k1m * kTRow] = yr[k1m] * vx - yi[k1m] * vy;
            ti[k1m * kTRow] = yr[k1m] * vy + yi[k1m] * vx;
          }
        }
      }
      __syncthreads();
      // the PCM span is consumed: fetch the next tile's span under passes 2..4
      if (next < total_tiles) kind_next = stage(b_next, tt_next);
      // ---------------- pass 2: thread (grp, k1) transforms over n2 -> Z[k1 + 20 k2], then the Hermitian split + power of
      // the two real frames (frame a = 2 grp in Re, frame b = 2 grp + 1 in Im).  A thread's own bins k = k1 + 20 i
      // (i = 0..10, k <= 200) are exactly its outputs k2 = i and never leave the registers; only the partners Z[400 - k]
      // — the upper half of the spectrum, k2 >= 10 — go through shared memory (half the stores and loads of a full
      // round trip).  The spectra are then written bin-major, pw[k][frame] with a row pitch of 33 floats, so that in the
      // mel pass lanes = frames read consecutive words (no bank conflicts) and the per-filter weight is a broadcast.
      {
        float yr[20], yi[20];
        const float* tr = &s.tr[grp * kTStride + sub * kTRow];
        const float* ti = &s.ti[grp * kTStride + sub * kTRow];
        {
          c32_t z[20], y[20];
#pragma unroll
          for (int n2 = 0; n2 < 20; ++n2) z[n2] = c_pack(tr[n2], ti[n2]);
          dft20_packed(z, y);
#pragma unroll
          for (int k2 = 0; k2 < 20; ++k2) c_unpack(y[k2], yr[k2], yi[k2]);
        }
        __syncthreads();  // everyone has read its transpose rows; the buffer now takes the upper half of Z
        {
          float* zr = &s.tr[grp * kTStride + sub];
          float* zi = &s.ti[grp * kTStride + sub];
#pragma unroll
          for (int k2 = 10; k2 < 20; ++k2) {
            zr[20 * k2] = yr[k2];
            zi[20 * k2] = yi[k2];
          }
        }
        __syncthreads();
        const float* zr = &s.tr[grp * kTStride];
        const float* zi = &s.ti[grp * kTStride];
        float pa[11], pb[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          const int k = sub + kRadix * i;
          pa[i] = 0.f;
          pb[i] = 0.f;
          if (k < kNfreq) {
            // partner Z[(400 - k) mod 400]; bin 0 is its own partner (and lies in the half that is not stored)
            const float wr = (k == 0) ? yr[0] : zr[kNfft - k];
            const float wi = (k == 0) ? yi[0] : zi[kNfft - k];
            split_power(yr[i], yi[i], wr, wi, pa[i], pb[i]);
          }
        }
        __syncthreads();  // all Z reads done; the (re) buffer now holds the power spectra
        // frame f of the tile sits in COLUMN pw_col(f) of its bin row: consecutive frame pairs are 20 columns apart (mod 32),
        // so the two pairs whose threads share a warp write disjoint banks; the mel pass reads column pw_col(lane)
        float* pw = &s.tr[pw_col(2 * grp)];
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          const int k = sub + kRadix * i;
          if (k < kNfreq) {
            pw[k * kPwPitch] = pa[i];
            pw[k * kPwPitch + 1] = pb[i];
          }
        }
      }
      __syncthreads();
      // ---------------- mel projection + log + store.  lane = frame within the tile; warp w owns a contiguous range of
      // filters and walks its frequency bins once with two accumulators (a bin feeds <= 2 adjacent triangles): the
      // program (bin, weight for the current filter, weight for the next one, filters completed) is built on the host.
      {
        const int lane = tid & 31;
        const int wrp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler: constant-bank indexing
        const bool live = (t0 + lane) < n_frames;
        // pitch of the bf16 staging rows (one per frame): the staging store is unconditional, so without a time-major
        // output the rows still must not overlap (they are simply never copied out)
        const int stg_ld = tmajor ? tmajor_ld : kMaxMels;
        if constexpr (MELB != 0) {
          // the bank is known at compile time: straight-line LDS + FFMA per bin, weights and offsets as immediates
          auto run = [&](auto o) {
            o.out = raw + static_cast<long long>(b) * MELB * n_frames + t0 + lane;
            o.nf = n_frames;
            o.stg = reinterpret_cast<uint32_t*>(&s.ti[0]) + lane * ((stg_ld >> 1) + 1);   // ti is idle by now
            o.vmax = -INFINITY;
            o.vmin = INFINITY;
            o.live = live;
            MelBaked<MELB>::run(wrp, &s.tr[pw_col(lane)], o);
            if (live) {
              tmax = fmaxf(tmax, o.vmax);
              tmin = fminf(tmin, o.vmin);
            }
          };
          if (raw == nullptr) run(MelEmit<0>{});
          else if (t0 + kTileFrames <= n_frames) run(MelEmit<1>{});
          else run(MelEmit<2>{});
        } else {
          int m = mel.m0[wrp];
          float* out_ptr = raw + (static_cast<long long>(b) * n_mels + m) * n_frames + t0 + lane;
          __nv_bfloat16* stg = reinterpret_cast<__nv_bfloat16*>(&s.ti[0]) + lane * (stg_ld + 2) + m;  // ti is idle by now
          const unsigned char* pw = reinterpret_cast<const unsigned char*>(&s.tr[pw_col(lane)]);
          // the per-filter completion path is kept free of per-lane branches and pointer tests: one store predicate, the
          // staging row written unconditionally (it is shared memory), min / max folded in after the loop
          const bool st_raw = live && raw != nullptr;
          float vmax = -INFINITY, vmin = INFINITY;
          float acc_a = 0.f, acc_b = 0.f;
          const int op_end = mel.op_off[wrp + 1];
          // (Software-pipelining this loop by one op — next descriptor and power value requested before the current FMAs —
          // was measured: 0.603 -> 0.618 ms per 256 chunks; the co-resident CTA already fills those stalls.)
          for (int oi = mel.op_off[wrp]; oi < op_end; ++oi) {
            const int4 op = mel.ops[oi];
            const float p = *reinterpret_cast<const float*>(pw + op.x);
            acc_a = fmaf(__int_as_float(op.y), p, acc_a);
            acc_b = fmaf(__int_as_float(op.z), p, acc_b);
            if (op.w) {  // filter complete
              const float v = mel_to_y(acc_a);
              if (st_raw) *out_ptr = v;
              vmax = fmaxf(vmax, v);
              vmin = fminf(vmin, v);
              *stg = __float2bfloat16_rn(v);
              acc_a = acc_b;
              acc_b = 0.f;
              out_ptr += n_frames;
              ++stg;
            }
          }
          if (live) {
            tmax = fmaxf(tmax, vmax);
            tmin = fminf(tmin, vmin);
          }
        }
      }
      __syncthreads();  // power spectra consumed before the next tile's pass 1 overwrites the buffer; staging complete
      if (tmajor) {
        // [t][m] bf16 rows, channels contiguous and zero-padded to tmajor_ld: 4-byte words, coalesced per frame
        const int wpr = tmajor_ld >> 1;                     // words per row
        const uint32_t* stw = reinterpret_cast<const uint32_t*>(&s.ti[0]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tmajor + (static_cast<long long>(b) * n_frames + t0) * tmajor_ld);
        const int nrows = min(kTileFrames, n_frames - t0);
        for (int f = tid >> 5; f < nrows; f += kThreads / 32) {      // one warp per frame row, lanes along the channels
          for (int wi = tid & 31; wi < wpr; wi += 32) {
            uint32_t v = stw[f * (wpr + 1) + wi];
            if (2 * wi + 1 >= n_mels) v = (2 * wi < n_mels) ? (v & 0xffffu) : 0u;
            dst[static_cast<long long>(f) * wpr + wi] = v;
          }
        }
        // (the next writer of ti is the next tile's pass 1, behind the block barrier of the tile-min reduction below)
      }
    }
    // ---------------- tile max -> chunk max (one atomic per warp); tile min -> table (one store per tile)
    // (one REDUX each on the order-preserving integer encoding instead of a five-step shuffle tree)
    tmax = dec_ordered(__reduce_max_sync(0xffffffffu, enc_ordered(tmax)));
    tmin = dec_ordered(__reduce_min_sync(0xffffffffu, enc_ordered(tmin)));
    // (double-buffered: a padding tile reaches this point without any block barrier, so the other warps may already
    // be writing the NEXT tile's minima while thread 0 still reads this tile's)
    if ((tid & 31) == 0) {
      if (tmax > -INFINITY) atomicMax(&chunk_max[b], enc_ordered(tmax));
      s.wmin[wbuf][tid >> 5] = tmin;
    }
    __syncthreads();
    if (tid == 0) {
      float m = s.wmin[wbuf][0];
#pragma unroll
      for (int w = 1; w < kThreads / 32; ++w) m = fminf(m, s.wmin[wbuf][w]);
      tile_min[static_cast<long long>(b) * tiles_per_chunk + tt] = m;
    }
    wbuf ^= 1;
    kind_cur = kind_next;
    b = b_next;
    tt = tt_next;
  }
}

// Second pass, normally a no-op: CTA (x, b) looks at kClampTiles tiles of chunk b and rewrites only those whose
// minimum is below y_max - 2 (clamp) or that kernel 1 left unwritten (all-padding tiles: constant fill).
constexpr int kClampTiles = 8;
__global__ void __launch_bounds__(256)
logmel_clamp_kernel(float* __restrict__ feats, const unsigned* __restrict__ chunk_max, const float* __restrict__ tile_min,
                    int n_frames, int n_mels, int tiles_per_chunk, __nv_bfloat16* __restrict__ tmajor, int tmajor_ld,
                    float clamp_y) {
  __shared__ float tile[kMaxMels][33];
  const int b = blockIdx.y;
  const float lo = dec_ordered(chunk_max[b]) - clamp_y;   // 8 decades of log10 = 2.0 after (x + 4) / 4; +inf = no clamp
  float* base = feats + static_cast<long long>(b) * n_mels * n_frames;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int wpr = tmajor_ld >> 1;  // 4-byte words per bf16 row
  const int tile_end = min((blockIdx.x + 1) * kClampTiles, tiles_per_chunk);
  for (int tt = blockIdx.x * kClampTiles; tt < tile_end; ++tt) {
    const float tmin = tile_min[static_cast<long long>(b) * tiles_per_chunk + tt];
    const bool unwritten = (tmin == -INFINITY);
    if (tmin >= lo && !unwritten) continue;         // the common case: nothing below the clamp
    const int t0 = tt * kTileFrames;
    const int t = t0 + lane;
    const float fillv = fmaxf(kYFloor, lo);
    for (int m = wrp; m < n_mels; m += 8) {         // lanes = frames: 128-byte segments along time
      float v = fillv;
      if (t < n_frames) {
        if (feats) {
          float* ptr = base + static_cast<long long>(m) * n_frames + t;
          if (!unwritten) v = fmaxf(*ptr, lo);
          *ptr = v;
        } else if (!unwritten) {
          // time-major-only run: the bf16 copy is the only output; rounding is monotone, so clamping the rounded value
          // and rounding the bound gives bf16(max(y, lo)) exactly
          v = fmaxf(__bfloat162float(tmajor[(static_cast<long long>(b) * n_frames + t) * tmajor_ld + m]), lo);
        }
      }
      tile[m][lane] = v;
    }
    if (tmajor) {                                   // rewrite the bf16 rows of the tile, channels contiguous
      __syncthreads();
      uint32_t* dst = reinterpret_cast<uint32_t*>(tmajor + (static_cast<long long>(b) * n_frames + t0) * tmajor_ld);
      const int nrows = min(kTileFrames, n_frames - t0);
      for (int idx = threadIdx.x; idx < nrows * wpr; idx += 256) {
        const int f = idx / wpr, mp = (idx - f * wpr) * 2;
        const float v0 = mp < n_mels ? tile[mp][f] : 0.f;
        const float v1 = mp + 1 < n_mels ? tile[mp + 1][f] : 0.f;
        dst[idx] = pack_bf16x2(v0, v1);
      }
      __syncthreads();                              // tile[] is reused by the next flagged tile
    }
  }
}

}  // namespace

size_t frontend_smem_bytes() { return sizeof(FrontSmem); }
unsigned long long frontend_baked_hash(int n_mels) { return n_mels == 80 ? MelBaked<80>::kHash : 0ull; }
int frontend_tiles(int n_samples) { return (n_samples / kHop + kTileFrames - 1) / kTileFrames; }

cudaError_t launch_logmel(const void* pcm, int pcm_is_i16, long long row_stride, const int* n_valid, int n_samples,
                          int n_mels, int batch, const FrontTables& tables, const MelProgram& mel, float* feats, unsigned* chunk_max,
                          float* tile_min, __nv_bfloat16* tmajor, int tmajor_ld, int num_sms, cudaStream_t stream,
                          float clamp_decades, int mel_baked) {
  const int n_frames = n_samples / kHop;
  const int tiles_per_chunk = (n_frames + kTileFrames - 1) / kTileFrames;
  const long long total = static_cast<long long>(batch) * tiles_per_chunk;
  if (total == 0) return cudaSuccess;
  const size_t smem = sizeof(FrontSmem);
  if (tmajor && static_cast<size_t>(kTileFrames) * (tmajor_ld + 2) * sizeof(__nv_bfloat16) > sizeof(FrontSmem::ti))
    return cudaErrorInvalidValue;
  if (mel_baked != 0 && mel_baked != n_mels) return cudaErrorInvalidValue;
  {
    cudaError_t e = cudaMemsetAsync(chunk_max, 0, sizeof(unsigned) * batch, stream);  // 0 < enc_ordered(any float)
    if (e != cudaSuccess) return e;
  }
  const int grid = static_cast<int>(total < 2LL * num_sms ? total : 2LL * num_sms);
  auto launch = [&](auto kern, PerDeviceOnce& once, const auto* in) -> cudaError_t {
    if (once.need()) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<grid, kThreads, smem, stream>>>(in, row_stride, n_valid, n_samples, n_frames, n_mels, batch, tables, mel, feats,
                                           chunk_max, tile_min, tmajor, tmajor_ld);
    return cudaSuccess;
  };
  static PerDeviceOnce once[4];
  cudaError_t le;
  const float* pf = static_cast<const float*>(pcm);
  const int16_t* pi = static_cast<const int16_t*>(pcm);
  if (mel_baked == 80)
    le = pcm_is_i16 ? launch(logmel_frames_kernel<int16_t, 80>, once[0], pi) : launch(logmel_frames_kernel<float, 80>, once[1], pf);
  else
    le = pcm_is_i16 ? launch(logmel_frames_kernel<int16_t, 0>, once[2], pi) : launch(logmel_frames_kernel<float, 0>, once[3], pf);
  if (le != cudaSuccess) return le;
  dim3 g2((tiles_per_chunk + kClampTiles - 1) / kClampTiles, batch);
  logmel_clamp_kernel<<<g2, 256, 0, stream>>>(feats, chunk_max, tile_min, n_frames, n_mels, tiles_per_chunk, tmajor,
                                              tmajor_ld, clamp_decades * 0.25f);
  return cudaGetLastError();
}

}  // namespace ttasr
