// Non-causal multi-head attention forward for the Whisper encoder on sm_100a (head_dim 64, no mask, scale folded
// into q):  out = softmax(q k^T) v  per (chunk, head), reference eager_attention_forward / WhisperAttention.forward
// (transformers/models/whisper/modeling_whisper.py:215-238,284-357).
//
// One persistent CTA per SM walks work items (chunk, head, 256 query rows).  Per item, two 128-row query tiles
// ping-pong over the K/V tiles so the tensor pipe and the exp pipe overlap:
//   warp 0      TMA producer : Q tiles once per item; K/V 128 x 64 tiles through a 4-stage ring (128B swizzle)
//   warp 1      MMA issuer   : S_t = Q_t K^T (tcgen05.mma SS, 128x128x64, fp32 in TMEM);
//                              O_t += P_t V (tcgen05.mma TS: P read from TMEM, V as an MN-major smem operand)
//   warp 2      TMEM allocator (512 columns: S0 S1 | P0 P1 (bf16 pairs) | O0 O1)
//   warps 4-7 / 8-11  softmax of query tile 0 / 1, one thread per query row.  The exp sweeps of the two warpgroups
//                              are forced into anti-phase by a token (two named barriers) so the 16-op/clk exp pipe
//                              always has exactly one warpgroup feeding it, while the other one waits for its next
//                              S tile, loads it from TMEM and takes the row max (thread-local, exact).  O is
//                              rescaled only when the max outgrows the exponent base by > 2^32.  exp2 on
//                              pre-scaled logits, fp32 row sums, P written back to TMEM as packed bf16; final
//                              O / l -> bf16 -> smem -> TMA store.
// The 1500 x 1500 score matrix never leaves the SM; keys beyond n_ctx in the last tile are masked to -inf.
#include "attention_sm100.h"

#include <stdlib.h>
#include <string.h>

#include "gemm_sm100.h"  // encode_tmap
#include "ptx_sm100.cuh"

namespace ttasr {
namespace {

constexpr int kTile = 128;          // query rows per tile, keys per KV tile
constexpr int kHeadDim = 64;
constexpr int kTileBytes = kTile * kHeadDim * 2;  // 16 KB
constexpr int kKvStages = 4;
constexpr int kAttnThreads = 384;
constexpr float kLog2e = 1.4426950408889634f;
#ifndef TTASR_ATTN_PRETOKEN
#define TTASR_ATTN_PRETOKEN 1
#endif
constexpr int kPreTokenChunks = TTASR_ATTN_PRETOKEN;  // quarters of the exp sweep done outside the token
constexpr float kRescaleThreshold = 32.0f;  // log2 units: P stays <= 2^32 (bf16 range 2^127, O and l are fp32)

// TMEM column map
constexpr uint32_t kColS = 0;     // S0 at 0, S1 at 128
constexpr uint32_t kColP = 256;   // P0 at 256, P1 at 320 (128 keys as bf16 pairs = 64 columns)
constexpr uint32_t kColO = 384;   // O0 at 384, O1 at 448

struct AttnParams {
  CUtensorMap tm_qkv;  // 3-D (3*d, n_ctx, batch), box (64, 128, 1)
  CUtensorMap tm_out;  // 3-D (d, n_ctx, batch), box (64, 128, 1)
  int n_ctx;
  int n_heads;
  int d_model;
  int q_blocks;   // ceil(n_ctx / 256)
  int kv_tiles;   // ceil(n_ctx / 128)
  int num_items;  // batch * heads * q_blocks
};

// Debug timeline (only with -DTTASR_ATTN_TRACE=1): CTA 0 records (tag, clock64) pairs per role into a global buffer
// of 4 regions x cap int64 (0: MMA thread, 1/2: softmax tile 0/1 (first thread), 3: TMA producer).
#ifndef TTASR_ATTN_TRACE
#define TTASR_ATTN_TRACE 0
#endif
#if TTASR_ATTN_TRACE
__device__ long long* g_trace_buf = nullptr;
__device__ int g_trace_cap = 0;
struct Tracer {
  long long* base;
  int n, cap;
  __device__ Tracer(int region, bool on) : base(nullptr), n(0), cap(0) {
    if (on && blockIdx.x == 0 && g_trace_buf) { base = g_trace_buf + static_cast<long long>(region) * g_trace_cap; cap = g_trace_cap; }
  }
  __device__ __forceinline__ void operator()(int tag) {
    if (base && n + 2 <= cap) { base[n] = tag; base[n + 1] = clock64(); n += 2; }
  }
};
#else
struct Tracer {
  __device__ Tracer(int, bool) {}
  __device__ __forceinline__ void operator()(int) {}
};
#endif

struct AttnSmem {
  uint8_t q[2][kTileBytes];
  uint8_t k[kKvStages][kTileBytes];
  uint8_t v[kKvStages][kTileBytes];
  uint8_t o[2][kTileBytes];
  unsigned long long q_full, q_free;
  unsigned long long kv_full[kKvStages], kv_free[kKvStages];
  unsigned long long s_full[2], s_free[2], p_ready[2], o_done[2];
  uint32_t tmem_ptr;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// max over 32 fp32 values held in registers (two chains of 3-input max)
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    m0 = fmaxf(m0, __uint_as_float(v[2 * i]));
    m1 = fmaxf(m1, __uint_as_float(v[2 * i + 1]));
  }
  return fmaxf(m0, m1);
}

// keys >= valid (tile-relative) of the 32-column chunk starting at col0 become -inf
__device__ __forceinline__ void mask_tail(uint32_t (&v)[32], int col0, int valid) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (col0 + i >= valid) v[i] = 0xff800000u;
}

// p_i = 2^(s_i*log2e - m_used), written back over the scores; the row sum and the bf16 packing (sum_pack) of a chunk
// are issued after the exponentials of the next chunk.
#ifndef TTASR_ATTN_F32X2
#define TTASR_ATTN_F32X2 1
#endif
#ifndef TTASR_ATTN_SETMAXNREG
#define TTASR_ATTN_SETMAXNREG 0
#endif
#ifndef TTASR_ATTN_POLY_Q0
#define TTASR_ATTN_POLY_Q0 0
#endif
#ifndef TTASR_ATTN_POLY_Q1
#define TTASR_ATTN_POLY_Q1 0
#endif
#ifndef TTASR_ATTN_POLY_Q2
#define TTASR_ATTN_POLY_Q2 0
#endif
#ifndef TTASR_ATTN_POLY_Q3
#define TTASR_ATTN_POLY_Q3 0
#endif
// 2^x for a PAIR on the FMA / ALU pipes instead of the 16-op/clk MUFU pipe (FlashAttention-4's trick): n = rint(x) by
// the magic-number add, f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial (max rel err 7.6e-5 — P is
// rounded to bf16, 3.9e-3, right after), exponent added into the bit pattern.  x <= 32 here (lazy rescale bound);
// x is clamped at -126, where 2^x is 1e-38 and indistinguishable from 0 in the row sum.
__device__ __forceinline__ void exp2_poly_pair(uint32_t& a, uint32_t& b) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
  const float x0 = fmaxf(__uint_as_float(a), -126.0f), x1 = fmaxf(__uint_as_float(b), -126.0f);
  const f32x2_t x = pack2(x0, x1);
  const f32x2_t t = add2(x, pack2(kMagic, kMagic));
  const f32x2_t n = add2(t, pack2(-kMagic, -kMagic));
  const f32x2_t f = add2(x, n ^ 0x8000000080000000ull);  // x - n
  f32x2_t p = fma2(pack2(0.05520551f, 0.05520551f), f, pack2(0.24261396f, 0.24261396f));
  p = fma2(p, f, pack2(0.69325476f, 0.69325476f));
  p = fma2(p, f, pack2(0.99992773f, 0.99992773f));
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  a = __float_as_uint(p0) + (__float_as_uint(t0) << 23);
  b = __float_as_uint(p1) + (__float_as_uint(t1) << 23);
}

// TTASR_ATTN_LATEMAX=1 (experiment, measured slower — DESIGN.md 4.3): only quarter 0's own maximum is taken before the
// sweep starts; the maximum of the other three quarters is computed on the ALU pipe BESIDE quarter 0's exponentials
// (which run against the previous base), the base is corrected afterwards in the rare case it has to move, and the
// o_done wait moves behind quarter 0.  The in-kernel timeline shows the same chain length: the steps it removes
// (row max 390 -> 207 cycles, o_done wait) come back as the longer quarter 0 and the extra checks.
#ifndef TTASR_ATTN_LATEMAX
#define TTASR_ATTN_LATEMAX 0
#endif
#ifndef TTASR_ATTN_SMSP_TOKEN
#define TTASR_ATTN_SMSP_TOKEN 0
#endif
#ifndef TTASR_ATTN_EARLY_RELEASE
#define TTASR_ATTN_EARLY_RELEASE 0
#endif
#ifndef TTASR_ATTN_FAKE_EXP
#define TTASR_ATTN_FAKE_EXP 0
#endif
#ifndef TTASR_ATTN_TWO_MMA
#define TTASR_ATTN_TWO_MMA 0
#endif
// TTASR_ATTN_ABLATE (diagnostic builds only, wrong results): bit 0 = no row max, bit 1 = no P store to TMEM,
// bit 2 = only the first quarter of S is loaded from TMEM, bit 3 = no bf16 packing / row sums
#ifndef TTASR_ATTN_ABLATE
#define TTASR_ATTN_ABLATE 0
#endif
// TTASR_ATTN_POLL=1 (experiment): the kernel's mbarrier waits poll with test_wait instead of the suspending try_wait
#ifndef TTASR_ATTN_POLL
#define TTASR_ATTN_POLL 0
#endif
#if TTASR_ATTN_POLL
#define ATTN_WAIT mbar_wait_poll
#else
#define ATTN_WAIT mbar_wait
#endif

// POLY8: of every 8 consecutive scores of the chunk, the first POLY8 (even) take the polynomial, the rest MUFU.EX2
template <int POLY8>
__device__ __forceinline__ void exp_inplace(uint32_t (&v)[32], float m_used) {
#if TTASR_ATTN_F32X2
  // packed FFMA2: one issue slot scales and shifts two scores (the sweep shares its scheduler with the MMA / TMA warps)
  const float neg_m = -m_used;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    asm("{ .reg .b64 t, u, w;\n\t"
        "mov.b64 t, {%0, %1};\n\t"
        "mov.b64 u, {%2, %2};\n\t"
        "mov.b64 w, {%3, %3};\n\t"
        "fma.rn.f32x2 t, t, u, w;\n\t"
        "mov.b64 {%0, %1}, t; }"
        : "+r"(v[i]), "+r"(v[i + 1])
        : "r"(__float_as_uint(kLog2e)), "r"(__float_as_uint(neg_m)));
    if ((i & 7) < POLY8) {
      exp2_poly_pair(v[i], v[i + 1]);
#if TTASR_ATTN_FAKE_EXP   // diagnostic builds only (wrong results): that many of every 8 exponentials cost nothing
    } else if ((i & 7) < TTASR_ATTN_FAKE_EXP) {
      v[i] &= 0x3fffffffu;
      v[i + 1] &= 0x3fffffffu;
#endif
    } else {
      v[i] = __float_as_uint(ex2(__uint_as_float(v[i])));
      v[i + 1] = __float_as_uint(ex2(__uint_as_float(v[i + 1])));
    }
  }
#else
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(ex2(fmaf(__uint_as_float(v[i]), kLog2e, -m_used)));
#endif
}
// exp_inplace over v, interleaved with the maximum over three other chunks (independent work for the ALU pipe while
// the exponentials occupy the MUFU pipe)
template <int POLY8>
__device__ __forceinline__ float exp_inplace_max3(uint32_t (&v)[32], float m_used, const uint32_t (&a)[32],
                                                  const uint32_t (&b)[32], const uint32_t (&c)[32]) {
  const float neg_m = -m_used;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    asm("{ .reg .b64 t, u, w;\n\t"
        "mov.b64 t, {%0, %1};\n\t"
        "mov.b64 u, {%2, %2};\n\t"
        "mov.b64 w, {%3, %3};\n\t"
        "fma.rn.f32x2 t, t, u, w;\n\t"
        "mov.b64 {%0, %1}, t; }"
        : "+r"(v[i]), "+r"(v[i + 1])
        : "r"(__float_as_uint(kLog2e)), "r"(__float_as_uint(neg_m)));
    if ((i & 7) < POLY8) {
      exp2_poly_pair(v[i], v[i + 1]);
    } else {
      v[i] = __float_as_uint(ex2(__uint_as_float(v[i])));
      v[i + 1] = __float_as_uint(ex2(__uint_as_float(v[i + 1])));
    }
    // three 3-input maxima per pair of exponentials, two dependency chains
    m0 = fmaxf(m0, fmaxf(__uint_as_float(a[i]), __uint_as_float(a[i + 1])));
    m1 = fmaxf(m1, fmaxf(__uint_as_float(b[i]), __uint_as_float(b[i + 1])));
    if (i & 2) m0 = fmaxf(m0, fmaxf(__uint_as_float(c[i]), __uint_as_float(c[i + 1])));
    else m1 = fmaxf(m1, fmaxf(__uint_as_float(c[i]), __uint_as_float(c[i + 1])));
  }
  return fmaxf(m0, m1);
}
__device__ __forceinline__ float sum_pack(const uint32_t (&v)[32], uint32_t (&pk)[16]) {
#if TTASR_ATTN_F32X2
  uint32_t s0 = 0u, s1 = 0u;  // two packed fp32 partial sums
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    asm("{ .reg .b64 t, u;\n\t"
        "mov.b64 t, {%0, %1};\n\t"
        "mov.b64 u, {%2, %3};\n\t"
        "add.rn.f32x2 t, t, u;\n\t"
        "mov.b64 {%0, %1}, t; }"
        : "+r"(s0), "+r"(s1)
        : "r"(v[2 * i]), "r"(v[2 * i + 1]));
    pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
  }
  return __uint_as_float(s0) + __uint_as_float(s1);
#else
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float p0 = __uint_as_float(v[2 * i]), p1 = __uint_as_float(v[2 * i + 1]);
    sum0 += p0;
    sum1 += p1;
    pk[i] = pack_bf16x2(p0, p1);
  }
  return sum0 + sum1;
#endif
}

__global__ void __launch_bounds__(kAttnThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t pad = ((smem0 + 1023u) & ~1023u) - smem0;
  AttnSmem& s = *reinterpret_cast<AttnSmem*>(smem_raw + pad);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler as well
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_qkv);
    prefetch_tmap(&p.tm_out);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(smem_u32(&s.q_full), 1);
    mbar_init(smem_u32(&s.q_free), TTASR_ATTN_TWO_MMA ? 2 : 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(smem_u32(&s.kv_full[i]), 1);
      mbar_init(smem_u32(&s.kv_free[i]), TTASR_ATTN_TWO_MMA ? 2 : 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&s.s_full[t]), 1);
      mbar_init(smem_u32(&s.s_free[t]), 4);
      mbar_init(smem_u32(&s.p_ready[t]), 4);
      mbar_init(smem_u32(&s.o_done[t]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(smem_u32(&s.tmem_ptr), 512);
    tmem_relinquish<1>();
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(&s.tmem_ptr), 0);
  pdl_wait();   // the QKV GEMM's output is visible from here on (ptx_sm100.cuh: programmatic dependent launch)

  auto item_coords = [&](int item, int& b, int& h, int& q0) {
    const int qb = item % p.q_blocks;
    const int bh = item / p.q_blocks;
    h = bh % p.n_heads;
    b = bh / p.n_heads;
    q0 = qb * 2 * kTile;
  };

  // register file: 384 threads x 168 at launch; with TTASR_ATTN_SETMAXNREG the four control warps give theirs to the two
  // softmax warpgroups, whose 128 live scores + pipelined exp / pack state otherwise sit exactly at the 168 ceiling
#if TTASR_ATTN_SETMAXNREG
#define TTASR_REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 56;")
#define TTASR_REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 224;")
#else
#define TTASR_REG_DEC()
#define TTASR_REG_INC()
#endif
  if (warp == 0) {
    // ===================================================== TMA producer (whole warp walks the loop, one elected
    // lane issues: uniform control flow keeps descriptors / addresses in uniform registers)
    TTASR_REG_DEC();
    {
      int stage = 0;
      uint32_t phase = 0, qphase = 0;
      Tracer tr(3, lane == 0);
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int b, h, q0;
        item_coords(item, b, h, q0);
        ATTN_WAIT(smem_u32(&s.q_free), qphase ^ 1);
        qphase ^= 1;
        if (elect_one()) {
          mbar_arrive_expect_tx(smem_u32(&s.q_full), 2 * kTileBytes);
          tma_load_3d(smem_u32(&s.q[0][0]), &p.tm_qkv, smem_u32(&s.q_full), h * kHeadDim, q0, b);
          tma_load_3d(smem_u32(&s.q[1][0]), &p.tm_qkv, smem_u32(&s.q_full), h * kHeadDim, q0 + kTile, b);
        }
        __syncwarp();
        for (int j = 0; j < p.kv_tiles; ++j) {
          tr(50);
          ATTN_WAIT(smem_u32(&s.kv_free[stage]), phase ^ 1);
          tr(51);
          if (elect_one()) {
            const uint32_t bar = smem_u32(&s.kv_full[stage]);
            mbar_arrive_expect_tx(bar, 2 * kTileBytes);
            tma_load_3d(smem_u32(&s.k[stage][0]), &p.tm_qkv, bar, p.d_model + h * kHeadDim, j * kTile, b);
            tma_load_3d(smem_u32(&s.v[stage][0]), &p.tm_qkv, bar, 2 * p.d_model + h * kHeadDim, j * kTile, b);
          }
          __syncwarp();
          if (++stage == kKvStages) { stage = 0; phase ^= 1; }
        }
      }
    }
#if TTASR_ATTN_TWO_MMA
  } else if (warp == 1 || warp == 3) {
    // ===================================================== one MMA issuer PER QUERY TILE (experiment): warp 1 serves
    // tile 0, warp 3 tile 1, each in its own order (S_t(j+1) when S_t(j) has been read out, PV_t(j) when P_t(j) is
    // ready), so neither warpgroup's hand-overs queue behind the other's in a shared program order.  K/V stages and Q are
    // released when both issuers have committed (barrier count 2).
    TTASR_REG_DEC();
    {
      const int t = (warp == 1) ? 0 : 1;
      constexpr uint32_t idesc_s = umma_idesc_bf16(kTile, kTile, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(kTile, kHeadDim, 0, 1);
      int stage = 0;            // stage of KV tile j+1 (S look-ahead)
      uint32_t phase = 0, qphase = 0, pphase = 0, fphase = 0;
      auto issue_s = [&](int st) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(&s.q[t][0]), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(&s.k[st][0]), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k)
            umma_ss<1>(tmem_base + kColS + t * kTile, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(smem_u32(&s.s_full[t]));
        }
        __syncwarp();
      };
      auto commit = [&](unsigned long long* bar) {
        if (elect_one()) umma_commit(smem_u32(bar));
        __syncwarp();
      };
      auto issue_pv = [&](int st, bool first) {
        const uint64_t vdesc = umma_desc_sw128(smem_u32(&s.v[st][0]), kTileBytes, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k)
            umma_ts(tmem_base + kColO + t * kHeadDim, tmem_base + kColP + t * 64 + k * 8, vdesc + 128 * k, idesc_o,
                    (first && k == 0) ? 0u : 1u);
          umma_commit(smem_u32(&s.o_done[t]));
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        ATTN_WAIT(smem_u32(&s.q_full), qphase);
        qphase ^= 1;
        ATTN_WAIT(smem_u32(&s.kv_full[stage]), phase);
        tc_fence_after();
        issue_s(stage);
        if (p.kv_tiles == 1) commit(&s.q_free);
        int cur = stage;
        if (++stage == kKvStages) { stage = 0; phase ^= 1; }
        for (int j = 0; j < p.kv_tiles; ++j) {
          const bool more = (j + 1 < p.kv_tiles);
          if (more) {
            ATTN_WAIT(smem_u32(&s.kv_full[stage]), phase);
            tc_fence_after();
          }
          ATTN_WAIT(smem_u32(&s.s_free[t]), fphase);
          fphase ^= 1;
          tc_fence_after();
          if (more) {
            issue_s(stage);
            if (j + 2 == p.kv_tiles) commit(&s.q_free);  // last S of the item issued
          }
          ATTN_WAIT(smem_u32(&s.p_ready[t]), pphase);
          pphase ^= 1;
          tc_fence_after();
          issue_pv(cur, j == 0);
          commit(&s.kv_free[cur]);  // this tile's MMAs on KV tile j are issued (the other issuer commits its own)
          if (more) {
            cur = stage;
            if (++stage == kKvStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
#else
  } else if (warp == 1) {
    // ===================================================== MMA issuer (uniform control flow, one elected lane
    // issues: each tcgen05.mma is then a single UTCHMMA on precomputed uniform registers instead of a per-instruction
    // elect/waterfall sequence that starves behind the softmax warps of the same scheduler)
    TTASR_REG_DEC();
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kTile, kTile, 0, 0);      // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(kTile, kHeadDim, 0, 1);   // P (tmem)   x V (MN-major)
      int stage = 0;            // stage of KV tile j+1 (S look-ahead)
      uint32_t phase = 0;
      uint32_t qphase = 0;
      uint32_t pphase[2] = {0, 0};
      auto issue_s = [&](int t, int st) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(&s.q[t][0]), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(&s.k[st][0]), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k)
            umma_ss<1>(tmem_base + kColS + t * kTile, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(smem_u32(&s.s_full[t]));
        }
        __syncwarp();
      };
      auto commit = [&](unsigned long long* bar) {
        if (elect_one()) umma_commit(smem_u32(bar));
        __syncwarp();
      };
      uint32_t fphase[2] = {0, 0};
      Tracer tr(0, lane == 0);
      // O_t (+)= P_t V : V tile is [128 keys][64] row-major = MN-major B operand, 16 keys per MMA
      auto issue_pv = [&](int t, int st, bool first, bool last) {
        const uint64_t vdesc = umma_desc_sw128(smem_u32(&s.v[st][0]), kTileBytes, 1024);
        if (elect_one()) {
          tr(60);
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k) {
            umma_ts(tmem_base + kColO + t * kHeadDim, tmem_base + kColP + t * 64 + k * 8, vdesc + 128 * k, idesc_o,
                    (first && k == 0) ? 0u : 1u);
            if (k == 0) tr(61);
            if (k == 3) tr(62);
          }
          tr(63);
          umma_commit(smem_u32(&s.o_done[t]));  // per-PV completion: P_t consumed, O_t quiescent
          tr(64);
        }
        __syncwarp();
        tr(65);
        (void)last;
      };
      auto wait_p = [&](int t) {
        tr(30 + t);
        ATTN_WAIT(smem_u32(&s.p_ready[t]), pphase[t]);
        pphase[t] ^= 1;
        tc_fence_after();
        tr(32 + t);
      };
      auto wait_sfree = [&](int t) {
        tr(40 + t);
        ATTN_WAIT(smem_u32(&s.s_free[t]), fphase[t]);
        fphase[t] ^= 1;
        tc_fence_after();
        tr(42 + t);
      };
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        ATTN_WAIT(smem_u32(&s.q_full), qphase);
        qphase ^= 1;
        ATTN_WAIT(smem_u32(&s.kv_full[stage]), phase);
        tc_fence_after();
        issue_s(0, stage);
        issue_s(1, stage);
        if (p.kv_tiles == 1) commit(&s.q_free);
        int cur = stage;   // stage of KV tile j
        int prev = stage;  // stage of KV tile j-1
        if (++stage == kKvStages) { stage = 0; phase ^= 1; }
        // Steady-state order of events in anti-phase operation (warpgroup 0 sweeps while warpgroup 1 loads its S):
        //   S0(j) read out -> S0(j+1);  P1(j-1) ready -> PV1(j-1);  S1(j) read out -> S1(j+1);  P0(j) ready -> PV0(j)
        // The next S tile of a warpgroup is thus produced a whole exp sweep before it is needed.
        for (int j = 0; j < p.kv_tiles; ++j) {
          const bool more = (j + 1 < p.kv_tiles);
          if (more) {
            tr(44);
            ATTN_WAIT(smem_u32(&s.kv_full[stage]), phase);
            tc_fence_after();
            tr(45);
            wait_sfree(0);
            issue_s(0, stage);
          } else {
            wait_sfree(0);
          }
          if (j > 0) {
            wait_p(1);
            issue_pv(1, prev, j == 1, false);
            commit(&s.kv_free[prev]);  // every MMA that read KV tile j-1 has been issued
          }
          if (more) {
            wait_sfree(1);
            issue_s(1, stage);
            if (j + 2 == p.kv_tiles) commit(&s.q_free);  // last S of the item issued
          } else {
            wait_sfree(1);
          }
          wait_p(0);
          issue_pv(0, cur, j == 0, !more);
          prev = cur;
          if (more) {
            cur = stage;
            if (++stage == kKvStages) { stage = 0; phase ^= 1; }
          }
        }
        wait_p(1);
        issue_pv(1, prev, p.kv_tiles == 1, true);
        commit(&s.kv_free[prev]);
      }
    }
#endif  // TTASR_ATTN_TWO_MMA
  } else if (warp >= 4) {
    // ===================================================== softmax + output, one thread per query row
    TTASR_REG_INC();
    const int t = (warp - 4) >> 2;          // query tile owned by this warpgroup
    const int wq = warp & 3;                // TMEM lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + kColS + t * kTile;
    const uint32_t p_addr = tmem_base + lane_base + kColP + t * 64;
    const uint32_t o_addr = tmem_base + lane_base + kColO + t * kHeadDim;
    const bool leader = (wq == 0 && lane == 0);
    // Token: warpgroup t may start its exp sweep (the other one has issued its last exponential).  The exp pipe is per
    // scheduler, and warp wq of either warpgroup sits on scheduler wq — so with TTASR_ATTN_SMSP_TOKEN the token is
    // passed between the two warps of each scheduler on their own (64-thread named barriers) instead of between whole
    // warpgroups (256 threads: every hand-over then waits for the slowest of the four schedulers).
#if TTASR_ATTN_SMSP_TOKEN
    const uint32_t kTokBar = 2 + 2 * wq;    // +t; ids 2..9
    constexpr uint32_t kTokThreads = 64;
    constexpr uint32_t kEpiBar = 10;        // +t: warpgroup-local barrier of the output staging
#else
    constexpr uint32_t kTokBar = 2;         // +t
    constexpr uint32_t kTokThreads = 256;
    constexpr uint32_t kEpiBar = 4;         // +t: warpgroup-local barrier of the output staging
#endif
    uint32_t sphase = 0, ophase = 0;
    Tracer tr(1 + t, wq == 0 && lane == 0);
    const int last_valid = p.n_ctx - (p.kv_tiles - 1) * kTile;  // valid keys in the last KV tile
    const bool last_masked = last_valid < kTile;
    if (t == 1) asm volatile("bar.arrive %0, %1;" ::"r"(kTokBar + 0), "r"(kTokThreads) : "memory");  // warpgroup 0 sweeps first
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int b, h, q0;
      item_coords(item, b, h, q0);
      float m_used = 0.f;  // base of the exponentials (log2 domain); lags the true row max by <= 2^32
      float l = 0.f;
      for (int j = 0; j < p.kv_tiles; ++j) {
        const bool masked = last_masked && (j == p.kv_tiles - 1);
        tr(10);
        ATTN_WAIT(smem_u32(&s.s_full[t]), sphase);
        sphase ^= 1;
        tc_fence_after();
        tr(11);
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld_32x32(s_addr, v0);
#if TTASR_ATTN_ABLATE & 4
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) { v1[i] = v0[i] ^ 1u; v2[i] = v0[i] ^ 2u; v3[i] = v0[i] ^ 3u; }
#else
        tmem_ld_32x32(s_addr + 32, v1);
        tmem_ld_32x32(s_addr + 64, v2);
        tmem_ld_32x32(s_addr + 96, v3);
        tmem_wait_ld();
#endif
        tr(12);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.s_free[t]));  // S_t is in registers: the next S_t may be produced now
        if (masked) {  // last KV tile only: keys >= n_ctx become -inf once, so max and sweep stay branch-free
          mask_tail(v0, 0, last_valid);
          mask_tail(v1, 32, last_valid);
          mask_tail(v2, 64, last_valid);
          mask_tail(v3, 96, last_valid);
        }
        // PV_t(j-1) (issued a sweep ago) must have consumed P_t and left O_t quiescent before either is written again.
        bool o_waited = (j == 0);
        auto wait_o = [&]() {
          if (!o_waited) {
            ATTN_WAIT(smem_u32(&s.o_done[t]), ophase);
            ophase ^= 1;
            tc_fence_after();
            o_waited = true;
          }
        };
        // move the exponent base to m_new: l, O (and, with fix_q0, the already exponentiated quarter 0) shrink by 2^(old - new)
        auto rebase = [&](float m_new, bool fix_q0) {
          const float factor = ex2(m_used - m_new);
          m_used = m_new;
          l *= factor;
          if (fix_q0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(__uint_as_float(v0[i]) * factor);
          }
          if (j > 0) {
            wait_o();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              uint32_t o[32];
              tmem_ld_32x32(o_addr + c * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
              tmem_st_32x32(o_addr + c * 32, o);
            }
          }
        };
#if TTASR_ATTN_LATEMAX
        // quarter 0's own maximum decides whether its exponentials may run against the current base (never above
        // 2^threshold); the rest of the row maximum is taken beside them (exp_inplace_max3) and checked afterwards
        const float m_q0 = chunk_max(v0) * kLog2e;
        tr(13);
        if (j == 0) {
          m_used = m_q0;
        } else if (__any_sync(0xffffffffu, (m_q0 - m_used) > kRescaleThreshold)) {  // rare
          rebase(fmaxf(m_used, m_q0), false);
        }
        tr(14);
#else
#if TTASR_ATTN_ABLATE & 1
        const float mx = __uint_as_float(v0[0]);
#else
        const float mx = fmaxf(fmaxf(chunk_max(v0), chunk_max(v1)), fmaxf(chunk_max(v2), chunk_max(v3)));
#endif
        const float m_tile = mx * kLog2e;
        tr(13);
        // (Deferring this wait into the sweep was measured slower: it then sits inside the token-exclusive section.)
        wait_o();
        tr(14);
        if (j == 0) {
          m_used = m_tile;
        } else if (__any_sync(0xffffffffu, (m_tile - m_used) > kRescaleThreshold)) {  // rare
          rebase(fmaxf(m_used, m_tile), false);
        }
#endif
        // ---- exp sweep.  The first kPreTokenChunks quarter(s) run beside the other warpgroup's sweep (one warp per
        // scheduler cannot quite saturate the 16-op/clk exp pipe); the rest is exclusive (token), so the pipe never
        // idles while this warpgroup waits for / reads its next S tile.  The exponentials go back over the scores in
        // place; the row sum / bf16 packing / P store of a quarter follow the exponentials of the next one.
        uint32_t pk[16];
        float lsum = 0.f;
        // TTASR_ATTN_POLY_Q<c>: how many of every 8 exponentials of quarter c run on the FMA pipe (polynomial)
        auto stage_a = [&](uint32_t (&v)[32], int c) {
          if (c == 0) exp_inplace<TTASR_ATTN_POLY_Q0>(v, m_used);
          else if (c == 1) exp_inplace<TTASR_ATTN_POLY_Q1>(v, m_used);
          else if (c == 2) exp_inplace<TTASR_ATTN_POLY_Q2>(v, m_used);
          else exp_inplace<TTASR_ATTN_POLY_Q3>(v, m_used);
        };
        auto stage_b = [&](const uint32_t (&v)[32], int c) {
#if TTASR_ATTN_ABLATE & 8        // no bf16 conversion (one LOP3 per pair keeps every exponential live), sums kept
          {
            uint32_t s0 = 0u, s1 = 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              asm("{ .reg .b64 t, u;\n\tmov.b64 t, {%0, %1};\n\tmov.b64 u, {%2, %3};\n\tadd.rn.f32x2 t, t, u;\n\tmov.b64 {%0, %1}, t; }"
                  : "+r"(s0), "+r"(s1) : "r"(v[2 * i]), "r"(v[2 * i + 1]));
              pk[i] = v[2 * i] ^ v[2 * i + 1];
            }
            lsum += __uint_as_float(s0) + __uint_as_float(s1);
          }
#elif TTASR_ATTN_ABLATE & 16     // no row sums, conversion kept
          {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            lsum += __uint_as_float(pk[c]);
          }
#else
          lsum += sum_pack(v, pk);
#endif
#if !(TTASR_ATTN_ABLATE & 2)
          tmem_st_32x16(p_addr + 16 * c, pk);
#else
          if (c == 3) tmem_st_32x16(p_addr + 16 * c, pk);
#endif
        };
        auto tok_acquire = [&]() { tr(15); asm volatile("bar.sync %0, %1;" ::"r"(kTokBar + t), "r"(kTokThreads) : "memory"); tr(16); };
        if (kPreTokenChunks == 0) tok_acquire();
#if TTASR_ATTN_LATEMAX
        {
          const float m_rest = exp_inplace_max3<TTASR_ATTN_POLY_Q0>(v0, m_used, v1, v2, v3) * kLog2e;
          if (__any_sync(0xffffffffu, (m_rest - m_used) > kRescaleThreshold)) rebase(fmaxf(m_used, m_rest), true);  // rare
          tr(19);
          wait_o();   // long complete by now: P_t may be overwritten from here on
        }
#else
        stage_a(v0, 0);
#endif
        if (kPreTokenChunks == 1) tok_acquire();
        stage_a(v1, 1);
        stage_b(v0, 0);
        if (kPreTokenChunks == 2) tok_acquire();
        stage_a(v2, 2);
        stage_b(v1, 1);
        if (kPreTokenChunks == 3) tok_acquire();
#if TTASR_ATTN_EARLY_RELEASE
        // hand the token over one quarter early: the other warpgroup's wake-up then overlaps the tail of this sweep
        asm volatile("bar.arrive %0, %1;" ::"r"(kTokBar + (1 - t)), "r"(kTokThreads) : "memory");
        stage_a(v3, 3);
#else
        stage_a(v3, 3);
        if (kPreTokenChunks == 4) tok_acquire();
        asm volatile("bar.arrive %0, %1;" ::"r"(kTokBar + (1 - t)), "r"(kTokThreads) : "memory");  // last exp issued
#endif
        stage_b(v2, 2);
        stage_b(v3, 3);
        l += lsum;
        tr(17);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.p_ready[t]));
        tr(18);
      }
      // ---- epilogue: O / l -> bf16 -> swizzled smem -> TMA store
      tr(20);
      ATTN_WAIT(smem_u32(&s.o_done[t]), ophase);
      tr(21);
      ophase ^= 1;
      tc_fence_after();
      if (leader) tma_store_wait_read<0>();  // staging tile of the previous item has been read out
      bar_sync(kEpiBar + t, 128);
      const float inv = 1.0f / l;
      const uint32_t o_row = smem_u32(&s.o[t][0]) + row * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(o_addr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = c * 4 + q;  // 16-byte chunk = 8 channels
          const uint32_t a0 = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
          const uint32_t a1 = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
          const uint32_t a2 = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
          const uint32_t a3 = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o_row + ((chunk ^ (row & 7)) << 4)), "r"(a0),
                       "r"(a1), "r"(a2), "r"(a3)
                       : "memory");
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      bar_sync(kEpiBar + t, 128);
      if (leader) {
        tma_store_3d(&p.tm_out, smem_u32(&s.o[t][0]), h * kHeadDim, q0 + t * kTile, b);
        tma_store_commit();
      }
      tr(22);
    }
    if (t == 0) asm volatile("bar.sync %0, %1;" ::"r"(kTokBar + 0), "r"(kTokThreads) : "memory");  // absorb the last token
    if (leader) tma_store_wait<0>();
  } else {
    TTASR_REG_DEC();  // warps 2-3: setmaxnreg is warpgroup-wide
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

#if TTASR_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int ttasr_debug_attention_trace(long long* buf_dev, int cap) {
  cudaError_t e = cudaMemcpyToSymbol(g_trace_buf, &buf_dev, sizeof(buf_dev));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
  return e == cudaSuccess ? 0 : -1;
}
#endif

// TTASR_ATTN_KERNEL selects the kernel: "2wg" (this file: two softmax warpgroups, 128-key tiles, exp sweeps in
// anti-phase) or "4wg" (attention4_sm100.cu: four softmax warpgroups alternating 64-key steps).
#ifndef TTASR_ATTN_DEFAULT_VARIANT
#define TTASR_ATTN_DEFAULT_VARIANT 2
#endif
int attention_variant() {
  static int v = [] {
    const char* e = getenv("TTASR_ATTN_KERNEL");
    if (e && !strcmp(e, "4wg")) return 4;
    if (e && !strcmp(e, "2wg")) return 2;
    return TTASR_ATTN_DEFAULT_VARIANT;
  }();
  return v;
}

cudaError_t attention_launch(const void* qkv, void* out, int batch, int n_ctx, int n_heads, int num_sms,
                             cudaStream_t stream, const char** why) {
  static const char* dummy;
  if (!why) why = &dummy;
  *why = nullptr;
  if (!qkv || !out) { *why = "attention: null operand"; return cudaErrorInvalidValue; }
  if (batch <= 0 || n_ctx <= 0 || n_heads <= 0) { *why = "attention: empty problem"; return cudaErrorInvalidValue; }
  if (attention_variant() == 4) return attention4_launch(qkv, out, batch, n_ctx, n_heads, num_sms, stream, why);
  const int d = n_heads * kHeadDim;
  AttnParams p{};
  p.n_ctx = n_ctx;
  p.n_heads = n_heads;
  p.d_model = d;
  p.q_blocks = (n_ctx + 2 * kTile - 1) / (2 * kTile);
  p.kv_tiles = (n_ctx + kTile - 1) / kTile;
  p.num_items = batch * n_heads * p.q_blocks;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(n_ctx), static_cast<uint64_t>(batch)};
    uint64_t str[2] = {dims[0] * 2, dims[0] * dims[1] * 2};
    uint32_t box[3] = {kHeadDim, kTile, 1};
    if (encode_tmap(&p.tm_qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B) !=
        CUDA_SUCCESS) {
      *why = "attention: cuTensorMapEncodeTiled(qkv) failed";
      return cudaErrorInvalidValue;
    }
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(n_ctx), static_cast<uint64_t>(batch)};
    uint64_t str[2] = {dims[0] * 2, dims[0] * dims[1] * 2};
    uint32_t box[3] = {kHeadDim, kTile, 1};
    if (encode_tmap(&p.tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B) !=
        CUDA_SUCCESS) {
      *why = "attention: cuTensorMapEncodeTiled(out) failed";
      return cudaErrorInvalidValue;
    }
  }
  const int smem = static_cast<int>(sizeof(AttnSmem)) + 1024;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  const int grid = p.num_items < num_sms ? p.num_items : num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_launch_attr(&attr[0]);
  return cudaLaunchKernelEx(&cfg, attention_kernel, p);
}

}  // namespace ttasr
