// Non-causal multi-head attention forward for the Whisper encoder on sm_100a, FOUR softmax warpgroups per CTA
// (head_dim 64, no mask, scale folded into q):  out = softmax(q k^T) v  per (chunk, head); reference
// eager_attention_forward / WhisperAttention.forward (transformers/models/whisper/modeling_whisper.py:215-238,284-357).
//
// Why a second kernel: at head_dim 64 the exponentials (16 MUFU.EX2 per clock per SM) cost twice the tensor time of a
// score tile, so the kernel is bound by how well the exp pipe is fed.  The two-warpgroup kernel (attention_sm100.cu)
// keeps it 77 % busy: each warpgroup walks a serial chain per tile (wait S, load, row max, wait PV, sweep, hand P over)
// and two chains cannot cover each other's ~1500 cycles of non-exp work.  Here sixteen softmax warps (four per scheduler)
// work on four independent chains, so the exp pipe always finds a warp with exponentials to issue:
//
//   work item = (chunk, head, 256 query rows) = two 128-row query tiles t = 0, 1
//   keys are walked in steps of 64; step j of tile t belongs to warpgroup (t, j & 1): the two warpgroups of a tile
//   ALTERNATE steps, so while one is in its sweep the other already holds the next scores
//   warp 0       TMA producer : Q tiles once per item; K/V 64 x 64 tiles through an 8-stage ring (128B swizzle)
//   warp 1 / 2   MMA issuer of tile 0 / 1: S = Q_t K_j^T (SS MMA 128x64x64, fp32 in TMEM) into the buffer of the
//                step's warpgroup; O_t += P V_j (TS MMA: P read from TMEM, V as an MN-major smem operand)
//   warp 3       TMEM allocator (512 columns: four S/P buffers of 64 | O0 O1 of 64 | spare)
//   warps 4-19   softmax, one thread per query row: tcgen05.ld of the 64 scores, row max, exp2 (MUFU, optionally a
//                share on the FMA pipe by polynomial), fp32 row sum, P written as packed bf16 OVER the scores (the
//                buffer is private to the warpgroup: the next S lands in it only after the PV that read P completed)
//   Row state: the exponent base m of a row is shared by the tile's two warpgroups through shared memory, published
//   step by step in step order (an mbarrier per parity); each warpgroup keeps its own partial row sum relative to the
//   base it last adopted, and the two partials are merged at the end of the item.  O is rescaled only when the running
//   max outgrows the base by > 2^32 (rare), by the warpgroup that sees it, after the PV of the previous step completed.
// The 1500 x 1500 score matrix never leaves the SM; keys beyond n_ctx in the last step are masked to -inf.
#include "attention_sm100.h"
#include "gemm_sm100.h"  // encode_tmap
#include "ptx_sm100.cuh"

namespace ttasr {
namespace {

constexpr int kQTile = 128;          // query rows per tile
constexpr int kStep = 64;            // keys per step
constexpr int kHeadDim = 64;
constexpr int kQBytes = kQTile * kHeadDim * 2;    // 16 KB
constexpr int kKvBytes = kStep * kHeadDim * 2;    // 8 KB per K or V tile
constexpr int kStages = 8;
constexpr int kThreads = 640;        // 4 control warps + 16 softmax warps
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 32.0f;

#ifndef TTASR_ATTN4_POLY8
#define TTASR_ATTN4_POLY8 0          // of every 8 exponentials, how many run on the FMA pipe (even; 0 = all MUFU)
#endif

// TMEM column map: S/P buffer of warpgroup (t, parity) at 64 * (2 t + parity); O_t at 256 + 64 t
constexpr uint32_t kColSP = 0;
constexpr uint32_t kColO = 256;

struct Attn4Params {
  CUtensorMap tm_q;    // 3-D (3*d, n_ctx, batch), box (64, 128, 1)
  CUtensorMap tm_kv;   // 3-D (3*d, n_ctx, batch), box (64, 64, 1)
  CUtensorMap tm_out;  // 3-D (d, n_ctx, batch), box (64, 128, 1)
  int n_ctx, n_heads, d_model;
  int q_blocks;    // ceil(n_ctx / 256)
  int steps;       // ceil(n_ctx / 64)
  int num_items;   // batch * heads * q_blocks
};

struct Attn4Smem {
  uint8_t q[2][kQBytes];
  uint8_t k[kStages][kKvBytes];
  uint8_t v[kStages][kKvBytes];
  uint8_t o[2][kQBytes];
  float m_row[2][kQTile];        // shared exponent base of each query row (log2 domain)
  float l_part[2][2][kQTile];    // partial row sums of the two warpgroups of a tile, relative to the final base
  unsigned long long q_full, q_free;
  unsigned long long kv_full[kStages], kv_free[kStages];
  unsigned long long s_full[2][2], p_ready[2][2], pv_done[2][2], m_pub[2][2];
  unsigned long long o_free[2];
  uint32_t tmem_ptr;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair on the FMA / ALU pipes (see attention_sm100.cu: exp2_poly_pair)
__device__ __forceinline__ void exp2_poly_pair4(float& a, float& b) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
  const float x0 = fmaxf(a, -126.0f), x1 = fmaxf(b, -126.0f);
  const f32x2_t x = pack2(x0, x1);
  const f32x2_t t = add2(x, pack2(kMagic, kMagic));
  const f32x2_t n = add2(t, pack2(-kMagic, -kMagic));
  const f32x2_t f = add2(x, n ^ 0x8000000080000000ull);
  f32x2_t p = fma2(pack2(0.05520551f, 0.05520551f), f, pack2(0.24261396f, 0.24261396f));
  p = fma2(p, f, pack2(0.69325476f, 0.69325476f));
  p = fma2(p, f, pack2(0.99992773f, 0.99992773f));
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  a = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  b = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

__global__ void __launch_bounds__(kThreads, 1) attention4_kernel(const __grid_constant__ Attn4Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t pad = ((smem0 + 1023u) & ~1023u) - smem0;
  Attn4Smem& s = *reinterpret_cast<Attn4Smem*>(smem_raw + pad);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler as well
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_q);
    prefetch_tmap(&p.tm_kv);
    prefetch_tmap(&p.tm_out);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(smem_u32(&s.q_full), 1);
    mbar_init(smem_u32(&s.q_free), 2);            // both MMA warps have issued their last S of the item
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&s.kv_full[i]), 1);
      mbar_init(smem_u32(&s.kv_free[i]), 2);      // both tiles' MMAs on the stage have been issued
    }
    for (int t = 0; t < 2; ++t) {
      for (int w = 0; w < 2; ++w) {
        mbar_init(smem_u32(&s.s_full[t][w]), 1);
        mbar_init(smem_u32(&s.p_ready[t][w]), 4);
        mbar_init(smem_u32(&s.pv_done[t][w]), 1);
        mbar_init(smem_u32(&s.m_pub[t][w]), 4);
      }
      mbar_init(smem_u32(&s.o_free[t]), 8);       // the 8 warps of the tile have read O of the finished item
    }
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc<1>(smem_u32(&s.tmem_ptr), 512);
    tmem_relinquish<1>();
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(&s.tmem_ptr), 0);
  pdl_wait();

  auto item_coords = [&](int item, int& b, int& h, int& q0) {
    const int qb = item % p.q_blocks;
    const int bh = item / p.q_blocks;
    h = bh % p.n_heads;
    b = bh / p.n_heads;
    q0 = qb * 2 * kQTile;
  };
  const int steps = p.steps;

  if (warp == 0) {
    // ===================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    int stage = 0;
    uint32_t phase = 0, qphase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int b, h, q0;
      item_coords(item, b, h, q0);
      mbar_wait(smem_u32(&s.q_free), qphase ^ 1);
      qphase ^= 1;
      if (elect_one()) {
        mbar_arrive_expect_tx(smem_u32(&s.q_full), 2 * kQBytes);
        tma_load_3d(smem_u32(&s.q[0][0]), &p.tm_q, smem_u32(&s.q_full), h * kHeadDim, q0, b);
        tma_load_3d(smem_u32(&s.q[1][0]), &p.tm_q, smem_u32(&s.q_full), h * kHeadDim, q0 + kQTile, b);
      }
      __syncwarp();
      for (int j = 0; j < steps; ++j) {
        mbar_wait(smem_u32(&s.kv_free[stage]), phase ^ 1);
        if (elect_one()) {
          const uint32_t bar = smem_u32(&s.kv_full[stage]);
          mbar_arrive_expect_tx(bar, 2 * kKvBytes);
          tma_load_3d(smem_u32(&s.k[stage][0]), &p.tm_kv, bar, p.d_model + h * kHeadDim, j * kStep, b);
          tma_load_3d(smem_u32(&s.v[stage][0]), &p.tm_kv, bar, 2 * p.d_model + h * kHeadDim, j * kStep, b);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================================================== MMA issuer of query tile t (uniform control flow, one
    // elected lane issues: every tcgen05.mma is then one UTCHMMA on uniform registers)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int t = warp - 1;
    constexpr uint32_t idesc_s = umma_idesc_bf16(kQTile, kStep, 0, 0);      // Q (K-major) x K (K-major) -> 128 x 64
    constexpr uint32_t idesc_o = umma_idesc_bf16(kQTile, kHeadDim, 0, 1);   // P (tmem)   x V (MN-major) -> 128 x 64
    int stage = 0;
    uint32_t phase = 0, qphase = 0, ophase = 0;
    uint32_t pv_phase[2] = {0, 0};   // phase of the NEXT completion of pv_done[t][w] this warp will wait for
    uint32_t pr_phase[2] = {0, 0};
    uint32_t used[2] = {0, 0};       // has buffer w been handed to a PV already (then its completion gates the next S)
    bool first_item = true;
    auto issue_s = [&](int w, int st) {
      const uint64_t adesc = umma_desc_sw128(smem_u32(&s.q[t][0]), 16, 1024);
      const uint64_t bdesc = umma_desc_sw128(smem_u32(&s.k[st][0]), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss<1>(tmem_base + kColSP + 64 * (2 * t + w), adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(smem_u32(&s.s_full[t][w]));
      }
      __syncwarp();
    };
    // O_t (+)= P_w V : V tile is [64 keys][64] row-major = MN-major B operand, 16 keys per MMA
    auto issue_pv = [&](int w, int st, bool first) {
      const uint64_t vdesc = umma_desc_sw128(smem_u32(&s.v[st][0]), kKvBytes, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kStep / 16; ++k)
          umma_ts(tmem_base + kColO + 64 * t, tmem_base + kColSP + 64 * (2 * t + w) + k * 8, vdesc + 128 * k, idesc_o,
                  (first && k == 0) ? 0u : 1u);
        umma_commit(smem_u32(&s.pv_done[t][w]));   // P_w consumed (its buffer may take the next S), O_t advanced
        umma_commit(smem_u32(&s.kv_free[st]));     // every MMA of this tile on the stage has been issued
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      mbar_wait(smem_u32(&s.q_full), qphase);
      qphase ^= 1;
      int prev_stage = stage;
      for (int j = 0; j < steps; ++j) {
        const int w = j & 1;
        mbar_wait(smem_u32(&s.kv_full[stage]), phase);
        if (used[w]) {  // the PV that read P_w (aliased over S_w) must have completed before S_w is overwritten
          mbar_wait(smem_u32(&s.pv_done[t][w]), pv_phase[w]);
          pv_phase[w] ^= 1;
          used[w] = 0;
        }
        tc_fence_after();
        issue_s(w, stage);
        if (j + 1 == steps) {   // last S of the item issued: Q may be replaced
          if (elect_one()) umma_commit(smem_u32(&s.q_free));
          __syncwarp();
        }
        if (j > 0) {
          const int wp = w ^ 1;
          mbar_wait(smem_u32(&s.p_ready[t][wp]), pr_phase[wp]);
          pr_phase[wp] ^= 1;
          if (j == 1 && !first_item) {   // first PV of the item overwrites O_t: the epilogue of the previous item has read it
            mbar_wait(smem_u32(&s.o_free[t]), ophase);
            ophase ^= 1;
          }
          tc_fence_after();
          issue_pv(wp, prev_stage, j == 1);
          used[wp] = 1;
        }
        prev_stage = stage;
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      {
        const int wp = (steps - 1) & 1;
        mbar_wait(smem_u32(&s.p_ready[t][wp]), pr_phase[wp]);
        pr_phase[wp] ^= 1;
        if (steps == 1 && !first_item) {
          mbar_wait(smem_u32(&s.o_free[t]), ophase);
          ophase ^= 1;
        }
        tc_fence_after();
        issue_pv(wp, prev_stage, steps == 1);
        used[wp] = 1;
      }
      first_item = false;
    }
  } else if (warp >= 4) {
    // ===================================================== softmax + output, one thread per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int wg = (warp - 4) >> 2;
    const int t = wg >> 1;                  // query tile
    const int pi = wg & 1;                  // parity of the steps this warpgroup owns
    const int wq = warp & 3;                // TMEM lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t sp_addr = tmem_base + lane_base + kColSP + 64 * (2 * t + pi);
    const uint32_t o_addr = tmem_base + lane_base + kColO + 64 * t;
    const uint32_t kEpiBar = 2 + t;         // named barrier of the tile's 256 softmax threads (item epilogue only)
    uint32_t sphase = 0, mphase = 0;        // s_full[t][pi]; partner's m_pub[t][pi ^ 1]
    const int n_pv0 = (steps + 1) / 2, n_pv1 = steps / 2;   // PVs (= completions of pv_done[t][w]) per item, by parity
    const int n_pv_partner = pi ? n_pv0 : n_pv1;
    int it = 0;                             // items this CTA has finished
    const int last_valid = p.n_ctx - (steps - 1) * kStep;   // valid keys in the last step
    const bool leader = (pi == 0 && wq == 0 && lane == 0);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int b, h, q0;
      item_coords(item, b, h, q0);
      float m_w = 0.f;   // base of this warpgroup's partial sum (log2 domain); set at its first step
      float l = 0.f;
      for (int j = pi; j < steps; j += 2) {
        mbar_wait(smem_u32(&s.s_full[t][pi]), sphase);
        sphase ^= 1;
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(sp_addr, v0);
        tmem_ld_32x32(sp_addr + 32, v1);
        tmem_wait_ld();
        if (j == steps - 1 && last_valid < kStep) {  // keys >= n_ctx become -inf once: max and sweep stay branch-free
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= last_valid) v0[i] = 0xff800000u;
            if (32 + i >= last_valid) v1[i] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v0[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v1[i]), __uint_as_float(v1[i + 1])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v1[i + 2]), __uint_as_float(v1[i + 3])));
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * kLog2e;
        // ---- the row's exponent base, in step order: adopt what the partner published for step j - 1
        float m_used;
        float o_factor = 1.0f;
        bool rescale_o = false;
        if (j == 0) {
          m_used = m_tile;
        } else {
          mbar_wait(smem_u32(&s.m_pub[t][pi ^ 1]), mphase);
          mphase ^= 1;
          m_used = s.m_row[t][row];
          if (__any_sync(0xffffffffu, (m_tile - m_used) > kRescaleThreshold)) {  // rare: the base moves up
            const float m_new = fmaxf(m_used, m_tile);
            o_factor = ex2f(m_used - m_new);
            m_used = m_new;
            rescale_o = true;
          }
        }
        s.m_row[t][row] = m_used;
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.m_pub[t][pi]));
        if (j == pi) {
          m_w = m_used;              // first step of this warpgroup in the item: l = 0, any base will do
        } else if (m_used != m_w) {  // this warpgroup's partial sum moves to the new base
          l *= ex2f(m_w - m_used);
          m_w = m_used;
        }
        // ---- exp sweep in place over the scores, P packed over the S buffer
        const f32x2_t sc = pack2(kLog2e, kLog2e), sh = pack2(-m_used, -m_used);
        f32x2_t sum2 = pack2(0.f, 0.f);
        uint32_t pk[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t (&v)[32] = half == 0 ? v0 : v1;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float a, bb;
              unpack2(fma2(pack2(__uint_as_float(v[16 * c + i]), __uint_as_float(v[16 * c + i + 1])), sc, sh), a, bb);
              if ((i & 7) < TTASR_ATTN4_POLY8) {
                exp2_poly_pair4(a, bb);
              } else {
                a = ex2f(a);
                bb = ex2f(bb);
              }
              sum2 = add2(sum2, pack2(a, bb));
              pk[8 * c + (i >> 1)] = pack_bf16x2(a, bb);
            }
          }
          // (the PV of step j - 2 read P from this buffer: it completed before the S of this step was issued, so the
          // store below cannot overtake it)
          tmem_st_32x16(sp_addr + 16 * half, pk);
        }
        float s0, s1;
        unpack2(sum2, s0, s1);
        l += s0 + s1;
        if (rescale_o) {
          // O_t holds the sum over steps < j in the OLD base: wait until the PV of step j - 1 (the partner's) has been
          // accumulated (ours of step j - 2 completed before this step's S was issued), then rescale our row of O_t
          // (completion number it * n + (j - 1) / 2 of that barrier; the one before it is known complete: it gated the
          // S of step j - 1, which the partner had loaded before it published the base this step has adopted)
          mbar_wait(smem_u32(&s.pv_done[t][pi ^ 1]), static_cast<uint32_t>(it * n_pv_partner + ((j - 1) >> 1)) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(o_addr + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * o_factor);
            tmem_st_32x32(o_addr + c * 32, o);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.p_ready[t][pi]));
      }
      // ---- end of the item: merge the two partial sums on the final base, O / l -> bf16 -> swizzled smem -> TMA store
      // (each warpgroup converts its half of the 64 output channels)
      const int last_pi = (steps - 1) & 1;
      const int my_steps = (steps - pi + 1) / 2;
      if (pi != last_pi) {   // the final base is the one the partner publishes for the last step (also keeps the phases of
                             // m_pub in step: every publication is consumed exactly once)
        mbar_wait(smem_u32(&s.m_pub[t][pi ^ 1]), mphase);
        mphase ^= 1;
      }
      const float m_final = s.m_row[t][row];
      s.l_part[t][pi][row] = my_steps > 0 ? l * ex2f(m_w - m_final) : 0.f;
      // every PV of the item has completed: wait for the last one of each parity (the one before it is known complete:
      // it gated the last S of that parity, whose step has published its base by now)
      mbar_wait(smem_u32(&s.pv_done[t][0]), static_cast<uint32_t>(it * n_pv0 + n_pv0 - 1) & 1u);
      if (n_pv1 > 0) mbar_wait(smem_u32(&s.pv_done[t][1]), static_cast<uint32_t>(it * n_pv1 + n_pv1 - 1) & 1u);
      ++it;
      tc_fence_after();
      if (leader) tma_store_wait_read<0>();  // staging tile of the previous item has been read out
      bar_sync(kEpiBar, 256);                // partials visible; staging free
      const float inv = 1.0f / (s.l_part[t][0][row] + s.l_part[t][1][row]);
      const uint32_t o_row = smem_u32(&s.o[t][0]) + row * 128;
      {
        uint32_t v[32];
        tmem_ld_32x32(o_addr + pi * 32, v);   // this warpgroup's 32 of the 64 channels
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s.o_free[t]));   // O_t has been read: the next item's PVs may overwrite it
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const int chunk = pi * 4 + qd;  // 16-byte chunk = 8 channels
          const uint32_t a0 = pack_bf16x2(__uint_as_float(v[8 * qd + 0]) * inv, __uint_as_float(v[8 * qd + 1]) * inv);
          const uint32_t a1 = pack_bf16x2(__uint_as_float(v[8 * qd + 2]) * inv, __uint_as_float(v[8 * qd + 3]) * inv);
          const uint32_t a2 = pack_bf16x2(__uint_as_float(v[8 * qd + 4]) * inv, __uint_as_float(v[8 * qd + 5]) * inv);
          const uint32_t a3 = pack_bf16x2(__uint_as_float(v[8 * qd + 6]) * inv, __uint_as_float(v[8 * qd + 7]) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o_row + ((chunk ^ (row & 7)) << 4)), "r"(a0),
                       "r"(a1), "r"(a2), "r"(a3)
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      bar_sync(kEpiBar, 256);
      if (leader) {
        tma_store_3d(&p.tm_out, smem_u32(&s.o[t][0]), h * kHeadDim, q0 + t * kQTile, b);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait<0>();
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // warp 3: setmaxnreg is warpgroup-wide
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

cudaError_t attention4_launch(const void* qkv, void* out, int batch, int n_ctx, int n_heads, int num_sms,
                              cudaStream_t stream, const char** why) {
  const int d = n_heads * kHeadDim;
  Attn4Params p{};
  p.n_ctx = n_ctx;
  p.n_heads = n_heads;
  p.d_model = d;
  p.q_blocks = (n_ctx + 2 * kQTile - 1) / (2 * kQTile);
  p.steps = (n_ctx + kStep - 1) / kStep;
  p.num_items = batch * n_heads * p.q_blocks;
  uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(n_ctx), static_cast<uint64_t>(batch)};
  uint64_t str[2] = {dims[0] * 2, dims[0] * dims[1] * 2};
  uint32_t box_q[3] = {kHeadDim, kQTile, 1}, box_kv[3] = {kHeadDim, kStep, 1};
  if (encode_tmap(&p.tm_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, str, box_q, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS ||
      encode_tmap(&p.tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, str, box_kv, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS) {
    *why = "attention: cuTensorMapEncodeTiled(qkv) failed";
    return cudaErrorInvalidValue;
  }
  {
    uint64_t odims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(n_ctx), static_cast<uint64_t>(batch)};
    uint64_t ostr[2] = {odims[0] * 2, odims[0] * odims[1] * 2};
    if (encode_tmap(&p.tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, odims, ostr, box_q, CU_TENSOR_MAP_SWIZZLE_128B) !=
        CUDA_SUCCESS) {
      *why = "attention: cuTensorMapEncodeTiled(out) failed";
      return cudaErrorInvalidValue;
    }
  }
  const int smem = static_cast<int>(sizeof(Attn4Smem)) + 1024;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    cudaError_t e = cudaFuncSetAttribute(attention4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  const int grid = p.num_items < num_sms ? p.num_items : num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_launch_attr(&attr[0]);
  return cudaLaunchKernelEx(&cfg, attention4_kernel, p);
}

}  // namespace ttasr
