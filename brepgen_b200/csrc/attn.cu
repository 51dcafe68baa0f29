// tcgen05 flash attention for the denoisers' self-attention (12 heads x 64, key-padding mask).
//
// Reference semantics: nn.MultiheadAttention inside nn.TransformerEncoderLayer with src_key_padding_mask
// (/root/reference/network.py:1119-1123, 1193-1197, 1279-1283, 1387-1390): softmax(q k^T / 8 + (-inf on padded keys)) v.
// Edge stages run ONE sequence of L = faces*edges <= 4000 tokens per sample (network.py:1265-1283), so this is an
// online-softmax (flash) kernel; the surface stages (L <= 100) use the same kernel with one key block.
//
// CTA = NT query tiles of 128 rows for one (sample, head):
//   warps [0, 4*NT)   : softmax warpgroups, one per query tile; thread r owns query row r == TMEM lane r
//   warp 4*NT         : TMA producer (Q once; K / V 128-key tiles through an ST-deep mbarrier ring)
//   warps 4*NT+1 ..   : one MMA-issuing thread per query tile (the first of these warps owns the TMEM allocation)
//   warps 12, 13      : (NT == 2) helper warps: wait on the S-ready / PV-done mbarriers ahead of time and release the
//                       softmax warpgroup of their tile through named barriers (an mbarrier probe costs ~100-200 cycles on
//                       the softmax critical path even when the phase completed long ago, a named-barrier sync ~15)
// per key block j and tile t:   S_t = Q_t K_j^T            4 x tcgen05.mma M128 N128 K16  (A,B K-major SW128)
//                               P_t = exp2(c (S_t - m))     softmax WG: TMEM -> regs -> fp16 -> swizzled smem
//                               O_t (+)= P_t V_j            2 x 4 tcgen05.mma M128 N64 K16 (B = V, MN-major SW128)
//
// Round-2 structure of the softmax loop (the round-1 loop is attn5.cu): the per-warp dependency chain of one key block was
// S-wait -> TMEM load -> row max (340 cyc) -> wait for the PV MMA that still reads the single P buffer (255) -> exponentials
// -> hand-over, ~3100 cycles for a chain whose pipes need < 1600.  Now
//   * P is produced and consumed in two 64-key HALVES with their own full / done barriers: the PV MMAs of the first half
//     run under the exponentials of the second, and a half buffer is free again long before the next block needs it
//     (double buffering without a second 64 KB of shared memory);
//   * the exponentials are SPECULATIVE against the running reference m_ref: they start right after the TMEM load, the row
//     max of the half is computed alongside (FMNMX3 on the ALU pipe, which has room) and only checked afterwards.  The
//     reference already lagged the true maximum by up to 2^8 (lazy rescale); a half whose maximum exceeds it by more takes
//     the slow path (rescale O in TMEM, redo the half) BEFORE its P is released to the tensor core.  Slow path = the first
//     half of a row and rare jumps of the maximum.
// MMA issue order per tile: QK(j+1), PV(j, half 0), PV(j, half 1).
// Fully padded key blocks are skipped through a per-sample block list (result-preserving: their p is exactly 0).
// Roofline: tensor-bound; 4*L*L*64 flop per (sample, head).
#include <math.h>
#include <stdlib.h>

#include "bg_internal.h"
#include "ptx.cuh"

namespace bg {

int launch_attention_v5(cudaStream_t st, const AttnArgs& a);   // attn5.cu (round-1 kernel, A/B only)

namespace {

constexpr int DH = 64;
constexpr int NHEAD = 12;
constexpr int DMODEL = 768;
constexpr int TILE_BYTES = 128 * DH * 2;   // 16 KB: Q / K / V tile, 128 rows x 128 B
constexpr int P_BYTES = 128 * 128 * 2;     // 32 KB: two K-major SW128 blocks of 64 keys (= the two halves)
constexpr int P_HALF = P_BYTES / 2;

template <int NT>
struct ACfg {
  static constexpr int ST = (NT == 2) ? 3 : 2;              // K / V ring depth
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = NT * TILE_BYTES;
  static constexpr int OFF_V = OFF_K + ST * TILE_BYTES;
  static constexpr int OFF_P = OFF_V + ST * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_P + NT * P_BYTES;
  static constexpr int OFF_MASKW = OFF_BAR + 512;          // invalid-key bit words: 4 per key block, MAX_KB blocks
  static constexpr int MAX_KB = 64;                        // L <= 8192
  static constexpr int SMEM_BYTES = OFF_MASKW + MAX_KB * 16 + 1024;
  static constexpr int TMEM_COLS = (NT == 2) ? 512 : 256;
  // NT == 2: four warpgroups (2 softmax, producer / MMA warps, helper warps) so setmaxnreg can move registers
  static constexpr int THREADS = (NT == 2) ? 512 : NT * 128 + 64;
  static constexpr bool HELPER = NT == 2;
  static constexpr int TILE_COLS = 256;   // per tile: S at +0 (128 columns), O at +128 (64)
};

struct AttnParams {
  __half* out;
  int ldo;
  int B, L, nkb;
  const uint8_t* key_mask;
  const int* blk_list;
  const int* blk_count;
  const uint32_t* blk_words;   // [B][nkb][4] invalid-key bit words of the listed blocks, list order (per forward), or null
  float scale_log2;   // log2(e) / sqrt(64)
};

// exponentials of one 64-key half against the reference m_use:  p = exp2(c s - c m_use), fp16 P into the K-major SW128
// layout (row r, 16-byte chunk j8 at ((j8 ^ (r & 7)) << 4)), row sum in packed f32x2 math.
// Software-pipelined by one 16-key group: the MUFU.EX2 of group g are issued back to back, and only then are the results
// of group g-1 summed, packed and stored -- a consumer placed right behind its MUFU would stall the (in-order) warp for the
// MUFU latency.  WITH_MAX: the row maximum of the half is computed alongside on the ALU pipe.
// PM: 4-bit mask over the 4 element pairs of each 8-key chunk whose exp2 runs as a polynomial on the FMA pipe instead of
// MUFU.EX2 (the XU pipe, 16 ex2/clk/SM, is the binding unit of d=64 attention on B200).
template <int PM, int H, bool WITH_MAX>
__device__ __forceinline__ void exp_store_half(const float (&s)[128], float c, float m_use, uint32_t sProw, uint32_t r7,
                                               float& hsum, float& mx) {
  const float2 c2 = make_float2(c, c);
  const float2 nmc2 = make_float2(-m_use * c, -m_use * c);
  float2 acc = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
  float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float2 ecur[8], eprev[8];
  auto exp_group = [&](int g, float2 (&e)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float x0 = s[64 * H + 16 * g + 2 * q], x1 = s[64 * H + 16 * g + 2 * q + 1];
      if (WITH_MAX) mxa[q & 3] = fmax3(mxa[q & 3], x0, x1);
      const float2 a = ffma2(make_float2(x0, x1), c2, nmc2);
      e[q] = ((PM >> (q & 3)) & 1) ? exp2_poly2(a) : make_float2(ex2(a.x), ex2(a.y));
    }
  };
  auto drain_group = [&](int g, const float2 (&e)[8]) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 v = e[4 * hh + q];
        if (q & 1) acc1 = fadd2(acc1, v); else acc = fadd2(acc, v);
        __half2 h2 = __floats2half2_rn(v.x, v.y);
        pk[q] = *reinterpret_cast<uint32_t*>(&h2);
      }
      const uint32_t j8 = 2 * g + hh;                       // 16-byte chunk (8 keys) within this half
      st_shared_v4(sProw + ((j8 ^ r7) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
  };
  exp_group(0, eprev);
#pragma unroll
  for (int g = 1; g < 4; ++g) {
    exp_group(g, ecur);
    drain_group(g - 1, eprev);
#pragma unroll
    for (int q = 0; q < 8; ++q) eprev[q] = ecur[q];
  }
  drain_group(3, eprev);
  hsum = (acc.x + acc.y) + (acc1.x + acc1.y);
  if (WITH_MAX) mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
}

// Slow path of one half (first half of a row, or the row maximum jumped by more than 2^8 over the reference): rescale
// O (TMEM read-modify-write, all lanes of the warp take part; f == 1 for rows that do not need it) and recompute the half
// against the new reference.  Out of line and reading the scores from a local-memory copy: it runs once per row plus on rare
// maximum jumps, and keeping it out of the loop body keeps the hot loop inside the instruction cache.
__device__ __noinline__ float slow_half(const float* sl, float c, float m_use, uint32_t sProw, uint32_t r7, uint32_t o_tmem,
                                        float f, int rescale) {
  if (rescale) {
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t rr[32];
      tmem_ld_32x32b_x32(o_tmem + hh * 32, rr);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * f);
      tmem_st_32x32b_x32(o_tmem + hh * 32, rr);
    }
    tmem_st_wait();
  }
  const float nmc = -m_use * c;
  float sum = 0.f;
#pragma unroll 1
  for (uint32_t j8 = 0; j8 < 8; ++j8) {
    uint32_t pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float e0 = ex2(fmaf(sl[8 * j8 + 2 * q], c, nmc)), e1 = ex2(fmaf(sl[8 * j8 + 2 * q + 1], c, nmc));
      sum += e0 + e1;
      __half2 h2 = __floats2half2_rn(e0, e1);
      pk[q] = *reinterpret_cast<uint32_t*>(&h2);
    }
    st_shared_v4(sProw + ((j8 ^ r7) << 4), pk[0], pk[1], pk[2], pk[3]);
  }
  return sum;
}

template <int NT, int PM>
__global__ void __launch_bounds__(ACfg<NT>::THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  using C = ACfg<NT>;
  constexpr bool HW = C::HELPER;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + C::ST;
  uint64_t* v_full = k_empty + C::ST;
  uint64_t* v_empty = v_full + C::ST;
  uint64_t* s_full = v_empty + C::ST;
  uint64_t* s_free = s_full + NT;
  uint64_t* p_full = s_free + NT;          // [t * 2 + half]
  uint64_t* pv_full = p_full + NT * 2;     // [t * 2 + half]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + NT * 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qgrp = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nblk = p.blk_count ? p.blk_count[b] : p.nkb;
  const int* blist = p.blk_list ? p.blk_list + (size_t)b * p.nkb : nullptr;

  constexpr int PRODUCER_WARP = NT * 4;
  constexpr int MMA_WARP = NT * 4 + 1;

  if (warp == PRODUCER_WARP && elect_one()) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < C::ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], NT);     // one tcgen05.commit per MMA-issuing thread (one thread per query tile)
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], NT);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&p_full[t * 2 + i], 128);
        mbar_init(&pv_full[t * 2 + i], 1);
      }
    }
    fence_barrier_init();
    // start the first loads right away (this thread initialised the barriers itself): they overlap with the TMEM
    // allocation, the mask-word construction and the CTA-wide synchronisation below
    mbar_arrive_expect_tx(q_full, NT * TILE_BYTES);
    for (int t = 0; t < NT; ++t)
      tma_load_3d(smem + C::OFF_Q + t * TILE_BYTES, &tmQKV, q_full, h * DH, (qgrp * NT + t) * 128, b);
    for (int it = 0; it < nblk && it < C::ST; ++it) {
      const int kb = blist ? blist[it] : it;
      mbar_arrive_expect_tx(&k_full[it], TILE_BYTES);
      tma_load_3d(smem + C::OFF_K + it * TILE_BYTES, &tmQKV, &k_full[it], DMODEL + h * DH, kb * 128, b);
      mbar_arrive_expect_tx(&v_full[it], TILE_BYTES);
      tma_load_3d(smem + C::OFF_V + it * TILE_BYTES, &tmQKV, &v_full[it], 2 * DMODEL + h * DH, kb * 128, b);
    }
  }
  if (warp == MMA_WARP) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  // invalid-key bit words for every key block this CTA will visit (padded key or key >= L), built once: keeps the
  // global mask bytes off the per-block critical path
  uint32_t* maskw = reinterpret_cast<uint32_t*>(smem + C::OFF_MASKW);
  if (warp < PRODUCER_WARP) {
    if (p.blk_words || !p.key_mask) {
      // one word per thread, no dependent global-load chain: either copied from the per-forward table or, without a
      // mask, computed (only keys >= L are invalid)
      for (int wi = threadIdx.x; wi < nblk * 4; wi += PRODUCER_WARP * 32) {
        uint32_t w;
        if (p.blk_words) {      // list order: entry wi >> 2 belongs to key block blist[wi >> 2]
          w = p.blk_words[((size_t)b * p.nkb + (wi >> 2)) * 4 + (wi & 3)];
        } else {
          const int base = (wi >> 2) * 128 + (wi & 3) * 32;
          w = base + 32 <= p.L ? 0u : (base >= p.L ? 0xffffffffu : (0xffffffffu << (p.L - base)));
        }
        maskw[wi] = w;
      }
    } else {
      for (int wi = warp; wi < nblk * 4; wi += PRODUCER_WARP) {
        const int kb = blist ? blist[wi >> 2] : (wi >> 2);
        const int key = kb * 128 + (wi & 3) * 32 + lane;
        bool bad = key >= p.L;
        if (!bad && p.key_mask) bad = p.key_mask[(size_t)b * p.L + key] != 0;
        const uint32_t w = __ballot_sync(0xffffffffu, bad);
        if (lane == 0) maskw[wi] = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // register rebalancing (NT == 2): the softmax warpgroups hold a 128-wide score row per thread; the other two warpgroups
  // (producer, MMA issuers, helpers, idle warps) give their registers away.
  if (warp >= PRODUCER_WARP) {
    if constexpr (NT == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (HW && warp >= 12 && warp < 12 + NT) {
      // helper warp of tile t: named barrier 3 + t = "S_t of this block is in TMEM", 5 + t / 7 + t = "the PV MMAs that read
      // half 0 / half 1 of P_t in the previous block are done".  Same order as the softmax warpgroup syncs on them.
      const int t = warp - 12;
      for (int it = 0; it < nblk; ++it) {
        mbar_wait(&s_full[t], it & 1);
        named_bar_arrive(3 + t, 160);
        if (it > 0) {
          mbar_wait(&pv_full[t * 2], (it - 1) & 1);
          named_bar_arrive(5 + t, 160);
          mbar_wait(&pv_full[t * 2 + 1], (it - 1) & 1);
          named_bar_arrive(7 + t, 160);
        }
      }
      if (nblk > 0) {
        mbar_wait(&pv_full[t * 2], (nblk - 1) & 1);
        named_bar_arrive(5 + t, 160);
        mbar_wait(&pv_full[t * 2 + 1], (nblk - 1) & 1);
        named_bar_arrive(7 + t, 160);
      }
    } else if (warp == PRODUCER_WARP) {
      if (elect_one()) {
        for (int it = C::ST; it < nblk; ++it) {       // the first ST blocks were issued before the CTA-wide sync
          const int kb = blist ? blist[it] : it;
          const int s = it % C::ST;
          const uint32_t par = ((it / C::ST) & 1) ^ 1;
          mbar_wait(&k_empty[s], par);
          mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
          tma_load_3d(smem + C::OFF_K + s * TILE_BYTES, &tmQKV, &k_full[s], DMODEL + h * DH, kb * 128, b);
          mbar_wait(&v_empty[s], par);
          mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
          tma_load_3d(smem + C::OFF_V + s * TILE_BYTES, &tmQKV, &v_full[s], 2 * DMODEL + h * DH, kb * 128, b);
        }
      }
    } else if (warp < MMA_WARP + NT) {
      // one MMA-issuing thread per query tile: the two softmax warpgroups are not coupled through one in-order issuer
      if (elect_one()) {
        const int t = warp - MMA_WARP;
        constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_f16(128, DH, 0, 1);   // B (= V) is MN-major
        const uint32_t q_addr = smem_u32(smem + C::OFF_Q + t * TILE_BYTES);
        const uint32_t p_addr = smem_u32(smem + C::OFF_P + t * P_BYTES);
        const uint32_t s_tm = tmem_base + t * C::TILE_COLS, o_tm = s_tm + 128;
        mbar_wait(q_full, 0);
        tc_fence_after();
        for (int it = 0; it <= nblk; ++it) {
          if (it < nblk) {
            const int s = it % C::ST;
            mbar_wait(&k_full[s], (it / C::ST) & 1);
            mbar_wait(&s_free[t], (it & 1) ^ 1);
            tc_fence_after();
            const uint32_t k_addr = smem_u32(smem + C::OFF_K + s * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_f16_ss(s_tm, make_sw128_desc(q_addr + k * 32), make_sw128_desc(k_addr + k * 32), idesc_qk, k > 0 ? 1u : 0u);
            umma_commit(&s_full[t]);
            umma_commit(&k_empty[s]);
          }
          if (it > 0) {
            const int i = it - 1;
            const int s = i % C::ST;
            const uint32_t v_addr = smem_u32(smem + C::OFF_V + s * TILE_BYTES);
            mbar_wait(&v_full[s], (i / C::ST) & 1);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              mbar_wait(&p_full[t * 2 + hf], i & 1);   // this half of P_t(i) is written (and O_t rescaled if needed)
              tc_fence_after();
#pragma unroll
              for (int k = 4 * hf; k < 4 * hf + 4; ++k)
                umma_f16_ss(o_tm, make_sw128_desc(p_addr + hf * P_HALF + (k & 3) * 32), make_sw128_desc(v_addr + k * 2048),
                            idesc_pv, (i | k) != 0 ? 1u : 0u);
              umma_commit(&pv_full[t * 2 + hf]);
            }
            umma_commit(&v_empty[s]);
          }
        }
      }
    }
  } else {
    if constexpr (NT == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ------------------------------------------------------------------ softmax warpgroup for query tile t
    // One thread per query row (== TMEM lane).  The whole 128-key score row lives in registers (one TMEM read, S is
    // released to the next QK^T right away); O accumulates in TMEM across key blocks and is rescaled lazily: the exponent
    // reference m_ref only moves when the maximum of a half exceeds it by more than 2^8 (p <= 256 is harmless in fp16 P /
    // fp32 accumulation), so the TMEM read-modify-write of O is rare after the first half.
    const int t = warp >> 2;
    const uint32_t r = threadIdx.x & 127;                  // query row in tile == TMEM lane
    const uint32_t r7 = r & 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + t * C::TILE_COLS;
    const uint32_t o_tmem = s_tmem + 128;
    const uint32_t sP0 = smem_u32(smem + C::OFF_P + t * P_BYTES) + r * 128;
    const float c = p.scale_log2;

    float m_ref = -INFINITY, l = 0.f;
    float s[128];

    for (int it = 0; it < nblk; ++it) {
      const uint4 iw = *reinterpret_cast<const uint4*>(maskw + it * 4);
      if (HW) named_bar_sync(3 + t, 160);
      else mbar_wait(&s_full[t], it & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(s_tmem + cc * 32, reinterpret_cast<uint32_t*>(s) + cc * 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);          // S_t may be overwritten by QK^T of the next block

      auto half = [&](auto Hc) {
        constexpr int H = decltype(Hc)::value;
        // the PV MMAs that read this half of P in the previous block have finished (normally long ago)
        if (it > 0) {
          if (HW) named_bar_sync((H ? 7 : 5) + t, 160);
          else mbar_wait(&pv_full[t * 2 + H], (it - 1) & 1);
        }
        const uint32_t w0 = H ? iw.z : iw.x, w1 = H ? iw.w : iw.y;
        if ((w0 | w1) != 0) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (((i < 32 ? w0 : w1) >> (i & 31)) & 1u) s[64 * H + i] = -INFINITY;
        }
        const uint32_t sProw = sP0 + H * P_HALF;
        const bool unset = m_ref == -INFINITY;
        const bool fast = !__any_sync(0xffffffffu, unset);
        float hsum = 0.f, mx = -INFINITY;
        if (fast) {
          exp_store_half<PM, H, true>(s, c, m_ref, sProw, r7, hsum, mx);
        } else {
          float mxa[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) mxa[j] = fmaxf(s[64 * H + 2 * j], s[64 * H + 2 * j + 1]);
#pragma unroll
          for (int i = 16; i < 64; i += 16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mxa[j] = fmax3(mxa[j], s[64 * H + i + 2 * j], s[64 * H + i + 2 * j + 1]);
          }
          mx = fmaxf(fmax3(mxa[0], mxa[1], mxa[2]), fmax3(fmax3(mxa[3], mxa[4], mxa[5]), mxa[6], mxa[7]));
        }
        const bool need = mx > -INFINITY && (unset || (mx - m_ref) * c > 8.f);
        if (!fast || __any_sync(0xffffffffu, need)) {
          const int rescale = (it > 0 || H == 1) ? 1 : 0;
          float f = 1.f;
          if (rescale) {
            // every PV MMA issued so far must have completed before O is touched; the last one issued read the other
            // half: (it, 0) when this is half 1, (it - 1, 1) when this is half 0.  Direct (non-consuming) mbarrier wait:
            // the named-barrier sequence of the warpgroup stays untouched.
            mbar_wait(&pv_full[t * 2 + (H ^ 1)], (H ? it : it - 1) & 1);
            tc_fence_after();
            if (need && !unset) f = ex2((m_ref - mx) * c);
            l *= f;
          }
          if (need) m_ref = mx;
          float sl[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) sl[i] = s[64 * H + i];
          hsum = slow_half(sl, c, m_ref == -INFINITY ? 0.f : m_ref, sProw, r7, o_tmem, f, rescale);
        }
        l += hsum;
        tc_fence_before();                // orders the (rare) O rescale before the PV MMA that p_full releases
        fence_proxy_async_smem();         // generic-proxy writes of P -> visible to the tensor core (async proxy)
        mbar_arrive(&p_full[t * 2 + H]);
      };
      half(std::integral_constant<int, 0>{});
      half(std::integral_constant<int, 1>{});
    }

    float o[DH];
    if (nblk > 0) {
      if (HW) {
        named_bar_sync(5 + t, 160);
        named_bar_sync(7 + t, 160);
      } else {
        mbar_wait(&pv_full[t * 2], (nblk - 1) & 1);
        mbar_wait(&pv_full[t * 2 + 1], (nblk - 1) & 1);
      }
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) tmem_ld_32x32b_x32(o_tmem + hh * 32, reinterpret_cast<uint32_t*>(o) + hh * 32);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < DH; ++i) o[i] = 0.f;
    }
    const int row = (qgrp * NT + t) * 128 + (int)r;
    if (row < p.L) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.L + row) * p.ldo + h * DH);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        __half2 h0 = __floats2half2_rn(o[8 * q] * inv, o[8 * q + 1] * inv);
        __half2 h1 = __floats2half2_rn(o[8 * q + 2] * inv, o[8 * q + 3] * inv);
        __half2 h2 = __floats2half2_rn(o[8 * q + 4] * inv, o[8 * q + 5] * inv);
        __half2 h3 = __floats2half2_rn(o[8 * q + 6] * inv, o[8 * q + 7] * inv);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        dst[q] = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// one CTA per sample: which 128-key blocks hold at least one valid key (blk_list / blk_count), and the invalid-key bit
// words of the LISTED blocks in list order (blk_words[b][i][4] belongs to key block blk_list[b][i]), so that the
// attention kernel indexes them with its loop counter and needs no dependent load
__global__ void block_list_kernel(const uint8_t* __restrict__ key_mask, int L, int nkb, int* __restrict__ blk_list,
                                  int* __restrict__ blk_count, uint32_t* __restrict__ blk_words) {
  extern __shared__ int sm_bl[];
  int* pos = sm_bl;                                           // list position of key block kb, or -1
  uint32_t* wds = reinterpret_cast<uint32_t*>(sm_bl + nkb);   // [nkb][4]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int kb = warp; kb < nkb; kb += nwarp) {
    bool any = false;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int key = kb * 128 + c * 32 + lane;
      const bool bad = key >= L || key_mask[(size_t)b * L + key] != 0;
      const uint32_t w = __ballot_sync(0xffffffffu, bad);
      any = any || w != 0xffffffffu;
      if (lane == 0) wds[kb * 4 + c] = w;
    }
    if (lane == 0) pos[kb] = any ? 0 : -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int kb = 0; kb < nkb; ++kb)
      if (pos[kb] == 0) {
        blk_list[(size_t)b * nkb + n] = kb;
        pos[kb] = n++;
      }
    blk_count[b] = n;
  }
  __syncthreads();
  if (blk_words)
    for (int i = threadIdx.x; i < nkb * 4; i += blockDim.x)
      if (pos[i >> 2] >= 0) blk_words[((size_t)b * nkb + pos[i >> 2]) * 4 + (i & 3)] = wds[i];
}

template <int NT, int PM>
int launch_nt(cudaStream_t st, const CUtensorMap& tm, const AttnParams& p) {
  using C = ACfg<NT>;
  BG_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_kernel<NT, PM>), C::SMEM_BYTES));
  const int nq = (p.L + 127) / 128;
  dim3 grid((nq + NT - 1) / NT, NHEAD, p.B);
  attn_kernel<NT, PM><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tm, p);
  return check_launch("attn_kernel launch");
}

}  // namespace

int launch_attention(cudaStream_t st, const AttnArgs& a) {
  BG_REQUIRE(a.qkv && a.out && a.B > 0 && a.L > 0, "attention: bad arguments");
  BG_REQUIRE(a.ldo % 8 == 0, "attention: output pitch must be a multiple of 8");
  BG_REQUIRE(a.L <= 128 * ACfg<2>::MAX_KB, "attention: sequence longer than 8192 tokens is not supported");
  BG_REQUIRE((a.blk_list == nullptr) == (a.blk_count == nullptr), "attention: blk_list and blk_count go together");
  static int version = -1, poly = -1;     // environment knobs, read once per process
  if (version < 0) {
    const char* e = getenv("BG_ATTN_V");      // 5: the round-1 kernel (attn5.cu), A/B timing only
    version = e ? atoi(e) : 6;
    e = getenv("BG_ATTN_POLY");               // share of the exponentials on the FMA pipe: 0 | 1 (25 %, default) | 2 (50 %)
    poly = e ? atoi(e) : 1;
  }
  if (version == 5) return launch_attention_v5(st, a);
  CUtensorMap tm;
  BG_TRY(make_tmap_3d_f16(&tm, a.qkv, (uint64_t)a.B, (uint64_t)a.L, 3 * DMODEL, 3 * DMODEL, 128));
  AttnParams p;
  p.out = a.out; p.ldo = a.ldo; p.B = a.B; p.L = a.L; p.nkb = (a.L + 127) / 128;
  p.key_mask = a.key_mask; p.blk_list = a.blk_list; p.blk_count = a.blk_count; p.blk_words = a.blk_words;
  p.scale_log2 = 1.4426950408889634f / 8.0f;
  if (a.L <= 128) return launch_nt<1, 0x0>(st, tm, p);
  if (poly == 0) return launch_nt<2, 0x0>(st, tm, p);
  if (poly == 2) return launch_nt<2, 0xA>(st, tm, p);
  return launch_nt<2, 0x8>(st, tm, p);
}

int launch_build_block_list(cudaStream_t st, const uint8_t* key_mask, int B, int L, int* blk_list, int* blk_count,
                            uint32_t* blk_words) {
  BG_REQUIRE(key_mask && blk_list && blk_count && B > 0 && L > 0, "block list: bad arguments");
  const int nkb = (L + 127) / 128;
  block_list_kernel<<<B, 128, nkb * 5 * sizeof(int), st>>>(key_mask, L, nkb, blk_list, blk_count, blk_words);
  return check_launch("block_list_kernel launch");
}

}  // namespace bg
