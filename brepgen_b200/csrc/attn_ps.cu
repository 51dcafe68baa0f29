// Persistent variant of the tcgen05 flash attention for L > 128 (see attn.cu for the reference semantics, the per-block
// pipeline and the roles).  One CTA per SM loops over work items (sample, head, pair of 128-row query tiles); every
// mbarrier keeps running across items (phases follow a per-role running block counter), so the tail of item n -- the last
// softmax block, the last PV MMA, the O read-out and the global store -- overlaps the Q / K / V loads and the first QK^T
// of item n+1.  In the one-item-per-CTA kernel the barrier setup, TMEM allocation, first-load latency, final drain and
// CTA turnover are ~15 % of each CTA's life (in-kernel clock64 trace).
//
// 512 threads:  warps 0-7   softmax warpgroups (tile 0 / tile 1; thread r owns query row r == TMEM lane r)
//               warp 8      TMA producer (Q per item after q_empty; K / V through the ST-deep ring, flat over items)
//               warps 9,10  MMA issuer of tile 0 / 1, flat order  QK(g) PV(g-1)  across item boundaries
//               warps 12,13 helper warps: mbarrier waits (S ready / PV done) -> named barriers 3+t / 5+t
// Hazards across items: Q smem is reloaded only after both issuers committed their last QK^T of the item (q_empty);
// O_t of item n is read by the softmax warpgroup before its 128 threads arrive on p_full for block 0 of item n+1, and
// PV(0, n+1) -- the MMA that overwrites O_t -- waits on that p_full; P_t is rewritten only after wait_pv(last block).
// Mask words come from the per-forward table in LIST order (block_list_kernel), one item ahead in registers.
#include <math.h>
#include <stdlib.h>

#include "bg_internal.h"
#include "ptx.cuh"

namespace bg {

namespace {

constexpr int DH = 64;
constexpr int NHEAD = 12;
constexpr int DMODEL = 768;
constexpr int TILE_BYTES = 128 * DH * 2;
constexpr int P_BYTES = 128 * 128 * 2;
constexpr int ST = 3;
constexpr int NT = 2;

constexpr int OFF_Q = 0;
constexpr int OFF_K = NT * TILE_BYTES;
constexpr int OFF_V = OFF_K + ST * TILE_BYTES;
constexpr int OFF_P = OFF_V + ST * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + NT * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr int TMEM_COLS = 512;
constexpr int TILE_COLS = 256;       // per tile: S at +0 (128 columns), O at +128 (64 columns)
constexpr int THREADS = 512;
constexpr int PRODUCER_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int HELPER_WARP = 12;

struct PsParams {
  __half* out;
  int ldo;
  int B, L, nkb;
  int nqg;      // query-tile pairs per (sample, head)
  int total;    // work items = B * 12 * nqg; item w -> qg = w % nqg, h = (w / nqg) % 12, b = w / (12 nqg)
  const int* blk_list;
  const int* blk_count;
  const uint32_t* blk_words;   // [B][nkb][4] in list order, or null (no mask: only keys >= L are invalid)
  float scale_log2;
  int pingpong;
  int one;      // always 1 (opaque to the compiler; see the TK form)
  int stagger;  // ns that warpgroup 1 sleeps once before its first block (anti-phase start without a token; 0 = off)
};

template <int PM, int SP, int TK>
__global__ void __launch_bounds__(THREADS, 1) attn_ps_kernel(const __grid_constant__ CUtensorMap tmQKV, const PsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* q_empty = q_full + 1;
  uint64_t* k_full = q_empty + 1;
  uint64_t* k_empty = k_full + ST;
  uint64_t* v_full = k_empty + ST;
  uint64_t* v_empty = v_full + ST;
  uint64_t* s_full = v_empty + ST;
  uint64_t* s_free = s_full + NT;
  uint64_t* p_full = s_free + NT;
  uint64_t* pv_full = p_full + NT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + NT);

  const int warp = threadIdx.x >> 5;
  const int stride = gridDim.x;
  const int per_sample = p.nqg * NHEAD;

  if (warp == PRODUCER_WARP && elect_one()) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, NT);            // one tcgen05.commit per MMA-issuing thread after its last QK^T of the item
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], NT);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], NT);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // key blocks of the sample that item w belongs to (0 = fully padded sample: every role skips the item)
  auto nblk_of = [&](int w) -> int { return p.blk_count ? __ldg(p.blk_count + w / per_sample) : p.nkb; };

  if (warp >= PRODUCER_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == HELPER_WARP || warp == HELPER_WARP + 1) {
      // helper warp of tile t: named barrier 3 + t = "S_t of this block is in TMEM", 5 + t = "PV_t of the previous block done"
      const int t = warp - HELPER_WARP;
      uint32_t g = 0;                                    // parity of the running block counter
      for (int w = blockIdx.x; w < p.total; w += stride) {
        const int nblk = nblk_of(w);
        if (nblk == 0) continue;
        for (int it = 0; it < nblk; ++it) {
          mbar_wait(&s_full[t], g);
          named_bar_arrive(3 + t, 160);
          if (it > 0) {
            mbar_wait(&pv_full[t], g ^ 1);
            named_bar_arrive(5 + t, 160);
          }
          g ^= 1;
        }
        mbar_wait(&pv_full[t], g ^ 1);
        named_bar_arrive(5 + t, 160);
      }
    } else if (warp == PRODUCER_WARP) {
      if (elect_one()) {
        int s = 0;
        uint32_t ph = 0, na = 0;                         // ring stage / phase of the running block counter; active items
        for (int w = blockIdx.x; w < p.total; w += stride) {
          const int nblk = nblk_of(w);
          if (nblk == 0) continue;
          const int qg = w % p.nqg, hb = w / p.nqg, h = hb % NHEAD, b = hb / NHEAD;
          const int* blist = p.blk_list ? p.blk_list + (size_t)b * p.nkb : nullptr;
          mbar_wait(q_empty, (na & 1) ^ 1);              // every QK^T of the previous item has read Q
          mbar_arrive_expect_tx(q_full, NT * TILE_BYTES);
          for (int t = 0; t < NT; ++t)
            tma_load_3d(smem + OFF_Q + t * TILE_BYTES, &tmQKV, q_full, h * DH, (qg * NT + t) * 128, b);
          for (int it = 0; it < nblk; ++it) {
            const int kb = blist ? __ldg(blist + it) : it;
            mbar_wait(&k_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
            tma_load_3d(smem + OFF_K + s * TILE_BYTES, &tmQKV, &k_full[s], DMODEL + h * DH, kb * 128, b);
            mbar_wait(&v_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
            tma_load_3d(smem + OFF_V + s * TILE_BYTES, &tmQKV, &v_full[s], 2 * DMODEL + h * DH, kb * 128, b);
            if (++s == ST) { s = 0; ph ^= 1; }
          }
          ++na;
        }
      }
    } else if (warp == MMA_WARP || warp == MMA_WARP + 1) {
      if (elect_one()) {
        const int t = warp - MMA_WARP;
        constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_f16(128, DH, 0, 1);   // B (= V) is MN-major
        const uint32_t q_addr = smem_u32(smem + OFF_Q + t * TILE_BYTES);
        const uint32_t p_addr = smem_u32(smem + OFF_P + t * P_BYTES);
        const uint32_t s_tmem = tmem_base + t * TILE_COLS;
        int s = 0;
        uint32_t ph = 0, g = 0, na = 0;
        // the PV MMA of the previous block (possibly of the previous item) is issued AFTER the QK^T of the current one
        bool pend = false, pend_first = false;
        int pend_s = 0;
        uint32_t pend_ph = 0, pend_g = 0;
        auto issue_pv = [&]() {
          mbar_wait(&v_full[pend_s], pend_ph);
          mbar_wait(&p_full[t], pend_g);      // P_t written, O_t rescaled (or, first block of an item, read out) by the softmax WG
          tc_fence_after();
          const uint32_t v_addr = smem_u32(smem + OFF_V + pend_s * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k)
            umma_f16_ss(s_tmem + 128, make_sw128_desc(p_addr + (k >> 2) * (P_BYTES / 2) + (k & 3) * 32),
                        make_sw128_desc(v_addr + k * 2048), idesc_pv, (pend_first && k == 0) ? 0u : 1u);
          umma_commit(&pv_full[t]);
          umma_commit(&v_empty[pend_s]);
        };
        for (int w = blockIdx.x; w < p.total; w += stride) {
          const int nblk = nblk_of(w);
          if (nblk == 0) continue;
          mbar_wait(q_full, na & 1);
          tc_fence_after();
          for (int it = 0; it < nblk; ++it) {
            mbar_wait(&k_full[s], ph);
            mbar_wait(&s_free[t], g ^ 1);
            tc_fence_after();
            const uint32_t k_addr = smem_u32(smem + OFF_K + s * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_f16_ss(s_tmem, make_sw128_desc(q_addr + k * 32), make_sw128_desc(k_addr + k * 32), idesc_qk,
                          k > 0 ? 1u : 0u);
            umma_commit(&s_full[t]);
            umma_commit(&k_empty[s]);
            if (it == nblk - 1) umma_commit(q_empty);
            if (pend) issue_pv();
            pend = true; pend_first = it == 0; pend_s = s; pend_ph = ph; pend_g = g;
            g ^= 1;
            if (++s == ST) { s = 0; ph ^= 1; }
          }
          ++na;
        }
        if (pend) issue_pv();
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ------------------------------------------------------------------ softmax warpgroup for query tile t
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;                       // query row in tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + t * TILE_COLS;
    const uint32_t o_tmem = s_tmem + 128;
    const uint32_t sP = smem_u32(smem + OFF_P + t * P_BYTES) + r * 128;
    const float c = p.scale_log2;
    const bool pingpong = p.pingpong != 0;

    auto ld_words = [&](int b, int it, uint32_t (&wd)[4]) {
      if (p.blk_words) {
        const uint32_t* src = p.blk_words + ((size_t)b * p.nkb + it) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) wd[q] = __ldg(src + q);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int base = it * 128 + q * 32;
          wd[q] = base + 32 <= p.L ? 0u : (base >= p.L ? 0xffffffffu : (0xffffffffu << (p.L - base)));
        }
      }
    };
    // rows of tile t of a fully padded item: the reference would give NaN; the cascade never reads them -> zeros
    auto zero_fill = [&](int w) {
      const int qg = w % p.nqg, hb = w / p.nqg, h = hb % NHEAD, b = hb / NHEAD;
      const int row = (qg * NT + t) * 128 + r;
      if (row < p.L) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.L + row) * p.ldo + h * DH);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    // first item at or after w with at least one key block; zero-fills the skipped ones
    auto next_active = [&](int w, int& nb) -> int {
      for (; w < p.total; w += stride) {
        nb = nblk_of(w);
        if (nb > 0) break;
        zero_fill(w);
      }
      return w;
    };

    int nblk = 0;
    int w = next_active(blockIdx.x, nblk);
    uint32_t nxt[4] = {0u, 0u, 0u, 0u};
    if (w < p.total) {
      ld_words(w / per_sample, 0, nxt);
      if (pingpong && t == 1) named_bar_arrive(1, 256);    // tile 0 owns the XU token first
    }
    if (t == 1 && p.stagger > 0) __nanosleep((unsigned)p.stagger);
    while (w < p.total) {
      const int qg = w % p.nqg, hb = w / p.nqg, h = hb % NHEAD, b = hb / NHEAD;
      // candidate next item: its block count is requested now and consumed in the last block of this item
      const int cand = w + stride;
      int w_next = cand;
      int nblk_next = cand < p.total ? nblk_of(cand) : 0;
      float m_ref = -INFINITY, l = 0.f;

      for (int it = 0; it < nblk; ++it) {
        const uint32_t inval[4] = {nxt[0], nxt[1], nxt[2], nxt[3]};
        if (it + 1 < nblk) {
          ld_words(b, it + 1, nxt);
        } else {
          // last block: settle which item comes next (rarely a scan over fully padded samples) and fetch its first words
          if (cand < p.total && nblk_next == 0) {
            zero_fill(cand);
            w_next = next_active(cand + stride, nblk_next);
          }
          if (w_next < p.total) ld_words(w_next / per_sample, 0, nxt);
        }

        named_bar_sync(3 + t, 160);       // S_t(it) is in TMEM (helper warp waited on s_full)
        tc_fence_after();
        float s[128];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(s_tmem + cc * 32, reinterpret_cast<uint32_t*>(s) + cc * 32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);          // S_t may be overwritten by the next QK^T (of this or of the next item)

        if ((inval[0] | inval[1] | inval[2] | inval[3]) != 0) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if ((inval[i >> 5] >> (i & 31)) & 1u) s[i] = -INFINITY;
        }
        auto row_max = [&]() -> float {
          float mx[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) mx[j] = fmaxf(s[2 * j], s[2 * j + 1]);
#pragma unroll
          for (int i = 16; i < 128; i += 16) {     // eight independent FMNMX3 chains
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[j] = fmax3(mx[j], s[i + 2 * j], s[i + 2 * j + 1]);
          }
          return fmaxf(fmax3(m_ref, fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5])), fmaxf(mx[6], mx[7]));
        };
        bool pv_ready = it == 0;          // it == 0: the previous item's last PV was awaited before its O read-out
        // lazy rescale: the exponent reference only moves when the running max grew by more than 2^8
        auto rescale_to = [&](float m_new) {
          const bool need = (m_new - m_ref) * c > 8.f;
          if (__any_sync(0xffffffffu, need)) {
            if (!pv_ready) named_bar_sync(5 + t, 160);   // O_t complete up to block it-1 before its read-modify-write
            pv_ready = true;
            tc_fence_after();
            const float f = need ? ex2((m_ref - m_new) * c) : 1.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t rr[32];
              tmem_ld_32x32b_x32(o_tmem + hh * 32, rr);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * f);
              tmem_st_32x32b_x32(o_tmem + hh * 32, rr);
            }
            tmem_st_wait();
            l *= f;
            if (need) m_ref = m_new;
          }
        };
        // SP ("speculative reference"): after the first block the row max is NOT computed.  The exponentials are taken
        // against the current m_ref; if any row's block sum reaches 2^15 (so a single p could approach the fp16 range)
        // the warp computes the true max, rescales and redoes the block.  This takes the 100-instruction FMNMX chain
        // (~340 cycles of dependent latency) off the per-block critical path; results differ from the eager form only
        // in which power of two P is scaled by.
        if (it == 0) {
          const float m_new = row_max();
          m_ref = (m_new == -INFINITY) ? 0.f : m_new;
        } else if (!SP) {
          rescale_to(row_max());
        }
        if (TK) {
          // token-dense form: the scale / subtract of all 128 scores happens BEFORE the XU token is taken (ordered asm,
          // so ptxas keeps it on this side of the barrier); what runs while holding the token is then a compact
          // MUFU.EX2 stream (8 XU cycles per warp instruction) with only the pack / sum / store of the previous group
          // in its shadow.  Intended regime: each warpgroup's non-XU work hides completely behind the other's token time.
          const float2 c2 = make_float2(c, c);
          const float2 nmc2 = make_float2(-m_ref * c, -m_ref * c);
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const float2 a = ffma2_ordered(make_float2(s[2 * i], s[2 * i + 1]), c2, nmc2);
            s[2 * i] = a.x;
            s[2 * i + 1] = a.y;
          }
        }
        if (!pv_ready) named_bar_sync(5 + t, 160);          // the PV MMA that read the P buffer has finished
        pv_ready = true;
        if (pingpong) named_bar_sync(1 + t, 256);

        // one pass over the 128 scores of the row: p = exp2(c s - c m_ref), row sum (packed f32x2 math), fp16 P into the
        // K-major SW128 layout; software-pipelined by one 16-key group so that the MUFU.EX2 of group g are in flight while
        // group g-1 is summed, packed and stored.  Returns the block's row sum.
        // the very last block of this CTA (tile 1, no next item) has nobody to hand the XU token to
        const bool hand_over = pingpong && !(t == 1 && it == nblk - 1 && w_next >= p.total);
        auto exp_pass = [&]() -> float {
          const float2 c2 = make_float2(c, c);
          const float2 nmc2 = make_float2(-m_ref * c, -m_ref * c);
          float2 acc = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          float2 ecur[8], eprev[8];
          auto exp_group = [&](int g, float2 (&e)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (TK) {   // s[] already holds c s - c m_ref; the polynomial share (FMA pipe) rides in the MUFU stream's shadow
                if ((PM >> (q & 3)) & 1) e[q] = exp2_poly2(make_float2(s[16 * g + 2 * q], s[16 * g + 2 * q + 1]));
                else e[q] = make_float2(ex2_ordered(s[16 * g + 2 * q]), ex2_ordered(s[16 * g + 2 * q + 1]));
                continue;
              }
              float2 a = ffma2(make_float2(s[16 * g + 2 * q], s[16 * g + 2 * q + 1]), c2, nmc2);
              if ((PM >> (q & 3)) & 1) {
                if (SP) { a.x = fminf(a.x, 126.f); a.y = fminf(a.y, 126.f); }   // keep the exponent add of the polynomial in range
                e[q] = exp2_poly2(a);
              } else {
                e[q] = make_float2(ex2(a.x), ex2(a.y));
              }
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
                __half2 hv = __floats2half2_rn(v.x, v.y);
                pk[q] = *reinterpret_cast<uint32_t*>(&hv);
              }
              const int j8 = 2 * g + hh;                         // 16-byte chunk (8 keys) index along the 128 keys
              st_shared_v4(sP + (j8 >> 3) * (P_BYTES / 2) + (((j8 & 7) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
          };
          exp_group(0, eprev);
#pragma unroll
          for (int g = 1; g < 8; ++g) {
            exp_group(g, ecur);
            drain_group(g - 1, eprev);
#pragma unroll
            for (int q = 0; q < 8; ++q) eprev[q] = ecur[q];
          }
          // TK: every MUFU.EX2 of this block has been issued -> the other warpgroup may start its stream; the pack / sum /
          // store of the last group happen outside the token (no memory semantics are attached to the token barrier)
          if (TK && hand_over) named_bar_arrive_relaxed(1 + (1 - t), 256);
          drain_group(7, eprev);
          return (acc.x + acc.y) + (acc1.x + acc1.y);
        };
        // SP: first pass against the current reference; the (rare) redo with the true max is a second, cold copy.
        // ptxas does not hoist this pass above the token barrier the way it hoists the eager form's (there it fills the
        // latency gaps of the FMNMX chain with FFMA2 / MUFU work), so with ping-pong on the two warpgroups' exponential
        // phases are strictly serial -- see profiles/README.md v8s.
        float bsum = 0.f;
        if (TK) {
          // ptxas schedules arithmetic across BAR.SYNC freely (even `asm volatile`), so the MUFU stream is put into a
          // loop with an opaque trip count of one: code inside a loop body is not hoisted above the barrier before it
#pragma unroll 1
          for (int rep = 0; rep < p.one; ++rep) bsum = exp_pass();
        } else {
          bsum = exp_pass();
        }
        if (SP && it > 0 && __any_sync(0xffffffffu, !(bsum < 32768.f))) {
          rescale_to(row_max());          // a score outgrew the reference by 2^15 / 128 or more
          bsum = exp_pass();
        }
        if (!TK && hand_over) named_bar_arrive(1 + (1 - t), 256);       // hand the XU token over
        tc_fence_before();                // orders the O rescale / the previous item's O read-out before the PV MMA
        fence_proxy_async_smem();         // generic-proxy writes of P -> visible to the tensor core (async proxy)
        mbar_arrive(&p_full[t]);
        l += bsum;
      }

      // read-out of O_t; the next item's first QK^T and its loads are already in flight
      float o[DH];
      named_bar_sync(5 + t, 160);         // PV of the last block done
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) tmem_ld_32x32b_x32(o_tmem + hh * 32, reinterpret_cast<uint32_t*>(o) + hh * 32);
      tmem_ld_wait();
      const int row = (qg * NT + t) * 128 + r;
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
      w = w_next;
      nblk = nblk_next;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int PM, int SP, int TK>
int launch_ps(cudaStream_t st, const CUtensorMap& tm, const PsParams& p, int ctas) {
  static bool configured = false;
  if (!configured) {
    BG_CUDA(cudaFuncSetAttribute(attn_ps_kernel<PM, SP, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  attn_ps_kernel<PM, SP, TK><<<ctas, THREADS, SMEM_BYTES, st>>>(tm, p);
  return check_launch("attn_ps_kernel launch");
}

}  // namespace

bool attention_persistent_supported(const AttnArgs& a) {
  if (a.L <= 128) return false;
  // mask words must come from the per-forward table (list order); a bare key_mask goes to the one-item-per-CTA kernel
  if (a.key_mask || a.blk_list) return a.blk_list && a.blk_count && a.blk_words;
  return true;
}

// L > 128 path of launch_attention (attn.cu) when BG_ATTN_PS != 0 and attention_persistent_supported(a)
int launch_attention_persistent(cudaStream_t st, const AttnArgs& a, int poly) {
  BG_REQUIRE(attention_persistent_supported(a), "persistent attention: unsupported argument combination");
  CUtensorMap tm;
  BG_TRY(make_tmap_3d_f16(&tm, a.qkv, (uint64_t)a.B, (uint64_t)a.L, 3 * DMODEL, 3 * DMODEL, 128));
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    BG_CUDA(cudaGetDevice(&dev));
    BG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  PsParams p;
  p.out = a.out; p.ldo = a.ldo; p.B = a.B; p.L = a.L; p.nkb = (a.L + 127) / 128;
  p.nqg = (p.nkb + NT - 1) / NT;
  const long long total = (long long)a.B * NHEAD * p.nqg;
  BG_REQUIRE(total < (1ll << 30), "persistent attention: too many work items");
  p.total = (int)total;
  p.blk_list = a.blk_list; p.blk_count = a.blk_count; p.blk_words = a.blk_words;
  p.scale_log2 = 1.4426950408889634f / 8.0f;
  const char* e1 = getenv("BG_ATTN_PP");
  p.pingpong = e1 ? atoi(e1) : 1;
  p.one = 1;
  const char* e2 = getenv("BG_ATTN_STAGGER");   // experiment: start warpgroup 1 this many ns late (use with BG_ATTN_PP=0)
  p.stagger = e2 ? atoi(e2) : 0;
  const int ctas = p.total < sms ? p.total : sms;
  static int spec = -1;                 // BG_ATTN_SPEC = 1: no row max after the first block (see SP in the kernel)
  if (spec < 0) {
    const char* e = getenv("BG_ATTN_SPEC");
    spec = e ? atoi(e) : 0;
  }
  static int dense = -1;                // BG_ATTN_TK = 1: scale / subtract before the XU token, MUFU-only stream inside it
  if (dense < 0) {
    const char* e = getenv("BG_ATTN_TK");
    dense = e ? atoi(e) : 0;
  }
  if (dense) return poly == 0 ? launch_ps<0x0, 0, 1>(st, tm, p, ctas) : launch_ps<0x8, 0, 1>(st, tm, p, ctas);   // eager row max
  if (spec) return poly == 0 ? launch_ps<0x0, 1, 0>(st, tm, p, ctas) : launch_ps<0x8, 1, 0>(st, tm, p, ctas);
  return poly == 0 ? launch_ps<0x0, 0, 0>(st, tm, p, ctas) : launch_ps<0x8, 0, 0>(st, tm, p, ctas);
}

}  // namespace bg
