// Row-split variant of the tcgen05 flash attention (see attn.cu for the reference semantics and the pipeline):
// every query row is handled by TWO threads (64 keys each) of two different warps, so that 16 softmax warps
// (4 per SM sub-partition instead of 2) hide the per-warp latencies of the TMEM load / max / mbarrier chain and keep
// the XU (MUFU.EX2) pipe, the binding unit of d = 64 attention, fed.  The two halves of a row exchange their partial
// row max through shared memory once per key block (named barrier per query tile) and their partial row sums once at
// the end.  CTA = 2 query tiles x 128 rows of one (sample, head); 640 threads:
//   warps 0-15 : softmax; tile = warp / 8, key half = (warp / 4) % 2, TMEM lane quarter = warp % 4
//   warp 16    : TMA producer,  warps 17 / 18 : MMA issuer of tile 0 / 1 (+ TMEM allocator), warp 19 idle
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
constexpr int MAX_KB = 64;

constexpr int OFF_Q = 0;
constexpr int OFF_K = 2 * TILE_BYTES;
constexpr int OFF_V = OFF_K + ST * TILE_BYTES;
constexpr int OFF_P = OFF_V + ST * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr int OFF_MASKW = OFF_BAR + 512;
constexpr int OFF_XCH = OFF_MASKW + MAX_KB * 16;           // [2 buffers][2 tiles][2 halves][128 rows] floats
constexpr int SMEM_BYTES = OFF_XCH + 2 * 2 * 2 * 128 * 4 + 1024;
constexpr int TMEM_COLS = 512;
constexpr int THREADS = 640;
constexpr int TILE_COLS = 192;                              // per tile: S at +0 (128 cols), O at +128 (64 cols)
constexpr int PRODUCER_WARP = 16;
constexpr int MMA_WARP = 17;

struct Params {
  __half* out;
  int ldo;
  int B, L, nkb;
  const uint8_t* key_mask;
  const int* blk_list;
  const int* blk_count;
  float scale_log2;
  int pingpong;
};

template <int PM>
__global__ void __launch_bounds__(THREADS, 1) attn_rs_kernel(const __grid_constant__ CUtensorMap tmQKV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + ST;
  uint64_t* v_full = k_empty + ST;
  uint64_t* v_empty = v_full + ST;
  uint64_t* s_full = v_empty + ST;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* pv_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 2);
  uint32_t* maskw = reinterpret_cast<uint32_t*>(smem + OFF_MASKW);
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qgrp = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nblk = p.blk_count ? p.blk_count[b] : p.nkb;
  const int* blist = p.blk_list ? p.blk_list + (size_t)b * p.nkb : nullptr;

  if (warp == PRODUCER_WARP && elect_one()) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 256);
      mbar_init(&p_full[t], 256);
      mbar_init(&pv_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(tmem_slot);
  if (warp < PRODUCER_WARP) {
    for (int wi = warp; wi < nblk * 4; wi += PRODUCER_WARP) {
      const int kb = blist ? blist[wi >> 2] : (wi >> 2);
      const int key = kb * 128 + (wi & 3) * 32 + lane;
      bool bad = key >= p.L;
      if (!bad && p.key_mask) bad = p.key_mask[(size_t)b * p.L + key] != 0;
      const uint32_t w = __ballot_sync(0xffffffffu, bad);
      if (lane == 0) maskw[wi] = w;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= PRODUCER_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == PRODUCER_WARP) {
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
        for (int t = 0; t < 2; ++t)
          tma_load_3d(smem + OFF_Q + t * TILE_BYTES, &tmQKV, q_full, h * DH, (qgrp * 2 + t) * 128, b);
        for (int it = 0; it < nblk; ++it) {
          const int kb = blist ? blist[it] : it;
          const int s = it % ST;
          const uint32_t par = ((it / ST) & 1) ^ 1;
          mbar_wait(&k_empty[s], par);
          mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
          tma_load_3d(smem + OFF_K + s * TILE_BYTES, &tmQKV, &k_full[s], DMODEL + h * DH, kb * 128, b);
          mbar_wait(&v_empty[s], par);
          mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
          tma_load_3d(smem + OFF_V + s * TILE_BYTES, &tmQKV, &v_full[s], 2 * DMODEL + h * DH, kb * 128, b);
        }
      }
    } else if (warp < MMA_WARP + 2) {
      if (elect_one()) {
        const int t = warp - MMA_WARP;
        constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_f16(128, DH, 0, 1);
        mbar_wait(q_full, 0);
        tc_fence_after();
        for (int it = 0; it <= nblk; ++it) {
          if (it < nblk) {
            const int s = it % ST;
            mbar_wait(&k_full[s], (it / ST) & 1);
            mbar_wait(&s_free[t], (it & 1) ^ 1);
            tc_fence_after();
            const uint32_t k_addr = smem_u32(smem + OFF_K + s * TILE_BYTES);
            const uint32_t q_addr = smem_u32(smem + OFF_Q + t * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_f16_ss(tmem_base + t * TILE_COLS, make_sw128_desc(q_addr + k * 32), make_sw128_desc(k_addr + k * 32),
                          idesc_qk, k > 0 ? 1u : 0u);
            umma_commit(&s_full[t]);
            umma_commit(&k_empty[s]);
          }
          if (it > 0) {
            const int i = it - 1;
            const int s = i % ST;
            mbar_wait(&v_full[s], (i / ST) & 1);
            mbar_wait(&p_full[t], i & 1);
            tc_fence_after();
            const uint32_t v_addr = smem_u32(smem + OFF_V + s * TILE_BYTES);
            const uint32_t p_addr = smem_u32(smem + OFF_P + t * P_BYTES);
#pragma unroll
            for (int k = 0; k < 128 / 16; ++k)
              umma_f16_ss(tmem_base + t * TILE_COLS + 128, make_sw128_desc(p_addr + (k >> 2) * (P_BYTES / 2) + (k & 3) * 32),
                          make_sw128_desc(v_addr + k * 2048), idesc_pv, (i | k) != 0 ? 1u : 0u);
            umma_commit(&pv_full[t]);
            umma_commit(&v_empty[s]);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int t = warp >> 3;                    // query tile
    const int hf = (warp >> 2) & 1;             // key half of the row this thread owns
    const int r = (warp & 3) * 32 + lane;       // query row in tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + t * TILE_COLS + hf * 64;
    const uint32_t o_tmem = tmem_base + lane_base + t * TILE_COLS + 128 + hf * 32;
    const uint32_t sP = smem_u32(smem + OFF_P + t * P_BYTES) + hf * (P_BYTES / 2) + r * 128;
    const float c = p.scale_log2;
    float m_ref = -INFINITY, l = 0.f;

    // XU token between the two query tiles (named barriers 3 / 4, 256 waiting + 256 arriving threads): the exponential
    // phases of tile 0 and tile 1 alternate, so the MUFU pipe sees a steady 2 warps per sub-partition instead of all
    // four warps hitting it in lock-step and then all leaving it idle during their TMEM-load / max / barrier phases.
    const bool pingpong = p.pingpong != 0;
    if (pingpong && t == 1 && nblk > 0) named_bar_arrive(3, 512);

    for (int it = 0; it < nblk; ++it) {
      const uint2 iw = *reinterpret_cast<const uint2*>(maskw + it * 4 + hf * 2);
      mbar_wait(&s_full[t], it & 1);
      tc_fence_after();
      float s[64];
      tmem_ld_32x32b_x32(s_tmem, reinterpret_cast<uint32_t*>(s));
      tmem_ld_32x32b_x32(s_tmem + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);
      if ((iw.x | iw.y) != 0) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (((i < 32 ? iw.x : iw.y) >> (i & 31)) & 1u) s[i] = -INFINITY;
      }
      float mx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) mx[j] = fmaxf(s[2 * j], s[2 * j + 1]);
#pragma unroll
      for (int i = 16; i < 64; i += 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mx[j] = fmax3(mx[j], s[i + 2 * j], s[i + 2 * j + 1]);
      }
      const float mxp = fmaxf(fmax3(mx[0], mx[1], mx[2]), fmaxf(fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      // exchange the partial row max with the thread that owns the other 64 keys of this row
      float* xb = xch + (it & 1) * 512 + t * 256;
      xb[hf * 128 + r] = mxp;
      named_bar_sync(1 + t, 256);
      const float m_new = fmax3(m_ref, mxp, xb[(1 - hf) * 128 + r]);

      if (it == 0) {
        m_ref = (m_new == -INFINITY) ? 0.f : m_new;
      } else {
        mbar_wait(&pv_full[t], (it - 1) & 1);     // O_t complete up to block it-1; sP no longer read by the tensor core
        tc_fence_after();
        const bool need = (m_new - m_ref) * c > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float f = need ? ex2((m_ref - m_new) * c) : 1.f;
          uint32_t rr[32];
          tmem_ld_32x32b_x32(o_tmem, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * f);
          tmem_st_32x32b_x32(o_tmem, rr);
          tmem_st_wait();
          l *= f;
          if (need) m_ref = m_new;
        }
      }
      const float2 c2 = make_float2(c, c);
      const float2 nmc2 = make_float2(-m_ref * c, -m_ref * c);
      float2 acc = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      if (pingpong) named_bar_sync(3 + t, 512);
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a = ffma2(make_float2(s[8 * j8 + 2 * q], s[8 * j8 + 2 * q + 1]), c2, nmc2);
          const float2 e = ((PM >> q) & 1) ? exp2_poly2(a) : make_float2(ex2(a.x), ex2(a.y));
          if (q & 1) acc1 = fadd2(acc1, e); else acc = fadd2(acc, e);
          __half2 hv = __floats2half2_rn(e.x, e.y);
          pk[q] = *reinterpret_cast<uint32_t*>(&hv);
        }
        st_shared_v4(sP + ((j8 ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
      if (pingpong && !(t == 1 && it == nblk - 1)) named_bar_arrive(3 + (1 - t), 512);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[t]);
      l += (acc.x + acc.y) + (acc1.x + acc1.y);
    }

    // row sum of the other half, then this thread normalises and stores its 32 output columns
    float* xb = xch + (nblk & 1) * 512 + t * 256;
    xb[hf * 128 + r] = l;
    named_bar_sync(1 + t, 256);
    const float ltot = l + xb[(1 - hf) * 128 + r];
    float o[32];
    if (nblk > 0) {
      mbar_wait(&pv_full[t], (nblk - 1) & 1);
      tc_fence_after();
      tmem_ld_32x32b_x32(o_tmem, reinterpret_cast<uint32_t*>(o));
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0.f;
    }
    const int row = (qgrp * 2 + t) * 128 + r;
    if (row < p.L) {
      const float inv = ltot > 0.f ? 1.f / ltot : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.L + row) * p.ldo + h * DH + hf * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
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
  if (warp == MMA_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int PM>
int launch_pm(cudaStream_t st, const CUtensorMap& tm, const Params& p) {
  static bool configured = false;
  if (!configured) {
    BG_CUDA(cudaFuncSetAttribute(attn_rs_kernel<PM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int nq = (p.L + 127) / 128;
  dim3 grid((nq + 1) / 2, NHEAD, p.B);
  attn_rs_kernel<PM><<<grid, THREADS, SMEM_BYTES, st>>>(tm, p);
  return check_launch("attn_rs_kernel launch");
}

}  // namespace

// L > 128 path of launch_attention (attn.cu) when BG_ATTN_RS != 0
int launch_attention_rowsplit(cudaStream_t st, const AttnArgs& a, int poly) {
  CUtensorMap tm;
  BG_TRY(make_tmap_3d_f16(&tm, a.qkv, (uint64_t)a.B, (uint64_t)a.L, 3 * DMODEL, 3 * DMODEL, 128));
  Params p;
  p.out = a.out; p.ldo = a.ldo; p.B = a.B; p.L = a.L; p.nkb = (a.L + 127) / 128;
  p.key_mask = a.key_mask; p.blk_list = a.blk_list; p.blk_count = a.blk_count;
  p.scale_log2 = 1.4426950408889634f / 8.0f;
  static int pp = -1;
  if (pp < 0) {
    const char* e = getenv("BG_ATTN_PP");
    pp = e ? atoi(e) : 1;
  }
  p.pingpong = pp;
  if (poly == 0) return launch_pm<0x0>(st, tm, p);
  if (poly == 2) return launch_pm<0xA>(st, tm, p);
  return launch_pm<0x8>(st, tm, p);
}

}  // namespace bg
