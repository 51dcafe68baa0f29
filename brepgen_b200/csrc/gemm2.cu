// 2-CTA (cta_group::2) tcgen05 GEMM:  out[M,N] = epi( A[M,K] * W[N,K]^T ), default for N % 256 == 0 (gemm.cu: N = 128 tiles).
//
// A CTA pair (thread-block cluster of 2, same TPC) computes one 256 x 256 tile: each CTA owns 128 rows of M (its own A
// tile and its own 128 x 256 fp32 accumulator in TMEM) and loads only HALF of the 256 weight rows; the leader CTA's single
// MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), which reads A from both CTAs' shared memory and
// the two halves of B from the two CTAs.  Per CTA and 64-wide k-block this moves 16 KB (A) + 16 KB (B/2) instead of
// 16 + 32 KB through L2 -> TMA -> smem -> tensor core, which is what caps the 1-CTA kernel at ~75 % tensor-pipe activity.
//
//   warp 0 (both CTAs): TMA producer (cp.async.bulk.tensor ... .cta_group::2, completion bytes land on the LEADER's
//                       full barrier); waits on its own CTA's empty barrier
//   warp 1 (leader)   : MMA issuer; tcgen05.commit ... multicast::cluster arrives on the empty / tmem-full barriers of
//                       BOTH CTAs
//   warp 2 (both)     : TMEM allocator (tcgen05.alloc.cta_group::2, 512 columns: accumulator double buffer)
//   warps 4-11 (both) : epilogue; all 512 epilogue threads of the pair arrive on the leader's tmem-empty barrier (the
//                       peer through a cluster-mapped address)
//
// Three epilogues (template parameter RES):
//   RES = 0  gemm_epilogue.cuh: fp16 outputs written directly, fp32 outputs transposed through shared memory; 6-stage ring.
//   RES = 1  the in-place fp32 residual GEMMs  X += A W^T + b  (out-proj, FFN2), which are HBM-bound by construction
//            (256 KB of residual in + result out per 128 x 256 tile against 2048 tensor cycles): each epilogue warp
//            TMA-loads the residual 32 x 32 sub-tile into a swizzled staging buffer one chunk ahead, adds accumulator + bias
//            there and TMA-stores it -- deep asynchronous queues instead of per-warp load / store bursts.  Measured at
//            M = 256 000 (B = 64): out-proj 665 -> 548 us (4.31 TB/s), FFN2 571 -> 426 us (4.93 TB/s).  5-stage ring, two
//            staging buffers per warp (4 + 3: 569 / 428 us).
//   RES = 2  fp16 outputs (q|k|v, FFN1): the direct form of RES = 0 writes 64 B per thread into 32 different rows per warp
//            instruction (32 separate sectors); at K = 768 a tile's main loop is only 6144 cycles and that store pattern paced
//            the whole kernel (timing experiment without the stores: q|k 602 -> 393 us, FFN1 388 -> 270 us).  Here every
//            epilogue warp converts 32 rows x 64 columns, writes them as 128-byte rows into a swizzled staging buffer and
//            TMA-stores the box (two buffers per warp, the store of chunk c overlaps the conversion of chunk c + 1).
#include <stdlib.h>

#include "bg_internal.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace bg {

namespace {

constexpr int BM_CTA = 128;     // rows per CTA; the pair covers 256
constexpr int BK = 64;
constexpr int NSTAGES = 6;
constexpr int A_BYTES = BM_CTA * BK * 2;
// BN = 256 (default) or 128 (N % 256 != 0: the Cout = 128 convolutions of the VAEs -- a 1-CTA 128 x 128 tile reads 32 KB
// of shared memory per 256 tensor cycles and sits at ~37 % tensor-pipe activity; the pair's 256 x 128 tile reads 24 KB)
template <int BN> struct TL {
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB (BN = 256) / 24 KB (BN = 128) per CTA
  static constexpr int TMEM_COLS = 2 * BN;                // accumulator double buffer
};
constexpr int BAR_BYTES = 256;
constexpr int XPOSE_BYTES = 8 * 32 * GEMM_XPOSE_PITCH * 4;
template <int BN> constexpr int smem_bytes_res0() { return NSTAGES * TL<BN>::STAGE_BYTES + BAR_BYTES + XPOSE_BYTES + 1024; }
// RES == 1 (TMA-staged residual epilogue): 5-stage ring, 1 KB of barriers, then 8 warps x 2 buffers of 32 rows x 128 B
// RES == 2 (fp16 TMA-store epilogue): 5-stage ring, two staging buffers per warp
constexpr int X_BUF_BYTES = 32 * 128;
#ifndef BG_RES1_STAGES        // measured at M = 256 000: 4 stages + 3 staging buffers 569 us (out-proj) / 428 (FFN2), 5 + 2: 548 / 426
#define BG_RES1_STAGES 5
#define BG_RES1_BUFS 2
#endif
template <int RES, int BN> struct XL {               // shared-memory layout of the staged-epilogue variants
  static constexpr int STAGES = RES == 2 ? 5 : BG_RES1_STAGES;
  static constexpr int BUFS = RES == 2 ? 2 : BG_RES1_BUFS;
  static constexpr int OFF_BAR = STAGES * TL<BN>::STAGE_BYTES;
  static constexpr int OFF_STG = OFF_BAR + 1024;
  static constexpr int SMEM_BYTES = OFF_STG + 8 * BUFS * X_BUF_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "staged epilogue does not fit in shared memory");
};

// extra kernel parameter of the RES == 1 variant only
template <int RES> struct EpiMaps {};
template <> struct EpiMaps<1> { CUtensorMap x; };     // fp32 [M][ldo] residual / output matrix, box 32 x 32, SWIZZLE_128B
template <> struct EpiMaps<2> { CUtensorMap x; };     // fp16 [M][ldo] output matrix, box 32 rows x 64 columns, SWIZZLE_128B

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;               // shared::cluster address of the same offset in the even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {     // arrives on `bar` (same offset) in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int RES, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm2_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p_in,
                 const __grid_constant__ EpiMaps<RES> em) {
  const GemmParams p = gemm_resolve(p_in);
  constexpr int STAGE_BYTES = TL<BN>::STAGE_BYTES, TMEM_COLS = TL<BN>::TMEM_COLS;
  constexpr int STAGES = RES ? XL<RES, BN>::STAGES : NSTAGES;
  constexpr int X_BUFS = XL<RES, BN>::BUFS, X_OFF_BAR = XL<RES, BN>::OFF_BAR, X_OFF_STG = XL<RES, BN>::OFF_STG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (RES != 0) tma_prefetch_desc(&em.x);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);          // leader: its own arrive.expect_tx; the bytes of both CTAs complete on it
      mbar_init(&empty[i], 1);         // one multicast tcgen05.commit per phase
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 512);      // leader: 256 epilogue threads of each CTA
    }
    if constexpr (RES == 1) {           // residual-load barriers: 3 per epilogue warp, one arrive.expect_tx each
      uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + X_OFF_BAR + 256);
      for (int i = 0; i < 8 * X_BUFS; ++i) mbar_init(&xbar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + 2 * BM_CTA - 1) / (2 * BM_CTA);
  const int num_n = p.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = p.K / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int nk = (n_blk * BN < p.n_short) ? p.k_short / BK : num_k;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          const uint32_t leader_full = smem_u32(&full[stage]) & PEER_MASK;
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
          if (p.conv_taps) {      // implicit convolution: this CTA's 128 output pixels, box shifted by the tap
            int c0, x, y, n;
            conv_coords(p, kb, m_blk * 2 * BM_CTA + (int)rank * BM_CTA, c0, x, y, n);
            tma_load_4d_2sm(sA, &tmA, leader_full, c0, x, y, n);
          } else {
            const int ka = p.a_kwrap ? (kb * BK) % p.a_kwrap : kb * BK;
            tma_load_2d_2sm(sA, &tmA, leader_full, ka, m_blk * 2 * BM_CTA + (int)rank * BM_CTA);
          }
          tma_load_2d_2sm(sB, &tmB, leader_full, kb * BK, n_blk * BN + (int)rank * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(2 * BM_CTA, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t accphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty[acc], accphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int nk = ((tile % num_n) * BN < p.n_short) ? p.k_short / BK : num_k;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss_2cta(d_tmem, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc,
                             (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    constexpr int CHUNKS = BN / 64;
    float* xp = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BAR_BYTES) + (warp - 4) * 32 * GEMM_XPOSE_PITCH;
    int acc = 0;
    uint32_t accphase = 0;
    // RES == 1: this warp's three 4 KB staging buffers (32 rows x 128 B, SWIZZLE_128B), their load barriers and the
    // running chunk counter q (buffer q % 3, barrier phase (q / 3) & 1)
    uint8_t* xstg = smem + X_OFF_STG + (warp - 4) * X_BUFS * X_BUF_BYTES;
    uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + X_OFF_BAR + 256) + (warp - 4) * X_BUFS;
    uint32_t q = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row0 = m_blk * 2 * BM_CTA + (int)rank * BM_CTA + ew * 32;
      const int colbase = n_blk * BN + half * (BN / 2);
      if constexpr (RES == 1) {
        // ---- TMA-staged residual epilogue: out = resid (in place) + acc + bias, fp32 ----
        // residual chunk c+1 is loaded (TMA, swizzled rows) while chunk c is combined in shared memory and stored (TMA):
        // the memory system sees deep asynchronous queues instead of per-warp load / store bursts
        auto issue_load = [&](uint32_t qq, int c) {      // lane 0 only
          bulk_wait_read<X_BUFS - 2>();                  // the store that last read buffer qq % BUFS is done with it
          const uint32_t b = qq % X_BUFS;
          mbar_arrive_expect_tx(&xbar[b], X_BUF_BYTES);
          tma_load_2d(xstg + b * X_BUF_BYTES, &em.x, &xbar[b], colbase + c * 32, row0);
        };
        if (lane == 0) issue_load(q, 0);                 // overlaps the wait for the accumulator
        mbar_wait(&tfull[acc], accphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + half * (BN / 2);
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
          if (c + 1 < CHUNKS && lane == 0) issue_load(q + 1, c + 1);
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          const uint32_t b = q % X_BUFS;
          mbar_wait(&xbar[b], (q / X_BUFS) & 1);         // residual chunk c is in shared memory
          tmem_ld_wait();
          const int col0 = colbase + c * 32;
          const uint32_t rowaddr = smem_u32(xstg + b * X_BUF_BYTES) + lane * 128;   // this thread's accumulator row
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = rowaddr + ((j ^ (lane & 7)) << 4);                   // 16-byte chunk j of the row, 128B swizzle
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
            v.x += __uint_as_float(r[4 * j]) + bb.x;
            v.y += __uint_as_float(r[4 * j + 1]) + bb.y;
            v.z += __uint_as_float(r[4 * j + 2]) + bb.z;
            v.w += __uint_as_float(r[4 * j + 3]) + bb.w;
            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          }
          fence_proxy_async_smem();                      // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&em.x, xstg + b * X_BUF_BYTES, col0, row0);    // rows >= M are clipped by the tensor map
            bulk_commit();
          }
          ++q;
        }
      } else if constexpr (RES == 2) {
        // ---- fp16 output through a swizzled staging buffer and TMA stores: out = fp16(relu(acc + bias)) ----
        mbar_wait(&tfull[acc], accphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + half * (BN / 2);
#pragma unroll 1
        for (int c = 0; c < BN / 2 / 64; ++c) {              // two chunks of 64 columns per warp
          uint32_t r[64];
          tmem_ld_32x32b_x32(taddr + c * 64, r);
          tmem_ld_32x32b_x32(taddr + c * 64 + 32, r + 32);
          const uint32_t b = q & 1;
          if (lane == 0) bulk_wait_read<1>();                // the store that last read buffer b (two chunks ago) is done
          __syncwarp();
          tmem_ld_wait();
          const int col0 = colbase + c * 64;
          const uint32_t rowaddr = smem_u32(xstg + b * X_BUF_BYTES) + lane * 128;   // this thread's output row (64 x fp16)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]);
            if (p.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + 2 * j);
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + 2 * j + 1);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
            __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
            st_shared_v4(rowaddr + ((j ^ (lane & 7)) << 4), *reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                         *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          }
          fence_proxy_async_smem();                          // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&em.x, xstg + b * X_BUF_BYTES, col0, row0);    // rows >= M are clipped by the tensor map
            bulk_commit();
          }
          ++q;
        }
      } else {
        mbar_wait(&tfull[acc], accphase);
        tc_fence_after();
        gemm_epilogue_tile<CHUNKS>(p, tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + half * (BN / 2), row0, colbase, xp, lane);
      }
      tc_fence_before();
      if (leader) mbar_arrive(&tempty[acc]);
      else mbar_arrive_cluster(&tempty[acc], 0);
      acc ^= 1;
      if (acc == 0) accphase ^= 1;
    }
    if constexpr (RES != 0) {
      if (lane == 0) bulk_wait_all();   // every TMA store of this thread has completed before the CTA may exit
    }
  }

  tc_fence_before();
  cluster_sync_all();       // the peer may still be signalling / reading this CTA's barriers and TMEM until here
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// CTA-pair path of launch_gemm_f16: 256-wide column tiles when N % 256 == 0, else 128-wide
template <int BN>
int launch_gemm2_bn(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p) {
  const int num_tiles = ((p.M + 255) / 256) * (p.N / BN);
  const int max_clusters = num_sms() / 2;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  if (!p.out_f16 && p.resid != nullptr && p.resid == p.out && p.ldr == p.ldo && p.rowvec == nullptr) {
    BG_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm2_f16_kernel<1, BN>), XL<1, BN>::SMEM_BYTES));
    EpiMaps<1> em;
    BG_TRY(make_tmap_2d_f32(&em.x, p.out, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldo, 32, 32));
    gemm2_f16_kernel<1, BN><<<2 * clusters, 384, XL<1, BN>::SMEM_BYTES, st>>>(tmA, tmB, p, em);
  } else if (p.out_f16 && p.ldo % 8 == 0) {
    BG_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm2_f16_kernel<2, BN>), XL<2, BN>::SMEM_BYTES));
    EpiMaps<2> em;
    BG_TRY(make_tmap_2d_f16(&em.x, p.out, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldo, 32, 64));
    gemm2_f16_kernel<2, BN><<<2 * clusters, 384, XL<2, BN>::SMEM_BYTES, st>>>(tmA, tmB, p, em);
  } else {
    BG_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&gemm2_f16_kernel<0, BN>), smem_bytes_res0<BN>()));
    gemm2_f16_kernel<0, BN><<<2 * clusters, 384, smem_bytes_res0<BN>(), st>>>(tmA, tmB, p, EpiMaps<0>{});
  }
  return check_launch("gemm2_f16_kernel launch");
}

}  // namespace

int launch_gemm2_f16(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p) {
  return p.N % 256 == 0 ? launch_gemm2_bn<256>(st, tmA, tmB, p) : launch_gemm2_bn<128>(st, tmA, tmB, p);
}

}  // namespace bg
