// Epilogue shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 GEMM kernels.
#pragma once
#include "bg_internal.h"
#include "ptx.cuh"

namespace bg {

struct GemmParams {
  int M, N, K;
  int a_kwrap;   // 0, or the K period of A: A column = k % a_kwrap (K-concatenated weights [W_hi | W_lo] reuse A)
  void* out;
  int ldo;
  int out_f16;
  int relu;
  const float* bias;
  const float* resid;
  int ldr;
  const float* rowvec;
  int rows_per_vec;
  int ldv;
  int n_short, k_short; // column tiles with n0 < n_short run only k_short / 64 k-blocks (partly split weight matrix)
  const int* m_dev;     // optional device int: the kernels work on min(M, *m_dev) rows (token compaction)
  const int* row_map;   // optional: rowvec row = row_map[row] / rows_per_vec
  // implicit-GEMM convolution (conv_taps > 0): see ConvGeom in bg_internal.h
  int conv_taps, conv_kw, conv_cpb /* C / 64 */, conv_C, conv_W, conv_HW, conv_pad_w, conv_pad_h, conv_lo_term;
};

// coordinates of the A box of k-block kb for the tile whose first row (output pixel) is row0: {channel, x, y, image}
__device__ __forceinline__ void conv_coords(const GemmParams& p, int kb, int row0, int& c0, int& x, int& y, int& n) {
  const int per_term = p.conv_taps * p.conv_cpb;
  const int term = kb / per_term, r = kb - term * per_term;
  const int tap = r / p.conv_cpb, cc = r - tap * p.conv_cpb;
  c0 = (term == p.conv_lo_term ? p.conv_C : 0) + cc * 64;
  x = tap % p.conv_kw - p.conv_pad_w;
  n = row0 / p.conv_HW;
  y = (row0 - n * p.conv_HW) / p.conv_W + tap / p.conv_kw - p.conv_pad_h;
}

// the kernels' working copy of the parameters with the row count resolved on the device
__device__ __forceinline__ GemmParams gemm_resolve(const GemmParams& p) {
  GemmParams q = p;
  if (p.m_dev) q.M = min(p.M, *p.m_dev);
  return q;
}


constexpr int GEMM_XPOSE_PITCH = 33;   // floats; +1 keeps both access patterns of the 32x32 staging tile conflict-free

// One epilogue warp, one accumulator tile: `taddr` = TMEM address of column 0 of this warp's column half (lane quarter
// already encoded), rows [row0, row0 + 32), columns [colbase, colbase + 32 * CHUNKS).  tcgen05.ld hands each thread one
// accumulator ROW; fp32 outputs are transposed through the warp's private shared-memory tile `xp` so that every residual
// / row-vector load and every store covers one contiguous 128-byte row segment; fp16 outputs (bias / ReLU only) are
// written directly, 64 contiguous bytes per thread (the transpose would compete with the MMA operand reads for nothing).
template <int CHUNKS>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, uint32_t taddr, int row0, int colbase, float* xp,
                                                   int lane) {
  constexpr int PITCH = GEMM_XPOSE_PITCH;
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = colbase + c * 32;
        const int col = col0 + lane;
        // additive terms in the transposed (lane == column) layout: issue the loads before touching TMEM
        const float bias = p.bias ? __ldg(p.bias + col) : 0.f;
        float add[32];
        if (p.resid && !p.out_f16) {
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const int row = row0 + rr;
            add[rr] = row < p.M ? p.resid[(size_t)row * p.ldr + col] : 0.f;
          }
        } else {
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) add[rr] = 0.f;
        }
        if (p.rowvec && !p.out_f16) {
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const int row = row0 + rr;
            if (row < p.M)
              add[rr] += __ldg(p.rowvec + (size_t)((p.row_map ? p.row_map[row] : row) / p.rows_per_vec) * p.ldv + col);
          }
        }
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (p.out_f16) {
          // fp16 outputs (bias / ReLU only): 64 contiguous bytes per thread, written directly -- the shared-memory
          // transpose would compete with the MMA operand reads (96 B/clk of the 128 B/clk smem bandwidth) for no gain
          const int row = row0 + lane;
          if (row < p.M) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            if (p.bias) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 t = __ldg(bp + i);
                v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + (size_t)row * p.ldo + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __half2 h0 = __floats2half2_rn(v[8 * i], v[8 * i + 1]);
              __half2 h1 = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
              __half2 h2 = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]);
              __half2 h3 = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
              uint4 u;
              u.x = *reinterpret_cast<uint32_t*>(&h0);
              u.y = *reinterpret_cast<uint32_t*>(&h1);
              u.z = *reinterpret_cast<uint32_t*>(&h2);
              u.w = *reinterpret_cast<uint32_t*>(&h3);
              op[i] = u;
            }
          }
          continue;
        }
        __syncwarp();                                        // previous chunk's reads of xp are done
#pragma unroll
        for (int i = 0; i < 32; ++i) xp[lane * PITCH + i] = __uint_as_float(r[i]);
        __syncwarp();
        {
          float* ob = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            float v = xp[rr * PITCH + lane] + bias + add[rr];
            if (p.relu) v = fmaxf(v, 0.f);
            const int row = row0 + rr;
            if (row < p.M) ob[(size_t)row * p.ldo + col] = v;
          }
        }
      }
}

}  // namespace bg
