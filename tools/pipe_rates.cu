// Micro-benchmark (round-2 groundwork): issue / pipe rates of the instructions on the attention softmax path, per SM
// sub-partition (SMSP), with 1, 2 and 4 resident warps per SMSP and with independent chains so that latency is hidden
// inside a warp.  Output: cycles per warp-instruction per SMSP.  The softmax warpgroups of attn.cu run exactly one
// warp of each warpgroup per SMSP, i.e. the "2 warps" column is the regime that matters.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates tools/pipe_rates.cu && ./pipe_rates
//
// Each kernel runs ITER iterations of UNROLL independent operations per thread; one CTA of (warps_per_smsp * 4) warps
// on one SM; time = clock64 delta of warp 0.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2000;
constexpr int UNROLL = 16;

enum Op { FFMA, FFMA2_RRR, FFMA2, FADD2, FMUL2, FMNMX3, F2FP, MUFU, MUFU_FFMA2_1_4, MUFU_FFMA_1_8, FFMA2_F2FP, FFMA2_FMNMX3,
          MUFU_H2, ADD_F32_F16, MUFU_3OTHER, MUFU_7OTHER, MUFU_H2_7OTHER, HADD2, FMNMX, MUFU_F2FP, MUFU_2F2FP };

template <int OP>
__global__ void rate_kernel(float* out, long long* cycles, float seed) {
  float x[UNROLL], y[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) {
    x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    y[i] = seed * 0.5f + i * 0.002f;
  }
  const float a = seed * 0.999f, b = seed * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (OP == FFMA) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
      } else if (OP == FFMA2_RRR) {           // three distinct 64-bit register operands (the polynomial's q = q * f + k form)
        asm volatile(
            "{\n\t.reg .b64 v, p, q;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 p, {%2, %3};\n\tmov.b64 q, {%3, %2};\n\t"
            "fma.rn.f32x2 v, v, p, q;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(x[i]), "+f"(x[(i + 1) % UNROLL])
            : "f"(y[i]), "f"(y[(i + 1) % UNROLL]));
      } else if (OP == FFMA2) {
        asm volatile(
            "{\n\t.reg .b64 v, p, q;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 p, {%2, %2};\n\tmov.b64 q, {%3, %3};\n\t"
            "fma.rn.f32x2 v, v, p, q;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(x[i]), "+f"(y[i])
            : "f"(a), "f"(b));
      } else if (OP == FADD2) {
        asm volatile(
            "{\n\t.reg .b64 v, q;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 q, {%2, %2};\n\t"
            "add.rn.f32x2 v, v, q;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(x[i]), "+f"(y[i])
            : "f"(b));
      } else if (OP == FMUL2) {
        asm volatile(
            "{\n\t.reg .b64 v, q;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 q, {%2, %2};\n\t"
            "mul.rn.f32x2 v, v, q;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(x[i]), "+f"(y[i])
            : "f"(a));
      } else if (OP == FMNMX3) {
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(b));
      } else if (OP == F2FP) {
        unsigned h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[i]), "f"(y[i]));
        x[i] = __uint_as_float(h & 0x3fffffffu);
      } else if (OP == MUFU) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (OP == MUFU_H2) {             // two fp16 exponentials per lane and instruction
        unsigned h = __float_as_uint(x[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        x[i] = __uint_as_float(h);
      } else if (OP == ADD_F32_F16) {         // mixed-precision add (PTX 8.6, sm_100): f32 accumulator += f16
        unsigned short hh = (unsigned short)(__float_as_uint(y[i]) & 0x3fffu);
        asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(x[i]) : "h"(hh));
      } else if (OP == HADD2) {
        unsigned h = __float_as_uint(x[i]) & 0x3fff3fffu, g = __float_as_uint(y[i]) & 0x3fff3fffu;
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h) : "r"(g));
        x[i] = __uint_as_float(h);
      } else if (OP == FMNMX) {
        asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y[i]));
      } else if (OP == MUFU_3OTHER || OP == MUFU_7OTHER || OP == MUFU_H2_7OTHER) {
        // one exponential followed by 3 / 7 independent single-cycle-class instructions (FFMA): can ONE warp keep the XU pipe
        // busy while it also issues its other work?
        if (OP == MUFU_H2_7OTHER) {
          unsigned h = __float_as_uint(x[i]);
          asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
          x[i] = __uint_as_float(h);
        } else {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
        constexpr int NO = OP == MUFU_3OTHER ? 3 : 7;
#pragma unroll
        for (int k = 0; k < NO; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[(i + k) % UNROLL]) : "f"(a), "f"(b));
      } else if (OP == MUFU_F2FP || OP == MUFU_2F2FP) {   // do MUFU.EX2 and F2FP share a pipe?  (8 + 4 = 12 if they do, 8 if not)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        unsigned h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[i]), "f"(y[(i + 1) % UNROLL]));
        y[i] = __uint_as_float(h & 0x3fffffffu);
        if (OP == MUFU_2F2FP) {
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[(i + 2) % UNROLL]), "f"(y[(i + 3) % UNROLL]));
          y[(i + 2) % UNROLL] = __uint_as_float(h & 0x3fffffffu);
        }
      } else if (OP == MUFU_FFMA2_1_4) {      // the attention pattern: 1 MUFU per 4 packed FMA-pipe instructions
        if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile(
            "{\n\t.reg .b64 v, p, q;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 p, {%2, %2};\n\tmov.b64 q, {%3, %3};\n\t"
            "fma.rn.f32x2 v, v, p, q;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(y[i]), "+f"(y[(i + 1) % UNROLL])
            : "f"(a), "f"(b));
      } else if (OP == MUFU_FFMA_1_8) {
        if ((i & 7) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(a), "f"(b));
      } else if (OP == FFMA2_F2FP) {          // FMA pipe + conversion: do they overlap?
        asm volatile(
            "{\n\t.reg .b64 v, p, q;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 p, {%2, %2};\n\tmov.b64 q, {%3, %3};\n\t"
            "fma.rn.f32x2 v, v, p, q;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(x[i]), "+f"(x[(i + 1) % UNROLL])
            : "f"(a), "f"(b));
        unsigned h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[i]), "f"(y[(i + 1) % UNROLL]));
        y[i] = __uint_as_float(h & 0x3fffffffu);
      } else if (OP == FFMA2_FMNMX3) {        // FMA pipe + ALU pipe
        asm volatile(
            "{\n\t.reg .b64 v, p, q;\n\tmov.b64 v, {%0, %1};\n\tmov.b64 p, {%2, %2};\n\tmov.b64 q, {%3, %3};\n\t"
            "fma.rn.f32x2 v, v, p, q;\n\tmov.b64 {%0, %1}, v;\n\t}"
            : "+f"(x[i]), "+f"(x[(i + 1) % UNROLL])
            : "f"(a), "f"(b));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(y[(i + 1) % UNROLL]), "f"(b));
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) acc += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_unroll_step) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  printf("%-28s", name);
  for (int wps : {1, 2, 4}) {
    const int threads = wps * 4 * 32;
    rate_kernel<OP><<<1, threads>>>(out, cyc, 1.0f);     // warm-up
    rate_kernel<OP><<<1, threads>>>(out, cyc, 1.0f);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    // warp-instructions issued per SMSP = wps * ITER * UNROLL * instr_per_unroll_step
    const double per = (double)c / ((double)wps * ITER * UNROLL * instr_per_unroll_step);
    printf("  %d warp/SMSP: %6.2f cyc/instr", wps, per);
  }
  printf("\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  printf("cycles per warp-instruction per SMSP (lower = faster); independent chains of %d per thread\n", UNROLL);
  run<FFMA>("FFMA (3-reg)", 1);
  run<FFMA2>("FFMA2 (scalar b, c)", 1);
  run<FFMA2_RRR>("FFMA2 (3 register pairs)", 1);
  run<FADD2>("FADD2 (add.rn.f32x2)", 1);
  run<FMUL2>("FMUL2 (mul.rn.f32x2)", 1);
  run<FMNMX3>("FMNMX3 (max.f32 a,b,c)", 1);
  run<F2FP>("F2FP.F16.F32.PACK_AB", 1);
  run<MUFU>("MUFU.EX2", 1);
  run<FMNMX>("FMNMX (max.f32 a,b)", 1);
  run<MUFU_H2>("MUFU.EX2.F16x2", 1);
  run<ADD_F32_F16>("add.rn.f32.f16 (mixed)", 1);
  run<HADD2>("HADD2", 1);
  printf("mixed (cycles per unroll step = the group in the name):\n");
  run<MUFU_F2FP>("1 MUFU + 1 F2FP per step", 1);
  run<MUFU_2F2FP>("1 MUFU + 2 F2FP per step", 1);
  run<MUFU_3OTHER>("1 MUFU + 3 FFMA per step", 1);
  run<MUFU_7OTHER>("1 MUFU + 7 FFMA per step", 1);
  run<MUFU_H2_7OTHER>("1 MUFU.F16x2 + 7 FFMA per step", 1);
  run<MUFU_FFMA2_1_4>("1/4 MUFU + 1 FFMA2 per step", 1);
  run<MUFU_FFMA_1_8>("1/8 MUFU + 1 FFMA per step", 1);
  run<FFMA2_F2FP>("FFMA2 + F2FP per step", 1);
  run<FFMA2_FMNMX3>("FFMA2 + FMNMX3 per step", 1);
  return 0;
}
