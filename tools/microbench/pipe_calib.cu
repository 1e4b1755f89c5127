// Calibrates the sm_100a issue model for the inlier-count kernel: register-only loops with a fixed
// instruction mix, 8 warps per SMSP, reports warp-instructions per cycle per SMSP and the SM clock.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipe_calib pipe_calib.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ int iadd(int a, int b) { int d; asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ void pred_inc(float d, float tau, int& cnt) {
  asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(cnt) : "f"(d), "f"(tau));
}
__device__ __forceinline__ void setp_only(float d, float tau, int& cnt) {
  asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p bra L%=; L%=: }" :: "f"(d), "f"(tau));
}

constexpr int ITERS = 4096;
constexpr int NC = 16;  // independent chains

template <int MODE>
__global__ void __launch_bounds__(256) k(float seed, int* out, long long* cycles) {
  float a = seed + threadIdx.x * 1e-7f, b = 0.999f;
  float f[NC]; u64 g[NC]; int c[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) { f[i] = seed * i; g[i] = (u64)__float_as_uint(seed * i) | ((u64)__float_as_uint(seed) << 32); c[i] = i; }
  u64 A = (u64)__float_as_uint(a) | ((u64)__float_as_uint(a) << 32), B = (u64)__float_as_uint(b) | ((u64)__float_as_uint(b) << 32);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if (MODE == 0) f[i] = ffma(f[i], b, a);                                   // FFMA only
      if (MODE == 1) g[i] = fma2(g[i], B, A);                                   // FFMA2 only
      if (MODE == 2) { f[i] = ffma(f[i], b, a); c[i] = iadd(c[i], it); }        // FFMA : IADD = 1:1
      if (MODE == 3) { g[i] = fma2(g[i], B, A); c[i] = iadd(c[i], it); }        // FFMA2 : IADD = 1:1
      if (MODE == 4) c[i] = iadd(c[i], it);                                     // IADD only
      if (MODE == 5) { g[i] = fma2(g[i], B, A); g[i] = fma2(g[i], B, A); g[i] = fma2(g[i], B, A);   // count mix:
                       pred_inc(__uint_as_float((unsigned)g[i]), a, c[i]); pred_inc(__uint_as_float((unsigned)(g[i] >> 32)), a, c[i]); }  // 3 FFMA2 + 2 FSETP + 2 @P IADD
      if (MODE == 6) { f[i] = ffma(f[i], b, a); f[i] = ffma(f[i], b, a); f[i] = ffma(f[i], b, a); pred_inc(f[i], a, c[i]); }  // 3 FFMA + FSETP + @P IADD
      if (MODE == 7) { pred_inc(f[i], a, c[i]); }                               // FSETP + @P IADD only
      if (MODE == 8) { g[i] = fma2(g[i], B, A); g[i] = fma2(g[i], B, A); g[i] = fma2(g[i], B, A); c[i] = iadd(c[i], it); }  // 3 FFMA2 + 1 IADD
    }
  }
  long long t1 = clock64();
  int s = 0; float fs = 0;
#pragma unroll
  for (int i = 0; i < NC; ++i) { s += c[i]; fs += f[i] + __uint_as_float((unsigned)g[i]); }
  if (fs == 123.456f) s++;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  int blocks = sms * 4;  // 4 blocks x 8 warps = 32 warps/SM = 8 per SMSP
  int* d_out; long long* d_cyc; cudaMalloc(&d_out, blocks * 256 * 4); cudaMalloc(&d_cyc, 8);
  const char* names[] = {"FFMA", "FFMA2", "FFMA+IADD", "FFMA2+IADD", "IADD", "3FFMA2+2FSETP+2@IADD", "3FFMA+FSETP+@IADD", "FSETP+@IADD", "3FFMA2+IADD"};
  const double instr_per_iter[] = {1, 1, 2, 2, 1, 7, 5, 2, 4};
  const double fma_per_iter[] = {1, 2, 1, 2, 0, 6, 3, 0, 6};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(M) { k<M><<<blocks, 256>>>(1.0001f, d_out, d_cyc); cudaDeviceSynchronize(); cudaEventRecord(e0); k<M><<<blocks, 256>>>(1.0001f, d_out, d_cyc); cudaEventRecord(e1); cudaEventSynchronize(e1); \
    float ms; cudaEventElapsedTime(&ms, e0, e1); long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost); \
    double wi = (double)ITERS * NC * instr_per_iter[M] * 8;  /* warp-instr per SMSP */ \
    double mhz = cyc / (ms * 1e-3) / 1e6; \
    printf("%-22s %8.3f ms  %9lld cyc  clk~%.0f MHz  %.3f warp-instr/cyc/SMSP  %.3f FMA-lane-ops/cyc/SMSP (32 = peak)  err=%s\n", names[M], ms, cyc, mhz, wi / cyc, (double)ITERS * NC * fma_per_iter[M] * 8 * 32 / cyc, cudaGetErrorString(cudaGetLastError())); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
  return 0;
}
