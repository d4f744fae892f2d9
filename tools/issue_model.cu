// Micro-benchmark: does a packed fma.rn.f32x2 (FFMA2, two FP32-pipe cycles per warp) also hold the scheduler's issue port for
// two cycles, or can another instruction (integer ALU) issue in its shadow?  Each loop iteration has NF independent FFMA2 (or
// FFMA) and NI independent LOP3-class integer ops on other registers; 16 warps per SM (4 per scheduler), one block per SM.
//   cycles / iteration / warp-per-scheduler:  FFMA2 only = 2 NF;  shadow issue => max(2 NF, NF + NI);  blocked => 2 NF + NI.
// nvcc -arch=sm_100a -O3 -o tools/_bin/issue_model tools/issue_model.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long F2;
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE, int NI>   // MODE 0: 16 FFMA2, 1: 32 FFMA (same FMA work), 2: no FP work
__global__ void k(float* out, int iters, float a, float b, unsigned m) {
  F2 acc2[16]; float acc[32]; unsigned z[16];
  for (int i = 0; i < 16; ++i) { acc2[i] = (F2)(threadIdx.x + i) * 0x100000001ull; z[i] = threadIdx.x * 7 + i; }
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
  const float x = a + threadIdx.x * 1e-6f, y = b;
  const F2 x2 = ((F2)__float_as_uint(x) << 32) | __float_as_uint(y);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) acc2[i] = fma2(x2, acc2[(i + 1) & 15], acc2[i]);
      if (MODE == 1) { acc[2 * i] = fmaf(x, acc[(2 * i + 2) & 31], acc[2 * i]); acc[2 * i + 1] = fmaf(y, acc[(2 * i + 3) & 31], acc[2 * i + 1]); }
      if (i < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(m), "r"(z[(i + 1) & 15]));
    }
  }
  float s = 0; unsigned zz = 0;
  for (int i = 0; i < 16; ++i) { s += __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32)); zz ^= z[i]; }
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + zz;
}
template <int MODE, int NI> void run(const char* name) {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, threads = 512;
  k<MODE, NI><<<148, threads>>>(out, 100, 1.0001f, 0.5f, 0x5a5a5a5au);
  cudaEventRecord(e0); k<MODE, NI><<<148, threads>>>(out, iters, 1.0001f, 0.5f, 0x5a5a5a5au); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  // cycles per iteration per scheduler with 4 warps on it (at the nominal clock; relative numbers are what matter)
  printf("%-34s %.3f ms  %.1f cycles/iteration/scheduler (4 warps)\n", name, ms, ms * 1e-3 * khz * 1e3 / iters);
  cudaFree(out);
}
int main() {
  run<0, 0>("16 FFMA2");            run<0, 8>("16 FFMA2 + 8 LOP3");   run<0, 16>("16 FFMA2 + 16 LOP3");
  run<1, 0>("32 FFMA");             run<1, 8>("32 FFMA + 8 LOP3");    run<1, 16>("32 FFMA + 16 LOP3");
  run<2, 8>("8 LOP3");              run<2, 16>("16 LOP3");
  return 0;
}
