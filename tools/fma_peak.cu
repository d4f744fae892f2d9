// Micro-benchmark: fp32 FMA issue rate per SM on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2), with 1..8 warps per
// scheduler.  Used to decide what bounds the 7x7 depthwise stencil.  nvcc -arch=sm_100a -O3 -o tools/_bin/fma_peak
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long F2;
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float acc[16]; F2 acc2[16];
  for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x + i; acc2[i] = (F2)(threadIdx.x + i) * 0x100000001ull; }
  float x = a + threadIdx.x, y = b;
  F2 x2 = ((F2)__float_as_uint(x) << 32) | __float_as_uint(y), y2 = ((F2)__float_as_uint(y) << 32) | __float_as_uint(x);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) acc[i] = fmaf(acc[i], x, y);          // 3 distinct registers, one reused pair
        else if (MODE == 1) acc[i] = fmaf(x, y, acc[i]);     // stencil form: shared multiplicands
        else acc2[i] = fma2(x2, y2, acc2[i]);
      }
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i] + __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int threads) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148, threads>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0); k<MODE><<<148, threads>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fmas = 148.0 * threads * iters * 64.0 * (MODE == 2 ? 2 : 1);
  printf("%-28s %4d thr/SM  %.3f ms  %.2f TFMA/s\n", name, threads, ms, fmas / ms * 1e-9);
  cudaFree(out);
}
int main() {
  for (int t : {128, 256, 512, 1024}) { run<0>("FFMA acc*x+y", t); run<1>("FFMA x*y+acc", t); run<2>("FFMA2 x*y+acc", t); }
  return 0;
}
