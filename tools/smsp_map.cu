// Micro-benchmark: how are the warps of a thread block distributed over the four SM sub-partitions (schedulers)?
// A pure FFMA loop saturates one scheduler's FP32 pipe with a single warp (16 independent chains), so the run time of
// B blocks x W warps per SM is proportional to the MAXIMUM number of warps on one scheduler.  If a block's warp w always goes
// to scheduler w % 4, 4 blocks x 3 warps (12 warps) load the schedulers 4/4/4/0 and take 4 units, 3 blocks x 4 warps take 3.
// nvcc -arch=sm_100a -O3 -o tools/_bin/smsp_map tools/smsp_map.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, float a, float b) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  const float x = a + threadIdx.x * 1e-6f, y = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(x, y, acc[i]);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
static void run(int blocks_per_sm, int threads) {
  float* out;
  cudaMalloc(&out, 148 * 32 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  // dynamic smem sized so that exactly blocks_per_sm blocks fit on an SM (one wave: grid = 148 * blocks_per_sm)
  const int smem = (220 * 1024) / blocks_per_sm - 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<148 * blocks_per_sm, threads, smem>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<<<148 * blocks_per_sm, threads, smem>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = 148.0 * blocks_per_sm * threads * iters * 64.0;
  printf("%d blocks/SM x %3d threads (%2d warps/SM)  %.3f ms  %.2f TFMA/s\n", blocks_per_sm, threads, blocks_per_sm * threads / 32, ms,
         fmas / ms * 1e-9);
  cudaFree(out);
}
int main() {
  run(1, 32); run(1, 64); run(1, 96); run(1, 128);
  run(4, 96); run(3, 128); run(4, 192); run(6, 128); run(2, 384); run(8, 96); run(4, 160); run(2, 96); run(2, 192);
  return 0;
}
