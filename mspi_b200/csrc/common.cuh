// Shared helpers for the mspi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/mspi_b200.h"

namespace mspi {

// ---- host side -----------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define MSPI_CHECK_ARG(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) return mspi::set_error(MSPI_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define MSPI_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return mspi::set_error(MSPI_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
                             cudaGetErrorString(e__), __FILE__, __LINE__);                 \
  } while (0)

#define MSPI_LAUNCH_CHECK()                  \
  do {                                       \
    mspi::count_launch();                    \
    MSPI_CUDA(cudaGetLastError());           \
  } while (0)

int num_sms();

// Programmatic dependent launch (MSPI_PDL, default on): a kernel launched with this attribute may become resident while the
// previous kernel of its stream is still draining — every CTA of that kernel has passed pdl_launch_dependents() or exited —
// and runs its prologue (barrier init, TMEM allocation, tensor-map prefetch) there; pdl_wait() then blocks until the
// previous kernel has COMPLETED and its writes are visible.  Only kernels that execute pdl_wait() before their first global
// access are launched this way.  Inside a captured graph the edge becomes a programmatic dependency.
bool pdl_enabled();
bool pdl_small_enabled();   // MSPI_PDL_SMALL: the same for the small elementwise / pooling / norm kernels (launch_pdl)
inline int pdl_attr(cudaLaunchAttribute* a) {   // fills *a and returns 1 when PDL is on, else 0
  if (!pdl_enabled()) return 0;
  a->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}
// kern<<<grid, block, smem, stream>>>(args...) with the attribute above; only for kernels that start with pdl_wait().
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_small_enabled() ? pdl_attr(&attr[0]) : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- device side ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  lo = __bfloat162float(v.x);
  hi = __bfloat162float(v.y);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// r -> (r / d, r % d).  Index decompositions run once per thread-element, and a 64-bit division costs ~100 instructions
// on the GPU (it made the pooling / upsampling / LayerNorm kernels issue bound): use the 32-bit unit whenever r fits.
__device__ __forceinline__ int divmod(long long& r, int d) {
  int rem;
  if (static_cast<unsigned long long>(r) <= 0xffffffffull) {
    const unsigned r32 = static_cast<unsigned>(r), q = r32 / static_cast<unsigned>(d);
    rem = static_cast<int>(r32 - q * static_cast<unsigned>(d));
    r = q;
  } else {
    rem = static_cast<int>(r % d);
    r /= d;
  }
  return rem;
}

// Packed fp32 pair in a 64-bit register: sm_100's {fma,add,mul}.f32x2 retire two fp32 operations per lane per issue slot,
// each half rounded exactly like the scalar instruction.
typedef unsigned long long F2;
__device__ __forceinline__ F2 pack2(float lo, float hi) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(F2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
  F2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
  F2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
  F2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MSPI_ACT_RELU: return fmaxf(v, 0.f);
    case MSPI_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case MSPI_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}
#endif  // __CUDACC__

}  // namespace mspi
