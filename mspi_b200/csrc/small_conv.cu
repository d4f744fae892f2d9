// (1,3,3) convolution with very few channels (Cin = Cout = 8 or 16) + folded BatchNorm + ReLU, on CUDA cores.
//
// SlowFast's fast pathway keeps 8 / 16 channels at 56x96 / 28x48 over 64 frames (sf.py / resnet_helper.py BottleneckTransform
// branch2.b with dim_inner = 8, 16).  As an implicit GEMM such a layer is 9 taps of K = 8: every (row, tap) of a tile is its own
// TMA request for 16 bytes, and the tcgen05 kernel ran it at 4 TFLOP/s (0.75 ms for a 176 MB tensor whose HBM floor is
// 0.06 ms; profiles/r02_small_kernels.md §4).  Here a thread owns NPX consecutive output pixels and all C output channels:
// the 3 x (NPX + 2) input pixels are read as 16-byte vectors, the [tap][ci][co] weights (BatchNorm scale folded in on the
// host, fp32) come from shared memory as broadcast float4 and are shared by the thread's NPX pixels.
#include <cstdlib>

#include "common.cuh"

namespace mspi {
namespace {

template <int C>
struct PixVec;   // one pixel's C bf16 channels in registers
template <>
struct PixVec<8> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
    unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
  }
};
template <>
struct PixVec<16> {
  uint4 v, u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    v = __ldg(reinterpret_cast<const uint4*>(p));
    u = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
  __device__ __forceinline__ void zero() { v = u = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float (&f)[16]) const {
    unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
    unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
    unpack_bf16x2(u.x, f[8], f[9]); unpack_bf16x2(u.y, f[10], f[11]);
    unpack_bf16x2(u.z, f[12], f[13]); unpack_bf16x2(u.w, f[14], f[15]);
  }
};

template <int C, int NPX>
__global__ void __launch_bounds__(256)
conv133_small_kernel(const __nv_bfloat16* __restrict__ x, long long xcs, const float* __restrict__ wp,
                     const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, long long ycs, long long items, int gpr,
                     int h, int wd, int act) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float4 s_w[9 * C * C / 4];   // [tap][ci][co], scale folded in
  __shared__ float s_shift[C];
  for (int i = threadIdx.x; i < 9 * C * C / 4; i += blockDim.x) s_w[i] = __ldg(reinterpret_cast<const float4*>(wp) + i);
  if (threadIdx.x < C) s_shift[threadIdx.x] = shift ? __ldg(shift + threadIdx.x) : 0.f;
  __syncthreads();
  for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = it;
    const int x0 = divmod(r, gpr) * NPX;
    const long long row = r;                 // (n * T + t) * H + yy
    long long r2 = row;
    const int yy = divmod(r2, h);
    float acc[NPX][C];
#pragma unroll
    for (int p = 0; p < NPX; ++p)
#pragma unroll
      for (int co = 0; co < C; ++co) acc[p][co] = s_shift[co];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ry = yy + kh - 1;
      if (ry < 0 || ry >= h) continue;
      const __nv_bfloat16* rp = x + ((row + kh - 1) * wd + x0 - 1) * xcs;
      PixVec<C> in[NPX + 2];
#pragma unroll
      for (int j = 0; j < NPX + 2; ++j) {
        const int px = x0 - 1 + j;
        if (px >= 0 && px < wd) in[j].load(rp + j * xcs); else in[j].zero();
      }
#pragma unroll
      for (int j = 0; j < NPX + 2; ++j) {     // input pixel j feeds output pixel p = j - kw for kw = 0..2
        float f[C];
        in[j].unpack(f);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int p = j - kw;
          if (p < 0 || p >= NPX) continue;
          const float4* wrow = s_w + ((kh * 3 + kw) * C) * (C / 4);
#pragma unroll
          for (int ci = 0; ci < C; ++ci) {
#pragma unroll
            for (int q = 0; q < C / 4; ++q) {
              const float4 w4 = wrow[ci * (C / 4) + q];
              acc[p][4 * q + 0] = fmaf(f[ci], w4.x, acc[p][4 * q + 0]);
              acc[p][4 * q + 1] = fmaf(f[ci], w4.y, acc[p][4 * q + 1]);
              acc[p][4 * q + 2] = fmaf(f[ci], w4.z, acc[p][4 * q + 2]);
              acc[p][4 * q + 3] = fmaf(f[ci], w4.w, acc[p][4 * q + 3]);
            }
          }
        }
      }
    }
    __nv_bfloat16* yp = y + (row * wd + x0) * ycs;
#pragma unroll
    for (int p = 0; p < NPX; ++p) {
      if (x0 + p >= wd) break;
#pragma unroll
      for (int q = 0; q < C / 8; ++q) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = act == MSPI_ACT_RELU ? fmaxf(acc[p][8 * q + e], 0.f) : acc[p][8 * q + e];
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        reinterpret_cast<uint4*>(yp + p * ycs)[q] = o;
      }
    }
  }
}

}  // namespace
}  // namespace mspi

using namespace mspi;

extern "C" int mspi_conv133_small(const void* x, int64_t x_cstride, const float* w_packed, const float* shift, void* y,
                                  int64_t y_cstride, int64_t planes, int h, int w, int c, int act, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && w_packed && y && planes > 0 && h > 0 && w > 0, "mspi_conv133_small: bad argument");
  MSPI_CHECK_ARG(c == 8 || c == 16, "mspi_conv133_small serves Cin = Cout = 8 or 16 (c = %d)", c);
  MSPI_CHECK_ARG(act == MSPI_ACT_NONE || act == MSPI_ACT_RELU, "mspi_conv133_small: activation %d", act);
  MSPI_CHECK_ARG(x_cstride % 8 == 0 && y_cstride % 8 == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w_packed)) & 15) == 0,
                 "mspi_conv133_small: pixels must be 16-byte aligned");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int npx = c == 8 ? 4 : 2;
  const int gpr = (w + npx - 1) / npx;
  const long long items = planes * h * gpr;
  long long blocks = (items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  if (c == 8)
    MSPI_CUDA(launch_pdl(conv133_small_kernel<8, 4>, static_cast<unsigned>(blocks), 256, 0, stream, xb, x_cstride, w_packed, shift, yb,
                         y_cstride, items, gpr, h, w, act));
  else
    MSPI_CUDA(launch_pdl(conv133_small_kernel<16, 2>, static_cast<unsigned>(blocks), 256, 0, stream, xb, x_cstride, w_packed, shift, yb,
                         y_cstride, items, gpr, h, w, act));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}
