// Saliency-map post-processing of the inference driver on the GPU (inference.py:65-91 of the reference):
//   cv2.GaussianBlur(log_map, (11,11), 0)  [sigma = 0.3*((11-1)*0.5-1)+0.8 = 2.0, BORDER_REFLECT_101]
//   -> exp -> cv2.resize(., (ow, oh)) [INTER_LINEAR, half-pixel centres, edge clamp] -> min-max normalise
//   -> round(255 * .) (round-half-to-even, np.round) -> uint8.
// Removes the device->host copy of the fp32 maps and the per-frame CPU cv2 stage.
#include "common.cuh"

namespace mspi {
namespace {

struct Gauss11 { float w[11]; };   // passed as a kernel argument: valid on whichever device the launch goes to

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// Separable 11-tap blur through a shared tile, then exp.  Block = 32x8 outputs.
__global__ void blur_exp_kernel(const float* __restrict__ x, float* __restrict__ y, int h, int w, const Gauss11 kG) {
  const float* kGauss11 = kG.w;
  constexpr int TW = 32, TH = 8, R = 5;
  __shared__ float tile[TH + 2 * R][TW + 2 * R];
  __shared__ float rows[TH + 2 * R][TW];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const float* src = x + static_cast<long long>(b) * h * w;
  for (int i = threadIdx.y * TW + threadIdx.x; i < (TH + 2 * R) * (TW + 2 * R); i += TW * TH) {
    const int ty = i / (TW + 2 * R), tx = i % (TW + 2 * R);
    tile[ty][tx] = src[static_cast<long long>(reflect101(y0 + ty - R, h)) * w + reflect101(x0 + tx - R, w)];
  }
  __syncthreads();
  for (int ty = threadIdx.y; ty < TH + 2 * R; ty += TH) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) s = fmaf(kGauss11[k], tile[ty][threadIdx.x + k], s);
    rows[ty][threadIdx.x] = s;
  }
  __syncthreads();
  const int ox = x0 + threadIdx.x, oy = y0 + threadIdx.y;
  if (ox < w && oy < h) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) s = fmaf(kGauss11[k], rows[threadIdx.y + k][threadIdx.x], s);
    y[(static_cast<long long>(b) * h + oy) * w + ox] = expf(s);
  }
}

// Bilinear resize (cv2 INTER_LINEAR) + per-image min / max (positive values: integer ordering of the bit patterns).
__global__ void resize_minmax_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned int* __restrict__ mm, int h,
                                     int w, int oh, int ow) {
  const int b = blockIdx.y;
  const float sy = static_cast<float>(h) / oh, sx = static_cast<float>(w) / ow;
  const float* src = x + static_cast<long long>(b) * h * w;
  float lo = INFINITY, hi = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < oh * ow; i += gridDim.x * blockDim.x) {
    const int oy = i / ow, ox = i % ow;
    float fy = (oy + 0.5f) * sy - 0.5f, fx = (ox + 0.5f) * sx - 0.5f;
    int iy = static_cast<int>(floorf(fy)), ix = static_cast<int>(floorf(fx));
    float wy = fy - iy, wx = fx - ix;
    if (iy < 0) { iy = 0; wy = 0.f; }
    if (ix < 0) { ix = 0; wx = 0.f; }
    if (iy >= h - 1) { iy = h - 1; wy = 0.f; }
    if (ix >= w - 1) { ix = w - 1; wx = 0.f; }
    const int iy1 = min(iy + 1, h - 1), ix1 = min(ix + 1, w - 1);
    const float v00 = src[iy * w + ix], v01 = src[iy * w + ix1], v10 = src[iy1 * w + ix], v11 = src[iy1 * w + ix1];
    const float v = (v00 * (1.f - wx) + v01 * wx) * (1.f - wy) + (v10 * (1.f - wx) + v11 * wx) * wy;
    y[static_cast<long long>(b) * oh * ow + i] = v;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 2 * b, __float_as_uint(lo));
    atomicMax(mm + 2 * b + 1, __float_as_uint(hi));
  }
}

__global__ void normalize_u8_kernel(const float* __restrict__ x, const unsigned int* __restrict__ mm, uint8_t* __restrict__ y,
                                    long long per_image, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per_image;
    const float lo = __uint_as_float(mm[2 * b]), hi = __uint_as_float(mm[2 * b + 1]);
    const float v = (x[i] - lo) / (hi - lo);
    y[i] = static_cast<uint8_t>(fminf(fmaxf(rintf(v * 255.f), 0.f), 255.f));
  }
}

}  // namespace
}  // namespace mspi

using namespace mspi;

extern "C" int mspi_postprocess_maps(const float* log_maps, uint8_t* out, float* work, int b, int h, int w, int oh, int ow,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(log_maps && out && work && b > 0 && h > 5 && w > 5 && oh > 0 && ow > 0, "mspi_postprocess_maps: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  // cv2.getGaussianKernel(11, sigma=2.0): exp(-(i-5)^2 / (2 sigma^2)), normalised to sum 1
  Gauss11 kG;
  {
    double g[11], s = 0.0;
    for (int i = 0; i < 11; ++i) { g[i] = exp(-(i - 5) * (i - 5) / 8.0); s += g[i]; }
    for (int i = 0; i < 11; ++i) kG.w[i] = static_cast<float>(g[i] / s);
  }
  // work: [b*h*w] blurred+exp | [b*oh*ow] resized | [2*b] min/max bit patterns
  float* blurred = work;
  float* resized = work + static_cast<long long>(b) * h * w;
  unsigned int* mm = reinterpret_cast<unsigned int*>(resized + static_cast<long long>(b) * oh * ow);
  {
    dim3 grid((w + 31) / 32, (h + 7) / 8, b), block(32, 8);
    blur_exp_kernel<<<grid, block, 0, stream>>>(log_maps, blurred, h, w, kG);
    MSPI_LAUNCH_CHECK();
  }
  MSPI_CUDA(cudaMemsetAsync(mm, 0xFF, sizeof(unsigned int) * 2 * b, stream));  // min <- 0xFFFFFFFF; max fixed below
  {
    // max starts at 0: positive floats order like unsigned ints
    for (int i = 0; i < b; ++i) MSPI_CUDA(cudaMemsetAsync(mm + 2 * i + 1, 0, sizeof(unsigned int), stream));
    dim3 grid(64, b);
    resize_minmax_kernel<<<grid, 256, 0, stream>>>(blurred, resized, mm, h, w, oh, ow);
    MSPI_LAUNCH_CHECK();
  }
  {
    const long long total = static_cast<long long>(b) * oh * ow;
    long long blocks = (total + 255) / 256;
    if (blocks > num_sms() * 8ll) blocks = num_sms() * 8ll;
    normalize_u8_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(resized, mm, out, static_cast<long long>(oh) * ow, total);
    MSPI_LAUNCH_CHECK();
  }
  return MSPI_OK;
}
