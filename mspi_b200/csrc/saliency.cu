// Saliency-map kernels: log-softmax over pixels, fused KLD/CC/SIM/NSS metrics + SalLoss, and the
// log power-spectrogram audio front end.  fp32 data, fp64 block accumulators where the reference's
// result depends on long sums (the maps have 86k pixels), warp-shuffle + shared-memory reductions.
#include "common.cuh"

namespace mspi {
namespace {

constexpr int kRedThreads = 1024;

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = lane < nw ? red[lane] : 0.0;
  return warp_sum(r);
}
__device__ __forceinline__ float block_max_f(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = lane < nw ? red[lane] : -INFINITY;
  return warp_max(r);
}
__device__ __forceinline__ float block_min_f(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = lane < nw ? red[lane] : INFINITY;
  return warp_min(r);
}

// ------------------------------------------------------------------------- log-softmax over pixels
__global__ void __launch_bounds__(kRedThreads) logsoftmax2d_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                  long long pixels) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float redf[32];
  __shared__ double redd[32];
  const float* xb = x + blockIdx.x * pixels;
  float* yb = y + blockIdx.x * pixels;
  float m = -INFINITY;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) m = fmaxf(m, __ldg(xb + i));
  m = block_max_f(m, redf);
  double s = 0.0;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) s += static_cast<double>(expf(__ldg(xb + i) - m));
  s = block_sum_d(s, redd);
  const float lse = m + static_cast<float>(log(s));
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) yb[i] = __ldg(xb + i) - lse;
}

// ------------------------------------------------------------------------- metrics
// One block per map.  Pass 1: sums, extrema.  Pass 2: everything that needs the pass-1 scalars.
// work[b*16 + {0:kld, 1:cc, 2:sim, 3:nss}]
__global__ void __launch_bounds__(kRedThreads) metrics_map_kernel(const float* __restrict__ pred, int pred_is_log,
                                                                 const float* __restrict__ gt,
                                                                 const float* __restrict__ fix,
                                                                 float* __restrict__ work, long long pixels) {
  __shared__ float redf[32];
  __shared__ double redd[32];
  const long long off = blockIdx.x * pixels;
  const float* sp = pred + off;
  const float* gp = gt + off;
  const float* fp = fix ? fix + off : nullptr;
  double sum_s = 0.0, sum_g = 0.0, sum_f = 0.0;
  float min_s = INFINITY, max_s = -INFINITY, min_g = INFINITY, max_g = -INFINITY;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    const float s = pred_is_log ? expf(__ldg(sp + i)) : __ldg(sp + i);
    const float g = __ldg(gp + i);
    sum_s += s; sum_g += g;
    min_s = fminf(min_s, s); max_s = fmaxf(max_s, s);
    min_g = fminf(min_g, g); max_g = fmaxf(max_g, g);
    if (fp) sum_f += __ldg(fp + i);
  }
  sum_s = block_sum_d(sum_s, redd);
  sum_g = block_sum_d(sum_g, redd);
  sum_f = block_sum_d(sum_f, redd);
  min_s = block_min_f(min_s, redf); max_s = block_max_f(max_s, redf);
  min_g = block_min_f(min_g, redf); max_g = block_max_f(max_g, redf);

  const float fsum_s = static_cast<float>(sum_s), fsum_g = static_cast<float>(sum_g);
  const float mean_s = static_cast<float>(sum_s / pixels), mean_g = static_cast<float>(sum_g / pixels);
  const float rng_s = max_s - min_s, rng_g = max_g - min_g;
  // sums of the min-max normalised maps (similarity re-normalises them to unit mass)
  const float nsum_s = static_cast<float>((sum_s - static_cast<double>(min_s) * pixels) / rng_s);
  const float nsum_g = static_cast<float>((sum_g - static_cast<double>(min_g) * pixels) / rng_g);
  const float eps = 2.2204e-16f;

  double kld = 0.0, sim = 0.0, ab = 0.0, aa = 0.0, bb = 0.0, sf = 0.0;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    const float s = pred_is_log ? expf(__ldg(sp + i)) : __ldg(sp + i);
    const float g = __ldg(gp + i);
    const float sn = s / fsum_s, gn = g / fsum_g;
    kld += static_cast<double>(gn * logf(eps + gn / (sn + eps)));
    const float s1 = ((s - min_s) / rng_s) / nsum_s, g1 = ((g - min_g) / rng_g) / nsum_g;
    sim += static_cast<double>(fminf(s1, g1));
    const float ds = s - mean_s, dg = g - mean_g;
    ab += static_cast<double>(ds * dg);
    aa += static_cast<double>(ds * ds);
    bb += static_cast<double>(dg * dg);
    if (fp) sf += static_cast<double>(ds * __ldg(fp + i));
  }
  kld = block_sum_d(kld, redd);
  sim = block_sum_d(sim, redd);
  ab = block_sum_d(ab, redd);
  aa = block_sum_d(aa, redd);
  bb = block_sum_d(bb, redd);
  sf = block_sum_d(sf, redd);
  if (threadIdx.x == 0) {
    float* w = work + blockIdx.x * 16;
    w[0] = static_cast<float>(kld);
    w[1] = static_cast<float>(ab / sqrt(aa * bb));
    w[2] = static_cast<float>(sim);
    float nss = 0.f;
    if (fp) {
      const float std_s = static_cast<float>(sqrt(aa / static_cast<double>(pixels - 1)));  // unbiased, torch.std
      nss = static_cast<float>(sf / (static_cast<double>(std_s) + 2.2204e-16) / sum_f);
    }
    w[3] = nss;
  }
}

__global__ void metrics_finalize_kernel(const float* __restrict__ work, float* __restrict__ out, int b, int has_fix) {
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int i = 0; i < b; ++i) s += work[i * 16 + threadIdx.x];
    out[threadIdx.x] = s / static_cast<float>(b);
  }
  __syncwarp();
  if (threadIdx.x == 0) out[4] = out[0] - out[1] - (has_fix ? 0.1f * out[3] : 0.f);
}

// ------------------------------------------------------------------------- log power spectrogram
// Block = (frame, batch).  512-point real DFT by direct summation against a shared twiddle table:
// 257 bins x 512 samples per frame is ~0.26 MFLOP, far below launch cost; what matters is one pass
// over the wave and fp32 results that track torch.stft.
constexpr int kNfft = 512, kHop = 160, kBins = 257, kSpecThreads = 288;

__global__ void __launch_bounds__(kSpecThreads) logspec_kernel(const float* __restrict__ wave, float* __restrict__ out,
                                                              int n, int frames, int frames_out) {
  __shared__ float xs[kNfft];
  __shared__ float cs[kNfft];
  __shared__ float sn[kNfft];
  __shared__ float red[32];
  const int f = blockIdx.x, b = blockIdx.y;
  float* ob = out + static_cast<long long>(b) * kBins * frames_out;
  if (f >= frames) {
    for (int k = threadIdx.x; k < kBins; k += blockDim.x) ob[static_cast<long long>(k) * frames_out + f] = 0.02f;
    return;
  }
  const float* wb = wave + static_cast<long long>(b) * n;
  for (int i = threadIdx.x; i < kNfft; i += blockDim.x) {
    int idx = f * kHop - kNfft / 2 + i;  // center=True
    if (idx < 0) idx = -idx;             // reflect padding (no edge repeat)
    if (idx >= n) idx = 2 * (n - 1) - idx;
    const float ang = 6.283185307179586f * static_cast<float>(i) / kNfft;
    float s, c;
    sincosf(ang, &s, &c);
    cs[i] = c;
    sn[i] = s;
    xs[i] = __ldg(wb + idx) * (0.5f - 0.5f * c);  // periodic Hann window
  }
  __syncthreads();
  const int k = threadIdx.x;
  float val = 0.f;
  if (k < kBins) {
    float re = 0.f, im = 0.f;
    int ph = 0;
#pragma unroll 8
    for (int i = 0; i < kNfft; ++i) {
      re = fmaf(xs[i], cs[ph], re);
      im = fmaf(xs[i], sn[ph], im);
      ph = (ph + k) & (kNfft - 1);
    }
    val = logf(re * re + im * im + 1e-6f);
  }
  const float mean = block_sum(k < kBins ? val : 0.f, red) / kBins;
  const float dv = k < kBins ? val - mean : 0.f;
  const float var = block_sum(dv * dv, red) / (kBins - 1);  // unbiased
  if (k < kBins) ob[static_cast<long long>(k) * frames_out + f] = dv / (sqrtf(var) + 1e-6f);
}

}  // namespace
}  // namespace mspi

using namespace mspi;

extern "C" int mspi_logsoftmax2d(const float* x, float* y, int b, int64_t pixels, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && y && b > 0 && pixels > 0, "mspi_logsoftmax2d: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  MSPI_CUDA(launch_pdl(logsoftmax2d_kernel, b, kRedThreads, 0, stream, x, y, pixels));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_saliency_metrics(const float* pred, int pred_is_log, const float* gt, const float* fix, float* out,
                                     float* work, int b, int64_t pixels, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(pred && gt && out && work && b > 0 && pixels > 1, "mspi_saliency_metrics: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  metrics_map_kernel<<<b, kRedThreads, 0, stream>>>(pred, pred_is_log, gt, fix, work, pixels);
  MSPI_LAUNCH_CHECK();
  metrics_finalize_kernel<<<1, 32, 0, stream>>>(work, out, b, fix != nullptr);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_logspec(const float* wave, float* out, int b, int n, int frames_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(wave && out && b > 0 && n > kNfft / 2 && frames_out > 0, "mspi_logspec: bad argument (n must exceed 256)");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int frames = n / kHop + 1;
  dim3 grid(frames_out, b);
  logspec_kernel<<<grid, kSpecThreads, 0, stream>>>(wave, out, n, frames, frames_out);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}
