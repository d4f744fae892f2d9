// Training-step kernels (row a20: engine_train.py:27-76 — model.train() + frozen_encoder(), loss.backward(), AdamW).
//
// Everything here is fp32, channels-last ([pixels][C] views with an element stride between pixels, so channel slices of
// concatenated buffers work in place).  The GEMM-shaped parts of the backward pass (data and weight gradients of every
// Conv3d / Linear) run on the tensor cores through mspi_conv_gemm / mspi_conv_wgrad; this file holds what surrounds them:
// batch-statistics BatchNorm (forward + backward), activation / bias gradients, max-pool and bilinear-upsample adjoints,
// LayerNorm / softmax / gating / depthwise-conv gradients, the saliency-loss and SimSiam gradients, a small strided
// batched GEMM for the attention backward, the weight (re)packing permute and the flat AdamW update.
// All are HBM-bound: coalesced 128-byte channel rows or float4 vectors, per-block partial sums, one atomic per block and
// channel.  Gradient outputs take an `accumulate` flag (0: overwrite, 1: add) unless they are scatter-adds.
#include <math.h>

#include "common.cuh"

namespace mspi {
namespace {

constexpr int kBlock = 256;

inline int grid_for(long long total, int block = kBlock) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// pixels handled by one block of the (32 channels x 8 pixel lanes) reduction kernels
inline int pixels_per_block(long long pixels, int cblocks) {
  long long want = static_cast<long long>(num_sms()) * 8 / (cblocks > 0 ? cblocks : 1);
  if (want < 1) want = 1;
  long long ppb = (pixels + want - 1) / want;
  if (ppb < 64) ppb = 64;
  if (ppb > 8192) ppb = 8192;
  return static_cast<int>(ppb);
}

__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void stf4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void red4(float* p, float4 v) {
  atomicAdd(reinterpret_cast<float4*>(p), v);  // sm_90+: one 16-byte reduction
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

// ------------------------------------------------------------------------- BatchNorm, batch statistics
// block (32, 8): 32 consecutive channels x 8 pixel lanes; double partial sums, one double atomic per block and channel
__global__ void bn_stats_kernel(const float* __restrict__ x, long long cs, long long pixels, int c, int ppb,
                                double* __restrict__ work) {
  const int ch = blockIdx.y * 32 + threadIdx.x;
  const long long p0 = static_cast<long long>(blockIdx.x) * ppb;
  const long long p1 = min(p0 + ppb, pixels);
  double s = 0.0, q = 0.0;
  if (ch < c)
    for (long long p = p0 + threadIdx.y; p < p1; p += 8) {
      const double v = x[p * cs + ch];
      s += v;
      q += v * v;
    }
  __shared__ double sh[2][8][32];
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    double S = 0.0, Q = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      S += sh[0][j][threadIdx.x];
      Q += sh[1][j][threadIdx.x];
    }
    atomicAdd(work + ch, S);
    atomicAdd(work + c + ch, Q);
  }
}

// Normalise + affine + ReLU.  Every block first derives (scale, shift) of all channels from the batch sums in `work`
// (mean, biased variance) into shared memory; block 0 also stores the saved statistics for the backward pass and moves
// the running buffers (torch.nn.BatchNorm: running_var takes the unbiased variance).  `work` is only read here: the
// caller clears it before the next step.
__global__ void bn_apply_kernel(const float* __restrict__ x, long long xcs, float* __restrict__ y, long long ycs,
                                long long total, int c4, int c, const double* __restrict__ work, double count,
                                const float* __restrict__ w, const float* __restrict__ b, float eps, float momentum,
                                float* __restrict__ rmean, float* __restrict__ rvar, long long* __restrict__ tracked,
                                float* __restrict__ save_mean, float* __restrict__ save_invstd, int relu) {
  extern __shared__ float sm[];  // scale[c], shift[c]
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const double mean = work[ch] / count;
    double var = work[c + ch] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double sc = static_cast<double>(w[ch]) * invstd;
    sm[ch] = static_cast<float>(sc);
    sm[c + ch] = static_cast<float>(static_cast<double>(b[ch]) - mean * sc);
    if (blockIdx.x == 0) {
      save_mean[ch] = static_cast<float>(mean);
      save_invstd[ch] = static_cast<float>(invstd);
      if (rmean) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        rmean[ch] = static_cast<float>((1.0 - momentum) * rmean[ch] + momentum * mean);
        rvar[ch] = static_cast<float>((1.0 - momentum) * rvar[ch] + momentum * unb);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && tracked) *tracked += 1;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i;
    const int cc = divmod(pix, c4) * 4;
    float4 v = ldf4(x + pix * xcs + cc);
    v.x = fmaf(v.x, sm[cc], sm[c + cc]);
    v.y = fmaf(v.y, sm[cc + 1], sm[c + cc + 1]);
    v.z = fmaf(v.z, sm[cc + 2], sm[c + cc + 2]);
    v.w = fmaf(v.w, sm[cc + 3], sm[c + cc + 3]);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    stf4(y + pix * ycs + cc, v);
  }
}

// sums over pixels of g and g * xhat, g = dy masked by the ReLU that followed the BatchNorm (y > 0)
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ x, long long xcs, const float* __restrict__ y, long long ycs,
                                     const float* __restrict__ dy, long long dcs, long long pixels, int c, int ppb,
                                     const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                     double* __restrict__ work) {
  const int ch = blockIdx.y * 32 + threadIdx.x;
  const long long p0 = static_cast<long long>(blockIdx.x) * ppb;
  const long long p1 = min(p0 + ppb, pixels);
  double s = 0.0, q = 0.0;
  if (ch < c) {
    const float m = mean[ch], is = invstd[ch];
    for (long long p = p0 + threadIdx.y; p < p1; p += 8) {
      float g = dy[p * dcs + ch];
      if (relu && !(y[p * ycs + ch] > 0.f)) g = 0.f;
      s += g;
      q += static_cast<double>(g * ((x[p * xcs + ch] - m) * is));
    }
  }
  __shared__ double sh[2][8][32];
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    double S = 0.0, Q = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      S += sh[0][j][threadIdx.x];
      Q += sh[1][j][threadIdx.x];
    }
    atomicAdd(work + ch, S);
    atomicAdd(work + c + ch, Q);
  }
}

// dx (+)= w*invstd * (g - mean(g) - xhat * mean(g*xhat)).  Every block derives the three per-channel coefficients from the
// sums in `work` into shared memory; block 0 adds the parameter gradients (dweight += sum g*xhat, dbias += sum g).
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, long long xcs, const float* __restrict__ y, long long ycs,
                                    const float* __restrict__ dy, long long dcs, float* __restrict__ dx, long long dxcs,
                                    long long total, int c4, int c, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ w,
                                    const double* __restrict__ work, double count, float* __restrict__ dweight,
                                    float* __restrict__ dbias, int relu, int acc) {
  extern __shared__ float sm[];  // k[c], a[c], b[c], mean[c], invstd[c]
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const double S = work[ch], Q = work[c + ch];
    sm[ch] = w[ch] * invstd[ch];
    sm[c + ch] = static_cast<float>(S / count);
    sm[2 * c + ch] = static_cast<float>(Q / count);
    sm[3 * c + ch] = mean[ch];
    sm[4 * c + ch] = invstd[ch];
    if (blockIdx.x == 0) {
      dweight[ch] += static_cast<float>(Q);
      dbias[ch] += static_cast<float>(S);
    }
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i;
    const int cc = divmod(pix, c4) * 4;
    const float4 xv = ldf4(x + pix * xcs + cc);
    float4 g = ldf4(dy + pix * dcs + cc);
    if (relu) {
      const float4 yv = ldf4(y + pix * ycs + cc);
      if (!(yv.x > 0.f)) g.x = 0.f;
      if (!(yv.y > 0.f)) g.y = 0.f;
      if (!(yv.z > 0.f)) g.z = 0.f;
      if (!(yv.w > 0.f)) g.w = 0.f;
    }
    const float xe[4] = {xv.x, xv.y, xv.z, xv.w}, ge[4] = {g.x, g.y, g.z, g.w};
    float oe[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = cc + j;
      oe[j] = sm[ch] * (ge[j] - sm[c + ch] - (xe[j] - sm[3 * c + ch]) * sm[4 * c + ch] * sm[2 * c + ch]);
    }
    float4 o = make_float4(oe[0], oe[1], oe[2], oe[3]);
    float* dp = dx + pix * dxcs + cc;
    if (acc) {
      const float4 old = ldf4(dp);
      o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
    }
    stf4(dp, o);
  }
}

// ------------------------------------------------------------------------- activations and bias gradients
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = ldf4(x + i * 4);
    if (act == MSPI_ACT_GELU) {
      v.x = gelu_exact(v.x); v.y = gelu_exact(v.y); v.z = gelu_exact(v.z); v.w = gelu_exact(v.w);
    } else if (act == MSPI_ACT_RELU) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    stf4(y + i * 4, v);
  }
}

// g = dy * act'(ref) (ref = the activation's OUTPUT for ReLU, its INPUT for GELU); dz = g (optional, may alias dy);
// dbias[ch] += sum over pixels of g (optional)
__global__ void act_bwd_kernel(const float* dy, long long dcs, const float* __restrict__ ref, long long rcs, float* dz,
                               long long zcs, long long pixels, int c, int ppb, int act, float* __restrict__ dbias) {
  const int ch = blockIdx.y * 32 + threadIdx.x;
  const long long p0 = static_cast<long long>(blockIdx.x) * ppb;
  const long long p1 = min(p0 + ppb, pixels);
  float s = 0.f;
  if (ch < c)
    for (long long p = p0 + threadIdx.y; p < p1; p += 8) {
      float g = dy[p * dcs + ch];
      if (act == MSPI_ACT_RELU) {
        if (!(ref[p * rcs + ch] > 0.f)) g = 0.f;
      } else if (act == MSPI_ACT_GELU) {
        g *= gelu_grad(ref[p * rcs + ch]);
      }
      if (dz) dz[p * zcs + ch] = g;
      s += g;
    }
  if (!dbias) return;
  __shared__ float sh[8][32];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float S = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) S += sh[j][threadIdx.x];
    atomicAdd(dbias + ch, S);
  }
}

// ------------------------------------------------------------------------- max pool (fp32) and its adjoint
__global__ void maxpool3d_f32_kernel(MspiPoolDesc d, const float* __restrict__ x, float* __restrict__ y, long long total,
                                     int c4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int cc = divmod(r, c4) * 4;
    const long long opix = r;
    const int ow = divmod(r, d.ow);
    const int oh = divmod(r, d.oh);
    const int ot = divmod(r, d.ot);
    const int n = static_cast<int>(r);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const int t0 = ot * d.st - d.pt, h0 = oh * d.sh - d.ph, w0 = ow * d.sw - d.pw;
    for (int kt = 0; kt < d.kt; ++kt) {
      const int it = t0 + kt;
      if (it < 0 || it >= d.t) continue;
      for (int kh = 0; kh < d.kh; ++kh) {
        const int ih = h0 + kh;
        if (ih < 0 || ih >= d.h) continue;
        for (int kw = 0; kw < d.kw; ++kw) {
          const int iw = w0 + kw;
          if (iw < 0 || iw >= d.w) continue;
          const long long pix = ((static_cast<long long>(n) * d.t + it) * d.h + ih) * d.w + iw;
          const float4 v = ldf4(x + pix * d.in_cstride + cc);
          m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
      }
    }
    stf4(y + opix * d.out_cstride + cc, m);
  }
}

// scatter: every output element sends its gradient to the FIRST maximal input of its window in (t, h, w) scan order
// (the index torch's max_pool3d forward records).  dx must hold the running sum (zero or other consumers' gradients).
__global__ void maxpool3d_bwd_kernel(MspiPoolDesc d, const float* __restrict__ x, const float* __restrict__ dy,
                                     long long dcs, float* __restrict__ dx, long long dxcs, long long total, int c4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int cc = divmod(r, c4) * 4;
    const long long opix = r;
    const int ow = divmod(r, d.ow);
    const int oh = divmod(r, d.oh);
    const int ot = divmod(r, d.ot);
    const int n = static_cast<int>(r);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    long long arg[4] = {-1, -1, -1, -1};
    const int t0 = ot * d.st - d.pt, h0 = oh * d.sh - d.ph, w0 = ow * d.sw - d.pw;
    for (int kt = 0; kt < d.kt; ++kt) {
      const int it = t0 + kt;
      if (it < 0 || it >= d.t) continue;
      for (int kh = 0; kh < d.kh; ++kh) {
        const int ih = h0 + kh;
        if (ih < 0 || ih >= d.h) continue;
        for (int kw = 0; kw < d.kw; ++kw) {
          const int iw = w0 + kw;
          if (iw < 0 || iw >= d.w) continue;
          const long long pix = ((static_cast<long long>(n) * d.t + it) * d.h + ih) * d.w + iw;
          const float4 v = ldf4(x + pix * d.in_cstride + cc);
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (e[j] > m[j] || arg[j] < 0) { m[j] = e[j]; arg[j] = pix; }
        }
      }
    }
    const float4 g = ldf4(dy + opix * dcs + cc);
    const float ge[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (arg[j] >= 0 && ge[j] != 0.f) atomicAdd(dx + arg[j] * dxcs + cc + j, ge[j]);
  }
}

// ------------------------------------------------------------------------- bilinear upsample adjoint (scatter)
__global__ void upsample_bwd_kernel(MspiUpDesc d, const float* __restrict__ dy, const float* __restrict__ y,
                                    float* __restrict__ dx, long long total, int c4) {
  const int oh_ = d.h * d.k, ow_ = d.w * d.k;
  const float inv = 1.f / static_cast<float>(d.k);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int cc = divmod(r, c4) * 4;
    const long long opix = r;
    const int ox = divmod(r, ow_);
    const int oy = divmod(r, oh_);
    const long long plane = r;
    float4 g = ldf4(dy + opix * d.out_cstride + cc);
    if (d.act == MSPI_ACT_RELU) {
      const float4 yv = ldf4(y + opix * d.out_cstride + cc);
      if (!(yv.x > 0.f)) g.x = 0.f;
      if (!(yv.y > 0.f)) g.y = 0.f;
      if (!(yv.z > 0.f)) g.z = 0.f;
      if (!(yv.w > 0.f)) g.w = 0.f;
    }
    const float sy = fmaxf((oy + 0.5f) * inv - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * inv - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = min(y0 + 1, d.h - 1), x1 = min(x0 + 1, d.w - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float wgt[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
    const long long base = plane * d.h * d.w;
    const long long idx[4] = {base + static_cast<long long>(y0) * d.w + x0, base + static_cast<long long>(y0) * d.w + x1,
                              base + static_cast<long long>(y1) * d.w + x0, base + static_cast<long long>(y1) * d.w + x1};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (wgt[j] == 0.f) continue;
      red4(dx + idx[j] * d.in_cstride + cc, make_float4(wgt[j] * g.x, wgt[j] * g.y, wgt[j] * g.z, wgt[j] * g.w));
    }
  }
}

// ------------------------------------------------------------------------- LayerNorm backward (rows)
// one warp per row; dw / db partial sums per block in shared memory, one atomic per block and channel
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, long long xrs, const float* __restrict__ dy, long long drs,
                                     long long rows_per_group, long long dgs, const float* __restrict__ y_relu,
                                     const float* __restrict__ w, float eps, float* __restrict__ dx, long long dxrs,
                                     long long rows, int c, int acc, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float sm[];  // [2][c]
  float* sdw = sm;
  float* sdb = sm + c;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (long long row = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * warps) {
    const float* xr = x + row * xrs;
    const long long g = row / rows_per_group, within = row - g * rows_per_group;
    const long long doff = g * dgs + within * drs;
    const float* dr = dy + doff;
    const float* yr = y_relu ? y_relu + doff : nullptr;
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += xr[i];
    const float mean = warp_sum(s) / c;
    float q = 0.f;
    for (int i = lane; i < c; i += 32) {
      const float a = xr[i] - mean;
      q += a * a;
    }
    const float rstd = rsqrtf(warp_sum(q) / c + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < c; i += 32) {
      float gv = dr[i];
      if (yr && !(yr[i] > 0.f)) gv = 0.f;
      const float xh = (xr[i] - mean) * rstd;
      const float gw = gv * __ldg(w + i);
      s1 += gw;
      s2 += gw * xh;
      atomicAdd(sdw + i, gv * xh);
      atomicAdd(sdb + i, gv);
    }
    s1 = warp_sum(s1) / c;
    s2 = warp_sum(s2) / c;
    float* dxr = dx + row * dxrs;
    for (int i = lane; i < c; i += 32) {
      float gv = dr[i];
      if (yr && !(yr[i] > 0.f)) gv = 0.f;
      const float xh = (xr[i] - mean) * rstd;
      float o = rstd * (gv * __ldg(w + i) - s1 - xh * s2);
      if (acc) o += dxr[i];
      dxr[i] = o;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    if (sdw[i] != 0.f) atomicAdd(dw + i, sdw[i]);
    if (sdb[i] != 0.f) atomicAdd(db + i, sdb[i]);
  }
}

// ------------------------------------------------------------------------- softmax backward (rows, in place on dP)
// dS = scale * P * (dP - sum_k dP*P)
__global__ void softmax_bwd_rows_kernel(const float* __restrict__ p, float* __restrict__ dp, long long rows, int n,
                                        long long stride, float scale) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (long long row = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * warps) {
    const float* pr = p + row * stride;
    float* dr = dp + row * stride;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += pr[i] * dr[i];
    s = warp_sum(s);
    for (int i = lane; i < n; i += 32) dr[i] = scale * pr[i] * (dr[i] - s);
  }
}

// ------------------------------------------------------------------------- small strided batched GEMM (fp32 FMA)
// C[b1][b0][i][j] (+)= alpha * sum_k A[b1][b0][i][k] * B[b1][b0][k][j], every stride free.  Serves the attention backward
// (372-token sequences, 4 heads: ~1 GFLOP per block) where each of the four products reads a differently transposed view.
struct SgemmArgs {
  int m, n, k, b0, b1, acc;
  float alpha;
  long long a_i, a_k, a_b0, a_b1;
  long long b_k, b_j, b_b0, b_b1;
  long long c_i, c_j, c_b0, c_b1;
};
__global__ void __launch_bounds__(256) sgemm_strided_kernel(SgemmArgs g, const float* __restrict__ A,
                                                            const float* __restrict__ B, float* __restrict__ C) {
  __shared__ float As[32][33], Bs[32][33];
  const int batch = blockIdx.z;
  const int b0 = batch % g.b0, b1 = batch / g.b0;
  A += b0 * g.a_b0 + b1 * g.a_b1;
  B += b0 * g.b_b0 + b1 * g.b_b1;
  C += b0 * g.c_b0 + b1 * g.c_b1;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < g.k; k0 += 32) {
    // As[i][k], Bs[k][j]; the lane index runs along whichever axis is contiguous in memory
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int u = ty + 8 * r;
      if (g.a_k == 1) {
        const int i = i0 + u, k = k0 + tx;
        As[u][tx] = (i < g.m && k < g.k) ? A[i * g.a_i + k] : 0.f;
      } else {
        const int i = i0 + tx, k = k0 + u;
        As[tx][u] = (i < g.m && k < g.k) ? A[i * g.a_i + k * g.a_k] : 0.f;
      }
      if (g.b_j == 1) {
        const int k = k0 + u, j = j0 + tx;
        Bs[u][tx] = (k < g.k && j < g.n) ? B[k * g.b_k + j] : 0.f;
      } else {
        const int k = k0 + tx, j = j0 + u;
        Bs[tx][u] = (k < g.k && j < g.n) ? B[k * g.b_k + j * g.b_j] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float bv = Bs[kk][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[ty + 8 * r][kk], bv, acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r, j = j0 + tx;
    if (i < g.m && j < g.n) {
      float* cp = C + i * g.c_i + j * g.c_j;
      const float v = g.alpha * acc[r];
      *cp = g.acc ? *cp + v : v;
    }
  }
}

// ------------------------------------------------------------------------- SA gate backward
// y = x * (1 + sigmoid(l)):  dx (+)= dy * (1 + s);  dl = s (1 - s) * sum_c dy * x.   one warp per pixel
__global__ void sa_gate_bwd_kernel(const float* __restrict__ x, long long xcs, const float* __restrict__ logit,
                                   const float* __restrict__ dy, long long dcs, float* __restrict__ dx, long long dxcs,
                                   float* __restrict__ dlogit, long long pixels, int c, int acc) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (long long p = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); p < pixels;
       p += static_cast<long long>(gridDim.x) * warps) {
    const float s = 1.f / (1.f + expf(-logit[p]));
    float dot = 0.f;
    for (int i = lane * 4; i < c; i += 128) {
      const float4 g = ldf4(dy + p * dcs + i), xv = ldf4(x + p * xcs + i);
      dot += g.x * xv.x + g.y * xv.y + g.z * xv.z + g.w * xv.w;
      float4 o = make_float4(g.x * (1.f + s), g.y * (1.f + s), g.z * (1.f + s), g.w * (1.f + s));
      float* dp = dx + p * dxcs + i;
      if (acc) {
        const float4 old = ldf4(dp);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      stf4(dp, o);
    }
    dot = warp_sum(dot);
    if (lane == 0) dlogit[p] = dot * s * (1.f - s);
  }
}

// ------------------------------------------------------------------------- Conv (1,3,3) 32 -> 1 backward
// (SA.conv_mask.2, model_utils.py:163; readout.12, model_utils.py:503).  x [planes][H][W][32] (pixel stride xcs),
// dy [planes][H][W] one channel.  Per pixel q and input channel ci (= lane):
//   dx[q][ci] (+)= sum_tap dy[q - off(tap)] W[ci][tap];   dW[ci][tap] += dy[q - off(tap)] x[q][ci];   db += dy[q]
__global__ void conv_c1_bwd_kernel(const float* __restrict__ x, long long xcs, const float* __restrict__ dy,
                                   const float* __restrict__ w, float* __restrict__ dx, long long dxcs, float* __restrict__ dw,
                                   float* __restrict__ db, long long pixels, int h, int wd, int ppb, int acc) {
  const int ci = threadIdx.x;
  float wr[9], aw[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    wr[t] = w[ci * 9 + t];
    aw[t] = 0.f;
  }
  float ab = 0.f;
  const long long p0 = static_cast<long long>(blockIdx.x) * ppb;
  const long long p1 = min(p0 + ppb, pixels);
  for (long long q = p0 + threadIdx.y; q < p1; q += 8) {
    long long qq = q;
    const int qx = divmod(qq, wd);
    const int qy = divmod(qq, h);
    const float xv = x[q * xcs + ci];
    float o = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int sy = qy - (kh - 1), sx = qx - (kw - 1);  // output position that read q through tap (kh, kw)
        if (sy < 0 || sy >= h || sx < 0 || sx >= wd) continue;
        const float g = __ldg(dy + q - static_cast<long long>(kh - 1) * wd - (kw - 1));
        o = fmaf(g, wr[kh * 3 + kw], o);
        aw[kh * 3 + kw] = fmaf(g, xv, aw[kh * 3 + kw]);
      }
    if (dx) {
      float* dp = dx + q * dxcs + ci;
      *dp = acc ? *dp + o : o;
    }
    if (ci == 0) ab += dy[q];
  }
  __shared__ float sh[8][32];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
    sh[threadIdx.y][ci] = aw[t];
    __syncthreads();
    if (threadIdx.y == 0) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += sh[j][ci];
      atomicAdd(dw + ci * 9 + t, s);
    }
  }
  __syncthreads();
  sh[threadIdx.y][ci] = ab;
  __syncthreads();
  if (threadIdx.y == 0 && ci == 0 && db) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][0];
    atomicAdd(db, s);
  }
}

// Forward of the same layer: y[q] = bias + sum_tap sum_ci x[q + off(tap)][ci] W[ci][tap].  One output channel cannot use a
// tensor-core tile (N = 1 of 16) and the implicit GEMM re-fetches the input once per tap; here each input row (32 channels =
// one 128-byte / 64-byte line per pixel) is read once per output row-neighbourhood from L1/L2 and reduced with shuffles.
// A warp produces 4 consecutive pixels of a row: lane = input channel.
template <typename TI>
__global__ void conv_c1_fwd_kernel(const TI* __restrict__ x, long long xcs, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ y, long long planes, int h, int wd,
                                   int strip) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  // one warp = a 4-pixel wide, `strip`-row tall column of one plane, walked top to bottom with the three input rows under
  // the current output row in registers (lane = channel): every input pixel is loaded 1.5 times instead of 9
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  float wr[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wr[t] = w[lane * 9 + t];
  const float b0 = bias ? bias[0] : 0.f;
  const int gpr = (wd + 3) / 4;                      // pixel groups per image row
  const int spp = (h + strip - 1) / strip;           // strips per plane
  const long long items = planes * spp * gpr;
  for (long long it = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); it < items;
       it += static_cast<long long>(gridDim.x) * warps) {
    long long r = it;
    const int x0 = divmod(r, gpr) * 4;
    const int y0 = divmod(r, spp) * strip;
    const long long plane = r;
    const int y1 = min(y0 + strip, h);
    const TI* xp = x + plane * h * wd * xcs + lane;
    float row[3][6];
    auto load_row = [&](int yy, float (&dst)[6]) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const int sx = x0 + c - 1;
        dst[c] = (yy >= 0 && yy < h && sx >= 0 && sx < wd)
                     ? static_cast<float>(xp[(static_cast<long long>(yy) * wd + sx) * xcs]) : 0.f;
      }
    };
    load_row(y0 - 1, row[0]);
    load_row(y0, row[1]);
    for (int yy = y0; yy < y1; ++yy) {
      load_row(yy + 1, row[2]);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) acc[j] = fmaf(row[kh][j + kw], wr[kh * 3 + kw], acc[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
      if (lane < 4 && x0 + lane < wd) y[(plane * h + yy) * wd + x0 + lane] = acc[lane] + b0;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        row[0][c] = row[1][c];
        row[1][c] = row[2][c];
      }
    }
  }
}

// ------------------------------------------------------------------------- depthwise conv weight / bias gradient
// dw[ch][tap] += sum_p dy[p][ch] x[p + off(tap)][ch];  db[ch] += sum_p dy[p][ch].   block (32 channels, 8 pixel lanes)
template <int KT, int KH, int KW>
__global__ void dw_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int n, int t, int h, int w, int c,
                                int ppb, float* __restrict__ dw, float* __restrict__ db) {
  constexpr int TAPS = KT * KH * KW;
  const int ch = blockIdx.y * 32 + threadIdx.x;
  const long long pixels = static_cast<long long>(n) * t * h * w;
  const long long p0 = static_cast<long long>(blockIdx.x) * ppb;
  const long long p1 = min(p0 + ppb, pixels);
  float a[TAPS];
#pragma unroll
  for (int i = 0; i < TAPS; ++i) a[i] = 0.f;
  float ab = 0.f;
  if (ch < c)
    for (long long p = p0 + threadIdx.y; p < p1; p += 8) {
      long long pp = p;
      const int pw = divmod(pp, w);
      const int ph = divmod(pp, h);
      const int pt = divmod(pp, t);
      const float g = dy[p * c + ch];
      ab += g;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const int it = pt + kt - KT / 2;
        if (it < 0 || it >= t) continue;
#pragma unroll
        for (int kh = 0; kh < KH; ++kh) {
          const int ih = ph + kh - KH / 2;
          if (ih < 0 || ih >= h) continue;
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) {
            const int iw = pw + kw - KW / 2;
            if (iw < 0 || iw >= w) continue;
            const long long q = p + (static_cast<long long>(kt - KT / 2) * h + (kh - KH / 2)) * w + (kw - KW / 2);
            a[(kt * KH + kh) * KW + kw] = fmaf(g, x[q * c + ch], a[(kt * KH + kh) * KW + kw]);
          }
        }
      }
    }
  __shared__ float sh[8][32];
#pragma unroll
  for (int i = 0; i <= TAPS; ++i) {
    __syncthreads();
    sh[threadIdx.y][threadIdx.x] = i < TAPS ? a[i < TAPS ? i : 0] : ab;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
      if (i < TAPS) atomicAdd(dw + static_cast<long long>(ch) * TAPS + i, s);
      else atomicAdd(db + ch, s);
    }
  }
}

// ------------------------------------------------------------------------- saliency loss backward
// loss = mean_b [ KL(gt || P) - CC(P, gt) ],  P = exp(logp), logp = logits - logsumexp(logits)
// (utils/loss.py:26-49 on utils/compute_saliency_metrics.py:9-31,75-92).  One block per sample.
//   dlogits = scale/B * ( h - P * sum(h) ),  h_j = P_j * dLoss_b/dP_j
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : 0.0;
  r = warp_sum(r);
  return r;
}

__global__ void __launch_bounds__(1024) salloss_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ gt,
                                                           float* __restrict__ dlogits, float* __restrict__ per_sample,
                                                           long long pixels, int bsz, float scale) {
  __shared__ double red[32];
  const double EPS = 2.2204e-16;
  const int b = blockIdx.x;
  const float* lp = logp + b * pixels;
  const float* g = gt + b * pixels;
  float* dl = dlogits + b * pixels;
  double sp = 0.0, sg = 0.0;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    sp += exp(static_cast<double>(lp[i]));
    sg += g[i];
  }
  sp = block_sum_d(sp, red);
  sg = block_sum_d(sg, red);
  const double mp = sp / pixels, mg = sg / pixels;
  double spp = 0.0, sgg = 0.0, spg = 0.0, kl = 0.0, sas = 0.0;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    const double P = exp(static_cast<double>(lp[i])), G = g[i];
    const double pc = P - mp, gc = G - mg;
    spp += pc * pc;
    sgg += gc * gc;
    spg += pc * gc;
    const double s = P / sp, gn = G / sg;
    const double ratio = gn / (s + EPS);
    kl += gn * log(EPS + ratio);
    // a = dKL/ds = gn * 1/(EPS + ratio) * (-gn / (s+EPS)^2)
    const double a = -gn * ratio / ((EPS + ratio) * (s + EPS));
    sas += a * s;
  }
  spp = block_sum_d(spp, red);
  sgg = block_sum_d(sgg, red);
  spg = block_sum_d(spg, red);
  kl = block_sum_d(kl, red);
  sas = block_sum_d(sas, red);
  const double D = sqrt(spp * sgg);
  const double r = spg / D;
  // h_j = P_j * [ (a_j - sas)/sp - (gc_j / D - r * pc_j / spp) ]
  double sh = 0.0;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    const double P = exp(static_cast<double>(lp[i])), G = g[i];
    const double s = P / sp, gn = G / sg;
    const double ratio = gn / (s + EPS);
    const double a = -gn * ratio / ((EPS + ratio) * (s + EPS));
    const double dP = (a - sas) / sp - ((G - mg) / D - r * (P - mp) / spp);
    sh += P * dP;
  }
  sh = block_sum_d(sh, red);
  const double k = static_cast<double>(scale) / bsz;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    const double P = exp(static_cast<double>(lp[i])), G = g[i];
    const double s = P / sp, gn = G / sg;
    const double ratio = gn / (s + EPS);
    const double a = -gn * ratio / ((EPS + ratio) * (s + EPS));
    const double dP = (a - sas) / sp - ((G - mg) / D - r * (P - mp) / spp);
    dl[i] = static_cast<float>(k * (P * dP - P * sh));
  }
  if (threadIdx.x == 0) {
    per_sample[2 * b] = static_cast<float>(kl);
    per_sample[2 * b + 1] = static_cast<float>(r);
  }
}

// out = {loss, kl, cc, loss_va}: loss = mean kl - mean cc + gamma * loss_va
__global__ void salloss_finalize_kernel(const float* __restrict__ per_sample, const float* __restrict__ loss_va, float gamma,
                                        int bsz, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double kl = 0.0, cc = 0.0;
  for (int b = 0; b < bsz; ++b) {
    kl += per_sample[2 * b];
    cc += per_sample[2 * b + 1];
  }
  kl /= bsz;
  cc /= bsz;
  const float va = loss_va ? loss_va[0] : 0.f;
  out[0] = static_cast<float>(kl - cc) + gamma * va;
  out[1] = static_cast<float>(kl);
  out[2] = static_cast<float>(cc);
  out[3] = va;
}

// ------------------------------------------------------------------------- SimSiam loss backward
// L = -0.5/B sum_b [cos(pv_b, za_b) + cos(pa_b, zv_b)], z detached (model_utils.py:285-290): gradients flow into p only.
// grid (B, 2): dp = scale * (-0.5/B) * ( z/(|p||z|) - cos * p/|p|^2 )
__global__ void simsiam_bwd_kernel(const float* __restrict__ pv, const float* __restrict__ za, const float* __restrict__ pa,
                                   const float* __restrict__ zv, float* __restrict__ dpv, float* __restrict__ dpa, int bsz, int c,
                                   float scale) {
  __shared__ float red[32];
  const int b = blockIdx.x, pair = blockIdx.y;
  const float* p = (pair == 0 ? pv : pa) + static_cast<long long>(b) * c;
  const float* z = (pair == 0 ? za : zv) + static_cast<long long>(b) * c;
  float* dp = (pair == 0 ? dpv : dpa) + static_cast<long long>(b) * c;
  float dot = 0.f, pp = 0.f, zz = 0.f;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    const float a = p[i], q = z[i];
    dot = fmaf(a, q, dot);
    pp = fmaf(a, a, pp);
    zz = fmaf(q, q, zz);
  }
  dot = block_sum(dot, red);
  pp = block_sum(pp, red);
  zz = block_sum(zz, red);
  const float inv = rsqrtf(fmaxf(pp * zz, 1e-16f));
  const float cosv = dot * inv;
  const float k = -0.5f * scale / static_cast<float>(bsz);
  for (int i = threadIdx.x; i < c; i += blockDim.x) dp[i] = k * (z[i] * inv - cosv * p[i] / fmaxf(pp, 1e-16f));
}

// ------------------------------------------------------------------------- token mean backward, strided row add
__global__ void token_mean_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int rows, int r0, int r1, int c) {
  const int b = blockIdx.z;
  const int r = r0 + blockIdx.y;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c || r >= r1) return;
  dx[(static_cast<long long>(b) * rows + r) * c + ch] += dy[static_cast<long long>(b) * c + ch] / static_cast<float>(r1 - r0);
}

__global__ void add_rows_kernel(const float* __restrict__ src, long long srs, long long sgs, float* __restrict__ dst,
                                long long drs, long long dgs, int rows, int c4, long long total, int acc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long g = i;
    const int cc = divmod(g, c4) * 4;
    const long long within = divmod(g, rows);
    float4 v = ldf4(src + g * sgs + within * srs + cc);
    float* dp = dst + g * dgs + within * drs + cc;
    if (acc) {
      const float4 o = ldf4(dp);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    stf4(dp, v);
  }
}

// ------------------------------------------------------------------------- 4-D strided permute copy (weight packing)
// dst[i0*d0 + i1*d1 + i2*d2 + i3*d3] = src[i0*s0 + i1*s1 + i2*s2 + i3*s3]   (fp32 -> fp32 / bf16; acc: dst += for fp32)
struct PermArgs {
  int n[4];
  long long s[4], d[4];
  int dst_bf16, acc;
};
__global__ void permute_copy_kernel(PermArgs a, const float* __restrict__ src, void* __restrict__ dst, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int i3 = static_cast<int>(r % a.n[3]); r /= a.n[3];
    const int i2 = static_cast<int>(r % a.n[2]); r /= a.n[2];
    const int i1 = static_cast<int>(r % a.n[1]); r /= a.n[1];
    const int i0 = static_cast<int>(r);
    const float v = src[i0 * a.s[0] + i1 * a.s[1] + i2 * a.s[2] + i3 * a.s[3]];
    const long long o = i0 * a.d[0] + i1 * a.d[1] + i2 * a.d[2] + i3 * a.d[3];
    if (a.dst_bf16) static_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
    else if (a.acc) static_cast<float*>(dst)[o] += v;
    else static_cast<float*>(dst)[o] = v;
  }
}

// All weight re-packs of one step in ONE launch: a device table of jobs (16 x int64 each: src, dst, n[4], src_strides[4],
// dst_strides[4], dst_is_bf16, total) and a block -> (job, chunk) map built once by the host.
constexpr int kPackChunk = 2048;   // elements per block
__global__ void __launch_bounds__(256) permute_copy_batched_kernel(const long long* __restrict__ table,
                                                                   const int* __restrict__ block_job,
                                                                   const int* __restrict__ block_chunk) {
  const long long* j = table + static_cast<long long>(block_job[blockIdx.x]) * 16;
  const float* src = reinterpret_cast<const float*>(j[0]);
  void* dst = reinterpret_cast<void*>(j[1]);
  const int n1 = static_cast<int>(j[3]), n2 = static_cast<int>(j[4]), n3 = static_cast<int>(j[5]);
  const long long s0 = j[6], s1 = j[7], s2 = j[8], s3 = j[9], d0 = j[10], d1 = j[11], d2 = j[12], d3 = j[13];
  const bool bf = j[14] != 0;
  const long long total = j[15];
  const long long base = static_cast<long long>(block_chunk[blockIdx.x]) * kPackChunk;
#pragma unroll
  for (int e = 0; e < kPackChunk / 256; ++e) {
    const long long i = base + e * 256 + threadIdx.x;
    if (i >= total) break;
    long long r = i;
    const int i3 = static_cast<int>(r % n3); r /= n3;
    const int i2 = static_cast<int>(r % n2); r /= n2;
    const int i1 = static_cast<int>(r % n1); r /= n1;
    const long long i0 = r;
    const float v = src[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3];
    const long long o = i0 * d0 + i1 * d1 + i2 * d2 + i3 * d3;
    if (bf) static_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
    else static_cast<float*>(dst)[o] = v;
  }
}

// ------------------------------------------------------------------------- AdamW over the flat parameter buffer
// torch.optim.AdamW (train.py:157-158): p *= 1 - lr*wd; m, v moments; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n4, float lr, float b1, float b2, float eps, float wd, int step_host,
                             const int* __restrict__ step_dev, float gscale) {
  const float step = static_cast<float>(step_dev ? *step_dev : step_host);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 P = ldf4(p + i * 4), G = ldf4(g + i * 4), M = ldf4(m + i * 4), V = ldf4(v + i * 4);
    float* pe[4] = {&P.x, &P.y, &P.z, &P.w};
    const float ge[4] = {G.x * gscale, G.y * gscale, G.z * gscale, G.w * gscale};
    float* me[4] = {&M.x, &M.y, &M.z, &M.w};
    float* ve[4] = {&V.x, &V.y, &V.z, &V.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float pj = *pe[j] * (1.f - lr * wd);
      const float mj = b1 * *me[j] + (1.f - b1) * ge[j];
      const float vj = b2 * *ve[j] + (1.f - b2) * ge[j] * ge[j];
      const float denom = sqrtf(vj) / bc2_sqrt + eps;
      pj -= (lr / bc1) * mj / denom;
      *pe[j] = pj;
      *me[j] = mj;
      *ve[j] = vj;
    }
    stf4(p + i * 4, P);
    stf4(m + i * 4, M);
    stf4(v + i * 4, V);
  }
}

}  // namespace
}  // namespace mspi

using namespace mspi;

#define MSPI_NEED_GPU() \
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device")
#define MSPI_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int mspi_bn_train_fwd(const float* x, int64_t x_cstride, float* y, int64_t y_cstride, int64_t pixels, int c,
                                 const float* weight, const float* bias, float eps, float momentum, float* running_mean,
                                 float* running_var, int64_t* num_batches_tracked, float* save_mean, float* save_invstd,
                                 double* work, int relu, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && y && weight && bias && save_mean && save_invstd && work, "mspi_bn_train_fwd: null argument");
  MSPI_CHECK_ARG(c % 4 == 0 && c <= 4096 && x_cstride % 4 == 0 && y_cstride % 4 == 0 && pixels > 0,
                 "channels / strides must be multiples of 4 (c <= 4096)");
  MSPI_CHECK_ARG(MSPI_ALIGNED16(x) && MSPI_ALIGNED16(y), "16-byte alignment");
  MSPI_NEED_GPU();
  const int cblocks = (c + 31) / 32;
  const int ppb = pixels_per_block(pixels, cblocks);
  dim3 grid(static_cast<unsigned>((pixels + ppb - 1) / ppb), cblocks), block(32, 8);
  bn_stats_kernel<<<grid, block, 0, stream>>>(x, x_cstride, pixels, c, ppb, work);
  MSPI_LAUNCH_CHECK();
  const long long total = pixels * (c / 4);
  bn_apply_kernel<<<grid_for(total), kBlock, 2 * c * sizeof(float), stream>>>(
      x, x_cstride, y, y_cstride, total, c / 4, c, work, static_cast<double>(pixels), weight, bias, eps, momentum, running_mean,
      running_var, reinterpret_cast<long long*>(num_batches_tracked), save_mean, save_invstd, relu);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_bn_train_bwd(const float* x, int64_t x_cstride, const float* y, int64_t y_cstride, const float* dy,
                                 int64_t dy_cstride, float* dx, int64_t dx_cstride, int64_t pixels, int c, const float* weight,
                                 const float* save_mean, const float* save_invstd, float* dweight, float* dbias, double* work,
                                 int relu, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && dy && dx && weight && save_mean && save_invstd && dweight && dbias && work && (y || !relu),
                 "mspi_bn_train_bwd: null argument");
  MSPI_CHECK_ARG(c % 4 == 0 && c <= 2048 && x_cstride % 4 == 0 && dy_cstride % 4 == 0 && dx_cstride % 4 == 0 &&
                     (!relu || y_cstride % 4 == 0),
                 "channels / strides must be multiples of 4 (c <= 2048)");
  MSPI_CHECK_ARG(MSPI_ALIGNED16(x) && MSPI_ALIGNED16(dy) && MSPI_ALIGNED16(dx) && (!relu || MSPI_ALIGNED16(y)), "16-byte alignment");
  MSPI_NEED_GPU();
  const int cblocks = (c + 31) / 32;
  const int ppb = pixels_per_block(pixels, cblocks);
  dim3 grid(static_cast<unsigned>((pixels + ppb - 1) / ppb), cblocks), block(32, 8);
  bn_bwd_reduce_kernel<<<grid, block, 0, stream>>>(x, x_cstride, y, y_cstride, dy, dy_cstride, pixels, c, ppb, save_mean,
                                                   save_invstd, relu, work);
  MSPI_LAUNCH_CHECK();
  const long long total = pixels * (c / 4);
  bn_bwd_apply_kernel<<<grid_for(total), kBlock, 5 * c * sizeof(float), stream>>>(
      x, x_cstride, y, y_cstride, dy, dy_cstride, dx, dx_cstride, total, c / 4, c, save_mean, save_invstd, weight, work,
      static_cast<double>(pixels), dweight, dbias, relu, accumulate);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_act_fwd(const float* x, float* y, int64_t n, int act, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && y && n % 4 == 0 && MSPI_ALIGNED16(x) && MSPI_ALIGNED16(y), "mspi_act_fwd: bad argument");
  MSPI_NEED_GPU();
  act_fwd_kernel<<<grid_for(n / 4), kBlock, 0, stream>>>(x, y, n / 4, act);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_act_bwd(const float* dy, int64_t dy_cstride, const float* ref, int64_t ref_cstride, float* dz,
                            int64_t dz_cstride, int64_t pixels, int c, int act, float* dbias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(dy && pixels > 0 && c > 0 && (ref || act == MSPI_ACT_NONE) && (dz || dbias), "mspi_act_bwd: bad argument");
  MSPI_CHECK_ARG(act == MSPI_ACT_NONE || act == MSPI_ACT_RELU || act == MSPI_ACT_GELU, "act %d", act);
  MSPI_NEED_GPU();
  const int cblocks = (c + 31) / 32;
  const int ppb = pixels_per_block(pixels, cblocks);
  dim3 grid(static_cast<unsigned>((pixels + ppb - 1) / ppb), cblocks), block(32, 8);
  act_bwd_kernel<<<grid, block, 0, stream>>>(dy, dy_cstride, ref, ref_cstride, dz, dz_cstride, pixels, c, ppb, act, dbias);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_maxpool3d_f32(const MspiPoolDesc* d, const float* x, float* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && y, "mspi_maxpool3d_f32: null argument");
  MSPI_CHECK_ARG(d->c % 4 == 0 && d->in_cstride % 4 == 0 && d->out_cstride % 4 == 0 && MSPI_ALIGNED16(x) && MSPI_ALIGNED16(y),
                 "channels must be multiples of 4, pointers 16-byte aligned");
  MSPI_NEED_GPU();
  const long long total = static_cast<long long>(d->n) * d->ot * d->oh * d->ow * (d->c / 4);
  maxpool3d_f32_kernel<<<grid_for(total), kBlock, 0, stream>>>(*d, x, y, total, d->c / 4);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_maxpool3d_bwd(const MspiPoolDesc* d, const float* x, const float* dy, int64_t dy_cstride, float* dx,
                                  int64_t dx_cstride, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && dy && dx, "mspi_maxpool3d_bwd: null argument");
  MSPI_CHECK_ARG(d->c % 4 == 0 && d->in_cstride % 4 == 0 && dy_cstride % 4 == 0 && MSPI_ALIGNED16(x) && MSPI_ALIGNED16(dy),
                 "channels must be multiples of 4, pointers 16-byte aligned");
  MSPI_NEED_GPU();
  const long long total = static_cast<long long>(d->n) * d->ot * d->oh * d->ow * (d->c / 4);
  maxpool3d_bwd_kernel<<<grid_for(total), kBlock, 0, stream>>>(*d, x, dy, dy_cstride, dx, dx_cstride, total, d->c / 4);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_upsample_bilinear_bwd(const MspiUpDesc* d, const float* dy, const float* y, float* dx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && dy && dx && d->k >= 1 && (y || d->act != MSPI_ACT_RELU), "mspi_upsample_bilinear_bwd: bad argument");
  MSPI_CHECK_ARG(d->c % 4 == 0 && d->in_cstride % 4 == 0 && d->out_cstride % 4 == 0 && MSPI_ALIGNED16(dy) && MSPI_ALIGNED16(dx),
                 "channels must be multiples of 4, pointers 16-byte aligned");
  MSPI_NEED_GPU();
  const long long total = static_cast<long long>(d->nt) * d->h * d->k * d->w * d->k * (d->c / 4);
  upsample_bwd_kernel<<<grid_for(total), kBlock, 0, stream>>>(*d, dy, y, dx, total, d->c / 4);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_layernorm_bwd(const float* x, int64_t x_rstride, const float* dy, int64_t dy_rstride,
                                  int64_t rows_per_group, int64_t dy_gstride, const float* y_relu, const float* w, float eps,
                                  float* dx, int64_t dx_rstride, int64_t rows, int c, int accumulate, float* dw, float* db,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && dy && w && dx && dw && db && rows > 0 && c > 0 && c <= 6000, "mspi_layernorm_bwd: bad argument");
  MSPI_NEED_GPU();
  if (rows_per_group <= 0) rows_per_group = rows;
  const int warps = 8;
  long long blocks = (rows + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (blocks > cap) blocks = cap;
  layernorm_bwd_kernel<<<static_cast<int>(blocks), warps * 32, 2 * c * sizeof(float), stream>>>(
      x, x_rstride, dy, dy_rstride, rows_per_group, dy_gstride, y_relu, w, eps, dx, dx_rstride, rows, c, accumulate, dw, db);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_softmax_bwd_rows(const float* p, float* dp, int64_t rows, int n, int64_t stride, float scale,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(p && dp && rows > 0 && n > 0, "mspi_softmax_bwd_rows: bad argument");
  MSPI_NEED_GPU();
  const int warps = 8;
  long long blocks = (rows + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  softmax_bwd_rows_kernel<<<static_cast<int>(blocks), warps * 32, 0, stream>>>(p, dp, rows, n, stride, scale);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_sgemm_strided(const MspiSgemmDesc* d, const float* a, const float* b, float* c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && a && b && c && d->m > 0 && d->n > 0 && d->k > 0 && d->batch0 > 0 && d->batch1 > 0,
                 "mspi_sgemm_strided: bad argument");
  MSPI_CHECK_ARG(static_cast<long long>(d->batch0) * d->batch1 <= 65535, "too many batches");
  MSPI_NEED_GPU();
  SgemmArgs g;
  g.m = d->m; g.n = d->n; g.k = d->k; g.b0 = d->batch0; g.b1 = d->batch1; g.acc = d->accumulate; g.alpha = d->alpha;
  g.a_i = d->a_strides[0]; g.a_k = d->a_strides[1]; g.a_b0 = d->a_strides[2]; g.a_b1 = d->a_strides[3];
  g.b_k = d->b_strides[0]; g.b_j = d->b_strides[1]; g.b_b0 = d->b_strides[2]; g.b_b1 = d->b_strides[3];
  g.c_i = d->c_strides[0]; g.c_j = d->c_strides[1]; g.c_b0 = d->c_strides[2]; g.c_b1 = d->c_strides[3];
  dim3 grid((d->n + 31) / 32, (d->m + 31) / 32, d->batch0 * d->batch1);
  sgemm_strided_kernel<<<grid, 256, 0, stream>>>(g, a, b, c);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_sa_gate_bwd(const float* x, int64_t x_cstride, const float* mask_logits, const float* dy,
                                int64_t dy_cstride, float* dx, int64_t dx_cstride, float* dlogits, int64_t pixels, int c,
                                int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && mask_logits && dy && dx && dlogits && c % 4 == 0 && x_cstride % 4 == 0 && dy_cstride % 4 == 0 &&
                     dx_cstride % 4 == 0,
                 "mspi_sa_gate_bwd: bad argument");
  MSPI_CHECK_ARG(MSPI_ALIGNED16(x) && MSPI_ALIGNED16(dy) && MSPI_ALIGNED16(dx), "16-byte alignment");
  MSPI_NEED_GPU();
  const int warps = 8;
  long long blocks = (pixels + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  sa_gate_bwd_kernel<<<static_cast<int>(blocks), warps * 32, 0, stream>>>(x, x_cstride, mask_logits, dy, dy_cstride, dx,
                                                                         dx_cstride, dlogits, pixels, c, accumulate);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_conv_c1_bwd(const float* x, int64_t x_cstride, const float* dy, const float* w, float* dx,
                                int64_t dx_cstride, float* dw, float* db, int64_t planes, int h, int wd, int cin,
                                int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && dy && w && dw && planes > 0 && h > 0 && wd > 0, "mspi_conv_c1_bwd: bad argument");
  MSPI_CHECK_ARG(cin == 32, "mspi_conv_c1_bwd serves the 32 -> 1 (1,3,3) convs (cin = %d)", cin);
  MSPI_NEED_GPU();
  const long long pixels = planes * h * wd;
  const int ppb = pixels_per_block(pixels, 1);
  dim3 block(32, 8);
  conv_c1_bwd_kernel<<<static_cast<unsigned>((pixels + ppb - 1) / ppb), block, 0, stream>>>(x, x_cstride, dy, w, dx, dx_cstride,
                                                                                           dw, db, pixels, h, wd, ppb, accumulate);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_conv_c1_fwd(const void* x, int x_dtype, int64_t x_cstride, const float* w, const float* bias, float* y,
                                int64_t planes, int h, int wd, int cin, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && w && y && planes > 0 && h > 0 && wd > 0, "mspi_conv_c1_fwd: bad argument");
  MSPI_CHECK_ARG(cin == 32, "mspi_conv_c1_fwd serves the 32 -> 1 (1,3,3) convs (cin = %d)", cin);
  MSPI_NEED_GPU();
  const int warps = 8;
  // strips tall enough to amortise the two halo rows, short enough to give every SM several warps
  int strip = 16;
  while (strip > 4 && planes * ((h + strip - 1) / strip) * ((wd + 3) / 4) < static_cast<long long>(num_sms()) * 64) strip /= 2;
  const long long items = planes * ((h + strip - 1) / strip) * ((wd + 3) / 4);
  long long blocks = (items + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (x_dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(conv_c1_fwd_kernel<__nv_bfloat16>, static_cast<int>(blocks), warps * 32, 0, stream, 
        static_cast<const __nv_bfloat16*>(x), x_cstride, w, bias, y, planes, h, wd, strip));
  else
    MSPI_CUDA(launch_pdl(conv_c1_fwd_kernel<float>, static_cast<int>(blocks), warps * 32, 0, stream, static_cast<const float*>(x), x_cstride, w,
                                                                                  bias, y, planes, h, wd, strip));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_dwconv_wgrad(const MspiDwDesc* d, const float* x, const float* dy, float* dw, float* db, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && dy && dw && db, "mspi_dwconv_wgrad: null argument");
  MSPI_NEED_GPU();
  const long long pixels = static_cast<long long>(d->n) * d->t * d->h * d->w;
  const int cblocks = (d->c + 31) / 32;
  const int ppb = pixels_per_block(pixels, cblocks);
  dim3 grid(static_cast<unsigned>((pixels + ppb - 1) / ppb), cblocks), block(32, 8);
  if (d->kt == 7 && d->kh == 1 && d->kw == 1)
    dw_wgrad_kernel<7, 1, 1><<<grid, block, 0, stream>>>(x, dy, d->n, d->t, d->h, d->w, d->c, ppb, dw, db);
  else if (d->kt == 1 && d->kh == 7 && d->kw == 7)
    dw_wgrad_kernel<1, 7, 7><<<grid, block, 0, stream>>>(x, dy, d->n, d->t, d->h, d->w, d->c, ppb, dw, db);
  else
    return set_error(MSPI_ERR_UNSUPPORTED, "mspi_dwconv_wgrad: kernel (%d,%d,%d) (ConvNextBlock uses (7,1,1) and (1,7,7))", d->kt,
                     d->kh, d->kw);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_salloss_bwd(const float* log_map, const float* gt, const float* loss_va, float gamma, float* dlogits,
                                float* out, float* work, int b, int64_t pixels, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(log_map && gt && dlogits && out && work && b > 0 && pixels > 1, "mspi_salloss_bwd: bad argument");
  MSPI_NEED_GPU();
  salloss_bwd_kernel<<<b, 1024, 0, stream>>>(log_map, gt, dlogits, work, pixels, b, scale);
  MSPI_LAUNCH_CHECK();
  salloss_finalize_kernel<<<1, 32, 0, stream>>>(work, loss_va, gamma, b, out);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_simsiam_bwd(const float* p_v, const float* z_a, const float* p_a, const float* z_v, float* dp_v,
                                float* dp_a, int b, int c, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(p_v && z_a && p_a && z_v && dp_v && dp_a && b > 0 && c > 0, "mspi_simsiam_bwd: bad argument");
  MSPI_NEED_GPU();
  simsiam_bwd_kernel<<<dim3(b, 2), 256, 0, stream>>>(p_v, z_a, p_a, z_v, dp_v, dp_a, b, c, scale);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_token_mean_bwd(const float* dy, float* dx, int b, int rows, int r0, int r1, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(dy && dx && b > 0 && r1 > r0 && r0 >= 0 && r1 <= rows && c > 0, "mspi_token_mean_bwd: bad argument");
  MSPI_NEED_GPU();
  token_mean_bwd_kernel<<<dim3((c + 127) / 128, r1 - r0, b), 128, 0, stream>>>(dy, dx, rows, r0, r1, c);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_add_rows(const float* src, int64_t src_rstride, int64_t src_gstride, float* dst, int64_t dst_rstride,
                             int64_t dst_gstride, int groups, int rows, int c, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(src && dst && groups > 0 && rows > 0 && c > 0 && c % 4 == 0, "mspi_add_rows: bad argument");
  MSPI_CHECK_ARG(src_rstride % 4 == 0 && src_gstride % 4 == 0 && dst_rstride % 4 == 0 && dst_gstride % 4 == 0 &&
                     MSPI_ALIGNED16(src) && MSPI_ALIGNED16(dst),
                 "strides must be multiples of 4, pointers 16-byte aligned");
  MSPI_NEED_GPU();
  const long long total = static_cast<long long>(groups) * rows * (c / 4);
  add_rows_kernel<<<grid_for(total), kBlock, 0, stream>>>(src, src_rstride, src_gstride, dst, dst_rstride, dst_gstride, rows,
                                                          c / 4, total, accumulate);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_permute_copy(const MspiPermDesc* d, const float* src, void* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && src && dst, "mspi_permute_copy: null argument");
  MSPI_NEED_GPU();
  PermArgs a;
  long long total = 1;
  for (int j = 0; j < 4; ++j) {
    MSPI_CHECK_ARG(d->n[j] >= 1, "extent %d = %d", j, d->n[j]);
    a.n[j] = d->n[j];
    a.s[j] = d->src_strides[j];
    a.d[j] = d->dst_strides[j];
    total *= d->n[j];
  }
  a.dst_bf16 = d->dst_dtype == MSPI_BF16;
  a.acc = d->accumulate;
  permute_copy_kernel<<<grid_for(total), kBlock, 0, stream>>>(a, src, dst, total);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_permute_copy_batched(const int64_t* table, const int32_t* block_job, const int32_t* block_chunk,
                                         int nblocks, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(table && block_job && block_chunk && nblocks > 0, "mspi_permute_copy_batched: bad argument");
  MSPI_NEED_GPU();
  permute_copy_batched_kernel<<<nblocks, 256, 0, stream>>>(reinterpret_cast<const long long*>(table), block_job, block_chunk);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                               float eps, float weight_decay, int step, const int32_t* step_dev, float grad_scale,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(p && g && m && v && n > 0 && n % 4 == 0 && (step >= 1 || step_dev),
                 "mspi_adamw_step: bad argument (n must be a multiple of 4)");
  MSPI_CHECK_ARG(MSPI_ALIGNED16(p) && MSPI_ALIGNED16(g) && MSPI_ALIGNED16(m) && MSPI_ALIGNED16(v), "16-byte alignment");
  MSPI_NEED_GPU();
  adamw_kernel<<<grid_for(n / 4), kBlock, 0, stream>>>(p, g, m, v, n / 4, lr, beta1, beta2, eps, weight_decay, step, step_dev,
                                                       grad_scale);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}
