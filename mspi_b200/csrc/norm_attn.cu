// Depthwise conv + LayerNorm, LayerNorm, small-sequence attention, token mean, SimSiam loss.
// One warp per pixel/row with the channel dimension spread over lanes (coalesced bf16x2 accesses),
// fp32 arithmetic, warp-shuffle reductions.
#include "common.cuh"

namespace mspi {
namespace {

// ------------------------------------------------------------------------- depthwise conv (+LN)
// MAXP = channel pairs per lane (C <= 64*MAXP).
__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16* row, int p) {
  return __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(row)[p]);
}
__device__ __forceinline__ float2 ld_pair(const float* row, int p) { return reinterpret_cast<const float2*>(row)[p]; }

template <int MAXP, typename TI>
__global__ void dwconv_ln_kernel(MspiDwDesc d, const TI* __restrict__ x, const float* __restrict__ wgt,
                                 const float* __restrict__ bias, const float* __restrict__ ln_w,
                                 const float* __restrict__ ln_b, void* __restrict__ y, long long pixels) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int c = d.c, pairs = c >> 1;
  const int pt = d.kt / 2, ph = d.kh / 2, pw = d.kw / 2;
  for (long long pix = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); pix < pixels;
       pix += static_cast<long long>(gridDim.x) * warps) {
    long long r = pix;
    const int ow = static_cast<int>(r % d.w); r /= d.w;
    const int oh = static_cast<int>(r % d.h); r /= d.h;
    const int ot = static_cast<int>(r % d.t); r /= d.t;
    const long long n = r;
    float acc[2 * MAXP];
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int p = lane + 32 * i;
      const float2 b = p < pairs ? __ldg(reinterpret_cast<const float2*>(bias) + p) : make_float2(0.f, 0.f);
      acc[2 * i] = b.x;
      acc[2 * i + 1] = b.y;
    }
    for (int kt = 0; kt < d.kt; ++kt) {
      const int it = ot + kt - pt;
      if (it < 0 || it >= d.t) continue;
      for (int kh = 0; kh < d.kh; ++kh) {
        const int ih = oh + kh - ph;
        if (ih < 0 || ih >= d.h) continue;
        for (int kw = 0; kw < d.kw; ++kw) {
          const int iw = ow + kw - pw;
          if (iw < 0 || iw >= d.w) continue;
          const int tap = (kt * d.kh + kh) * d.kw + kw;
          const TI* xp = x + (((n * d.t + it) * d.h + ih) * d.w + iw) * c;
          const float2* wp = reinterpret_cast<const float2*>(wgt + static_cast<long long>(tap) * c);
#pragma unroll
          for (int i = 0; i < MAXP; ++i) {
            const int p = lane + 32 * i;
            if (p < pairs) {
              const float2 xv = ld_pair(xp, p);
              const float2 wv = __ldg(wp + p);
              acc[2 * i] = fmaf(xv.x, wv.x, acc[2 * i]);
              acc[2 * i + 1] = fmaf(xv.y, wv.y, acc[2 * i + 1]);
            }
          }
        }
      }
    }
    if (ln_w != nullptr) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXP; ++i)
        if (lane + 32 * i < pairs) s += acc[2 * i] + acc[2 * i + 1];
      const float mean = warp_sum(s) / c;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < MAXP; ++i)
        if (lane + 32 * i < pairs) {
          const float a = acc[2 * i] - mean, b = acc[2 * i + 1] - mean;
          q += a * a + b * b;
        }
      const float rstd = rsqrtf(warp_sum(q) / c + d.ln_eps);
#pragma unroll
      for (int i = 0; i < MAXP; ++i) {
        const int p = lane + 32 * i;
        if (p < pairs) {
          const float2 g = __ldg(reinterpret_cast<const float2*>(ln_w) + p);
          const float2 b = __ldg(reinterpret_cast<const float2*>(ln_b) + p);
          acc[2 * i] = (acc[2 * i] - mean) * rstd * g.x + b.x;
          acc[2 * i + 1] = (acc[2 * i + 1] - mean) * rstd * g.y + b.y;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int p = lane + 32 * i;
      if (p < pairs) {
        if (d.out_dtype == MSPI_BF16)
          reinterpret_cast<__nv_bfloat162*>(static_cast<__nv_bfloat16*>(y) + pix * c)[p] =
              __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
        else
          reinterpret_cast<float2*>(static_cast<float*>(y) + pix * c)[p] = make_float2(acc[2 * i], acc[2 * i + 1]);
      }
    }
  }
}

// ------------------------------------------------------------------------- LayerNorm rows
template <typename TI>
__device__ __forceinline__ float ld_elem(const TI* p, long long i) { return static_cast<float>(p[i]); }

template <typename TI, typename TO>
__global__ void layernorm_kernel(MspiLnDesc d, const TI* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ b, const float* __restrict__ pos, TO* __restrict__ y) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (long long row = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); row < d.rows;
       row += static_cast<long long>(gridDim.x) * warps) {
    const TI* xr = x + row * d.in_rstride;
    float s = 0.f;
    for (int i = lane; i < d.c; i += 32) s += ld_elem(xr, i);
    const float mean = warp_sum(s) / d.c;
    float q = 0.f;
    for (int i = lane; i < d.c; i += 32) {
      const float a = ld_elem(xr, i) - mean;
      q += a * a;
    }
    const float rstd = rsqrtf(warp_sum(q) / d.c + d.eps);
    const long long g = row / d.rows_per_group, within = row - g * d.rows_per_group;
    TO* yr = y + g * d.out_gstride + within * d.out_rstride;
    const float* pr = d.pos_rows > 0 ? pos + (within % d.pos_rows) * d.c : nullptr;
    for (int i = lane; i < d.c; i += 32) {
      float v = (ld_elem(xr, i) - mean) * rstd * __ldg(w + i) + __ldg(b + i);
      if (d.relu) v = fmaxf(v, 0.f);
      if (pr) v += __ldg(pr + i);
      yr[i] = static_cast<TO>(v);
    }
  }
}

// Vectorised variant for C % 4 == 0, C <= 128*NI: a warp owns a row, each lane keeps its 4-element chunks in
// registers (one global read of the row), 8/16-byte accesses, two-pass statistics in registers.
template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  float4 f;
  unpack_bf16x2(u.x, f.x, f.y);
  unpack_bf16x2(u.y, f.z, f.w);
  return f;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// G rows per warp iteration: all of their loads are issued before the first reduction, so a warp keeps G x C elements
// in flight (one row per warp left the kernel latency bound at ~1.8 TB/s for C = 96).
template <typename TI, typename TO, int NI, int G>
__global__ void layernorm_vec_kernel(MspiLnDesc d, const TI* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ b, const float* __restrict__ pos, TO* __restrict__ y) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int c4 = d.c >> 2;
  const float inv_c = 1.f / d.c;
  for (long long row0 = (static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5)) * G; row0 < d.rows;
       row0 += static_cast<long long>(gridDim.x) * warps * G) {
    float4 v[G][NI];
    float s[G], sq[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long row = row0 + g;
      const TI* xr = x + row * d.in_rstride;
      s[g] = 0.f;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int q = lane + 32 * i;
        v[g][i] = (q < c4 && row < d.rows) ? ld4<TI>(xr + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        s[g] += (v[g][i].x + v[g][i].y) + (v[g][i].z + v[g][i].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      s[g] *= inv_c;  // mean
      sq[g] = 0.f;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (lane + 32 * i < c4) {
          const float a0 = v[g][i].x - s[g], a1 = v[g][i].y - s[g], a2 = v[g][i].z - s[g], a3 = v[g][i].w - s[g];
          sq[g] += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) sq[g] += __shfl_xor_sync(0xffffffffu, sq[g], o);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long row = row0 + g;
      if (row >= d.rows) break;
      const float mean = s[g];
      const float rstd = rsqrtf(sq[g] * inv_c + d.eps);
      long long grp = 0, within = row;   // (the 64-bit divisions below cost more than the rest of a C = 96 row)
      if (d.rows_per_group < d.rows) {
        grp = row;
        within = divmod(grp, static_cast<int>(d.rows_per_group));
      }
      TO* yr = y + grp * d.out_gstride + within * d.out_rstride;
      const float* pr = d.pos_rows > 0 ? pos + (within % d.pos_rows) * d.c : nullptr;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int q = lane + 32 * i;
        if (q < c4) {
          const float4 gw = __ldg(reinterpret_cast<const float4*>(w) + q);
          const float4 gb = __ldg(reinterpret_cast<const float4*>(b) + q);
          float4 o;
          o.x = (v[g][i].x - mean) * rstd * gw.x + gb.x;
          o.y = (v[g][i].y - mean) * rstd * gw.y + gb.y;
          o.z = (v[g][i].z - mean) * rstd * gw.z + gb.z;
          o.w = (v[g][i].w - mean) * rstd * gw.w + gb.w;
          if (d.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (pr) {
            const float4 pp = __ldg(reinterpret_cast<const float4*>(pr) + q);
            o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w;
          }
          st4(yr + 4 * q, o);
        }
      }
    }
  }
}

// Row LayerNorm for channel counts that are multiples of 8, the ConvNeXt stem / downsample / stage-3 shapes.  The 4-element
// kernel above needs ~70 issue slots per 4 x 24 lanes of a 96-channel row, which caps it near 2 TB/s (the SMs run out of
// issue slots before HBM runs out of bandwidth).  Here a lane owns 8 consecutive channels (one 16-byte bf16 load), a row is
// spread over LPR = 8 / 16 / 32 lanes so that 32 / LPR rows sit side by side in a warp and share every shuffle, the
// arithmetic is packed f32x2, and the affine parameters stay in registers across rows.
template <typename T>
struct V8;
template <>
struct V8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, F2 (&v)[4]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = pack2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
    v[1] = pack2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    v[2] = pack2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
    v[3] = pack2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const F2 (&v)[4]) {
    uint4 o;
    float a, b;
    unpack2(v[0], a, b); o.x = pack_bf16x2(a, b);
    unpack2(v[1], a, b); o.y = pack_bf16x2(a, b);
    unpack2(v[2], a, b); o.z = pack_bf16x2(a, b);
    unpack2(v[3], a, b); o.w = pack_bf16x2(a, b);
    *reinterpret_cast<uint4*>(p) = o;
  }
};
template <>
struct V8<float> {
  static __device__ __forceinline__ void load(const float* p, F2 (&v)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = pack2(a.x, a.y); v[1] = pack2(a.z, a.w); v[2] = pack2(b.x, b.y); v[3] = pack2(b.z, b.w);
  }
  static __device__ __forceinline__ void store(float* p, const F2 (&v)[4]) {
    float4 a, b;
    unpack2(v[0], a.x, a.y); unpack2(v[1], a.z, a.w); unpack2(v[2], b.x, b.y); unpack2(v[3], b.z, b.w);
    reinterpret_cast<float4*>(p)[0] = a;
    reinterpret_cast<float4*>(p)[1] = b;
  }
};

template <typename TI, typename TO, int LPR, int NV, int G>
__global__ void __launch_bounds__(256)
layernorm_rows8_kernel(MspiLnDesc d, const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                       const float* __restrict__ pos, TO* __restrict__ y) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  constexpr int RPW = 32 / LPR;   // rows side by side in a warp
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const int warps = blockDim.x >> 5;
  const float inv_c = 1.f / d.c;
  F2 gw[NV][4], gb[NV][4];
  bool act[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = 8 * (sl + LPR * i);
    act[i] = e < d.c;
    if (act[i]) {
      V8<float>::load(w + e, gw[i]);
      V8<float>::load(b + e, gb[i]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) gw[i][k] = gb[i][k] = 0ull;
    }
  }
  // 32-bit row indices and strides (the host takes this kernel only when they fit): the 64-bit index arithmetic of the
  // generic kernel was a third of the instructions of a 96-channel row
  const int rows = static_cast<int>(d.rows), istride = static_cast<int>(d.in_rstride), ostride = static_cast<int>(d.out_rstride);
  const bool grouped = d.rows_per_group < d.rows;
  const int step = gridDim.x * warps * (RPW * G);
  for (int row0 = (blockIdx.x * warps + (threadIdx.x >> 5)) * (RPW * G); row0 < rows; row0 += step) {
    F2 v[G][NV][4];
    float s[G], sq[G];
    // Loads are unconditional so that all G x NV of them are issued before the first use: rows past the end re-read the
    // last row (their results are dropped below) and lanes past the channel count re-read the row start (their partial
    // sums are discarded by a select).
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int row = min(row0 + g * RPW + sub, rows - 1);
      const TI* xr = x + static_cast<size_t>(static_cast<unsigned>(row)) * static_cast<unsigned>(istride);
#pragma unroll
      for (int i = 0; i < NV; ++i) V8<TI>::load(xr + (act[i] ? 8 * (sl + LPR * i) : 0), v[g][i]);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      s[g] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const F2 part = add2(add2(v[g][i][0], v[g][i][1]), add2(v[g][i][2], v[g][i][3]));
        float lo, hi;
        unpack2(part, lo, hi);
        s[g] += act[i] ? lo + hi : 0.f;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float nm = -s[g] * inv_c;
      const F2 nm2 = pack2(nm, nm);
      sq[g] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        F2 acc = 0ull;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[g][i][k] = add2(v[g][i][k], nm2);   // centred
          acc = fma2(v[g][i][k], v[g][i][k], acc);
        }
        float lo, hi;
        unpack2(acc, lo, hi);
        sq[g] += act[i] ? lo + hi : 0.f;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) sq[g] += __shfl_xor_sync(0xffffffffu, sq[g], o);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int row = row0 + g * RPW + sub;
      if (row >= rows) continue;
      const float rstd = rsqrtf(sq[g] * inv_c + d.eps);
      const F2 rs2 = pack2(rstd, rstd);
      unsigned grp = 0, within = static_cast<unsigned>(row);
      if (grouped) {
        grp = within / static_cast<unsigned>(d.rows_per_group);
        within -= grp * static_cast<unsigned>(d.rows_per_group);
      }
      TO* yr = y + grp * d.out_gstride + static_cast<size_t>(within) * static_cast<unsigned>(ostride) + 8 * sl;
      const float* pr = d.pos_rows > 0 ? pos + static_cast<size_t>(within % static_cast<unsigned>(d.pos_rows)) * d.c + 8 * sl : nullptr;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (!act[i]) continue;
        F2 o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = fma2(mul2(v[g][i][k], rs2), gw[i][k], gb[i][k]);
        if (d.relu) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float lo, hi;
            unpack2(o[k], lo, hi);
            o[k] = pack2(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
          }
        }
        if (pr) {
          F2 pp[4];
          V8<float>::load(pr + 8 * LPR * i, pp);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = add2(o[k], pp[k]);
        }
        V8<TO>::store(yr + 8 * LPR * i, o);
      }
    }
  }
}

// ------------------------------------------------------------------------- attention
// qkv: [B][N][3][H][HD] bf16.  Block = (b*H + h, query tile of QT rows).  Scores for the whole
// key range live in shared memory (N <= 1024), softmax in fp32, then P·V.
constexpr int kQT = 16;
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
  unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int n, int heads, int hd, float scale, int n_pad) {
  extern __shared__ float sm[];
  float* q_s = sm;                 // [kQT][hd]
  float* s_s = sm + kQT * hd;      // [kQT][n_pad]
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int q0 = blockIdx.y * kQT;
  const int tid = threadIdx.x;
  const long long row_stride = 3ll * heads * hd;
  const T* base = qkv + static_cast<long long>(b) * n * row_stride + h * hd;
  for (int i = tid; i < kQT * hd; i += kAttnThreads) {
    const int qi = i / hd, dd = i % hd;
    const int qr = q0 + qi;
    q_s[i] = qr < n ? static_cast<float>(base[static_cast<long long>(qr) * row_stride + dd]) * scale : 0.f;
  }
  __syncthreads();
  // phase 1: scores
  for (int k = tid; k < n; k += kAttnThreads) {
    const T* kp = base + static_cast<long long>(k) * row_stride + heads * hd;
    float acc[kQT];
#pragma unroll
    for (int qi = 0; qi < kQT; ++qi) acc[qi] = 0.f;
    for (int d8 = 0; d8 < hd / 8; ++d8) {
      float kf[8];
      load8(kp + d8 * 8, kf);
#pragma unroll
      for (int qi = 0; qi < kQT; ++qi) {
        const float* qq = q_s + qi * hd + d8 * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[qi] = fmaf(qq[e], kf[e], acc[qi]);
      }
    }
#pragma unroll
    for (int qi = 0; qi < kQT; ++qi) s_s[qi * n_pad + k] = acc[qi];
  }
  __syncthreads();
  // phase 2: softmax per query row (one warp per row, 4 warps -> 4 rows at a time)
  const int warp = tid >> 5, lane = tid & 31;
  for (int qi = warp; qi < kQT; qi += kAttnThreads / 32) {
    float* sr = s_s + qi * n_pad;
    float m = -INFINITY;
    for (int k = lane; k < n; k += 32) m = fmaxf(m, sr[k]);
    m = warp_max(m);
    float sum = 0.f;
    for (int k = lane; k < n; k += 32) {
      const float e = __expf(sr[k] - m);
      sr[k] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    for (int k = lane; k < n; k += 32) sr[k] *= inv;
  }
  __syncthreads();
  // phase 3: O = P V ; thread owns output dim(s) dd, all kQT queries
  for (int dd = tid; dd < hd; dd += kAttnThreads) {
    float acc[kQT];
#pragma unroll
    for (int qi = 0; qi < kQT; ++qi) acc[qi] = 0.f;
    const T* vp = base + 2ll * heads * hd + dd;
    for (int k = 0; k < n; ++k) {
      const float v = static_cast<float>(vp[static_cast<long long>(k) * row_stride]);
#pragma unroll
      for (int qi = 0; qi < kQT; ++qi) acc[qi] = fmaf(s_s[qi * n_pad + k], v, acc[qi]);
    }
#pragma unroll
    for (int qi = 0; qi < kQT; ++qi) {
      const int qr = q0 + qi;
      if (qr < n) out[(static_cast<long long>(b) * n + qr) * (heads * hd) + h * hd + dd] = static_cast<T>(acc[qi]);
    }
  }
}

// ------------------------------------------------------------------------- attention as two batched GEMMs
// (mspi_conv_gemm with batched weights computes scores = Q K^T * scale and out = P V on the tensor cores; these two
//  kernels are the glue: the row softmax and the K-major copy of V the second GEMM needs.)
// One warp per score row: max, exp / sum, normalise; fp32 in place.
__global__ void softmax_rows_kernel(float* __restrict__ s, long long rows, int n, long long stride) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (long long row = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * warps) {
    float* r = s + row * stride;
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, r[i]);
    m = warp_max(m);
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float e = __expf(r[i] - m);
      r[i] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    for (int i = lane; i < n; i += 32) r[i] *= inv;
  }
}

// qkv [B][N][3][H][HD] (v = index 2) -> vt [B][H][HD][n_pad] (keys contiguous), 32x32 tiles through shared memory.
__global__ void transpose_v_kernel(const float* __restrict__ qkv, float* __restrict__ vt, int n, int heads, int hd, int n_pad) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float tile[32][33];
  const int bh = blockIdx.z, b = bh / heads, h = bh % heads;
  const int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const long long row_stride = 3ll * heads * hd;
  const float* src = qkv + static_cast<long long>(b) * n * row_stride + 2ll * heads * hd + h * hd;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, dd = d0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < n && dd < hd) ? src[static_cast<long long>(k) * row_stride + dd] : 0.f;
  }
  __syncthreads();
  float* dst = vt + (static_cast<long long>(bh) * hd) * n_pad;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int dd = d0 + r, k = k0 + threadIdx.x;
    if (dd < hd && k < n_pad) dst[static_cast<long long>(dd) * n_pad + k] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------- token mean
template <typename T>
__global__ void token_mean_kernel(const T* __restrict__ x, float* __restrict__ y, int rows, int r0, int r1, int c) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int b = blockIdx.y;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const T* p = x + (static_cast<long long>(b) * rows + r0) * c + ch;
  float s = 0.f;
  for (int r = r0; r < r1; ++r, p += c) s += static_cast<float>(*p);
  y[static_cast<long long>(b) * c + ch] = s / static_cast<float>(r1 - r0);
}

// ------------------------------------------------------------------------- SimSiam loss
// grid (B, 2): one block per (sample, pair); out must be zero on entry (the host wrapper clears it)
__global__ void simsiam_kernel(const float* __restrict__ pv, const float* __restrict__ za, const float* __restrict__ pa,
                               const float* __restrict__ zv, float* __restrict__ out, int bsz, int c) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float red[32];
  const int b = blockIdx.x, pair = blockIdx.y;
  const float* p = (pair == 0 ? pv : pa) + static_cast<long long>(b) * c;
  const float* z = (pair == 0 ? za : zv) + static_cast<long long>(b) * c;
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
  // F.cosine_similarity: x.y / sqrt(max(|x|^2 |y|^2, eps^2)), eps = 1e-8
  if (threadIdx.x == 0) atomicAdd(out, -0.5f * dot * rsqrtf(fmaxf(pp * zz, 1e-16f)) / static_cast<float>(bsz));
}

}  // namespace
}  // namespace mspi

namespace mspi {
int dwconv_fast_path(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias, const float* ln_w,
                     const float* ln_b, void* y, cudaStream_t stream);  // dwconv.cu
}
using namespace mspi;

extern "C" int mspi_dwconv_ln(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias,
                              const float* ln_w, const float* ln_b, void* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && wgt && bias && y, "mspi_dwconv_ln: null argument");
  MSPI_CHECK_ARG(d->c % 2 == 0 && d->c <= 1024, "channels %d unsupported", d->c);
  MSPI_CHECK_ARG((d->kt & 1) && (d->kh & 1) && (d->kw & 1), "kernel extents must be odd");
  MSPI_CHECK_ARG((ln_w == nullptr) == (ln_b == nullptr), "ln_w / ln_b must both be given or both null");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  {
    const int rc = dwconv_fast_path(d, x, wgt, bias, ln_w, ln_b, y, stream);
    if (rc <= 0) return rc;  // handled (or failed) by a specialised kernel; 1 = not covered, use the generic one
  }
  const long long pixels = static_cast<long long>(d->n) * d->t * d->h * d->w;
  const int threads = 256, warps = threads / 32;
  long long blocks = (pixels + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  const int pairs = d->c / 2;
  const int g = static_cast<int>(blocks);
#define MSPI_DW_LAUNCH(MAXP)                                                                                         \
  do {                                                                                                                 \
    if (d->in_dtype == MSPI_BF16)                                                                                      \
      MSPI_CUDA(launch_pdl(dwconv_ln_kernel<MAXP, __nv_bfloat16>, g, threads, 0, stream, *d, static_cast<const __nv_bfloat16*>(x), wgt,  \
                                                                       bias, ln_w, ln_b, y, pixels));                   \
    else                                                                                                               \
      MSPI_CUDA(launch_pdl(dwconv_ln_kernel<MAXP, float>, g, threads, 0, stream, *d, static_cast<const float*>(x), wgt, bias, ln_w,      \
                                                               ln_b, y, pixels));                                       \
  } while (0)
  if (pairs <= 96) MSPI_DW_LAUNCH(3);
  else if (pairs <= 192) MSPI_DW_LAUNCH(6);
  else if (pairs <= 384) MSPI_DW_LAUNCH(12);
  else MSPI_DW_LAUNCH(16);
#undef MSPI_DW_LAUNCH
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_layernorm(const MspiLnDesc* d, const void* x, const float* w, const float* b, const float* pos,
                              void* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && w && b && y, "mspi_layernorm: null argument");
  MSPI_CHECK_ARG(d->pos_rows == 0 || pos, "pos table missing");
  MSPI_CHECK_ARG(d->rows_per_group > 0, "rows_per_group");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int threads = 256, warps = threads / 32;
  long long blocks = (d->rows + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  using bf = __nv_bfloat16;
  const int g = static_cast<int>(blocks);
  // vector kernel: G rows per warp iteration (8 / 4 / 2 / 1 for C <= 128 / 256 / 512 / 1024)
  const int rows_per_iter = d->c <= 128 ? 8 : (d->c <= 256 ? 4 : (d->c <= 512 ? 2 : 1));
  long long vblocks = (d->rows + static_cast<long long>(warps) * rows_per_iter - 1) / (static_cast<long long>(warps) * rows_per_iter);
  const long long vcap = static_cast<long long>(num_sms()) * 16;
  if (vblocks > vcap) vblocks = vcap;
  if (vblocks < 1) vblocks = 1;
  const int gv = static_cast<int>(vblocks);
  const int ies = d->in_dtype == MSPI_BF16 ? 2 : 4, oes = d->out_dtype == MSPI_BF16 ? 2 : 4;
  const bool vec_ok = d->c % 4 == 0 && d->c <= 1024 && (d->in_rstride * ies) % (4 * ies) == 0 &&
                      (d->out_rstride * oes) % (4 * oes) == 0 && (d->out_gstride * oes) % (4 * oes) == 0 &&
                      (reinterpret_cast<uintptr_t>(x) % (4 * ies)) == 0 && (reinterpret_cast<uintptr_t>(y) % (4 * oes)) == 0 &&
                      (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
                      (pos == nullptr || (reinterpret_cast<uintptr_t>(pos) & 15) == 0);
  const bool rows8_ok = d->c % 8 == 0 && d->c <= 1024 && d->rows < (1ll << 30) && d->in_rstride < (1ll << 31) && d->out_rstride < (1ll << 31) &&
                        d->rows * d->in_rstride < (1ll << 62) && d->in_rstride % 8 == 0 && d->out_rstride % 8 == 0 && d->out_gstride % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(b) & 15) == 0 && (pos == nullptr || (reinterpret_cast<uintptr_t>(pos) & 15) == 0);
  static const bool rows8_on = [] { const char* e = getenv("MSPI_LN_ROWS8"); return !e || atoi(e) != 0; }();
  if (rows8_ok && rows8_on) {
#define MSPI_LN_R8(TI, TO, LPR, NV, G)                                                                                   \
  do {                                                                                                                   \
    const long long per = static_cast<long long>(warps) * (32 / LPR) * G;                                                \
    long long nb = (d->rows + per - 1) / per;                                                                            \
    if (nb > vcap) nb = vcap;                                                                                            \
    if (nb < 1) nb = 1;                                                                                                  \
    MSPI_CUDA(launch_pdl(layernorm_rows8_kernel<TI, TO, LPR, NV, G>, static_cast<int>(nb), threads, 0, stream,                             \
        *d, static_cast<const TI*>(x), w, b, pos, static_cast<TO*>(y)));                                                  \
  } while (0)
#define MSPI_LN_R8_C(TI, TO)                                                                                             \
  do {                                                                                                                   \
    if (d->c <= 64) MSPI_LN_R8(TI, TO, 8, 1, 4);                                                                         \
    else if (d->c <= 128) MSPI_LN_R8(TI, TO, 16, 1, 4);                                                                  \
    else if (d->c <= 256) MSPI_LN_R8(TI, TO, 32, 1, 4);                                                                  \
    else if (d->c <= 512) MSPI_LN_R8(TI, TO, 32, 2, 2);                                                                  \
    else if (d->c <= 768) MSPI_LN_R8(TI, TO, 32, 3, 2);                                                                  \
    else MSPI_LN_R8(TI, TO, 32, 4, 1);                                                                                   \
  } while (0)
    if (d->in_dtype == MSPI_BF16 && d->out_dtype == MSPI_BF16) MSPI_LN_R8_C(bf, bf);
    else if (d->in_dtype == MSPI_BF16) MSPI_LN_R8_C(bf, float);
    else if (d->out_dtype == MSPI_BF16) MSPI_LN_R8_C(float, bf);
    else MSPI_LN_R8_C(float, float);
#undef MSPI_LN_R8_C
#undef MSPI_LN_R8
    MSPI_LAUNCH_CHECK();
    return MSPI_OK;
  }
  if (vec_ok) {
#define MSPI_LN_VEC(TI, TO)                                                                                              \
  do {                                                                                                                   \
    if (d->c <= 128)                                                                                                     \
      MSPI_CUDA(launch_pdl(layernorm_vec_kernel<TI, TO, 1, 8>, gv, threads, 0, stream, *d, static_cast<const TI*>(x), w, b, pos, static_cast<TO*>(y))); \
    else if (d->c <= 256)                                                                                                \
      MSPI_CUDA(launch_pdl(layernorm_vec_kernel<TI, TO, 2, 4>, gv, threads, 0, stream, *d, static_cast<const TI*>(x), w, b, pos, static_cast<TO*>(y))); \
    else if (d->c <= 512)                                                                                                \
      MSPI_CUDA(launch_pdl(layernorm_vec_kernel<TI, TO, 4, 2>, gv, threads, 0, stream, *d, static_cast<const TI*>(x), w, b, pos, static_cast<TO*>(y))); \
    else                                                                                                                 \
      MSPI_CUDA(launch_pdl(layernorm_vec_kernel<TI, TO, 8, 1>, gv, threads, 0, stream, *d, static_cast<const TI*>(x), w, b, pos, static_cast<TO*>(y))); \
  } while (0)
    if (d->in_dtype == MSPI_BF16 && d->out_dtype == MSPI_BF16) MSPI_LN_VEC(bf, bf);
    else if (d->in_dtype == MSPI_BF16) MSPI_LN_VEC(bf, float);
    else if (d->out_dtype == MSPI_BF16) MSPI_LN_VEC(float, bf);
    else MSPI_LN_VEC(float, float);
#undef MSPI_LN_VEC
    MSPI_LAUNCH_CHECK();
    return MSPI_OK;
  }
  if (d->in_dtype == MSPI_BF16 && d->out_dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(layernorm_kernel<bf, bf>, g, threads, 0, stream, *d, static_cast<const bf*>(x), w, b, pos, static_cast<bf*>(y)));
  else if (d->in_dtype == MSPI_BF16 && d->out_dtype == MSPI_F32)
    MSPI_CUDA(launch_pdl(layernorm_kernel<bf, float>, g, threads, 0, stream, *d, static_cast<const bf*>(x), w, b, pos, static_cast<float*>(y)));
  else if (d->in_dtype == MSPI_F32 && d->out_dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(layernorm_kernel<float, bf>, g, threads, 0, stream, *d, static_cast<const float*>(x), w, b, pos, static_cast<bf*>(y)));
  else
    MSPI_CUDA(launch_pdl(layernorm_kernel<float, float>, g, threads, 0, stream, *d, static_cast<const float*>(x), w, b, pos, static_cast<float*>(y)));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_attention(const void* qkv, void* out, int dtype, int b, int n, int heads, int hd, float scale,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(qkv && out && b > 0 && n > 0 && heads > 0, "mspi_attention: bad argument");
  MSPI_CHECK_ARG(hd % 8 == 0 && n <= 4096, "hd %d / n %d unsupported", hd, n);
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int n_pad = n + 1;  // odd-ish stride: rows of the score tile land in different banks
  const size_t smem = static_cast<size_t>(kQT) * (hd + n_pad) * sizeof(float);
  MSPI_CHECK_ARG(smem <= 100 * 1024, "attention tile needs %zu bytes of shared memory", smem);
  static bool attr = false;
  if (!attr) {
    MSPI_CUDA(cudaFuncSetAttribute(attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    MSPI_CUDA(cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  dim3 grid(b * heads, (n + kQT - 1) / kQT);
  if (dtype == MSPI_BF16)
    attention_kernel<__nv_bfloat16><<<grid, kAttnThreads, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), n, heads, hd, scale, n_pad);
  else
    attention_kernel<float><<<grid, kAttnThreads, smem, stream>>>(static_cast<const float*>(qkv),
                                                                  static_cast<float*>(out), n, heads, hd, scale, n_pad);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_softmax_rows(float* s, int64_t rows, int n, int64_t stride, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(s && rows > 0 && n > 0 && stride >= n, "mspi_softmax_rows: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int threads = 256, warps = threads / 32;
  long long blocks = (rows + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  MSPI_CUDA(launch_pdl(softmax_rows_kernel, static_cast<int>(blocks), threads, 0, stream, s, rows, n, stride));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_transpose_v(const float* qkv, float* vt, int b, int n, int heads, int hd, int n_pad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(qkv && vt && b > 0 && n > 0 && heads > 0 && hd > 0 && n_pad >= n, "mspi_transpose_v: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  dim3 grid((n_pad + 31) / 32, (hd + 31) / 32, b * heads), block(32, 8);
  MSPI_CUDA(launch_pdl(transpose_v_kernel, grid, block, 0, stream, qkv, vt, n, heads, hd, n_pad));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_token_mean(const void* x, int x_dtype, float* y, int b, int rows, int r0, int r1, int c,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && y && b > 0 && 0 <= r0 && r0 < r1 && r1 <= rows && c > 0, "mspi_token_mean: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  dim3 grid((c + 127) / 128, b);
  if (x_dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(token_mean_kernel<__nv_bfloat16>, grid, 128, 0, stream, static_cast<const __nv_bfloat16*>(x), y, rows, r0, r1, c));
  else
    MSPI_CUDA(launch_pdl(token_mean_kernel<float>, grid, 128, 0, stream, static_cast<const float*>(x), y, rows, r0, r1, c));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_simsiam_loss(const float* p_v, const float* z_a, const float* p_a, const float* z_v, float* out,
                                 int b, int c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(p_v && z_a && p_a && z_v && out && b > 0 && c > 0, "mspi_simsiam_loss: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  MSPI_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  MSPI_CUDA(launch_pdl(simsiam_kernel, dim3(b, 2), 256, 0, stream, p_v, z_a, p_a, z_v, out, b, c));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}
