// Layout changes, patch gather, max-pool, bilinear upsample and SA gating — HBM-bound kernels,
// 16-byte vectorised along the contiguous channel dimension of the NDHWC layout.
#include "common.cuh"

namespace mspi {
namespace {

constexpr int kBlock = 256;
constexpr int kGateMaxW = 1024;   // widest row the scale-ladder gate kernel stages gates for

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[VEC]) {
  if constexpr (sizeof(T) == 2) {
    static_assert(VEC == 8, "bf16 vectors are 8 wide");
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    unpack_bf16x2(u.x, v[0], v[1]); unpack_bf16x2(u.y, v[2], v[3]);
    unpack_bf16x2(u.z, v[4], v[5]); unpack_bf16x2(u.w, v[6], v[7]);
  } else {
#pragma unroll
    for (int q = 0; q < VEC / 4; ++q) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(p) + q);
      v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
    }
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[VEC]) {
  if constexpr (sizeof(T) == 2) {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = o;
  } else {
#pragma unroll
    for (int q = 0; q < VEC / 4; ++q)
      reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
}

// ------------------------------------------------------------------------- patch gather
// bf16 NDHWC source, C % 8 == 0: one thread moves 8 channels of one tap (16 B).
__global__ void patch_gather_nhwc_kernel(MspiPatchDesc d, const __nv_bfloat16* __restrict__ src,
                                         __nv_bfloat16* __restrict__ dst, long long total_chunks, int chunks_per_row,
                                         int c8) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / chunks_per_row;
    const int ch = static_cast<int>(i - m * chunks_per_row);
    uint4 v = make_uint4(0, 0, 0, 0);
    const int tap = ch / c8, cc = ch - tap * c8;
    if (tap < d.kt * d.kh * d.kw) {
      long long r = m;
      const int ow = static_cast<int>(r % d.ow); r /= d.ow;
      const int oh = static_cast<int>(r % d.oh); r /= d.oh;
      const int ot = static_cast<int>(r % d.ot); r /= d.ot;
      const int n = static_cast<int>(r);
      const int kw = tap % d.kw, kh = (tap / d.kw) % d.kh, kt = tap / (d.kw * d.kh);
      const int it = ot * d.st - d.pt + kt, ih = oh * d.sh - d.ph + kh, iw = ow * d.sw - d.pw + kw;
      if (it >= 0 && it < d.t && ih >= 0 && ih < d.h && iw >= 0 && iw < d.w) {
        const long long pix = ((static_cast<long long>(n) * d.t + it) * d.h + ih) * d.w + iw;
        v = ldg16(src + pix * d.src_cstride + cc * 8);
      }
    }
    *reinterpret_cast<uint4*>(dst + m * d.k_pad + static_cast<long long>(ch) * 8) = v;
  }
}

// Generic element-wise gather: fp32 NCDHW (src_layout 0) or bf16 NDHWC with any C.  One thread
// produces 8 consecutive K elements of one row (one 16 B store).
__global__ void patch_gather_generic_kernel(MspiPatchDesc d, const void* __restrict__ src,
                                            __nv_bfloat16* __restrict__ dst, long long total_chunks, int chunks_per_row) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int ktot = d.kt * d.kh * d.kw * d.c;
  const long long thw = static_cast<long long>(d.t) * d.h * d.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / chunks_per_row;
    const int ch = static_cast<int>(i - m * chunks_per_row);
    long long r = m;
    const int ow = static_cast<int>(r % d.ow); r /= d.ow;
    const int oh = static_cast<int>(r % d.oh); r /= d.oh;
    const int ot = static_cast<int>(r % d.ot); r /= d.ot;
    const int n = static_cast<int>(r);
    float vals[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ch * 8 + e;
      float v = 0.f;
      if (k < ktot) {
        const int c = k % d.c;
        int tap = k / d.c;
        const int kw = tap % d.kw; tap /= d.kw;
        const int kh = tap % d.kh;
        const int kt = tap / d.kh;
        const int it = ot * d.st - d.pt + kt, ih = oh * d.sh - d.ph + kh, iw = ow * d.sw - d.pw + kw;
        if (it >= 0 && it < d.t && ih >= 0 && ih < d.h && iw >= 0 && iw < d.w) {
          const long long sp = (static_cast<long long>(it) * d.h + ih) * d.w + iw;
          if (d.src_layout == 0)
            v = __ldg(static_cast<const float*>(src) + (static_cast<long long>(n) * d.c + c) * thw + sp);
          else
            v = bf2f(static_cast<const __nv_bfloat16*>(src)[(static_cast<long long>(n) * thw + sp) * d.src_cstride + c]);
        }
      }
      vals[e] = v;
    }
    uint4 o;
    o.x = pack_bf16x2(vals[0], vals[1]); o.y = pack_bf16x2(vals[2], vals[3]);
    o.z = pack_bf16x2(vals[4], vals[5]); o.w = pack_bf16x2(vals[6], vals[7]);
    *reinterpret_cast<uint4*>(dst + m * d.k_pad + static_cast<long long>(ch) * 8) = o;
  }
}

// ------------------------------------------------------------------------- layout changes
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n, int c,
                                      long long thw, long long cstride) {
  const long long total = static_cast<long long>(n) * thw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / thw, sp = i - b * thw;
    for (int ch = 0; ch < c; ++ch)
      dst[i * cstride + ch] = __float2bfloat16_rn(__ldg(src + (b * c + ch) * thw + sp));
  }
}

// fp32 NCDHW clip [N][3][T][H][W] -> bf16 zero-padded frames [N*T][Hp][Wp][4] (channel 3 = 0).  A thread converts 4
// consecutive pixels of a row: three coalesced float4 loads (one per colour plane), two 16-byte stores.  The borders
// of the destination are never written (the buffer is zeroed once when it is allocated).
struct ClipMap {  // destination frame j of clip b <- source frame map[j]; dst frame index = b*dst_fpc + dst_f0 + j
  int t_out, dst_fpc, dst_f0;
  int map[32];
};
__global__ void clip_to_padded_nhwc4_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n, int t,
                                            int h, int w, int pad_t, int pad_l, int hp, int wp, ClipMap m) {
  const int w4 = w >> 2;
  const long long total = static_cast<long long>(n) * m.t_out * h * w4;
  const long long plane = static_cast<long long>(t) * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xq = static_cast<int>(i % w4);
    long long r = i / w4;
    const int yy = static_cast<int>(r % h); r /= h;
    const int j = static_cast<int>(r % m.t_out);
    const long long b = r / m.t_out;
    const int tt = m.map[j];
    const float* sp = src + (b * 3 * t + tt) * static_cast<long long>(h) * w + static_cast<long long>(yy) * w + 4 * xq;
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(sp));
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(sp + plane));
    const float4 c2 = __ldg(reinterpret_cast<const float4*>(sp + 2 * plane));
    uint4 o0, o1;
    o0.x = pack_bf16x2(c0.x, c1.x); o0.y = pack_bf16x2(c2.x, 0.f);
    o0.z = pack_bf16x2(c0.y, c1.y); o0.w = pack_bf16x2(c2.y, 0.f);
    o1.x = pack_bf16x2(c0.z, c1.z); o1.y = pack_bf16x2(c2.z, 0.f);
    o1.z = pack_bf16x2(c0.w, c1.w); o1.w = pack_bf16x2(c2.w, 0.f);
    const long long f = b * m.dst_fpc + m.dst_f0 + j;
    __nv_bfloat16* dp = dst + ((f * hp + (yy + pad_t)) * static_cast<long long>(wp) + pad_l + 4 * xq) * 4;
    reinterpret_cast<uint4*>(dp)[0] = o0;
    reinterpret_cast<uint4*>(dp)[1] = o1;
  }
}

// uint8 frames [N][T][H][W][3] (what a decoder hands over; inference.py:154-165 turns them into fp32 on the host) -> the
// same bf16 zero-padded frames, with ToTensor's /255 and Normalize's (x - mean) / std done here in fp32 with IEEE division,
// i.e. bit for bit what converting the reference's normalised fp32 clip gives.  A thread converts 4 pixels: 12 bytes in.
struct NormParams { float mean[3], std[3]; };
__global__ void clip_u8_to_padded_nhwc4_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n, int t,
                                               int h, int w, int pad_t, int pad_l, int hp, int wp, ClipMap m, NormParams np) {
  const int w4 = w >> 2;
  const long long total = static_cast<long long>(n) * m.t_out * h * w4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xq = static_cast<int>(i % w4);
    long long r = i / w4;
    const int yy = static_cast<int>(r % h); r /= h;
    const int j = static_cast<int>(r % m.t_out);
    const long long b = r / m.t_out;
    const int tt = m.map[j];
    const uint8_t* sp = src + (((b * t + tt) * static_cast<long long>(h) + yy) * w + 4 * xq) * 3;   // 12 bytes, 4-byte aligned
    const uint32_t u0 = __ldg(reinterpret_cast<const uint32_t*>(sp));
    const uint32_t u1 = __ldg(reinterpret_cast<const uint32_t*>(sp) + 1);
    const uint32_t u2 = __ldg(reinterpret_cast<const uint32_t*>(sp) + 2);
    const uint32_t by[12] = {u0 & 255u, (u0 >> 8) & 255u, (u0 >> 16) & 255u, u0 >> 24, u1 & 255u, (u1 >> 8) & 255u,
                             (u1 >> 16) & 255u, u1 >> 24, u2 & 255u, (u2 >> 8) & 255u, (u2 >> 16) & 255u, u2 >> 24};
    float f[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int c = k % 3;
      f[k] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(by[k]), 255.f), np.mean[c]), np.std[c]);
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(f[0], f[1]); o0.y = pack_bf16x2(f[2], 0.f);
    o0.z = pack_bf16x2(f[3], f[4]); o0.w = pack_bf16x2(f[5], 0.f);
    o1.x = pack_bf16x2(f[6], f[7]); o1.y = pack_bf16x2(f[8], 0.f);
    o1.z = pack_bf16x2(f[9], f[10]); o1.w = pack_bf16x2(f[11], 0.f);
    const long long fr = b * m.dst_fpc + m.dst_f0 + j;
    __nv_bfloat16* dp = dst + ((fr * hp + (yy + pad_t)) * static_cast<long long>(wp) + pad_l + 4 * xq) * 4;
    reinterpret_cast<uint4*>(dp)[0] = o0;
    reinterpret_cast<uint4*>(dp)[1] = o1;
  }
}

// dst row i = src row idx[i] (rows of row_vecs 16-byte vectors): per-window views of per-frame feature maps.
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ idx, uint4* __restrict__ dst,
                                   long long n_rows, int row_vecs, int src_rows) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const long long total = n_rows * row_vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row_vecs;
    const int v = static_cast<int>(i - r * row_vecs);
    int s = __ldg(idx + r);
    s = s < 0 ? 0 : (s >= src_rows ? src_rows - 1 : s);   // clamped: an index outside the cache must not fault
    dst[i] = __ldg(src + static_cast<long long>(s) * row_vecs + v);
  }
}

// [N][THW][C] -> [N][C][THW] through a 32x33 shared tile.
template <typename T>
__global__ void ndhwc_to_ncdhw_kernel(const T* __restrict__ src, long long cstride, float* __restrict__ dst, int c,
                                      long long thw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long sp0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long sp = sp0 + r;
    const int ch = c0 + threadIdx.x;
    float v = 0.f;
    if (sp < thw && ch < c) v = static_cast<float>(src[(static_cast<long long>(n) * thw + sp) * cstride + ch]);
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ch = c0 + r;
    const long long sp = sp0 + threadIdx.x;
    if (sp < thw && ch < c) dst[(static_cast<long long>(n) * c + ch) * thw + sp] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------- max pool
__global__ void maxpool3d_kernel(MspiPoolDesc d, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                 long long total, int c8) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int cc = divmod(r, c8);
    const long long opix = r;
    const int ow = divmod(r, d.ow);
    const int oh = divmod(r, d.oh);
    const int ot = divmod(r, d.ot);
    const int n = static_cast<int>(r);
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
    __nv_bfloat162 m0 = ninf, m1 = ninf, m2 = ninf, m3 = ninf;
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
          const uint4 v = ldg16(x + pix * d.in_cstride + cc * 8);
          m0 = __hmax2(m0, *reinterpret_cast<const __nv_bfloat162*>(&v.x));
          m1 = __hmax2(m1, *reinterpret_cast<const __nv_bfloat162*>(&v.y));
          m2 = __hmax2(m2, *reinterpret_cast<const __nv_bfloat162*>(&v.z));
          m3 = __hmax2(m3, *reinterpret_cast<const __nv_bfloat162*>(&v.w));
        }
      }
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m0); o.y = *reinterpret_cast<uint32_t*>(&m1);
    o.z = *reinterpret_cast<uint32_t*>(&m2); o.w = *reinterpret_cast<uint32_t*>(&m3);
    *reinterpret_cast<uint4*>(y + opix * d.out_cstride + cc * 8) = o;
  }
}

// Fixed window sizes (the S3D / ResNet pools: (1,3,3), (3,3,3), (2,2,2), (1,2,2)): the window is unrolled and out-of-range taps
// are clamped into the window's valid part (a duplicate does not change a maximum), so all KT*KH*KW loads are independent
// and in flight together; the generic kernel's run-time loops with `continue` issue them one at a time.
template <int KT, int KH, int KW>
__global__ void maxpool3d_fixed_kernel(MspiPoolDesc d, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                       long long total, int c8) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int cc = divmod(r, c8);
    const long long opix = r;
    const int ow = divmod(r, d.ow);
    const int oh = divmod(r, d.oh);
    const int ot = divmod(r, d.ot);
    const int n = static_cast<int>(r);
    const int t0 = ot * d.st - d.pt, h0 = oh * d.sh - d.ph, w0 = ow * d.sw - d.pw;
    const __nv_bfloat16* xb = x + cc * 8;
    uint4 v[KT * KH * KW];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const int it = min(max(t0 + kt, 0), d.t - 1);
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) {
        const int ih = min(max(h0 + kh, 0), d.h - 1);
        const long long rowp = ((static_cast<long long>(n) * d.t + it) * d.h + ih) * d.w;
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
          const int iw = min(max(w0 + kw, 0), d.w - 1);
          v[(kt * KH + kh) * KW + kw] = ldg16(xb + (rowp + iw) * d.in_cstride);
        }
      }
    }
    __nv_bfloat162 m0 = *reinterpret_cast<const __nv_bfloat162*>(&v[0].x), m1 = *reinterpret_cast<const __nv_bfloat162*>(&v[0].y);
    __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&v[0].z), m3 = *reinterpret_cast<const __nv_bfloat162*>(&v[0].w);
#pragma unroll
    for (int k = 1; k < KT * KH * KW; ++k) {
      m0 = __hmax2(m0, *reinterpret_cast<const __nv_bfloat162*>(&v[k].x));
      m1 = __hmax2(m1, *reinterpret_cast<const __nv_bfloat162*>(&v[k].y));
      m2 = __hmax2(m2, *reinterpret_cast<const __nv_bfloat162*>(&v[k].z));
      m3 = __hmax2(m3, *reinterpret_cast<const __nv_bfloat162*>(&v[k].w));
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m0); o.y = *reinterpret_cast<uint32_t*>(&m1);
    o.z = *reinterpret_cast<uint32_t*>(&m2); o.w = *reinterpret_cast<uint32_t*>(&m3);
    *reinterpret_cast<uint4*>(y + opix * d.out_cstride + cc * 8) = o;
  }
}

// (3,3,3) / stride 1 / padding 1 (the Inception pooling branch, s3d.py:134): a thread owns 8 channels of one (n,t,h) row and
// walks along w keeping the maxima of the last three columns (each over its 3x3 (t,h) neighbourhood): 9 loads per output
// instead of 27.
__global__ void maxpool333_kernel(MspiPoolDesc d, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                  long long total, int c8) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int cc = divmod(r, c8);
    const long long row = r;                       // (n*T + t)*H + h
    const int hh = divmod(r, d.h);
    const int tt = divmod(r, d.t);
    // Border handling by clamping: a duplicated row / column does not change a maximum, so every column is always 9
    // unrolled loads from 9 row pointers computed once per thread, all in flight together (the loops with run-time
    // bounds this replaces issued one load at a time).
    const long long plane0 = (r * d.t) * d.h;      // first row of sample n
    const __nv_bfloat16* rp[9];
#pragma unroll
    for (int dt = 0; dt < 3; ++dt)
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int t2 = min(max(tt + dt - 1, 0), d.t - 1), h2 = min(max(hh + dh - 1, 0), d.h - 1);
        rp[dt * 3 + dh] = x + ((plane0 + static_cast<long long>(t2) * d.h + h2) * d.w) * d.in_cstride + cc * 8;
      }
    __nv_bfloat162 cm[3][4];                       // column maxima at w-1, w, w+1 (4 x bf16x2 = 8 channels)
    auto column = [&](int ww, __nv_bfloat162 (&m)[4]) {
      const long long off = static_cast<long long>(min(max(ww, 0), d.w - 1)) * d.in_cstride;
      uint4 v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = ldg16(rp[k] + off);
      m[0] = *reinterpret_cast<const __nv_bfloat162*>(&v[0].x);
      m[1] = *reinterpret_cast<const __nv_bfloat162*>(&v[0].y);
      m[2] = *reinterpret_cast<const __nv_bfloat162*>(&v[0].z);
      m[3] = *reinterpret_cast<const __nv_bfloat162*>(&v[0].w);
#pragma unroll
      for (int k = 1; k < 9; ++k) {
        m[0] = __hmax2(m[0], *reinterpret_cast<const __nv_bfloat162*>(&v[k].x));
        m[1] = __hmax2(m[1], *reinterpret_cast<const __nv_bfloat162*>(&v[k].y));
        m[2] = __hmax2(m[2], *reinterpret_cast<const __nv_bfloat162*>(&v[k].z));
        m[3] = __hmax2(m[3], *reinterpret_cast<const __nv_bfloat162*>(&v[k].w));
      }
    };
    column(-1, cm[0]);
    column(0, cm[1]);
    for (int ww = 0; ww < d.w; ++ww) {
      column(ww + 1, cm[2]);
      uint4 o;
      __nv_bfloat162 m;
      m = __hmax2(__hmax2(cm[0][0], cm[1][0]), cm[2][0]); o.x = *reinterpret_cast<uint32_t*>(&m);
      m = __hmax2(__hmax2(cm[0][1], cm[1][1]), cm[2][1]); o.y = *reinterpret_cast<uint32_t*>(&m);
      m = __hmax2(__hmax2(cm[0][2], cm[1][2]), cm[2][2]); o.z = *reinterpret_cast<uint32_t*>(&m);
      m = __hmax2(__hmax2(cm[0][3], cm[1][3]), cm[2][3]); o.w = *reinterpret_cast<uint32_t*>(&m);
      *reinterpret_cast<uint4*>(y + (row * d.w + ww) * d.out_cstride + cc * 8) = o;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        cm[0][q] = cm[1][q];
        cm[1][q] = cm[2][q];
      }
    }
  }
}

// Same pooling from a shared-memory tile (the default): a block owns one sample, a strip of `hs` output rows, 16 channels
// (one 32-byte sector per pixel) and ALL T <= 8 frames.  The (hs + 2) x W x T input tile is fetched once with cp.async — every
// load of the block is in flight at the same time — and each thread then reduces 9 neighbours per frame from shared memory
// and slides the 3-frame maximum over t in registers.  The sliding-window kernel above walks a row serially (9 loads, wait,
// reduce, next column): ncu showed 14-18 long-scoreboard stalls per issue at 22 % occupancy, 1.1-2.1 TB/s, and its L2 -> L1
// traffic is 9x the tensor (profiles/r02_small_kernels.md).  Here the tensor is read 1 + 2 / hs times.
constexpr int kPoolTileMaxT = 8;
constexpr int kPoolTileSmem = 74 * 1024;
__global__ void __launch_bounds__(256) maxpool333_tile_kernel(MspiPoolDesc d, const __nv_bfloat16* __restrict__ x,
                                                              __nv_bfloat16* __restrict__ y, int hs, int strips, int cgroups,
                                                              int lanes) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  extern __shared__ uint4 pool_tile[];   // [T][hs + 2][W][2]
  const int T = d.t, H = d.h, W = d.w, R = hs + 2;
  int r = blockIdx.x;
  const int cgp = r % cgroups;
  r /= cgroups;
  const int strip = r % strips, n = r / strips;
  const int h0 = strip * hs, rows = min(hs, H - h0);
  const int xw = threadIdx.x % (2 * W), ry = threadIdx.x / (2 * W);   // (column, 8-channel half) and row lane of this thread
  const int w = xw >> 1, v = xw & 1;
  const long long coff = static_cast<long long>(cgp * 2 + v) * 8;
  if (ry < lanes) {
    for (int q = ry; q < T * R; q += lanes) {
      const int t = q / R, rr = q - t * R;
      const int hsrc = min(max(h0 - 1 + rr, 0), H - 1);   // a duplicated row does not change a maximum
      const __nv_bfloat16* src = x + ((static_cast<long long>(n * T + t) * H + hsrc) * W + w) * d.in_cstride + coff;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(&pool_tile[(q * W + w) * 2 + v]));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (ry >= lanes) return;
  const int wl = max(w - 1, 0), wr = min(w + 1, W - 1);
  for (int hl = ry; hl < rows; hl += lanes) {
    __nv_bfloat162 m[kPoolTileMaxT][4];
#pragma unroll
    for (int t = 0; t < kPoolTileMaxT; ++t) {
      if (t < T) {
        const uint4* base = pool_tile + static_cast<size_t>(t * R + hl) * W * 2 + v;
        uint4 a = base[wl * 2];
        m[t][0] = *reinterpret_cast<const __nv_bfloat162*>(&a.x);
        m[t][1] = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
        m[t][2] = *reinterpret_cast<const __nv_bfloat162*>(&a.z);
        m[t][3] = *reinterpret_cast<const __nv_bfloat162*>(&a.w);
#pragma unroll
        for (int k = 1; k < 9; ++k) {
          const int kh = k / 3, kw = k - 3 * kh;
          const uint4 b = base[(kh * W + (kw == 0 ? wl : (kw == 1 ? w : wr))) * 2];
          m[t][0] = __hmax2(m[t][0], *reinterpret_cast<const __nv_bfloat162*>(&b.x));
          m[t][1] = __hmax2(m[t][1], *reinterpret_cast<const __nv_bfloat162*>(&b.y));
          m[t][2] = __hmax2(m[t][2], *reinterpret_cast<const __nv_bfloat162*>(&b.z));
          m[t][3] = __hmax2(m[t][3], *reinterpret_cast<const __nv_bfloat162*>(&b.w));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < kPoolTileMaxT; ++t) {
      if (t < T) {
        constexpr int kNext = kPoolTileMaxT - 1;
        const int ta = t > 0 ? t - 1 : 0, tn = t < kNext ? t + 1 : t;   // compile-time indices (t is unrolled)
        const bool has_next = t + 1 < T;
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162 nx = has_next ? m[tn][j] : m[t][j];
          const __nv_bfloat162 q = __hmax2(__hmax2(m[ta][j], m[t][j]), nx);
          ow[j] = *reinterpret_cast<const uint32_t*>(&q);
        }
        const uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        *reinterpret_cast<uint4*>(y + ((static_cast<long long>(n * T + t) * H + h0 + hl) * W + w) * d.out_cstride + coff) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------- bilinear upsample
// align_corners=False, integer scale k: src = (dst + 0.5)/k - 0.5 clamped at 0 (PyTorch's
// area_pixel_compute_source_index), second index clamped at size-1.
template <typename TI, typename TO, int VEC>
__global__ void upsample_kernel(MspiUpDesc d, const TI* __restrict__ x, TO* __restrict__ y, long long total, int cv) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int oh_ = d.h * d.k, ow_ = d.w * d.k;
  const float inv = 1.f / static_cast<float>(d.k);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int cc = divmod(r, cv);
    const long long opix = r;
    const int ox = divmod(r, ow_);
    const int oy = divmod(r, oh_);
    const long long plane = r;
    float sy = fmaxf((oy + 0.5f) * inv - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * inv - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = min(y0 + 1, d.h - 1), x1 = min(x0 + 1, d.w - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const long long base = plane * d.h * d.w;
    float a[VEC], b[VEC], c[VEC], e[VEC], v[VEC];
    load_vec<TI, VEC>(x + (base + static_cast<long long>(y0) * d.w + x0) * d.in_cstride + cc * VEC, a);
    load_vec<TI, VEC>(x + (base + static_cast<long long>(y0) * d.w + x1) * d.in_cstride + cc * VEC, b);
    load_vec<TI, VEC>(x + (base + static_cast<long long>(y1) * d.w + x0) * d.in_cstride + cc * VEC, c);
    load_vec<TI, VEC>(x + (base + static_cast<long long>(y1) * d.w + x1) * d.in_cstride + cc * VEC, e);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = w00 * a[j] + w01 * b[j] + w10 * c[j] + w11 * e[j];
    TO* yp = y + opix * d.out_cstride + cc * VEC;
    if (d.accumulate) {
      float o[VEC];
      load_vec<TO, VEC>(yp, o);
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] += o[j];
    }
    if (d.act == MSPI_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    store_vec<TO, VEC>(yp, v);
  }
}

// ------------------------------------------------------------------------- SA gate, add
template <typename T>
__global__ void sa_gate_kernel(const T* __restrict__ x, long long xcs, const float* __restrict__ m, T* __restrict__ y,
                               long long ycs, long long total, int c8) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pix = i;
    const int cc = divmod(pix, c8);
    const float g = 1.f + 1.f / (1.f + expf(-__ldg(m + pix)));  // x*mask + x
    float f[8];
    load_vec<T, 8>(x + pix * xcs + cc * 8, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= g;
    store_vec<T, 8>(y + pix * ycs + cc * 8, f);
  }
}

// SA gate + the top-down sums in one pass (model_utils.py:167-170,566-568):
//   y = x * (1 + sigmoid(l)) + sum_i up_{k_i}(src_i),  fp32, up = bilinear (1,k,k), align_corners=False.
// The unfused form writes y once and then reads + rewrites it for every upsampled term (4.2 GB for s0 at B=32); here y is
// written once and the (4x .. 64x smaller) sources are gathered from L2.
struct GateSrc {
  const float* p[3];
  long long cs[3];
  int k[3];
  int sh[3], sw[3];   // source height / width = h / k, w / k (computed on the host: no division in the kernel)
  float inv[3];       // 1 / k
  int n;
};
// One output ROW (plane, oy) per block iteration: the row decomposition and everything that depends on oy only (source rows,
// vertical weights, row base pointers) is block-uniform and computed once per row; a thread only splits its element index
// into (ox, channel group) with one multiply-high.  The first version decomposed a linear index per element (three
// divisions, six more for the source sizes): ~700 instructions per 32 output bytes, issue bound at 2.5 TB/s.
template <int NS>   // number of top-down sources (compile time: the loads of all sources are issued before the first blend)
__global__ void __launch_bounds__(256)
sa_gate_fused_kernel(const float* x, long long xcs, const float* __restrict__ m, float* y, long long ycs, int rows, int c8,
                     int h, int w, unsigned c8_magic, GateSrc s) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int per_row = w * c8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int plane = row / h, oy = row - plane * h;
    const float* r0[NS > 0 ? NS : 1];
    const float* r1[NS > 0 ? NS : 1];
    float ly[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const float sy = fmaxf((oy + 0.5f) * s.inv[j] - 0.5f, 0.f);
      const int y0 = static_cast<int>(sy), y1 = min(y0 + 1, s.sh[j] - 1);
      ly[j] = sy - y0;
      const long long base = static_cast<long long>(plane) * s.sh[j];
      r0[j] = s.p[j] + (base + y0) * s.sw[j] * s.cs[j];
      r1[j] = s.p[j] + (base + y1) * s.sw[j] * s.cs[j];
    }
    const long long pix0 = static_cast<long long>(row) * w;
    for (int e = threadIdx.x; e < per_row; e += blockDim.x) {
      const int ox = static_cast<int>(__umulhi(static_cast<unsigned>(e), c8_magic));
      const int cc = (e - ox * c8) * 8;
      const long long pix = pix0 + ox;
      const float g = m != nullptr ? 1.f + 1.f / (1.f + expf(-__ldg(m + pix))) : 1.f;   // no mask: plain y = x + sum up(src)
      float f[8];
      {
        const float4* xp = reinterpret_cast<const float4*>(x + pix * xcs + cc);   // x may alias y (in place): no __ldg
        const float4 a0 = xp[0], a1 = xp[1];
        f[0] = a0.x; f[1] = a0.y; f[2] = a0.z; f[3] = a0.w; f[4] = a1.x; f[5] = a1.y; f[6] = a1.z; f[7] = a1.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] *= g;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const float sx = fmaxf((ox + 0.5f) * s.inv[j] - 0.5f, 0.f);
        const int x0 = static_cast<int>(sx), x1 = min(x0 + 1, s.sw[j] - 1);
        const float lx = sx - x0;
        const float w00 = (1.f - ly[j]) * (1.f - lx), w01 = (1.f - ly[j]) * lx, w10 = ly[j] * (1.f - lx), w11 = ly[j] * lx;
        const long long o0 = x0 * s.cs[j] + cc, o1 = x1 * s.cs[j] + cc;
        float a[8], b[8], c[8], e2[8];
        load_vec<float, 8>(r0[j] + o0, a);
        load_vec<float, 8>(r0[j] + o1, b);
        load_vec<float, 8>(r1[j] + o0, c);
        load_vec<float, 8>(r1[j] + o1, e2);
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] += w00 * a[q] + w01 * b[q] + w10 * c[q] + w11 * e2[q];
      }
      store_vec<float, 8>(y + pix * ycs + cc, f);
    }
  }
}

// Same operation for the decoder's scale ladder (sources at 1/2, 1/4, 1/8 of the output: model_utils.py:566-568 and the
// commuted readout.0), four output pixels x four channels per thread.  The row kernel above is bound by L1 bandwidth (ncu:
// l1tex 90 %, issue 29 %): 24 gathered float4 per 32 output bytes.  Adjacent output pixels blend the SAME source pixels, and
// for a power-of-two scale the horizontal phase of a 4-aligned pixel group is fixed, so a thread loads each source column of
// its group once (4 + 3 + 2 columns x 2 rows for the three levels: 18 float4 per 64 output bytes, 2.4x less L1 traffic),
// blends the two rows, then blends columns with compile-time weights.  align_corners=False with its clamps is the same as
// replicating the border column, which is what clamping the column INDEX does.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) {   // (1 - t) a + t b
  const float u = 1.f - t;
  return make_float4(fmaf(t, b.x, u * a.x), fmaf(t, b.y, u * a.y), fmaf(t, b.z, u * a.z), fmaf(t, b.w, u * a.w));
}
__device__ __forceinline__ void acc_lerp4(float4& f, float4 a, float4 b, float t) {
  const float4 v = lerp4(a, b, t);
  f.x += v.x; f.y += v.y; f.z += v.z; f.w += v.w;
}
template <int K>
__device__ __forceinline__ void topdown_add4(float4 (&f)[4], const float* r0, const float* r1, float ly, long long cs, int sw,
                                             int q) {
  auto col = [&](int c) { return static_cast<long long>(min(max(c, 0), sw - 1)) * cs; };
  if constexpr (K == 2) {
    float4 v[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const long long o = col(2 * q - 1 + d);
      v[d] = lerp4(ld4(r0 + o), ld4(r1 + o), ly);
    }
    acc_lerp4(f[0], v[0], v[1], 0.75f);
    acc_lerp4(f[1], v[1], v[2], 0.25f);
    acc_lerp4(f[2], v[1], v[2], 0.75f);
    acc_lerp4(f[3], v[2], v[3], 0.25f);
  } else if constexpr (K == 4) {
    float4 v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const long long o = col(q - 1 + d);
      v[d] = lerp4(ld4(r0 + o), ld4(r1 + o), ly);
    }
    acc_lerp4(f[0], v[0], v[1], 0.625f);
    acc_lerp4(f[1], v[0], v[1], 0.875f);
    acc_lerp4(f[2], v[1], v[2], 0.125f);
    acc_lerp4(f[3], v[1], v[2], 0.375f);
  } else {
    static_assert(K == 8, "scale ladder 2, 4, 8");
    const int par = q & 1, r = q >> 1;
    const long long o0 = col(r - 1 + par), o1 = col(r + par);
    const float4 v0 = lerp4(ld4(r0 + o0), ld4(r1 + o0), ly), v1 = lerp4(ld4(r0 + o1), ld4(r1 + o1), ly);
    const float l0 = par ? 0.0625f : 0.5625f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc_lerp4(f[i], v0, v1, l0 + 0.125f * i);
  }
}
template <int NS>
__global__ void __launch_bounds__(256)
sa_gate_ladder_kernel(const float* x, long long xcs, const float* __restrict__ m, float* y, long long ycs, int rows, int c4,
                      int h, int w, unsigned c4_magic, GateSrc s) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int per_row = (w >> 2) * c4;
  __shared__ float s_gate[kGateMaxW];   // 1 + sigmoid(mask) of the row's pixels: once per pixel, not once per channel quad
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int plane = row / h, oy = row - plane * h;
    if (m != nullptr) {
      __syncthreads();   // the previous row's gates have been read
      for (int i = threadIdx.x; i < w; i += blockDim.x)
        s_gate[i] = 1.f + 1.f / (1.f + expf(-__ldg(m + static_cast<long long>(row) * w + i)));
      __syncthreads();
    }
    const float* r0[NS > 0 ? NS : 1];
    const float* r1[NS > 0 ? NS : 1];
    float ly[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const float sy = fmaxf((oy + 0.5f) * s.inv[j] - 0.5f, 0.f);
      const int y0 = static_cast<int>(sy), y1 = min(y0 + 1, s.sh[j] - 1);
      ly[j] = sy - y0;
      const long long base = static_cast<long long>(plane) * s.sh[j];
      r0[j] = s.p[j] + (base + y0) * s.sw[j] * s.cs[j];
      r1[j] = s.p[j] + (base + y1) * s.sw[j] * s.cs[j];
    }
    const long long pix0 = static_cast<long long>(row) * w;
    for (int e = threadIdx.x; e < per_row; e += blockDim.x) {
      const int q = static_cast<int>(__umulhi(static_cast<unsigned>(e), c4_magic));
      const int cc = (e - q * c4) * 4;
      const long long pix = pix0 + 4 * q;
      float4 f[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[i] = *reinterpret_cast<const float4*>(x + (pix + i) * xcs + cc);   // x may alias y (in place): no __ldg
        if (m != nullptr) {
          const float g = s_gate[4 * q + i];
          f[i].x *= g; f[i].y *= g; f[i].z *= g; f[i].w *= g;
        }
      }
      if constexpr (NS > 0) topdown_add4<2>(f, r0[0] + cc, r1[0] + cc, ly[0], s.cs[0], s.sw[0], q);
      if constexpr (NS > 1) topdown_add4<4>(f, r0[1] + cc, r1[1] + cc, ly[1], s.cs[1], s.sw[1], q);
      if constexpr (NS > 2) topdown_add4<8>(f, r0[2] + cc, r1[2] + cc, ly[2], s.cs[2], s.sw[2], q);
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(y + (pix + i) * ycs + cc) = f[i];
    }
  }
}

__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ y, long long n8) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = ldg16(a + i * 8), v = ldg16(b + i * 8);
    float f[8], g[8];
    unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
    unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
    unpack_bf16x2(v.x, g[0], g[1]); unpack_bf16x2(v.y, g[2], g[3]);
    unpack_bf16x2(v.z, g[4], g[5]); unpack_bf16x2(v.w, g[6], g[7]);
    uint4 o;
    o.x = pack_bf16x2(f[0] + g[0], f[1] + g[1]); o.y = pack_bf16x2(f[2] + g[2], f[3] + g[3]);
    o.z = pack_bf16x2(f[4] + g[4], f[5] + g[5]); o.w = pack_bf16x2(f[6] + g[6], f[7] + g[7]);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  }
}

template <typename TI, typename TO>
__global__ void cast_rows_kernel(const TI* __restrict__ src, long long srs, long long sgs, TO* __restrict__ dst,
                                 long long drs, long long dgs, int rows, int c, long long total) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const long long rr = i / c;
    const int r = static_cast<int>(rr % rows);
    const long long g = rr / rows;
    dst[g * dgs + r * drs + ch] = static_cast<TO>(static_cast<float>(src[g * sgs + r * srs + ch]));
  }
}

inline int grid_for(long long work) {
  const long long blocks = (work + kBlock - 1) / kBlock;
  const long long cap = static_cast<long long>(num_sms()) * 16;  // a few waves of resident CTAs
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace mspi

using namespace mspi;

extern "C" int mspi_patch_gather(const MspiPatchDesc* d, const void* src, void* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && src && dst, "mspi_patch_gather: null argument");
  MSPI_CHECK_ARG(d->k_pad % 8 == 0 && d->k_pad >= d->kt * d->kh * d->kw * d->c, "k_pad %d", d->k_pad);
  MSPI_CHECK_ARG(d->src_layout == 0 || d->src_layout == 1, "src_layout");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const long long m = static_cast<long long>(d->n) * d->ot * d->oh * d->ow;
  const int cpr = d->k_pad / 8;
  const long long total = m * cpr;
  if (d->src_layout == 1 && d->c % 8 == 0 && d->src_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    MSPI_CUDA(launch_pdl(patch_gather_nhwc_kernel, grid_for(total), kBlock, 0, stream, 
        *d, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), total, cpr, d->c / 8));
  } else {
    MSPI_CUDA(launch_pdl(patch_gather_generic_kernel, grid_for(total), kBlock, 0, stream, *d, src, static_cast<__nv_bfloat16*>(dst),
                                                                        total, cpr));
  }
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_ncdhw_to_ndhwc(const float* src, void* dst, int n, int c, int thw, int64_t dst_cstride,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(src && dst && n > 0 && c > 0 && thw > 0 && dst_cstride >= c, "mspi_ncdhw_to_ndhwc: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  ncdhw_to_ndhwc_kernel<<<grid_for(static_cast<long long>(n) * thw), kBlock, 0, stream>>>(
      src, static_cast<__nv_bfloat16*>(dst), n, c, thw, dst_cstride);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

static int clip_to_padded(const float* src, void* dst, int n, int t, int h, int w, int pad_t, int pad_l, int hp, int wp,
                          const ClipMap& m, cudaStream_t stream) {
  MSPI_CHECK_ARG(src && dst && n > 0 && t > 0 && h > 0 && w > 0, "mspi_clip_to_padded_nhwc4: bad argument");
  MSPI_CHECK_ARG(w % 4 == 0 && pad_l % 2 == 0 && wp % 2 == 0 && hp >= h + pad_t && wp >= w + pad_l,
                 "mspi_clip_to_padded_nhwc4: w %% 4, even pad_l / wp and hp >= h+pad_t, wp >= w+pad_l required");
  MSPI_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "16-byte alignment");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const long long total = static_cast<long long>(n) * m.t_out * h * (w / 4);
  clip_to_padded_nhwc4_kernel<<<grid_for(total), kBlock, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), n, t, h, w,
                                                                      pad_t, pad_l, hp, wp, m);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_clip_to_padded_nhwc4(const float* src, void* dst, int n, int t, int h, int w, int pad_t, int pad_l,
                                         int hp, int wp, void* stream_) {
  MSPI_CHECK_ARG(t <= 32, "at most 32 frames per clip");
  ClipMap m;
  m.t_out = t; m.dst_fpc = t; m.dst_f0 = 0;
  for (int j = 0; j < 32; ++j) m.map[j] = j < t ? j : 0;
  return clip_to_padded(src, dst, n, t, h, w, pad_t, pad_l, hp, wp, m, static_cast<cudaStream_t>(stream_));
}

extern "C" int mspi_clip_frames_to_padded_nhwc4(const float* src, void* dst, int n, int t, int h, int w, int pad_t, int pad_l,
                                                int hp, int wp, const int32_t* frame_map, int t_out,
                                                int dst_frames_per_clip, int dst_frame0, void* stream_) {
  MSPI_CHECK_ARG(frame_map && t_out >= 1 && t_out <= 32 && dst_frame0 >= 0 && dst_frame0 + t_out <= dst_frames_per_clip,
                 "mspi_clip_frames_to_padded_nhwc4: bad frame map");
  ClipMap m;
  m.t_out = t_out; m.dst_fpc = dst_frames_per_clip; m.dst_f0 = dst_frame0;
  for (int j = 0; j < 32; ++j) {
    m.map[j] = j < t_out ? frame_map[j] : 0;
    MSPI_CHECK_ARG(m.map[j] >= 0 && m.map[j] < t, "frame_map[%d] = %d outside the clip", j, m.map[j]);
  }
  return clip_to_padded(src, dst, n, t, h, w, pad_t, pad_l, hp, wp, m, static_cast<cudaStream_t>(stream_));
}

extern "C" int mspi_clip_u8_to_padded_nhwc4(const void* src, void* dst, int n, int t, int h, int w, int pad_t, int pad_l,
                                            int hp, int wp, const int32_t* frame_map, int t_out, int dst_frames_per_clip,
                                            int dst_frame0, const float* mean3, const float* std3, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(src && dst && mean3 && std3 && n > 0 && t > 0 && t <= 32 && h > 0 && w > 0, "mspi_clip_u8_to_padded_nhwc4: bad argument");
  MSPI_CHECK_ARG(w % 4 == 0 && pad_l % 2 == 0 && wp % 2 == 0 && hp >= h + pad_t && wp >= w + pad_l,
                 "mspi_clip_u8_to_padded_nhwc4: w %% 4, even pad_l / wp and hp >= h+pad_t, wp >= w+pad_l required");
  MSPI_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "alignment");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  ClipMap m;
  if (frame_map == nullptr) {
    m.t_out = t; m.dst_fpc = t; m.dst_f0 = 0;
    for (int j = 0; j < 32; ++j) m.map[j] = j < t ? j : 0;
  } else {
    MSPI_CHECK_ARG(t_out >= 1 && t_out <= 32 && dst_frame0 >= 0 && dst_frame0 + t_out <= dst_frames_per_clip, "bad frame map");
    m.t_out = t_out; m.dst_fpc = dst_frames_per_clip; m.dst_f0 = dst_frame0;
    for (int j = 0; j < 32; ++j) {
      m.map[j] = j < t_out ? frame_map[j] : 0;
      MSPI_CHECK_ARG(m.map[j] >= 0 && m.map[j] < t, "frame_map[%d] = %d outside the clip", j, m.map[j]);
    }
  }
  NormParams np;
  for (int c = 0; c < 3; ++c) { np.mean[c] = mean3[c]; np.std[c] = std3[c]; }
  const long long total = static_cast<long long>(n) * m.t_out * h * (w / 4);
  clip_u8_to_padded_nhwc4_kernel<<<grid_for(total), kBlock, 0, stream>>>(static_cast<const uint8_t*>(src),
                                                                         static_cast<__nv_bfloat16*>(dst), n, t, h, w, pad_t,
                                                                         pad_l, hp, wp, m, np);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_gather_rows(const void* src, const int32_t* idx, void* dst, int64_t n_rows, int64_t row_bytes,
                                int64_t src_rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(src && idx && dst && n_rows > 0 && row_bytes > 0 && src_rows > 0, "mspi_gather_rows: bad argument");
  MSPI_CHECK_ARG(row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
                 "mspi_gather_rows: rows must be multiples of 16 bytes, 16-byte aligned");
  MSPI_CHECK_ARG(row_bytes / 16 < (1ll << 31) && src_rows < (1ll << 31), "mspi_gather_rows: row too long");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int row_vecs = static_cast<int>(row_bytes / 16);
  MSPI_CUDA(launch_pdl(gather_rows_kernel, grid_for(n_rows * row_vecs), kBlock, 0, stream, static_cast<const uint4*>(src), idx,
                                                                          static_cast<uint4*>(dst), n_rows, row_vecs,
                                                                          static_cast<int>(src_rows)));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_ndhwc_to_ncdhw(const void* src, int src_dtype, int64_t src_cstride, float* dst, int n, int c,
                                   int thw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(src && dst && n > 0 && c > 0 && thw > 0, "mspi_ndhwc_to_ncdhw: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  dim3 grid((thw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  if (src_dtype == MSPI_BF16)
    ndhwc_to_ncdhw_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(src),
                                                                     src_cstride, dst, c, thw);
  else
    ndhwc_to_ncdhw_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(src), src_cstride, dst, c, thw);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_maxpool3d(const MspiPoolDesc* d, const void* x, void* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && y, "mspi_maxpool3d: null argument");
  MSPI_CHECK_ARG(d->c % 8 == 0 && d->in_cstride % 8 == 0 && d->out_cstride % 8 == 0, "channels must be multiples of 8");
  MSPI_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "16-byte alignment");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int c8 = d->c / 8;
  const long long total = static_cast<long long>(d->n) * d->ot * d->oh * d->ow * c8;
  if (d->kt == 3 && d->kh == 3 && d->kw == 3 && d->st == 1 && d->sh == 1 && d->sw == 1 && d->pt == 1 && d->ph == 1 &&
      d->pw == 1 && d->ot == d->t && d->oh == d->h && d->ow == d->w) {
    // shared-memory tile kernel (MSPI_POOL_TILE, default on): T <= 8 frames, 16-channel groups, rows of up to 128 pixels
    static const bool tile_on = [] { const char* e = getenv("MSPI_POOL_TILE"); return !e || atoi(e) != 0; }();
    if (tile_on && d->t <= kPoolTileMaxT && c8 % 2 == 0 && 2 * d->w <= 256) {
      const size_t row_bytes = static_cast<size_t>(d->t) * d->w * 32;     // one tile row: all frames, 16 channels
      int hs = static_cast<int>(kPoolTileSmem / row_bytes) - 2;
      if (hs > d->h) hs = d->h;
      if (hs >= 2 || hs == d->h) {
        const int lanes = (256 / (2 * d->w)) < 1 ? 1 : 256 / (2 * d->w);
        const int threads = 2 * d->w * lanes;
        const int strips = (d->h + hs - 1) / hs, cgroups = c8 / 2;
        const size_t smem = row_bytes * (hs + 2);
        const long long blocks = static_cast<long long>(d->n) * strips * cgroups;
        MSPI_CHECK_ARG(blocks < (1ll << 31), "mspi_maxpool3d: too many tiles");
        static bool attr_set = false;
        if (!attr_set) {
          MSPI_CUDA(cudaFuncSetAttribute(maxpool333_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoolTileSmem));
          attr_set = true;
        }
        MSPI_CUDA(launch_pdl(maxpool333_tile_kernel, static_cast<unsigned>(blocks), threads, smem, stream, *d,
                             static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), hs, strips, cgroups, lanes));
        MSPI_LAUNCH_CHECK();
        return MSPI_OK;
      }
    }
    const long long rows = static_cast<long long>(d->n) * d->t * d->h * c8;
    MSPI_CUDA(launch_pdl(maxpool333_kernel, grid_for(rows), kBlock, 0, stream, *d, static_cast<const __nv_bfloat16*>(x),
                                                             static_cast<__nv_bfloat16*>(y), rows, c8));
    MSPI_LAUNCH_CHECK();
    return MSPI_OK;
  }
  // clamping is only valid if every window starts inside the tensor (true for pad < kernel and the usual output sizes)
  const bool starts_in = (d->ot - 1) * d->st - d->pt < d->t && (d->oh - 1) * d->sh - d->ph < d->h && (d->ow - 1) * d->sw - d->pw < d->w &&
                         d->pt < d->kt && d->ph < d->kh && d->pw < d->kw;
#define MSPI_POOL_FIXED(KT, KH, KW)                                                                                       \
  if (starts_in && d->kt == KT && d->kh == KH && d->kw == KW) {                                                             \
    MSPI_CUDA(launch_pdl(maxpool3d_fixed_kernel<KT, KH, KW>, grid_for(total), kBlock, 0, stream, *d, static_cast<const __nv_bfloat16*>(x),   \
                                                                                static_cast<__nv_bfloat16*>(y), total, c8)); \
    MSPI_LAUNCH_CHECK();                                                                                                  \
    return MSPI_OK;                                                                                                       \
  }
  MSPI_POOL_FIXED(1, 3, 3)
  MSPI_POOL_FIXED(3, 3, 3)
  MSPI_POOL_FIXED(2, 2, 2)
  MSPI_POOL_FIXED(1, 2, 2)
#undef MSPI_POOL_FIXED
  MSPI_CUDA(launch_pdl(maxpool3d_kernel, grid_for(total), kBlock, 0, stream, *d, static_cast<const __nv_bfloat16*>(x),
                                                           static_cast<__nv_bfloat16*>(y), total, c8));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_upsample_bilinear(const MspiUpDesc* d, const void* x, void* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && y && d->k >= 1, "mspi_upsample_bilinear: bad argument");
  MSPI_CHECK_ARG(d->c % 8 == 0 && d->in_cstride % 8 == 0 && d->out_cstride % 8 == 0,
                 "channels and pixel strides must be multiples of 8");
  MSPI_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "16-byte alignment");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  constexpr int VEC = 8;
  const int cv = d->c / VEC;
  const long long total = static_cast<long long>(d->nt) * d->h * d->k * d->w * d->k * cv;
  const int g = grid_for(total);
  using bf = __nv_bfloat16;
  if (d->in_dtype == MSPI_BF16 && d->out_dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(upsample_kernel<bf, bf, VEC>, g, kBlock, 0, stream, *d, static_cast<const bf*>(x), static_cast<bf*>(y), total, cv));
  else if (d->in_dtype == MSPI_BF16 && d->out_dtype == MSPI_F32)
    MSPI_CUDA(launch_pdl(upsample_kernel<bf, float, VEC>, g, kBlock, 0, stream, *d, static_cast<const bf*>(x), static_cast<float*>(y), total, cv));
  else if (d->in_dtype == MSPI_F32 && d->out_dtype == MSPI_F32)
    MSPI_CUDA(launch_pdl(upsample_kernel<float, float, VEC>, g, kBlock, 0, stream, *d, static_cast<const float*>(x), static_cast<float*>(y), total, cv));
  else
    MSPI_CUDA(launch_pdl(upsample_kernel<float, bf, VEC>, g, kBlock, 0, stream, *d, static_cast<const float*>(x), static_cast<bf*>(y), total, cv));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_sa_gate(const void* x, int64_t x_cstride, const float* mask_logits, void* y, int64_t y_cstride,
                            int64_t pixels, int c, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && mask_logits && y && c % 8 == 0 && x_cstride % 8 == 0 && y_cstride % 8 == 0,
                 "mspi_sa_gate: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const long long total = pixels * (c / 8);
  if (dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(sa_gate_kernel<__nv_bfloat16>, grid_for(total), kBlock, 0, stream, 
        static_cast<const __nv_bfloat16*>(x), x_cstride, mask_logits, static_cast<__nv_bfloat16*>(y), y_cstride, total, c / 8));
  else
    MSPI_CUDA(launch_pdl(sa_gate_kernel<float>, grid_for(total), kBlock, 0, stream, static_cast<const float*>(x), x_cstride, mask_logits,
                                                                  static_cast<float*>(y), y_cstride, total, c / 8));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_sa_gate_fused(const float* x, int64_t x_cstride, const float* mask_logits, float* y, int64_t y_cstride,
                                  int nt, int h, int w, int c, int nsrc, const float* const* srcs, const int64_t* src_cstrides,
                                  const int32_t* src_scales, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && y && c % 8 == 0 && x_cstride % 8 == 0 && y_cstride % 8 == 0 && nsrc >= 0 && nsrc <= 3,
                 "mspi_sa_gate_fused: bad argument");
  GateSrc s;
  s.n = nsrc;
  for (int j = 0; j < nsrc; ++j) {
    MSPI_CHECK_ARG(srcs[j] && src_scales[j] >= 1 && h % src_scales[j] == 0 && w % src_scales[j] == 0 && src_cstrides[j] % 8 == 0,
                   "source %d: scale %d must divide %dx%d", j, src_scales[j], h, w);
    s.p[j] = srcs[j];
    s.cs[j] = src_cstrides[j];
    s.k[j] = src_scales[j];
    s.sh[j] = h / src_scales[j];
    s.sw[j] = w / src_scales[j];
    s.inv[j] = 1.f / static_cast<float>(src_scales[j]);
  }
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const long long rows_ll = static_cast<long long>(nt) * h;
  const int c8 = c / 8;
  MSPI_CHECK_ARG(rows_ll < (1ll << 31) && static_cast<long long>(w) * c8 < (1ll << 31) / c8,
                 "mspi_sa_gate_fused: %lld rows of %d x %d elements exceed the row kernel's index range", rows_ll, w, c8);
  const int rows = static_cast<int>(rows_ll);
  if (rows == 0) return MSPI_OK;
  const unsigned magic = static_cast<unsigned>(((1ull << 32) + c8 - 1) / c8);   // e / c8 == umulhi(e, magic) for e * c8 < 2^32
  const int grid = rows < num_sms() * 8 ? rows : num_sms() * 8;
  bool ladder = nsrc > 0 && w % 4 == 0 && w <= kGateMaxW && x_cstride % 4 == 0 && y_cstride % 4 == 0 &&
                ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  for (int j = 0; j < nsrc; ++j)
    ladder = ladder && src_scales[j] == (2 << j) && src_cstrides[j] % 4 == 0 && (reinterpret_cast<uintptr_t>(srcs[j]) & 15) == 0;
  static const bool ladder_on = [] { const char* e = getenv("MSPI_GATE_LADDER"); return !e || atoi(e) != 0; }();
  if (ladder && ladder_on) {
    const int c4 = c / 4;
    const unsigned magic4 = static_cast<unsigned>(((1ull << 32) + c4 - 1) / c4);
    switch (nsrc) {
      case 1: MSPI_CUDA(launch_pdl(sa_gate_ladder_kernel<1>, grid, kBlock, 0, stream, x, x_cstride, mask_logits, y, y_cstride, rows, c4, h, w, magic4, s)); break;
      case 2: MSPI_CUDA(launch_pdl(sa_gate_ladder_kernel<2>, grid, kBlock, 0, stream, x, x_cstride, mask_logits, y, y_cstride, rows, c4, h, w, magic4, s)); break;
      default: MSPI_CUDA(launch_pdl(sa_gate_ladder_kernel<3>, grid, kBlock, 0, stream, x, x_cstride, mask_logits, y, y_cstride, rows, c4, h, w, magic4, s)); break;
    }
    MSPI_LAUNCH_CHECK();
    return MSPI_OK;
  }
  switch (nsrc) {
    case 0: MSPI_CUDA(launch_pdl(sa_gate_fused_kernel<0>, grid, kBlock, 0, stream, x, x_cstride, mask_logits, y, y_cstride, rows, c8, h, w, magic, s)); break;
    case 1: MSPI_CUDA(launch_pdl(sa_gate_fused_kernel<1>, grid, kBlock, 0, stream, x, x_cstride, mask_logits, y, y_cstride, rows, c8, h, w, magic, s)); break;
    case 2: MSPI_CUDA(launch_pdl(sa_gate_fused_kernel<2>, grid, kBlock, 0, stream, x, x_cstride, mask_logits, y, y_cstride, rows, c8, h, w, magic, s)); break;
    default: MSPI_CUDA(launch_pdl(sa_gate_fused_kernel<3>, grid, kBlock, 0, stream, x, x_cstride, mask_logits, y, y_cstride, rows, c8, h, w, magic, s)); break;
  }
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(a && b && y && n % 8 == 0, "mspi_add_bf16: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  MSPI_CUDA(launch_pdl(add_bf16_kernel, grid_for(n / 8), kBlock, 0, stream, static_cast<const __nv_bfloat16*>(a),
                                                          static_cast<const __nv_bfloat16*>(b),
                                                          static_cast<__nv_bfloat16*>(y), n / 8));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_cast_rows(const void* src, int src_dtype, int64_t src_rstride, int64_t src_gstride, void* dst,
                              int dst_dtype, int64_t dst_rstride, int64_t dst_gstride, int groups, int rows, int c,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(src && dst && groups > 0 && rows > 0 && c > 0, "mspi_cast_rows: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const long long total = static_cast<long long>(groups) * rows * c;
  const int g = grid_for(total);
  using bf = __nv_bfloat16;
  if (src_dtype == MSPI_F32 && dst_dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(cast_rows_kernel<float, bf>, g, kBlock, 0, stream, static_cast<const float*>(src), src_rstride, src_gstride,
                                                          static_cast<bf*>(dst), dst_rstride, dst_gstride, rows, c, total));
  else if (src_dtype == MSPI_BF16 && dst_dtype == MSPI_F32)
    MSPI_CUDA(launch_pdl(cast_rows_kernel<bf, float>, g, kBlock, 0, stream, static_cast<const bf*>(src), src_rstride, src_gstride,
                                                          static_cast<float*>(dst), dst_rstride, dst_gstride, rows, c, total));
  else if (src_dtype == MSPI_BF16)
    MSPI_CUDA(launch_pdl(cast_rows_kernel<bf, bf>, g, kBlock, 0, stream, static_cast<const bf*>(src), src_rstride, src_gstride,
                                                       static_cast<bf*>(dst), dst_rstride, dst_gstride, rows, c, total));
  else
    MSPI_CUDA(launch_pdl(cast_rows_kernel<float, float>, g, kBlock, 0, stream, static_cast<const float*>(src), src_rstride, src_gstride,
                                                             static_cast<float*>(dst), dst_rstride, dst_gstride, rows, c, total));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}
