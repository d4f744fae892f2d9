// Kernels specific to the X3D-L motion encoder (backbones/X3D.py, SlowFast/resnet_helper.py:213-351):
//   * depthwise 3x3x3 convolution with spatial stride, folded BatchNorm and optional ReLU / Swish,
//   * the Squeeze-Excitation path: per-(sample, channel) mean, the two tiny FC layers + sigmoid,
//     and the gate * Swish applied to the activation,
//   * the stem's depthwise temporal (5,1,1) conv + BN + ReLU for 16-frame clips.
// All are HBM-bound elementwise / stencil / reduction kernels: channels-last bf16, 16-byte accesses,
// fp32 arithmetic.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mspi {
namespace {

__device__ __forceinline__ float act_f(float v, int act) {
  if (act == MSPI_ACT_RELU) return fmaxf(v, 0.f);
  // approximate division (MUFU.RCP + multiply, 2 ulp) instead of the IEEE sequence (~15 instructions): the results are rounded
  // to bf16, and with full division the Swish epilogue was 60 % of the depthwise kernel's instructions (1.34 vs 0.74 ms)
  if (act == MSPI_ACT_SWISH) return __fdividef(v, 1.f + __expf(-v));
  if (act == MSPI_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-v));
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  unpack_bf16x2(u.x, f[0], f[1]); unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]); unpack_bf16x2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// thread = (output position, 8 channels).  wgt: fp32 [kt*kh*kw][C] with the BatchNorm scale folded in, shift fp32 [C].
__global__ void dw3d_kernel(MspiDw3dDesc d, const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt,
                            const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, long long total, int c8) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int pt = d.kt / 2, ph = d.kh / 2, pw = d.kw / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    long long r = i / c8;
    const int ow = static_cast<int>(r % d.ow); r /= d.ow;
    const int oh = static_cast<int>(r % d.oh); r /= d.oh;
    const int ot = static_cast<int>(r % d.t);
    const long long n = r / d.t;
    float acc[8];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg), b = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg + 1);
      acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; acc[3] = a.w; acc[4] = b.x; acc[5] = b.y; acc[6] = b.z; acc[7] = b.w;
    }
    for (int kt = 0; kt < d.kt; ++kt) {
      const int it = ot + kt - pt;
      if (it < 0 || it >= d.t) continue;
      for (int kh = 0; kh < d.kh; ++kh) {
        const int ih = oh * d.sh + kh - ph;
        if (ih < 0 || ih >= d.h) continue;
        for (int kw = 0; kw < d.kw; ++kw) {
          const int iw = ow * d.sw + kw - pw;
          if (iw < 0 || iw >= d.w) continue;
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((n * d.t + it) * d.h + ih) * d.w + iw) * d.in_cstride) + cg);
          float f[8];
          unpack8(u, f);
          const float4* wp = reinterpret_cast<const float4*>(wgt + static_cast<long long>((kt * d.kh + kh) * d.kw + kw) * d.c) + 2 * cg;
          const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
          acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]);
          acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
          acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]);
          acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = act_f(acc[j], d.act);
    reinterpret_cast<uint4*>(y + (((n * d.t + ot) * d.oh + oh) * d.ow + ow) * d.out_cstride)[cg] = pack8(acc);
  }
}

// Register-tiled 3x3x3 variant: a thread owns 8 channels of a strip of P consecutive outputs along W, so each input
// vector it loads feeds up to three accumulators and the nine (kt,kh) weight rows are loaded once per strip instead of
// once per output (the plain kernel issues 81 loads per 216 FMA; this one (P*S+2+6)*9 per 216*P).
template <int P, int S, int ACT>
__global__ void dw3d_strip_kernel(MspiDw3dDesc d, const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt,
                                  const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, long long total, int c8,
                                  int strips) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  constexpr int NIN = (P - 1) * S + 3;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = static_cast<int>(i % c8);
  long long r = i / c8;
  const int ow0 = static_cast<int>(r % strips) * P; r /= strips;
  const int oh = static_cast<int>(r % d.oh); r /= d.oh;
  const int ot = static_cast<int>(r % d.t);
  const long long n = r / d.t;
  float acc[P][8];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg), b = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg + 1);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      acc[p][0] = a.x; acc[p][1] = a.y; acc[p][2] = a.z; acc[p][3] = a.w;
      acc[p][4] = b.x; acc[p][5] = b.y; acc[p][6] = b.z; acc[p][7] = b.w;
    }
  }
  const int iw0 = ow0 * S - 1;
#pragma unroll 1
  for (int kt = 0; kt < 3; ++kt) {
    const int it = ot + kt - 1;
    if (it < 0 || it >= d.t) continue;
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * S + kh - 1;
      if (ih < 0 || ih >= d.h) continue;
      const __nv_bfloat16* xrow = x + ((n * d.t + it) * d.h + ih) * static_cast<long long>(d.w) * d.in_cstride + 8 * cg;
      uint4 raw[NIN];
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        const int iw = min(max(iw0 + j, 0), d.w - 1);  // clamped address, masked below: loads stay unconditional
        raw[j] = __ldg(reinterpret_cast<const uint4*>(xrow + static_cast<long long>(iw) * d.in_cstride));
      }
      float w[3][8];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wp = reinterpret_cast<const float4*>(wgt + static_cast<long long>((kt * 3 + kh) * 3 + kw) * d.c) + 2 * cg;
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        w[kw][0] = w0.x; w[kw][1] = w0.y; w[kw][2] = w0.z; w[kw][3] = w0.w;
        w[kw][4] = w1.x; w[kw][5] = w1.y; w[kw][6] = w1.z; w[kw][7] = w1.w;
      }
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        float f[8];
        unpack8(raw[j], f);
        const int iw = iw0 + j;
        if (iw < 0 || iw >= d.w) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int q = j - kw;  // = p * S for the output p this input feeds through tap kw
          if (q >= 0 && q % S == 0 && q / S < P) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[q / S][e] = fmaf(f[e], w[kw][e], acc[q / S][e]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p) {
    if (ow0 + p < d.ow) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[p][e] = act_f(acc[p][e], ACT);
      reinterpret_cast<uint4*>(y + (((n * d.t + ot) * d.oh + oh) * d.ow + ow0 + p) * d.out_cstride)[cg] = pack8(acc[p]);
    }
  }
}

// Shared-memory tiled 3x3x3 depthwise convolution (stride 1).  The strip kernel above re-reads every input vector from
// L1 / L2 nine times (once per (kt, kh) of the outputs it feeds) and spends a third of its instructions on address math and
// border predicates.  Here a block owns a tile of TR rows x 8*SC columns of ONE output frame for a group of CG8 channel
// octets; the three input frames of the tile (with their one-pixel halo, (TR+2) x (8*SC+2) pixels each) arrive as ONE TMA
// box of the 5-D view (c, w, h, t, n): coordinates outside the tensor — spatial borders and the frames before / after the
// clip — are zero-filled by the copy engine, which is exactly the convolution's padding.  A thread owns 8 channels of a strip
// of 8 output pixels (64 accumulators as fp32 pairs); per (kt, kh) it reads 10 vectors from shared memory for 96 packed FMAs.
constexpr int kSeSlots = 16;
// mean[n][c] = sum over the kSeSlots partial means work[n][slot][c] the tiled depthwise kernel accumulated
__global__ void slot_sum_kernel(const float* __restrict__ work, float* __restrict__ mean, int n, int c) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int s = i / c, ch = i - s * c;
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < kSeSlots; ++k) t += work[(static_cast<long long>(s) * kSeSlots + k) * c + ch];
  mean[i] = t;
}

template <int ACT, int CG8T>   // CG8T: channel octets per group at compile time (shared-memory offsets become immediates), 0 = run time
__global__ void __launch_bounds__(160)
dw3d_tile_kernel(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ wgt, const float* __restrict__ shift,
                 __nv_bfloat16* __restrict__ y, int T, int H, int W, int C, long long out_cstride, int TR, int SC, int cg8_arg,
                 int tiles_x, int nthreads, float* __restrict__ mean_out, float mean_scale) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  constexpr int P = 8;
  // mean_out: the SE block's per-(sample, channel) mean of this layer's output (resnet_helper.py:47-73), accumulated here —
  // block-wide partial sums in shared memory, one atomic per channel and block — instead of by a separate pass over y
  __shared__ __align__(16) float s_part[160 * 8];   // per-thread partial sums (shared atomics: 16-way contention, +30 % kernel time)
  const int CG8 = CG8T > 0 ? CG8T : cg8_arg;
  extern __shared__ __align__(128) uint8_t dw3_smem[];
  __shared__ __align__(8) unsigned long long bar_mem;
  const int IW = P * SC + 2, IH = TR + 2, CG = CG8 * 8;
  const int grp = blockIdx.x / tiles_x, tx = blockIdx.x - grp * tiles_x, ty = blockIdx.y;
  const int n = blockIdx.z / T, ot = blockIdx.z - n * T;
  const int x0 = tx * P * SC, y0 = ty * TR, c0 = grp * CG;
  const uint32_t bar = tc::smem_u32(&bar_mem);
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc::mbar_expect_tx(bar, static_cast<uint32_t>(3 * IH * IW * CG * 2));
    tc::tma_load_5d(tc::smem_u32(dw3_smem), &map_x, bar, c0, x0 - 1, y0 - 1, ot - 1, n);
  }
  const int tid = threadIdx.x;
  const bool active = tid < nthreads;
  const int cg = tid % CG8;
  const int st = tid / CG8;          // strip index inside the tile
  const int sc = st % SC, r = st / SC;
  F2 acc[P][4];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(shift + c0) + 2 * cg), b = __ldg(reinterpret_cast<const float4*>(shift + c0) + 2 * cg + 1);
#pragma unroll
    for (int p = 0; p < P; ++p) { acc[p][0] = pack2(a.x, a.y); acc[p][1] = pack2(a.z, a.w); acc[p][2] = pack2(b.x, b.y); acc[p][3] = pack2(b.z, b.w); }
  }
  __syncthreads();
  tc::mbar_wait(bar, 0);
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
  const __nv_bfloat16* tile = reinterpret_cast<const __nv_bfloat16*>(dw3_smem);
  const float4* wq = reinterpret_cast<const float4*>(wgt + c0) + 2 * cg;
  const int wstride = C / 4;   // float4 per tap row
#pragma unroll 1
  for (int kt = 0; kt < 3; ++kt) {
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      F2 w[3][4];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4 w0 = __ldg(wq + ((kt * 3 + kh) * 3 + kw) * wstride), w1 = __ldg(wq + ((kt * 3 + kh) * 3 + kw) * wstride + 1);
        w[kw][0] = pack2(w0.x, w0.y); w[kw][1] = pack2(w0.z, w0.w); w[kw][2] = pack2(w1.x, w1.y); w[kw][3] = pack2(w1.z, w1.w);
      }
      const __nv_bfloat16* row = tile + ((static_cast<size_t>(kt) * IH + r + kh) * IW + sc * P) * CG + 8 * cg;
#pragma unroll
      for (int j = 0; j < P + 2; ++j) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + static_cast<size_t>(j) * CG);
        F2 v[4];
        v[0] = pack2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
        v[1] = pack2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
        v[2] = pack2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
        v[3] = pack2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int p = j - kw;
          if (p >= 0 && p < P) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[p][e] = fma2(v[e], w[kw][e], acc[p][e]);
          }
        }
      }
    }
  }
  const int oy = y0 + r;
  if (oy < H) {
  __nv_bfloat16* yrow = y + ((static_cast<long long>(n) * T + ot) * H + oy) * static_cast<long long>(W) * out_cstride + c0 + 8 * cg;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int ox = x0 + sc * P + p;
    if (ox < W) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) unpack2(acc[p][e], f[2 * e], f[2 * e + 1]);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = act_f(f[e], ACT);
      *reinterpret_cast<uint4*>(yrow + static_cast<long long>(ox) * out_cstride) = pack8(f);
#pragma unroll
      for (int e = 0; e < 8; ++e) csum[e] += f[e];
    }
  }
  }
  }   // active
  if (mean_out != nullptr) {   // block-uniform
    if (active) {
      float4* q = reinterpret_cast<float4*>(s_part + tid * 8);
      q[0] = make_float4(csum[0], csum[1], csum[2], csum[3]);
      q[1] = make_float4(csum[4], csum[5], csum[6], csum[7]);
    }
    __syncthreads();
    if (tid < CG) {            // thread = channel: add up the strips' partial sums (thread st * CG8 + cg holds channels 8 cg ..)
      const int cgi = tid >> 3, e = tid & 7;
      float t = 0.f;
      for (int st2 = cgi; st2 < nthreads; st2 += CG8) t += s_part[st2 * 8 + e];
      // kSeSlots partial means per (sample, channel): the 672 blocks of a sample would otherwise queue on one L2 atomic unit
      // per channel (measured: +35 us per launch); slot_sum_kernel adds the slots up
      const int slot = (blockIdx.x + blockIdx.y * 3 + blockIdx.z * 5) % kSeSlots;
      atomicAdd(mean_out + (static_cast<long long>(n) * kSeSlots + slot) * C + c0 + tid, t * mean_scale);
    }
  }
}

// Returns 1 when the shape is not covered (the caller falls back to the strip kernel).
static int launch_dw3d_tile(const MspiDw3dDesc* d, const void* x, const float* wgt, const float* shift, void* y,
                            cudaStream_t stream, float* mean_out = nullptr) {
  static const bool on = [] { const char* e = getenv("MSPI_DW3D_TILE"); return !e || atoi(e) != 0; }();
  if (!on || d->sh != 1 || d->sw != 1 || d->in_cstride != d->c) return 1;
  const int c8 = d->c / 8;
  int cg8 = 0;
  if (c8 <= 10) cg8 = c8;
  else
    for (int g = 10; g >= 5; --g)
      if (c8 % g == 0) { cg8 = g; break; }
  if (cg8 == 0) return 1;
  int sc = d->w % 16 == 0 ? 2 : (d->w <= 24 ? (d->w + 7) / 8 : 1);
  if (sc > 3) sc = 3;
  int tr = 128 / (cg8 * sc);
  if (tr > d->h) tr = d->h;
  if (tr < 1) tr = 1;
  auto smem_of = [&](int r) { return static_cast<size_t>(3) * (r + 2) * (8 * sc + 2) * cg8 * 16; };
  while (tr > 1 && smem_of(tr) > 72 * 1024) --tr;
  const size_t smem = smem_of(tr);
  const int nthreads = cg8 * sc * tr;
  if (smem > 100 * 1024 || nthreads > 160 || cg8 * 8 > 256) return 1;
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) return 1;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const cuuint64_t C = static_cast<cuuint64_t>(d->c);
  cuuint64_t gdim[5] = {C, static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->t),
                        static_cast<cuuint64_t>(d->n)};
  cuuint64_t gstr[4] = {C * 2, static_cast<cuuint64_t>(d->w) * C * 2, static_cast<cuuint64_t>(d->h) * d->w * C * 2,
                        static_cast<cuuint64_t>(d->t) * d->h * d->w * C * 2};
  cuuint32_t bdim[5] = {static_cast<cuuint32_t>(cg8 * 8), static_cast<cuuint32_t>(8 * sc + 2), static_cast<cuuint32_t>(tr + 2), 3, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  const int tiles_x = (d->w + 8 * sc - 1) / (8 * sc), tiles_y = (d->h + tr - 1) / tr, groups = c8 / cg8;
  const long long frames = static_cast<long long>(d->n) * d->t;
  if (frames > 65535 || tiles_y > 65535) return 1;
  const int block = (nthreads + 31) / 32 * 32;
  const float mean_scale = 1.f / (static_cast<float>(d->t) * d->h * d->w);
  if (mean_out != nullptr && (nthreads > 160 || block < cg8 * 8)) return 1;   // the block reduces its channels through s_part[160][8]
  dim3 grid(tiles_x * groups, tiles_y, static_cast<unsigned>(frames));
  auto yb = static_cast<__nv_bfloat16*>(y);
#define MSPI_DW3T_(A, G)                                                                                                     \
  {                                                                                                                          \
    MSPI_CUDA(cudaFuncSetAttribute(dw3d_tile_kernel<A, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));        \
    MSPI_CUDA(launch_pdl(dw3d_tile_kernel<A, G>, grid, block, smem, stream, map, wgt, shift, yb, d->t, d->h, d->w, d->c, d->out_cstride, tr, sc,   \
                                                          cg8, tiles_x, nthreads, mean_out, mean_scale));                    \
  }
#define MSPI_DW3T(A)                                                                                                         \
  {                                                                                                                          \
    if (cg8 == 7) MSPI_DW3T_(A, 7) else if (cg8 == 9) MSPI_DW3T_(A, 9) else MSPI_DW3T_(A, 0)                                 \
  }
  switch (d->act) {
    case MSPI_ACT_NONE: MSPI_DW3T(MSPI_ACT_NONE) break;
    case MSPI_ACT_RELU: MSPI_DW3T(MSPI_ACT_RELU) break;
    case MSPI_ACT_SWISH: MSPI_DW3T(MSPI_ACT_SWISH) break;
    default: return 1;
  }
#undef MSPI_DW3T
#undef MSPI_DW3T_
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

// Depthwise (kt,1,1) conv over up to 16 frames + shift + activation: a thread owns 8 channels of one (h,w) position for
// all frames of a sample, every input is read once.
template <int MAXT>
__global__ void dwt_bn_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ shift,
                              __nv_bfloat16* __restrict__ y, long long n_hw, int T, int HW, int C, int kt, int act) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int c8 = C >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_hw * c8) return;
  const int cg = static_cast<int>(idx % c8);
  const long long pos = idx / c8;
  const long long n = pos / HW, hw = pos % HW;
  const long long base = (n * T * HW + hw) * C + 8 * cg;
  const long long tstride = static_cast<long long>(HW) * C;
  uint4 v[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t)
    if (t < T) v[t] = __ldg(reinterpret_cast<const uint4*>(x + base + t * tstride));
  float sh[8];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg), b = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg + 1);
    sh[0] = a.x; sh[1] = a.y; sh[2] = a.z; sh[3] = a.w; sh[4] = b.x; sh[5] = b.y; sh[6] = b.z; sh[7] = b.w;
  }
  const int pt = kt / 2;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    if (t >= T) break;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sh[j];
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      const int k = u - t + pt;
      if (u < T && k >= 0 && k < kt) {
        float f[8];
        unpack8(v[u], f);
        const float4* wp = reinterpret_cast<const float4*>(wgt + static_cast<long long>(k) * C) + 2 * cg;
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]);
        acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
        acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]);
        acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = act_f(acc[j], act);
    *reinterpret_cast<uint4*>(y + base + t * tstride) = pack8(acc);
  }
}

// The same layer as a sliding window over t: a thread keeps the KT frames under the current output frame (KT x 4 registers)
// and the two frames it will need next, instead of all T frames of its position.  dwt_bn_kernel<16> holds 16 frames x 8
// channels plus the unrolled tap weights in 231 registers: two 128-thread blocks per SM, 12 % of the warp slots, 1.15 ms for
// the 1.06 GB of the X3D stem's temporal conv (ncu, profiles/r02_small_kernels.md §5).  Weights come from L1 per use.
template <int KT>
__global__ void __launch_bounds__(128)
dwt_bn_slide_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ shift,
                    __nv_bfloat16* __restrict__ y, long long n_hw, int T, int HW, int C, int act) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  constexpr int PT = KT / 2, AHEAD = 2;
  const int c8 = C >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_hw * c8) return;
  long long r = idx;
  const int cg = divmod(r, c8);
  const long long hw = divmod(r, HW), n = r;
  const long long base = (n * T * HW + hw) * C + 8 * cg;
  const long long tstride = static_cast<long long>(HW) * C;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  auto frame = [&](int f) { return (f >= 0 && f < T) ? __ldg(reinterpret_cast<const uint4*>(x + base + f * tstride)) : zero; };
  uint4 ring[KT + AHEAD];   // ring[j] = input frame t - PT + j
#pragma unroll
  for (int j = 0; j < KT + AHEAD; ++j) ring[j] = frame(j - PT);
  float sh[8];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg), b = __ldg(reinterpret_cast<const float4*>(shift) + 2 * cg + 1);
    sh[0] = a.x; sh[1] = a.y; sh[2] = a.z; sh[3] = a.w; sh[4] = b.x; sh[5] = b.y; sh[6] = b.z; sh[7] = b.w;
  }
  for (int t = 0; t < T; ++t) {
    const uint4 nxt = frame(t + 1 - PT + KT + AHEAD - 1);   // the frame that enters the ring after this step
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sh[j];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      float f[8];
      unpack8(ring[k], f);
      const float4* wp = reinterpret_cast<const float4*>(wgt + static_cast<long long>(k) * C) + 2 * cg;
      const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
      acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]);
      acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
      acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]);
      acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = act_f(acc[j], act);
    *reinterpret_cast<uint4*>(y + base + t * tstride) = pack8(acc);
#pragma unroll
    for (int j = 0; j + 1 < KT + AHEAD; ++j) ring[j] = ring[j + 1];
    ring[KT + AHEAD - 1] = nxt;
  }
}

// Per-(sample, channel) sum over `rows` positions: grid (chunks, N); a block reduces its chunk of rows (threads: channel
// pair x row lane) and adds its partial sums to out[n][c] (fp32 atomics; out is zeroed by the caller's memset node).
__global__ void channel_sum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int c,
                                   long long cstride, long long rows_per_block, float scale) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  extern __shared__ float part[];  // [blockDim.y][c]
  const int n = blockIdx.y;
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  const int pairs = c >> 1;
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(x + (static_cast<long long>(n) * rows + r) * cstride)[p];
      s0 += __bfloat162float(v.x);
      s1 += __bfloat162float(v.y);
    }
    part[threadIdx.y * c + 2 * p] = s0;
    part[threadIdx.y * c + 2 * p + 1] = s1;
  }
  __syncthreads();
  for (int ch = threadIdx.y * blockDim.x + threadIdx.x; ch < c; ch += blockDim.x * blockDim.y) {
    float s = 0.f;
    for (int j = 0; j < blockDim.y; ++j) s += part[j * c + ch];
    atomicAdd(out + static_cast<long long>(n) * c + ch, s * scale);
  }
}

// SE gate: gate[n][c] = sigmoid(W2 relu(W1 mean[n] + b1) + b2).  One block per sample.  resnet_helper.py:47-73
__global__ void se_gate_kernel(const float* __restrict__ mean, const float* __restrict__ w1, const float* __restrict__ b1,
                               const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ gate, int c,
                               int cfc) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  extern __shared__ float sm[];  // mean[c] + hidden[cfc]
  float* m_s = sm;
  float* h_s = sm + c;
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < c; i += blockDim.x) m_s[i] = mean[static_cast<long long>(n) * c + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < cfc; j += nw) {
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s = fmaf(w1[static_cast<long long>(j) * c + i], m_s[i], s);
    s = warp_sum(s);
    if (lane == 0) h_s[j] = fmaxf(s + b1[j], 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float s = b2[i];
    for (int j = 0; j < cfc; ++j) s = fmaf(w2[static_cast<long long>(i) * cfc + j], h_s[j], s);
    gate[static_cast<long long>(n) * c + i] = 1.f / (1.f + __expf(-s));
  }
}

// y = act(x * gate[n][c]) over [N][rows][C] bf16 (in place allowed).
__global__ void scale_act_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gate, __nv_bfloat16* __restrict__ y,
                                 long long rows, int c8, long long total, int act) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    const long long n = (i / c8) / rows;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + i);
    float f[8];
    unpack8(u, f);
    const float4* gp = reinterpret_cast<const float4*>(gate + n * (8ll * c8)) + 2 * cg;
    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
    f[0] = act_f(f[0] * g0.x, act); f[1] = act_f(f[1] * g0.y, act); f[2] = act_f(f[2] * g0.z, act); f[3] = act_f(f[3] * g0.w, act);
    f[4] = act_f(f[4] * g1.x, act); f[5] = act_f(f[5] * g1.y, act); f[6] = act_f(f[6] * g1.z, act); f[7] = act_f(f[7] * g1.w, act);
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
}

inline int grid_for(long long total, int block = 256) {
  long long b = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace mspi

using namespace mspi;

extern "C" int mspi_dwconv3d_bn(const MspiDw3dDesc* d, const void* x, const float* wgt, const float* shift, void* y,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && wgt && shift && y, "mspi_dwconv3d_bn: null argument");
  MSPI_CHECK_ARG(d->c % 8 == 0 && d->in_cstride % 8 == 0 && d->out_cstride % 8 == 0, "channels / strides must be multiples of 8");
  MSPI_CHECK_ARG((d->kt & 1) && (d->kh & 1) && (d->kw & 1) && d->sh >= 1 && d->sw >= 1, "odd kernel, positive strides");
  MSPI_CHECK_ARG(d->oh == (d->h + 2 * (d->kh / 2) - d->kh) / d->sh + 1 && d->ow == (d->w + 2 * (d->kw / 2) - d->kw) / d->sw + 1,
                 "output extents do not match the geometry");
  MSPI_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(wgt) |
                   reinterpret_cast<uintptr_t>(shift)) & 15) == 0, "16-byte alignment");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int c8 = d->c / 8;
  const long long total = static_cast<long long>(d->n) * d->t * d->oh * d->ow * c8;
  if (d->kh == 1 && d->kw == 1 && d->sh == 1 && d->sw == 1 && d->t <= 16 && d->in_cstride == d->c && d->out_cstride == d->c) {
    const int HW = d->h * d->w;
    const long long n_hw = static_cast<long long>(d->n) * HW;
    const long long blocks = (n_hw * c8 + 127) / 128;
    MSPI_CHECK_ARG(blocks < (1ll << 31), "grid out of range");
    static const bool slide = [] { const char* e = getenv("MSPI_DWT_SLIDE"); return !e || atoi(e) != 0; }();
    if (slide && d->kt == 5)
      MSPI_CUDA(launch_pdl(dwt_bn_slide_kernel<5>, static_cast<int>(blocks), 128, 0, stream, static_cast<const __nv_bfloat16*>(x), wgt,
                           shift, static_cast<__nv_bfloat16*>(y), n_hw, d->t, HW, d->c, d->act));
    else if (slide && d->kt == 3)
      MSPI_CUDA(launch_pdl(dwt_bn_slide_kernel<3>, static_cast<int>(blocks), 128, 0, stream, static_cast<const __nv_bfloat16*>(x), wgt,
                           shift, static_cast<__nv_bfloat16*>(y), n_hw, d->t, HW, d->c, d->act));
    else
    MSPI_CUDA(launch_pdl(dwt_bn_kernel<16>, static_cast<int>(blocks), 128, 0, stream, static_cast<const __nv_bfloat16*>(x), wgt, shift,
                                                                    static_cast<__nv_bfloat16*>(y), n_hw, d->t, HW, d->c,
                                                                    d->kt, d->act));
  } else if (d->kt == 3 && d->kh == 3 && d->kw == 3 && d->sh == d->sw && (d->sh == 1 || d->sh == 2)) {
    if (d->sh == 1) {   // shared-memory tiled kernel; 1 = shape not covered
      const int rc = launch_dw3d_tile(d, x, wgt, shift, y, stream);
      if (rc != 1) return rc;
    }
    const int P = (d->ow % 8 == 0 && d->sh == 1) ? 8 : 4;  // <8,2> would spill (17 input vectors + 64 accumulators)
    const int strips = (d->ow + P - 1) / P;
    const long long threads = static_cast<long long>(d->n) * d->t * d->oh * strips * c8;
    const long long blocks = (threads + 127) / 128;
    MSPI_CHECK_ARG(blocks < (1ll << 31), "grid out of range");
    auto xb = static_cast<const __nv_bfloat16*>(x);
    auto yb = static_cast<__nv_bfloat16*>(y);
    const int g = static_cast<int>(blocks);
#define MSPI_DW3D(PP, SS)                                                                                                   \
    switch (d->act) {                                                                                                        \
      case MSPI_ACT_NONE: MSPI_CUDA(launch_pdl(dw3d_strip_kernel<PP, SS, MSPI_ACT_NONE>, g, 128, 0, stream, *d, xb, wgt, shift, yb, threads, c8, strips)); break;   \
      case MSPI_ACT_RELU: MSPI_CUDA(launch_pdl(dw3d_strip_kernel<PP, SS, MSPI_ACT_RELU>, g, 128, 0, stream, *d, xb, wgt, shift, yb, threads, c8, strips)); break;   \
      case MSPI_ACT_SWISH: MSPI_CUDA(launch_pdl(dw3d_strip_kernel<PP, SS, MSPI_ACT_SWISH>, g, 128, 0, stream, *d, xb, wgt, shift, yb, threads, c8, strips)); break; \
      default: return set_error(MSPI_ERR_ARG, "mspi_dwconv3d_bn: activation %d", d->act);                                  \
    }
    if (P == 8) { MSPI_DW3D(8, 1) }
    else if (d->sh == 1) { MSPI_DW3D(4, 1) }
    else { MSPI_DW3D(4, 2) }
#undef MSPI_DW3D
  } else {
    MSPI_CUDA(launch_pdl(dw3d_kernel, grid_for(total), 256, 0, stream, *d, static_cast<const __nv_bfloat16*>(x), wgt, shift,
                                                     static_cast<__nv_bfloat16*>(y), total, c8));
  }
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_channel_mean(const void* x, float* out, int n, int64_t rows, int c, int64_t cstride, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && out && n > 0 && rows > 0 && c > 0 && c % 2 == 0 && cstride >= c && cstride % 2 == 0 && c <= 2048,
                 "mspi_channel_mean: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  MSPI_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n * c, stream));
  const int ty = 8, tx = 32;
  long long chunks = (static_cast<long long>(num_sms()) * 4 + n - 1) / n;  // ~4 blocks per SM over the batch
  const long long max_chunks = (rows + 63) / 64;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const long long rpb = (rows + chunks - 1) / chunks;
  dim3 grid(static_cast<unsigned>((rows + rpb - 1) / rpb), n), block(tx, ty);
  MSPI_CUDA(launch_pdl(channel_sum_kernel, grid, block, sizeof(float) * ty * c, stream, static_cast<const __nv_bfloat16*>(x), out, rows, c,
                                                                      cstride, rpb, 1.f / static_cast<float>(rows)));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_se_gate(const float* mean, const float* w1, const float* b1, const float* w2, const float* b2,
                            float* gate, int n, int c, int cfc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(mean && w1 && b1 && w2 && b2 && gate && n > 0 && c > 0 && cfc > 0 && c + cfc <= 8192, "mspi_se_gate: bad argument");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  MSPI_CUDA(launch_pdl(se_gate_kernel, n, 256, sizeof(float) * (c + cfc), stream, mean, w1, b1, w2, b2, gate, c, cfc));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_scale_act(const void* x, const float* gate, void* y, int n, int64_t rows, int c, int act, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && gate && y && n > 0 && rows > 0 && c > 0 && c % 8 == 0, "mspi_scale_act: bad argument");
  MSPI_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gate)) & 15) == 0,
                 "16-byte alignment");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int c8 = c / 8;
  const long long total = static_cast<long long>(n) * rows * c8;
  MSPI_CUDA(launch_pdl(scale_act_kernel, grid_for(total), 256, 0, stream, static_cast<const __nv_bfloat16*>(x), gate,
                                                        static_cast<__nv_bfloat16*>(y), rows, c8, total, act));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

// Depthwise conv + shift + activation AND the per-(sample, channel) mean of its output (the SE block's squeeze,
// resnet_helper.py:47-73): the shared-memory tile kernel accumulates the mean while it stores y; layers it does not cover
// (stride 2, odd shapes) run the plain kernel and the separate reduction.  mean_out: fp32 [n][c]; work: fp32 [n][16][c] scratch.
extern "C" int mspi_dwconv3d_bn_mean(const MspiDw3dDesc* d, const void* x, const float* wgt, const float* shift, void* y,
                                     float* mean_out, float* work, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && wgt && shift && y && mean_out && work, "mspi_dwconv3d_bn_mean: null argument");
  static const bool fuse = [] { const char* e = getenv("MSPI_SE_MEAN_FUSED"); return !e || atoi(e) != 0; }();
  if (fuse && d->kt == 3 && d->kh == 3 && d->kw == 3 && d->sh == 1 && d->sw == 1 && d->c % 8 == 0 && d->in_cstride % 8 == 0 &&
      d->out_cstride % 8 == 0 && d->oh == d->h && d->ow == d->w &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(wgt) |
        reinterpret_cast<uintptr_t>(shift)) & 15) == 0 && num_sms() > 0) {
    MSPI_CUDA(cudaMemsetAsync(work, 0, sizeof(float) * d->n * kSeSlots * d->c, stream));
    const int rc = launch_dw3d_tile(d, x, wgt, shift, y, stream, work);
    if (rc == MSPI_OK) {
      const int total = d->n * d->c;
      MSPI_CUDA(launch_pdl(slot_sum_kernel, (total + 255) / 256, 256, 0, stream, static_cast<const float*>(work), mean_out, d->n, d->c));
      MSPI_LAUNCH_CHECK();
      return MSPI_OK;
    }
    if (rc != 1) return rc;
  }
  const int rc = mspi_dwconv3d_bn(d, x, wgt, shift, y, stream_);
  if (rc != MSPI_OK) return rc;
  return mspi_channel_mean(y, mean_out, d->n, static_cast<int64_t>(d->t) * d->oh * d->ow, d->c, d->out_cstride, stream_);
}
