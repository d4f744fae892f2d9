// Depthwise 7x7 convolution + LayerNorm over channels (ConvNeXt block front half), and the (7,1,1)
// temporal depthwise convolution of the decoder's ConvNextBlock.  Channels-last, fp32 arithmetic.
//
// Replaces timm ConvNeXt conv_dw + norm (model_utils.py:361) and ConvNextBlock.dwconv_t / dwconv_s +
// LayerNorm3d (model_utils.py:293-303,321-323).
//
// The 7x7 stencil is FMA bound (49 FMA per output element against 4 bytes of HBM traffic), so the
// kernel is built around on-chip reuse:
//   1. the block copies its input tile, (S+6) x (P+6) pixels x C channels with a zero halo, into
//      shared memory with 16-byte cp.async (every load of the tile in flight at once);
//   2. a thread owns 4 channels of a horizontal strip of P output pixels: each input value it reads
//      from shared memory feeds up to 7 accumulators, the 7 taps of the current filter row sit in
//      registers (49 x 4 FMA per 1/P.. of a load);
//   3. the strip results (+bias) are parked in shared memory as fp32 (re-using the tile buffer) and
//      warps normalise one pixel at a time (two-pass mean / variance in registers), writing whole
//      channel rows.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mspi {
namespace {

template <typename T>
struct Ld4;
template <>
struct Ld4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 cvt(uint2 u) {
    // bf16 -> fp32 is a 16-bit left shift: one shift for the low half, one mask for the high half (4 ALU ops per 4 values;
    // the ALU pipe is the co-bottleneck of the stencil next to the FMA pipe)
    float4 f;
    f.x = __uint_as_float(u.x << 16);
    f.y = __uint_as_float(u.x & 0xffff0000u);
    f.z = __uint_as_float(u.y << 16);
    f.w = __uint_as_float(u.y & 0xffff0000u);
    return f;
  }
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) { return cvt(__ldg(reinterpret_cast<const uint2*>(p))); }
  static __device__ __forceinline__ float4 lds(const __nv_bfloat16* p) { return cvt(*reinterpret_cast<const uint2*>(p)); }
};
template <>
struct Ld4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ float4 lds(const float* p) { return *reinterpret_cast<const float4*>(p); }
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = pred ? 16 : 0;  // src-size 0: the 16 destination bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(bytes) : "memory");
}

// LayerNorm over the C channels of the S*P pixels parked (fp32) in out_s, written to y.  LPP lanes share a pixel (16 for
// C = 96, else 32), each owning channel pairs sl, sl + LPP, ...; a lane group handles G pixels per iteration so that G
// independent reduction chains are in flight.  Mean and sum of squared deviations are reduced TOGETHER by Chan's pairwise
// merge (per lane: local mean / M2 of its 2*NP values; per butterfly step: m' = (m + mo) / 2, M2' = M2 + M2o + (mo - m)^2 n/2):
// one dependent chain of log2(LPP) shuffle steps instead of two (mean, then variance), without the cancellation of the
// E[x^2] - mean^2 form.  Measured with the instrumented kernel (tools/prof_dw_phases.py): this phase took 6100 cycles per warp
// of a 17600-cycle block with the two-butterfly / two-pixel form — latency, not instruction count.
template <int C, int NPIX, int P, int NWARPS, int G, bool TWO_PASS = false>
__device__ __forceinline__ void dw_layernorm_store(const float* __restrict__ out_s, const float* __restrict__ ln_w,
                                                   const float* __restrict__ ln_b, void* __restrict__ y, int out_bf16,
                                                   long long tile_base, int y0, int x0, int H, int W, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int pairs = C / 2;
  constexpr int LPP = (pairs % 32 == 0) ? 32 : 16;
  static_assert(pairs % LPP == 0, "channel pairs must tile the lane group");
  constexpr int NP = pairs / LPP;
  constexpr int GRP = 32 / LPP;
  static_assert(NPIX % (G * GRP) == 0, "pixel groups");
  const int sub = lane / LPP, sl = lane % LPP;
  float2 gam[NP], bet[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    gam[i] = __ldg(reinterpret_cast<const float2*>(ln_w) + sl + LPP * i);
    bet[i] = __ldg(reinterpret_cast<const float2*>(ln_b) + sl + LPP * i);
  }
  for (int p0 = (warp * GRP + sub) * G; p0 < NPIX; p0 += NWARPS * GRP * G) {
    float2 v[G][NP];
    float mean[G], m2[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float2* src = reinterpret_cast<const float2*>(out_s + static_cast<size_t>(p0 + g) * C);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[g][i] = src[sl + LPP * i];
        sum += v[g][i].x + v[g][i].y;
      }
      mean[g] = sum * (1.f / (2 * NP));
      m2[g] = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float a = v[g][i].x - mean[g], b = v[g][i].y - mean[g];
        m2[g] = fmaf(a, a, fmaf(b, b, m2[g]));
      }
    }
    if constexpr (TWO_PASS) {   // study variant (MSPI_DW_LNV=0): mean butterfly, then variance butterfly
#pragma unroll
      for (int g = 0; g < G; ++g) mean[g] *= static_cast<float>(2 * NP);
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) mean[g] += __shfl_xor_sync(0xffffffffu, mean[g], o);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        mean[g] *= (1.f / C);
        m2[g] = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const float a = v[g][i].x - mean[g], b = v[g][i].y - mean[g];
          m2[g] += a * a + b * b;
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) m2[g] += __shfl_xor_sync(0xffffffffu, m2[g], o);
      }
    } else {
    float half_n = static_cast<float>(NP);   // n / 2 with n = 2 * NP values per lane
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float mo = __shfl_xor_sync(0xffffffffu, mean[g], o);
        const float m2o = __shfl_xor_sync(0xffffffffu, m2[g], o);
        const float d = mo - mean[g];
        mean[g] = 0.5f * (mean[g] + mo);
        m2[g] = fmaf(d * d, half_n, m2[g] + m2o);
      }
      half_n *= 2.f;
    }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float rs = rsqrtf(m2[g] * (1.f / C) + eps), ms = -mean[g] * rs;
      const int pix = p0 + g;
      const int ly = pix / P, lx = pix % P;   // P is a compile-time constant
      if (y0 + ly >= H || x0 + lx >= W) continue;
      const unsigned off = static_cast<unsigned>(ly * W + lx) * static_cast<unsigned>(C);   // within the tile: 32-bit
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int p = sl + LPP * i;
        // (v - mean) * rstd * gamma + beta as two FMAs
        const float a = fmaf(fmaf(v[g][i].x, rs, ms), gam[i].x, bet[i].x);
        const float b = fmaf(fmaf(v[g][i].y, rs, ms), gam[i].y, bet[i].y);
        if (out_bf16)
          reinterpret_cast<__nv_bfloat162*>(static_cast<__nv_bfloat16*>(y) + tile_base + off)[p] = __floats2bfloat162_rn(a, b);
        else
          reinterpret_cast<float2*>(static_cast<float*>(y) + tile_base + off)[p] = make_float2(a, b);
      }
    }
  }
}

// CQ = C/4 threads per strip, S strips (rows) per block, P output pixels per strip.
// Warp geometry.  fp32 tiles: a warp's lanes are QW channel quads x SW strips, so the SW strips' lanes load the SAME
// weight vector (one L1 wavefront instead of SW; the weights are per channel) and each strip's lanes read whole 128-byte
// rows of the tile (0.60 -> 0.59 ms on the 192-channel lateral).  bf16 tiles keep one strip per CQ consecutive threads: the
// 4-strip mapping needs an odd stored row width (64-byte strip segments must alternate bank halves), which costs the
// 192-channel stage its 4th resident block and gains nothing at 96 channels (measured: L1 wavefronts are not the limiter,
// the FMA pipe is ~75% busy while a block is in its stencil phase; the rest is load / LayerNorm phases of other blocks).
template <typename TI, int CQ, int S, int P>
struct DwGeom {
  static constexpr int C = 4 * CQ;
  static constexpr int SW = sizeof(TI) != 4 ? 1 : (S % 4 == 0 && CQ % 8 == 0) ? 4 : (S % 2 == 0 && CQ % 16 == 0) ? 2 : 1;   // strips per warp
  static constexpr int QW = 32 / SW;                                                                  // channel quads per warp
  static constexpr bool kPad = SW == 4 && sizeof(TI) == 2 && C % 96 == 0;
  static constexpr int TW = P + 6, TH = S + 6;
  static constexpr int TWP = kPad ? (TW | 1) : TW;                       // stored row width (pixels)
  static constexpr int kBoxC = kPad ? 96 : (C > 256 ? C / 2 : C);        // channels per TMA box (box dimensions are <= 256)
  static constexpr int kBoxes = C / kBoxC;                               // the tile is stored [box][TH][TWP][kBoxC]
  static constexpr size_t tile_bytes = static_cast<size_t>(TH) * TWP * C * sizeof(TI);
  static_assert(!kPad || (TWP * kBoxC * sizeof(TI)) % 128 == 64, "row stride must alternate 64-byte bank halves");
  static_assert(SW == 1 || kBoxC % (4 * QW) == 0, "a warp's channel group must lie inside one box");
};
#ifndef MSPI_DW_MINB
#define MSPI_DW_MINB 1
#endif
#ifndef MSPI_DW_MINB8
#define MSPI_DW_MINB8 4
#endif
template <typename TI, int CQ, int S, int P, bool GROUPED>
__global__ void __launch_bounds__(CQ * S, (CQ * S <= 192 && P == 16) ? MSPI_DW_MINB : (CQ * S <= 192 && P == 8 && CQ <= 48 && sizeof(TI) == 2) ? MSPI_DW_MINB8 : (CQ * S == 384 && !GROUPED && sizeof(TI) == 2) ? 2 : 1)
dw7x7_ln_kernel(const __grid_constant__ CUtensorMap map_x, const TI* __restrict__ x, const float* __restrict__ wgt,
                const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                void* __restrict__ y, int out_bf16, int H, int W, int tiles_x, int n0, float eps, int cs_arg, int use_tma) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  // cs: channels of the tensor (pixel stride).  cs == C (GROUPED false, a compile-time stride: the weight and output addresses
  // become immediates): the block owns whole pixels and can normalise them.  cs > C (GROUPED): the blocks of grid.y each own a
  // group of C channels (the stencil is per channel) and LayerNorm runs as a second kernel.
  constexpr int C = 4 * CQ;
  const int cs = GROUPED ? cs_arg : C;
  // grid = (tiles_x * groups, tiles_y, frames): no per-block integer divisions for the whole-pixel kernels
  const int grp = GROUPED ? blockIdx.x / tiles_x : 0;
  const int c0 = grp * C;
  using Geo = DwGeom<TI, CQ, S, P>;
  constexpr int TW = Geo::TW, TH = Geo::TH, TWP = Geo::TWP, kBoxC = Geo::kBoxC, kBoxes = Geo::kBoxes;
  constexpr int kThreads = CQ * S;
  constexpr int kVec = 16 / sizeof(TI);            // elements per 16-byte copy
  constexpr int kRowVecs = C / kVec;               // 16-byte copies per pixel
  extern __shared__ __align__(128) uint8_t dw_smem[];
  __shared__ __align__(8) unsigned long long tma_bar;
  TI* tile_s = reinterpret_cast<TI*>(dw_smem);        // [box][TH][TWP][kBoxC] input tile, zero halo
  float* out_s = reinterpret_cast<float*>(dw_smem);   // [S*P][C] results; re-uses the tile buffer after the stencil

  const int tx = GROUPED ? blockIdx.x - grp * tiles_x : blockIdx.x;
  const int ty = blockIdx.y;
  const int n = n0 + blockIdx.z;   // frame
  const int x0 = tx * P, y0 = ty * S;
  const TI* xin = x + static_cast<long long>(n) * H * W * cs + c0;

  if (use_tma) {
    // the whole (S+6) x (P+6) x C tile, zero halo included, is one bulk tensor load per <=256-channel box: TMA fills coordinates
    // outside the image with zeros, which is exactly the convolution's padding (no per-thread address math, no predicates)
    const uint32_t bar = tc::smem_u32(&tma_bar);
    if (threadIdx.x == 0) {
      tc::mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      tc::mbar_expect_tx(bar, static_cast<uint32_t>(Geo::tile_bytes));
#pragma unroll
      for (int b = 0; b < kBoxes; ++b)
        tc::tma_load_4d(tc::smem_u32(tile_s + static_cast<size_t>(b) * TH * TWP * kBoxC), &map_x, bar, c0 + b * kBoxC, x0 - 3,
                        y0 - 3, n);
    }
  } else {
  for (int i = threadIdx.x; i < TH * TW * kRowVecs; i += kThreads) {
    const int cv = i % kRowVecs, pix = i / kRowVecs;
    const int tc = pix % TW, tr = pix / TW;
    const int gy = y0 + tr - 3, gx = x0 + tc - 3;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const TI* src = xin + (static_cast<long long>(in ? gy : 0) * W + (in ? gx : 0)) * cs + cv * kVec;
    const int ch = cv * kVec, box = ch / kBoxC;
    cp_async16_zfill(tile_s + (static_cast<size_t>(box) * TH * TWP + tr * TWP + tc) * kBoxC + (ch - box * kBoxC), src, in);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  }

  int q, s;
  if (Geo::SW == 1) {
    q = threadIdx.x % CQ;
    s = threadIdx.x / CQ;
  } else {
    constexpr int NQG = CQ / Geo::QW;   // channel groups
    const int wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
    q = (wp % NQG) * Geo::QW + ln % Geo::QW;
    s = (wp / NQG) * Geo::SW + ln / Geo::QW;
  }
  // Accumulators, taps and inputs are kept as packed fp32 pairs: sm_100's fma.rn.f32x2 retires two FMAs per lane per
  // issue slot (each half rounds exactly like fmaf), which is what bounds this kernel (49 FMAs per output element).
  F2 acc[P][2];
  {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0) + q);
#pragma unroll
    for (int j = 0; j < P; ++j) { acc[j][0] = pack2(b.x, b.y); acc[j][1] = pack2(b.z, b.w); }
  }
  if (use_tma) {
    __syncthreads();   // the barrier init is visible to every waiter
    tc::mbar_wait(tc::smem_u32(&tma_bar), 0);
  } else {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }

  const float4* wq = reinterpret_cast<const float4*>(wgt + c0) + q;
#pragma unroll 1
  for (int kh = 0; kh < 7; ++kh) {
    F2 w[7][2];
#pragma unroll
    for (int kw = 0; kw < 7; ++kw) {
      const float4 t = __ldg(wq + (kh * 7 + kw) * (cs / 4));
      w[kw][0] = pack2(t.x, t.y);
      w[kw][1] = pack2(t.z, t.w);
    }
    const TI* trow = tile_s + (static_cast<size_t>((4 * q) / kBoxC) * TH + s + kh) * TWP * kBoxC + (4 * q) % kBoxC;
#pragma unroll
    for (int ix = 0; ix < TW; ++ix) {
      const float4 v = Ld4<TI>::lds(trow + ix * kBoxC);
      const F2 v0 = pack2(v.x, v.y), v1 = pack2(v.z, v.w);
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const int j = ix - kw;  // output pixel fed by this input through tap kw
        if (j >= 0 && j < P) {
          acc[j][0] = fma2(v0, w[kw][0], acc[j][0]);
          acc[j][1] = fma2(v1, w[kw][1], acc[j][1]);
        }
      }
    }
  }
  __syncthreads();  // everyone is done reading the tile: its memory becomes the result buffer
#pragma unroll
  for (int j = 0; j < P; ++j) {
    float4 o;
    unpack2(acc[j][0], o.x, o.y);
    unpack2(acc[j][1], o.z, o.w);
    reinterpret_cast<float4*>(out_s + static_cast<size_t>(s * P + j) * C)[q] = o;
  }
  __syncthreads();

  if constexpr (!GROUPED) {
    if (ln_w != nullptr) {   // whole pixels + LayerNorm: the shared single-butterfly implementation
      const long long tb = ((static_cast<long long>(n) * H + y0) * W + x0) * C;
      // G = 4 pixels in flight per lane group for the fp32 tiles (0.580 -> 0.545 ms on the 192-channel lateral); the bf16
      // kernels sit at their register budget and lose more in the stencil than the LayerNorm phase gains (measured)
      dw_layernorm_store<C, S * P, P, kThreads / 32, sizeof(TI) == 4 ? 4 : 2>(out_s, ln_w, ln_b, y, out_bf16, tb, y0, x0, H, W, eps);
      return;
    }
  }
  // ---- LayerNorm over C.  LPP lanes share a pixel (16 for C = 96, else 32), each owning channel pairs lane, lane+LPP, ...;
  // a warp handles G pixels per lane group at a time so the butterfly reductions of different pixels overlap, and the
  // affine parameters live in registers for the whole tile.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = kThreads / 32;
  constexpr int pairs = C / 2;
  constexpr int LPP = (pairs % 32 == 0) ? 32 : 16;   // lanes per pixel
  static_assert(pairs % LPP == 0, "channel pairs must tile the lane group");
  constexpr int NP = pairs / LPP;                     // pairs per lane
  constexpr int GRP = 32 / LPP;                       // pixels handled side by side in one warp
#ifndef MSPI_DW_LN_G
#define MSPI_DW_LN_G 2
#endif
  constexpr int G = ((S * P) % (MSPI_DW_LN_G * GRP) == 0) ? MSPI_DW_LN_G : 4;  // pixels per lane group per iteration
  static_assert((S * P) % (G * GRP) == 0, "pixel groups");
  const int sub = lane / LPP, sl = lane % LPP;
  float2 gam[NP], bet[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    gam[i] = ln_w != nullptr ? __ldg(reinterpret_cast<const float2*>(ln_w) + sl + LPP * i) : make_float2(1.f, 1.f);
    bet[i] = ln_w != nullptr ? __ldg(reinterpret_cast<const float2*>(ln_b) + sl + LPP * i) : make_float2(0.f, 0.f);
  }
  const long long tile_base = ((static_cast<long long>(n) * H + y0) * W + x0) * cs + c0;
  for (int p0 = (warp * GRP + sub) * G; p0 < S * P; p0 += nwarps * GRP * G) {
    float2 v[G][NP];
    float sum[G], sq[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float2* src = reinterpret_cast<const float2*>(out_s + static_cast<size_t>(p0 + g) * C);
      sum[g] = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[g][i] = src[sl + LPP * i];
        sum[g] += v[g][i].x + v[g][i].y;
      }
    }
    if (ln_w != nullptr) {
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) sum[g] += __shfl_xor_sync(0xffffffffu, sum[g], o);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        sum[g] *= (1.f / C);  // mean
        sq[g] = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const float a = v[g][i].x - sum[g], b = v[g][i].y - sum[g];
          sq[g] += a * a + b * b;
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) sq[g] += __shfl_xor_sync(0xffffffffu, sq[g], o);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) sq[g] = rsqrtf(sq[g] * (1.f / C) + eps);  // rstd
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) { sum[g] = 0.f; sq[g] = 1.f; }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int pix = p0 + g;
      const int ly = pix / P, lx = pix % P;   // P is a compile-time constant
      if (y0 + ly >= H || x0 + lx >= W) continue;
      const unsigned off = static_cast<unsigned>(ly * W + lx) * static_cast<unsigned>(cs);   // within the tile: 32-bit
      const float rs = sq[g], ms = -sum[g] * sq[g];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int p = sl + LPP * i;
        // (v - mean) * rstd * gamma + beta as two FMAs
        const float a = fmaf(fmaf(v[g][i].x, rs, ms), gam[i].x, bet[i].x);
        const float b = fmaf(fmaf(v[g][i].y, rs, ms), gam[i].y, bet[i].y);
        if (out_bf16)
          reinterpret_cast<__nv_bfloat162*>(static_cast<__nv_bfloat16*>(y) + tile_base + off)[p] = __floats2bfloat162_rn(a, b);
        else
          reinterpret_cast<float2*>(static_cast<float*>(y) + tile_base + off)[p] = make_float2(a, b);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent, double-buffered variant for the whole-pixel bf16 tiles (ConvNeXt stages 0 and 1).
//
// The one-tile-per-block kernel above spends a block's life in three serial phases (wait for the tile, stencil, LayerNorm)
// and relies on the other resident blocks to keep the FMA pipe busy; ncu showed the pipe 53 % busy overall against ~75 %
// while a block is inside its stencil (profiles/r01_dw7x7_full.csv).  Here a CTA stays resident (grid = MINB CTAs per SM),
// walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... and keeps TWO tile buffers: while the threads run the stencil and the
// LayerNorm of tile i out of buffer i & 1, one thread has already asked TMA for tile i + 1 in the other buffer (the request
// is issued before the stencil starts, so it has a whole tile's worth of arithmetic to land).  The load phase disappears
// from the critical path, the per-block launch / barrier-init / bias+parameter prologue is paid once per CTA instead of once
// per tile, and the LayerNorm parameters stay in registers across tiles.  The stencil results are parked (fp32) in the
// buffer the tile was read from, exactly as above.
template <int CQ, int S, int P, int MINB>
__global__ void __launch_bounds__(CQ * S, MINB)
dw7x7_ln_persist_kernel(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ wgt, const float* __restrict__ bias,
                        const float* __restrict__ ln_w, const float* __restrict__ ln_b, void* __restrict__ y, int out_bf16, int H,
                        int W, int tiles_x, int tiles_y, int total_tiles, float eps) {
  using TI = __nv_bfloat16;
  constexpr int C = 4 * CQ;
  using Geo = DwGeom<TI, CQ, S, P>;
  constexpr int TW = Geo::TW, TH = Geo::TH, TWP = Geo::TWP, kBoxC = Geo::kBoxC, kBoxes = Geo::kBoxes;
  static_assert(Geo::SW == 1, "bf16 tiles use one strip per CQ consecutive threads");
  constexpr int kThreads = CQ * S;
  constexpr uint32_t kBufBytes = (static_cast<uint32_t>(Geo::tile_bytes) + 127u) & ~127u;
  static_assert(static_cast<size_t>(S) * P * C * sizeof(float) <= Geo::tile_bytes, "the result tile re-uses the input tile's buffer");
  extern __shared__ __align__(128) uint8_t dw_smem[];
  __shared__ __align__(8) unsigned long long tma_bar[2];

  const int q = threadIdx.x % CQ, s = threadIdx.x / CQ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar0 = tc::smem_u32(&tma_bar[0]);
  const uint32_t smem0 = tc::smem_u32(dw_smem);
  const int per_frame = tiles_x * tiles_y;

  auto issue = [&](int tile, int buf) {   // one thread: the (S+6) x (P+6) x C tile with its zero halo, one TMA box per <=256 channels
    const int n = tile / per_frame, r = tile - n * per_frame;
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    const uint32_t bar = bar0 + 8u * buf;
    tc::mbar_expect_tx(bar, static_cast<uint32_t>(Geo::tile_bytes));
#pragma unroll
    for (int b = 0; b < kBoxes; ++b)
      tc::tma_load_4d(smem0 + buf * kBufBytes + static_cast<uint32_t>(b) * TH * TWP * kBoxC * sizeof(TI), &map_x, bar, b * kBoxC,
                      tx * P - 3, ty * S - 3, n);
  };

  if (threadIdx.x == 0) {
    tc::mbar_init(bar0, 1);
    tc::mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (static_cast<int>(blockIdx.x) < total_tiles) issue(blockIdx.x, 0);
  }
  // per-thread constants that survive the tile loop: bias (accumulator seed), LayerNorm affine parameters
  const float4 bq = __ldg(reinterpret_cast<const float4*>(bias) + q);
  constexpr int nwarps = kThreads / 32;
  constexpr int pairs = C / 2;
  constexpr int LPP = (pairs % 32 == 0) ? 32 : 16;   // lanes per pixel
  static_assert(pairs % LPP == 0, "channel pairs must tile the lane group");
  constexpr int NP = pairs / LPP;                     // pairs per lane
  constexpr int GRP = 32 / LPP;                       // pixels handled side by side in one warp
  constexpr int G = ((S * P) % (2 * GRP) == 0) ? 2 : 4;
  static_assert((S * P) % (G * GRP) == 0, "pixel groups");
  const int sub = lane / LPP, sl = lane % LPP;
  float2 gam[NP], bet[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    gam[i] = __ldg(reinterpret_cast<const float2*>(ln_w) + sl + LPP * i);
    bet[i] = __ldg(reinterpret_cast<const float2*>(ln_b) + sl + LPP * i);
  }
  const float4* wq = reinterpret_cast<const float4*>(wgt) + q;
  __syncthreads();   // barrier inits are visible to every waiter

  int it = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    // the other buffer was last read by the LayerNorm of the previous tile (closed by the __syncthreads at the loop's end)
    if (threadIdx.x == 0 && tile + static_cast<int>(gridDim.x) < total_tiles) issue(tile + gridDim.x, buf ^ 1);
    const TI* tile_s = reinterpret_cast<const TI*>(dw_smem + buf * kBufBytes);
    float* out_s = reinterpret_cast<float*>(dw_smem + buf * kBufBytes);
    const int n = tile / per_frame, r = tile - n * per_frame;
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    const int x0 = tx * P, y0 = ty * S;

    F2 acc[P][2];
#pragma unroll
    for (int j = 0; j < P; ++j) { acc[j][0] = pack2(bq.x, bq.y); acc[j][1] = pack2(bq.z, bq.w); }
    tc::mbar_wait(bar0 + 8u * buf, static_cast<uint32_t>(it >> 1) & 1u);

#pragma unroll 1
    for (int kh = 0; kh < 7; ++kh) {
      F2 w[7][2];
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const float4 t = __ldg(wq + (kh * 7 + kw) * CQ);
        w[kw][0] = pack2(t.x, t.y);
        w[kw][1] = pack2(t.z, t.w);
      }
      const TI* trow = tile_s + (static_cast<size_t>((4 * q) / kBoxC) * TH + s + kh) * TWP * kBoxC + (4 * q) % kBoxC;
#pragma unroll
      for (int ix = 0; ix < TW; ++ix) {
        const float4 v = Ld4<TI>::lds(trow + ix * kBoxC);
        const F2 v0 = pack2(v.x, v.y), v1 = pack2(v.z, v.w);
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
          const int j = ix - kw;
          if (j >= 0 && j < P) {
            acc[j][0] = fma2(v0, w[kw][0], acc[j][0]);
            acc[j][1] = fma2(v1, w[kw][1], acc[j][1]);
          }
        }
      }
    }
    __syncthreads();  // everyone is done reading the tile: its memory becomes the result buffer
#pragma unroll
    for (int j = 0; j < P; ++j) {
      float4 o;
      unpack2(acc[j][0], o.x, o.y);
      unpack2(acc[j][1], o.z, o.w);
      reinterpret_cast<float4*>(out_s + static_cast<size_t>(s * P + j) * C)[q] = o;
    }
    __syncthreads();

    const long long tile_base = ((static_cast<long long>(n) * H + y0) * W + x0) * C;
    for (int p0 = (warp * GRP + sub) * G; p0 < S * P; p0 += nwarps * GRP * G) {
      float2 v[G][NP];
      float sum[G], sq[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float2* src = reinterpret_cast<const float2*>(out_s + static_cast<size_t>(p0 + g) * C);
        sum[g] = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          v[g][i] = src[sl + LPP * i];
          sum[g] += v[g][i].x + v[g][i].y;
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) sum[g] += __shfl_xor_sync(0xffffffffu, sum[g], o);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        sum[g] *= (1.f / C);  // mean
        sq[g] = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const float a = v[g][i].x - sum[g], b = v[g][i].y - sum[g];
          sq[g] += a * a + b * b;
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) sq[g] += __shfl_xor_sync(0xffffffffu, sq[g], o);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float rs = rsqrtf(sq[g] * (1.f / C) + eps), ms = -sum[g] * rs;
        const int pix = p0 + g;
        const int ly = pix / P, lx = pix % P;
        if (y0 + ly >= H || x0 + lx >= W) continue;
        const unsigned off = static_cast<unsigned>(ly * W + lx) * static_cast<unsigned>(C);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int p = sl + LPP * i;
          const float a = fmaf(fmaf(v[g][i].x, rs, ms), gam[i].x, bet[i].x);
          const float b = fmaf(fmaf(v[g][i].y, rs, ms), gam[i].y, bet[i].y);
          if (out_bf16)
            reinterpret_cast<__nv_bfloat162*>(static_cast<__nv_bfloat16*>(y) + tile_base + off)[p] = __floats2bfloat162_rn(a, b);
          else
            reinterpret_cast<float2*>(static_cast<float*>(y) + tile_base + off)[p] = make_float2(a, b);
        }
      }
    }
    // this buffer is refilled by TMA (async proxy) at the top of the next iteration: order the generic-proxy accesses first
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
}

template <int CQ, int S, int P, int MINB>
int launch_dw7x7_persist(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias, const float* ln_w,
                         const float* ln_b, void* y, cudaStream_t stream) {
  using TI = __nv_bfloat16;
  constexpr int C = 4 * CQ;
  using Geo = DwGeom<TI, CQ, S, P>;
  constexpr size_t buf_bytes = (Geo::tile_bytes + 127) & ~static_cast<size_t>(127);
  constexpr size_t smem = 2 * buf_bytes;
  static_assert(MINB * (smem + 1024 + 64) <= 228 * 1024, "MINB persistent CTAs per SM must fit in shared memory");
  const int tiles_x = (d->w + P - 1) / P, tiles_y = (d->h + S - 1) / S;
  const long long frames = static_cast<long long>(d->n) * d->t;
  const long long total = frames * tiles_x * tiles_y;
  MSPI_CHECK_ARG(total < (1ll << 31), "dwconv 7x7: %lld tiles", total);
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) return 1;
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const cuuint64_t es = sizeof(TI);
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h),
                        static_cast<cuuint64_t>(frames)};
  cuuint64_t gstr[3] = {C * es, static_cast<cuuint64_t>(d->w) * C * es, static_cast<cuuint64_t>(d->h) * d->w * C * es};
  cuuint32_t bdim[4] = {static_cast<cuuint32_t>(Geo::kBoxC), static_cast<cuuint32_t>(Geo::TWP), static_cast<cuuint32_t>(S + 6), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;   // fall back to the one-tile-per-block kernel
  auto kern = dw7x7_ln_persist_kernel<CQ, S, P, MINB>;
  MSPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  long long grid = static_cast<long long>(MINB) * num_sms();
  if (grid > total) grid = total;
  kern<<<static_cast<unsigned>(grid), CQ * S, smem, stream>>>(map, wgt, bias, ln_w, ln_b, y, d->out_dtype == MSPI_BF16 ? 1 : 0, d->h,
                                                             d->w, tiles_x, tiles_y, static_cast<int>(total), d->ln_eps);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-row register blocking (bf16 whole-pixel tiles).  In the kernels above a thread owns ONE output row strip: per filter
// row it issues 14 LDS.64 + 56 unpack ops + 7 weight LDG.128 for 112 FFMA2, i.e. about one non-FMA instruction per packed
// FMA, and ncu shows every pipe of the SM around 55-70 % busy at once (FMA 53 %, LSU wavefronts 71 %, issue 60 %,
// profiles/r01_dw7x7_full.csv): the kernel is bound by its instruction mix, not by one pipe.  Here a thread owns the strips
// of TWO consecutive output rows (same 4 channels, same P columns): input row r of its window feeds filter row r of the upper
// strip and filter row r-1 of the lower one, so each loaded + unpacked input value and each weight vector is used twice —
// 8 x (14 LDS + 56 unpack) + 49 LDG per 1568 FFMA2 instead of 14 x (...) + 98.  The 8-step input-row loop is fully unrolled
// so that the two live filter rows rotate through registers by renaming.
__device__ unsigned long long g_dw_phase_cycles[8];   // profiling aid (mspi_debug_dw_phase_cycles): load wait, stencil, LayerNorm, blocks

template <int CQ, int SP, int P, int MINB, bool DBG = false, int LNV = 1, bool PACKED = true>
__global__ void __launch_bounds__(CQ * SP, MINB)
dw7x7_ln_r2_kernel(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ wgt, const float* __restrict__ bias,
                   const float* __restrict__ ln_w, const float* __restrict__ ln_b, void* __restrict__ y, int out_bf16, int H, int W,
                   int n0, float eps) {
  using TI = __nv_bfloat16;
  constexpr int C = 4 * CQ, S = 2 * SP;
  using Geo = DwGeom<TI, CQ, S, P>;
  constexpr int TW = Geo::TW, TH = Geo::TH, TWP = Geo::TWP, kBoxC = Geo::kBoxC, kBoxes = Geo::kBoxes;
  constexpr int kThreads = CQ * SP;
  static_assert(kThreads % 32 == 0, "whole warps");
  static_assert(static_cast<size_t>(S) * P * C * sizeof(float) <= Geo::tile_bytes, "the result tile re-uses the input tile's buffer");
  extern __shared__ __align__(128) uint8_t dw_smem[];
  __shared__ __align__(8) unsigned long long tma_bar;
  const TI* tile_s = reinterpret_cast<const TI*>(dw_smem);
  float* out_s = reinterpret_cast<float*>(dw_smem);
  const int tx = blockIdx.x, ty = blockIdx.y, n = n0 + blockIdx.z;
  const int x0 = tx * P, y0 = ty * S;
  const uint32_t bar = tc::smem_u32(&tma_bar);
  pdl_launch_dependents();
  pdl_wait();   // programmatic dependent launch (common.cuh): the blocks of the first wave are resident before the previous kernel ends
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc::mbar_expect_tx(bar, static_cast<uint32_t>(Geo::tile_bytes));
#pragma unroll
    for (int b = 0; b < kBoxes; ++b)
      tc::tma_load_4d(tc::smem_u32(tile_s + static_cast<size_t>(b) * TH * TWP * kBoxC), &map_x, bar, b * kBoxC, x0 - 3, y0 - 3, n);
  }
  const int q = threadIdx.x % CQ, sp = threadIdx.x / CQ;
  // PACKED: accumulators / taps / inputs as fp32 pairs on fma.rn.f32x2; otherwise scalar FFMA.  A packed FMA whose two
  // non-reused operands are distinct 64-bit register pairs (tap, accumulator) occupies the scheduler ~3.2 cycles against
  // ~1.3 for a scalar FFMA with one reused operand (tools/issue_model.cu), so two scalar FMAs can be the cheaper form.
  F2 acc[2][P][2];
  float accs[2][P][4];
  const float4 bq = __ldg(reinterpret_cast<const float4*>(bias) + q);
#pragma unroll
  for (int j = 0; j < P; ++j) {
    if constexpr (PACKED) {
      acc[0][j][0] = acc[1][j][0] = pack2(bq.x, bq.y);
      acc[0][j][1] = acc[1][j][1] = pack2(bq.z, bq.w);
    } else {
      accs[0][j][0] = accs[1][j][0] = bq.x; accs[0][j][1] = accs[1][j][1] = bq.y;
      accs[0][j][2] = accs[1][j][2] = bq.z; accs[0][j][3] = accs[1][j][3] = bq.w;
    }
  }
  const float4* wq = reinterpret_cast<const float4*>(wgt) + q;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (DBG) t0 = clock64();
  __syncthreads();   // the barrier init is visible to every waiter
  tc::mbar_wait(bar, 0);
  if (DBG) t1 = clock64();

  const TI* tcol = tile_s + (static_cast<size_t>((4 * q) / kBoxC) * TH + 2 * sp) * TWP * kBoxC + (4 * q) % kBoxC;
  F2 wa[7][2], wb[7][2];   // filter row r (upper strip) / r - 1 (lower strip)
  float4 was[7], wbs[7];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (r < 7) {
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const float4 t = __ldg(wq + (r * 7 + kw) * CQ);
        if constexpr (PACKED) {
          wa[kw][0] = pack2(t.x, t.y);
          wa[kw][1] = pack2(t.z, t.w);
        } else {
          was[kw] = t;
        }
      }
    }
    const TI* trow = tcol + static_cast<size_t>(r) * TWP * kBoxC;
#pragma unroll
    for (int ix = 0; ix < TW; ++ix) {
      const float4 v = Ld4<TI>::lds(trow + ix * kBoxC);
      const F2 v0 = pack2(v.x, v.y), v1 = pack2(v.z, v.w);
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const int j = ix - kw;
        if (j >= 0 && j < P) {
          if constexpr (PACKED) {
            if (r < 7) {
              acc[0][j][0] = fma2(v0, wa[kw][0], acc[0][j][0]);
              acc[0][j][1] = fma2(v1, wa[kw][1], acc[0][j][1]);
            }
            if (r > 0) {
              acc[1][j][0] = fma2(v0, wb[kw][0], acc[1][j][0]);
              acc[1][j][1] = fma2(v1, wb[kw][1], acc[1][j][1]);
            }
          } else {
            if (r < 7) {
              accs[0][j][0] = fmaf(v.x, was[kw].x, accs[0][j][0]); accs[0][j][1] = fmaf(v.y, was[kw].y, accs[0][j][1]);
              accs[0][j][2] = fmaf(v.z, was[kw].z, accs[0][j][2]); accs[0][j][3] = fmaf(v.w, was[kw].w, accs[0][j][3]);
            }
            if (r > 0) {
              accs[1][j][0] = fmaf(v.x, wbs[kw].x, accs[1][j][0]); accs[1][j][1] = fmaf(v.y, wbs[kw].y, accs[1][j][1]);
              accs[1][j][2] = fmaf(v.z, wbs[kw].z, accs[1][j][2]); accs[1][j][3] = fmaf(v.w, wbs[kw].w, accs[1][j][3]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int kw = 0; kw < 7; ++kw) {
      if constexpr (PACKED) { wb[kw][0] = wa[kw][0]; wb[kw][1] = wa[kw][1]; } else { wbs[kw] = was[kw]; }
    }
  }
  long long t2b = 0, t2c = 0;
  if (DBG) t2 = clock64();
  __syncthreads();  // everyone is done reading the tile: its memory becomes the result buffer
  if (DBG) t2b = clock64();
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int j = 0; j < P; ++j) {
      float4 o;
      if constexpr (PACKED) {
        unpack2(acc[rr][j][0], o.x, o.y);
        unpack2(acc[rr][j][1], o.z, o.w);
      } else {
        o = make_float4(accs[rr][j][0], accs[rr][j][1], accs[rr][j][2], accs[rr][j][3]);
      }
      reinterpret_cast<float4*>(out_s + static_cast<size_t>((2 * sp + rr) * P + j) * C)[q] = o;
    }
  __syncthreads();
  if (DBG) t2c = clock64();

  const long long tile_base = ((static_cast<long long>(n) * H + y0) * W + x0) * C;
#ifndef MSPI_DW_R2_LN_G
#define MSPI_DW_R2_LN_G 4
#endif
  if constexpr (LNV == 0) dw_layernorm_store<C, S * P, P, kThreads / 32, 2, true>(out_s, ln_w, ln_b, y, out_bf16, tile_base, y0, x0, H, W, eps);
  else if constexpr (LNV == 2) dw_layernorm_store<C, S * P, P, kThreads / 32, 2>(out_s, ln_w, ln_b, y, out_bf16, tile_base, y0, x0, H, W, eps);
  else dw_layernorm_store<C, S * P, P, kThreads / 32, MSPI_DW_R2_LN_G>(out_s, ln_w, ln_b, y, out_bf16, tile_base, y0, x0, H, W, eps);
  if (DBG && (threadIdx.x & 31) == 0) {   // one sample per warp
    const long long t3 = clock64();
    atomicAdd(&g_dw_phase_cycles[0], static_cast<unsigned long long>(t1 - t0));     // tile wait
    atomicAdd(&g_dw_phase_cycles[1], static_cast<unsigned long long>(t2 - t1));     // stencil
    atomicAdd(&g_dw_phase_cycles[2], static_cast<unsigned long long>(t3 - t2c));    // LayerNorm + stores
    atomicAdd(&g_dw_phase_cycles[3], 1ull);
    atomicAdd(&g_dw_phase_cycles[4], static_cast<unsigned long long>(t2b - t2));    // barrier after the stencil
    atomicAdd(&g_dw_phase_cycles[5], static_cast<unsigned long long>(t2c - t2b));   // result tile store + barrier
  }
}

template <int CQ, int SP, int P, int MINB>
int launch_dw7x7_r2(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias, const float* ln_w, const float* ln_b,
                    void* y, cudaStream_t stream) {
  using TI = __nv_bfloat16;
  constexpr int C = 4 * CQ, S = 2 * SP;
  using Geo = DwGeom<TI, CQ, S, P>;
  constexpr size_t smem = Geo::tile_bytes;
  static_assert(MINB * (smem + 1024 + 64) <= 228 * 1024, "MINB blocks per SM must fit in shared memory");
  const int tiles_x = (d->w + P - 1) / P, tiles_y = (d->h + S - 1) / S;
  const long long frames = static_cast<long long>(d->n) * d->t;
  MSPI_CHECK_ARG(frames < (1ll << 31) && tiles_y <= 65535, "dwconv 7x7: grid out of range (%lld frames)", frames);
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) return 1;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const cuuint64_t es = sizeof(TI);
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h),
                        static_cast<cuuint64_t>(frames)};
  cuuint64_t gstr[3] = {C * es, static_cast<cuuint64_t>(d->w) * C * es, static_cast<cuuint64_t>(d->h) * d->w * C * es};
  cuuint32_t bdim[4] = {static_cast<cuuint32_t>(Geo::kBoxC), static_cast<cuuint32_t>(Geo::TWP), static_cast<cuuint32_t>(S + 6), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;   // fall back to the one-row kernel
  static const bool dbg = getenv("MSPI_DW_DEBUG") != nullptr;   // per-phase cycle counters of block-thread 0 (study aid)
  // LayerNorm variant (study aid): 0 two butterflies, 1 merged butterfly with 4 pixels in flight (slower: the stencil loses
  // registers), 2 (default) merged butterfly with 2 pixels in flight
  static const int lnv = [] { const char* e = getenv("MSPI_DW_LNV"); return e ? atoi(e) : 2; }();
  static const bool scalar = [] { const char* e = getenv("MSPI_DW_SCALAR"); return e && atoi(e) != 0; }();   // scalar FFMA stencil
  auto kern = dbg ? (lnv == 0 ? dw7x7_ln_r2_kernel<CQ, SP, P, MINB, true, 0> : dw7x7_ln_r2_kernel<CQ, SP, P, MINB, true, 2>)
                  : scalar ? dw7x7_ln_r2_kernel<CQ, SP, P, MINB, false, 2, false>
                  : lnv == 0 ? dw7x7_ln_r2_kernel<CQ, SP, P, MINB, false, 0>
                  : lnv == 2 ? dw7x7_ln_r2_kernel<CQ, SP, P, MINB, false, 2> : dw7x7_ln_r2_kernel<CQ, SP, P, MINB, false, 1>;
  MSPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  for (long long n0 = 0; n0 < frames; n0 += 65535) {
    const unsigned nz = static_cast<unsigned>(frames - n0 < 65535 ? frames - n0 : 65535);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles_x, tiles_y, nz);
    cfg.blockDim = dim3(CQ * SP, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(&attr[0]);
    MSPI_CUDA(cudaLaunchKernelEx(&cfg, kern, map, wgt, bias, ln_w, ln_b, y, d->out_dtype == MSPI_BF16 ? 1 : 0, d->h, d->w,
                                 static_cast<int>(n0), d->ln_eps));
    MSPI_LAUNCH_CHECK();
  }
  return MSPI_OK;
}

// (kt,1,1) temporal depthwise conv, T <= 8 frames: a thread owns 4 channels of one (h,w) position for all T
// frames of one sample; every input is read once.  HBM bound.
template <typename TI, int MAXT>
__global__ void dwt_kernel(const TI* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                           void* __restrict__ y, int out_bf16, long long n_hw, int T, int HW, int C, int kt) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int cq = C >> 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_hw * cq) return;
  // 32-bit divisions whenever the index fits (common.cuh divmod): four 64-bit divisions were half of this kernel's instructions
  long long r = idx;
  const int q = divmod(r, cq);
  const long long hw = divmod(r, HW);   // r: n*HW + hw -> n
  const long long n = r;
  const long long base = (n * T * HW + hw) * C + 4 * q;
  const long long tstride = static_cast<long long>(HW) * C;
  float4 v[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t)
    if (t < T) v[t] = Ld4<TI>::ld(x + base + t * tstride);
  const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + q);
  const int pt = kt / 2;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    if (t >= T) break;
    float4 a = b;
#pragma unroll
    for (int u = 0; u < MAXT; ++u) {
      const int k = u - t + pt;  // tap that connects input frame u to output frame t
      if (u < T && k >= 0 && k < kt) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wgt + k * C) + q);
        a.x = fmaf(v[u].x, w.x, a.x);
        a.y = fmaf(v[u].y, w.y, a.y);
        a.z = fmaf(v[u].z, w.z, a.z);
        a.w = fmaf(v[u].w, w.w, a.w);
      }
    }
    if (out_bf16) {
      uint2 o;
      o.x = pack_bf16x2(a.x, a.y);
      o.y = pack_bf16x2(a.z, a.w);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + base + t * tstride) = o;
    } else {
      *reinterpret_cast<float4*>(static_cast<float*>(y) + base + t * tstride) = a;
    }
  }
}

template <typename TI, int CQ, int S, int P, bool GROUPED = false>
int launch_dw7x7(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias, const float* ln_w,
                 const float* ln_b, void* y, cudaStream_t stream, int groups = 1) {
  constexpr int C = 4 * CQ;
  using Geo = DwGeom<TI, CQ, S, P>;
  constexpr size_t tile_bytes = Geo::tile_bytes;
  constexpr size_t out_bytes = static_cast<size_t>(S) * P * C * sizeof(float);
  constexpr size_t smem = tile_bytes > out_bytes ? tile_bytes : out_bytes;
  static_assert(smem <= 113 * 1024, "two blocks per SM must fit");
  const int tiles_x = (d->w + P - 1) / P, tiles_y = (d->h + S - 1) / S;
  const long long frames = static_cast<long long>(d->n) * d->t;
  MSPI_CHECK_ARG(frames < (1ll << 31) && tiles_y <= 65535, "dwconv 7x7: grid out of range (%lld frames)", frames);
  auto kern = dw7x7_ln_kernel<TI, CQ, S, P, GROUPED>;
  MSPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // tile load by TMA (one 4-D box [C, P+6, S+6, 1] per block, out-of-image coordinates zero-filled) where the channel
  // group fits a box; MSPI_DW_TMA=0 keeps the per-thread cp.async path
  static const bool tma_on = [] { const char* e = getenv("MSPI_DW_TMA"); return !e || atoi(e) != 0; }();
  const int ctot = C * groups;
  constexpr int kBoxC = Geo::kBoxC;
  int use_tma = tma_on && kBoxC <= 256 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ctot * sizeof(TI)) % 16 == 0;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (use_tma) {
    tc::EncodeTiledFn encode = tc::get_encode_fn();
    if (!encode) {
      use_tma = 0;
    } else {
      const cuuint64_t es = sizeof(TI);
      cuuint64_t gdim[4] = {static_cast<cuuint64_t>(ctot), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h),
                            static_cast<cuuint64_t>(d->n) * d->t};
      cuuint64_t gstr[3] = {ctot * es, static_cast<cuuint64_t>(d->w) * ctot * es,
                            static_cast<cuuint64_t>(d->h) * d->w * ctot * es};
      cuuint32_t bdim[4] = {static_cast<cuuint32_t>(kBoxC), static_cast<cuuint32_t>(Geo::TWP), static_cast<cuuint32_t>(S + 6), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&map, sizeof(TI) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                          const_cast<void*>(x), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) use_tma = 0;
    }
  }
  for (long long n0 = 0; n0 < frames; n0 += 65535) {   // grid.z limit
    const unsigned nz = static_cast<unsigned>(frames - n0 < 65535 ? frames - n0 : 65535);
    MSPI_CUDA(launch_pdl(kern, dim3(tiles_x * groups, tiles_y, nz), CQ * S, smem, stream,
                         map, static_cast<const TI*>(x), wgt, bias, ln_w, ln_b, y, d->out_dtype == MSPI_BF16 ? 1 : 0, d->h, d->w, tiles_x,
                         static_cast<int>(n0), d->ln_eps, C * groups, use_tma));
    MSPI_LAUNCH_CHECK();
  }
  return MSPI_OK;
}

// Wide, small maps (ConvNeXt stages 2 and 3: 384 ch at 14x24, 768 ch at 7x12 for the default clip).  A block that owned
// whole pixels could only hold a 2x8-pixel tile (7x halo redundancy, latency bound: 0.38 ms for a 132 MB tensor), so here
// blocks own 192-channel groups of 7x12-pixel tiles (2.8x redundancy, 2352 FMA per thread per tile load) and write the
// raw convolution; LayerNorm over the full channel vector follows as its own (in-place, HBM-bound) kernel.
template <typename TI>
int launch_dw7x7_grouped(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias, const float* ln_w,
                         const float* ln_b, void* y, cudaStream_t stream) {
  const int rc = launch_dw7x7<TI, 48, 7, 12, true>(d, x, wgt, bias, nullptr, nullptr, y, stream, d->c / 192);
  if (rc != MSPI_OK || ln_w == nullptr) return rc;
  MspiLnDesc ln;
  ln.rows = static_cast<int64_t>(d->n) * d->t * d->h * d->w;
  ln.c = d->c;
  ln.in_rstride = ln.out_rstride = d->c;
  ln.in_dtype = ln.out_dtype = d->out_dtype;
  ln.eps = d->ln_eps;
  ln.relu = 0;
  ln.pos_rows = 0;
  ln.rows_per_group = ln.rows;
  ln.out_gstride = 0;
  return mspi_layernorm(&ln, y, ln_w, ln_b, nullptr, y, stream);
}

}  // namespace

// Fast paths of mspi_dwconv_ln (norm_attn.cu holds the generic kernel).  Returns 1 if the shape is not covered.
int dwconv_fast_path(const MspiDwDesc* d, const void* x, const float* wgt, const float* bias, const float* ln_w,
                     const float* ln_b, void* y, cudaStream_t stream) {
  const bool aligned = d->c % 8 == 0 && d->c <= 768;
  if (!aligned) return 1;
  if (d->kt == 1 && d->kh == 7 && d->kw == 7) {
    using bf = __nv_bfloat16;
    // strips of 8 pixels: 78 registers -> 4 blocks (24 warps) per SM.  Strips of 16 feed more FMAs per shared-memory load
    // (P+6 loads for 7P FMA groups) but need 118 registers (2 blocks per SM), and since the stencil runs on packed FFMA2 the
    // kernel is bound by how well the load / stencil / LayerNorm phases of different blocks overlap, not by instruction count
    // (stage 0: 0.95 ms with 16, 0.85 ms with 8).  MSPI_DW_P16=1 restores the wide strips (tuning aid).
    static const bool allow16 = getenv("MSPI_DW_P16") != nullptr;
    const bool wide = d->w % 16 == 0 && allow16;
    // MSPI_DW_R2 (default 1): two output rows per thread (half the loads / unpacks / weight fetches per FMA)
    static const int r2 = [] { const char* e = getenv("MSPI_DW_R2"); return e ? atoi(e) : 5; }();   // bit 2: C = 384 as well
    if (r2 && d->in_dtype == MSPI_BF16 && ln_w != nullptr && ln_b != nullptr && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      int rc = 1;
      if (d->c == 96) rc = (r2 & 2) ? launch_dw7x7_r2<24, 4, 8, 5>(d, x, wgt, bias, ln_w, ln_b, y, stream)
                                    : launch_dw7x7_r2<24, 4, 8, 4>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      else if (d->c == 192) rc = launch_dw7x7_r2<48, 2, 8, 4>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      else if (d->c == 384 && (r2 & 4)) rc = launch_dw7x7_r2<96, 2, 8, 2>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      if (rc != 1) return rc;
    }
    // MSPI_DW_PERSIST=1: persistent double-buffered kernel (measured slower: 18 instead of 24 warps per SM and 225 KB of
    // shared memory leave no L1 for the weight vectors; 0.79 -> 0.99 ms at stage 0) — kept as a study switch, default off
    static const bool persist = [] { const char* e = getenv("MSPI_DW_PERSIST"); return e && atoi(e) != 0; }();
    if (persist && d->in_dtype == MSPI_BF16 && ln_w != nullptr && ln_b != nullptr && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      int rc = 1;
      if (d->c == 96) rc = launch_dw7x7_persist<24, 8, 8, 3>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      else if (d->c == 192) rc = launch_dw7x7_persist<48, 4, 8, 2>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      if (rc != 1) return rc;
    }
    if (d->in_dtype == MSPI_BF16) {
      if (d->c == 96) return wide ? launch_dw7x7<bf, 24, 8, 16>(d, x, wgt, bias, ln_w, ln_b, y, stream)
                                  : launch_dw7x7<bf, 24, 8, 8>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      if (d->c == 192) return wide ? launch_dw7x7<bf, 48, 4, 16>(d, x, wgt, bias, ln_w, ln_b, y, stream)
                                   : launch_dw7x7<bf, 48, 4, 8>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      // C = 384: the fused 2x8-tile kernel (0.38 ms at 14x24) still beats grouped stencil + LayerNorm kernel (0.45 ms);
      // C = 768 has no whole-pixel tile that fits: grouped (0.50 -> 0.21 ms at 7x12)
      static const bool grp384 = getenv("MSPI_DW_GROUP384") != nullptr;  // tuning aid
      // C = 384 whole-pixel tiles: 4x8 pixels on 384 threads (107 KB tile, two blocks = 24 warps per SM, 4.4 loaded pixels
      // per output) or, MSPI_DW_384_S4=0, 2x8 pixels on 192 threads (86 KB, 12 warps per SM, 7 loaded pixels per output)
      static const bool s4_384 = [] { const char* e = getenv("MSPI_DW_384_S4"); return !e || atoi(e) != 0; }();
      if (d->c == 384 && !grp384 && s4_384) return launch_dw7x7<bf, 96, 4, 8>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      if (d->c == 384 && !grp384) return launch_dw7x7<bf, 96, 2, 8>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      if ((d->c == 384 || d->c == 768) && d->out_dtype == MSPI_BF16)
        return launch_dw7x7_grouped<bf>(d, x, wgt, bias, ln_w, ln_b, y, stream);
    } else {
      if (d->c == 192) return launch_dw7x7<float, 48, 4, 8>(d, x, wgt, bias, ln_w, ln_b, y, stream);
    }
    return 1;
  }
  if (d->kh == 1 && d->kw == 1 && d->t <= 8 && ln_w == nullptr) {
    const int HW = d->h * d->w;
    const long long n_hw = static_cast<long long>(d->n) * HW;
    const long long total = n_hw * (d->c / 4);
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    MSPI_CHECK_ARG(blocks < (1ll << 31), "dwconv_t: grid out of range");
    const int ob = d->out_dtype == MSPI_BF16 ? 1 : 0;
    // T <= 4 (the decoder's laterals at the default clip length): the 4-frame instance keeps half the registers of the
    // 8-frame one (82 registers left 22 % of the warp slots occupied on a kernel that only waits for HBM)
    if (d->in_dtype == MSPI_BF16) {
      if (d->t <= 4)
        MSPI_CUDA(launch_pdl(dwt_kernel<__nv_bfloat16, 4>, static_cast<int>(blocks), threads, 0, stream,
                             static_cast<const __nv_bfloat16*>(x), wgt, bias, y, ob, n_hw, d->t, HW, d->c, d->kt));
      else
        MSPI_CUDA(launch_pdl(dwt_kernel<__nv_bfloat16, 8>, static_cast<int>(blocks), threads, 0, stream,
                             static_cast<const __nv_bfloat16*>(x), wgt, bias, y, ob, n_hw, d->t, HW, d->c, d->kt));
    } else {
      if (d->t <= 4)
        MSPI_CUDA(launch_pdl(dwt_kernel<float, 4>, static_cast<int>(blocks), threads, 0, stream, static_cast<const float*>(x), wgt,
                             bias, y, ob, n_hw, d->t, HW, d->c, d->kt));
      else
        MSPI_CUDA(launch_pdl(dwt_kernel<float, 8>, static_cast<int>(blocks), threads, 0, stream, static_cast<const float*>(x), wgt,
                             bias, y, ob, n_hw, d->t, HW, d->c, d->kt));
    }
    MSPI_LAUNCH_CHECK();
    return MSPI_OK;
  }
  return 1;
}

}  // namespace mspi

extern "C" int mspi_debug_dw_phase_cycles(uint64_t* out4, int reset) {  // out4: 8 entries
  using namespace mspi;
  unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  MSPI_CUDA(cudaDeviceSynchronize());
  if (out4 != nullptr) {
    MSPI_CUDA(cudaMemcpyFromSymbol(h, g_dw_phase_cycles, sizeof(h)));
    for (int i = 0; i < 8; ++i) out4[i] = h[i];
  }
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    MSPI_CUDA(cudaMemcpyToSymbol(g_dw_phase_cycles, z, sizeof(z)));
  }
  return MSPI_OK;
}
