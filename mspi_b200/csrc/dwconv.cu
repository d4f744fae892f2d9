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
__global__ void __launch_bounds__(CQ * S, (CQ * S <= 192 && P == 16) ? MSPI_DW_MINB : (CQ * S <= 192 && P == 8 && CQ <= 48 && sizeof(TI) == 2) ? MSPI_DW_MINB8 : 1)
dw7x7_ln_kernel(const __grid_constant__ CUtensorMap map_x, const TI* __restrict__ x, const float* __restrict__ wgt,
                const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                void* __restrict__ y, int out_bf16, int H, int W, int tiles_x, int n0, float eps, int cs_arg, int use_tma) {
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

// (kt,1,1) temporal depthwise conv, T <= 8 frames: a thread owns 4 channels of one (h,w) position for all T
// frames of one sample; every input is read once.  HBM bound.
template <typename TI, int MAXT>
__global__ void dwt_kernel(const TI* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                           void* __restrict__ y, int out_bf16, long long n_hw, int T, int HW, int C, int kt) {
  const int cq = C >> 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_hw * cq) return;
  const int q = static_cast<int>(idx % cq);
  const long long pos = idx / cq;  // n*HW + hw
  const long long n = pos / HW, hw = pos % HW;
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
    kern<<<dim3(tiles_x * groups, tiles_y, nz), CQ * S, smem, stream>>>(
        map, static_cast<const TI*>(x), wgt, bias, ln_w, ln_b, y, d->out_dtype == MSPI_BF16 ? 1 : 0, d->h, d->w, tiles_x,
        static_cast<int>(n0), d->ln_eps, C * groups, use_tma);
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
    if (d->in_dtype == MSPI_BF16) {
      if (d->c == 96) return wide ? launch_dw7x7<bf, 24, 8, 16>(d, x, wgt, bias, ln_w, ln_b, y, stream)
                                  : launch_dw7x7<bf, 24, 8, 8>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      if (d->c == 192) return wide ? launch_dw7x7<bf, 48, 4, 16>(d, x, wgt, bias, ln_w, ln_b, y, stream)
                                   : launch_dw7x7<bf, 48, 4, 8>(d, x, wgt, bias, ln_w, ln_b, y, stream);
      // C = 384: the fused 2x8-tile kernel (0.38 ms at 14x24) still beats grouped stencil + LayerNorm kernel (0.45 ms);
      // C = 768 has no whole-pixel tile that fits: grouped (0.50 -> 0.21 ms at 7x12)
      static const bool grp384 = getenv("MSPI_DW_GROUP384") != nullptr;  // tuning aid
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
    if (d->in_dtype == MSPI_BF16)
      dwt_kernel<__nv_bfloat16, 8><<<static_cast<int>(blocks), threads, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(x), wgt, bias, y, ob, n_hw, d->t, HW, d->c, d->kt);
    else
      dwt_kernel<float, 8><<<static_cast<int>(blocks), threads, 0, stream>>>(static_cast<const float*>(x), wgt, bias, y,
                                                                            ob, n_hw, d->t, HW, d->c, d->kt);
    MSPI_LAUNCH_CHECK();
    return MSPI_OK;
  }
  return 1;
}

}  // namespace mspi
