// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (conv_gemm.cu, fused_mlp.cu).  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace mspi {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// Bounded wait: a protocol bug must surface as a CUDA error (trap), never as a hung GPU.  The common case (phase already
// complete) is one try_wait; the clock is only read once a wait has actually failed.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > (1ll << 32)) __trap();
  }
}
// One lane of a fully converged warp (warp-uniform control flow around tcgen05 / TMA issue, instead of `if (lane == 0)`
// which makes ptxas wrap every uniform-datapath instruction in a per-lane loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (KIND == MSPI_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// GELU for bf16 results: the tanh form 0.5x(1 + tanh(sqrt(2/pi)(x + 0.044715x^3))).  Against the exact erf form it differs
// by < 5e-4 absolute (worst near |x| = 2.7), i.e. below the bf16 half-ulp of any result of magnitude > 0.25 and far below
// the bf16 noise of the fc2 sum the hidden activation feeds.
// The GELU epilogues are MUFU bound: with one scalar tanh.approx.f32 per element every variant measured 3.9 elements/clk/SM
// whatever the warp count (ex2 + rcp instead of tanh: 3.2).  gelu_pair therefore evaluates TWO elements per MUFU op with
// tanh.approx.f16x2 (argument and result in fp16: 2^-11 on a value in [-1,1], again < 5e-4 absolute on the result);
// everything else stays fp32.
// FMA-pipe-only alternative (MSPI_GELU_POLY): gelu(x) = x * (0.5 + h(x)), h(x) = 0.5 erf(x / sqrt 2) ~ xc * Q(xc^2) with
// xc = clamp(x, -4, 4), Q a degree-6 minimax polynomial (|error| < 1.9e-4 on the whole axis, less than half of the tanh
// form's); no MUFU and no f16 conversions, 16 FP32-pipe instructions per element.  Measured SLOWER than the f16x2-tanh form
// (s0.fc1 0.96 -> 1.25 ms, fused MLP 0.95 -> 1.05 ms): the GELU epilogues are bound by issue slots, not by the MUFU pipe.
__device__ __forceinline__ float gelu_poly(float x) {
  const float xc = fminf(fmaxf(x, -4.f), 4.f);
  const float t = xc * xc;
  float q = 2.278128089e-08f;
  q = fmaf(q, t, -1.598594353e-06f);
  q = fmaf(q, t, 4.79553645e-05f);
  q = fmaf(q, t, -0.0008140140042f);
  q = fmaf(q, t, 0.00877238132f);
  q = fmaf(q, t, -0.06457309567f);
  q = fmaf(q, t, 0.3978833387f);
  float h = xc * q;
  h = fabsf(x) >= 4.f ? copysignf(0.5f, x) : h;
  return fmaf(x, h, 0.5f * x);
}
#ifdef MSPI_GELU_POLY
__device__ __forceinline__ float gelu_bf16(float x) { return gelu_poly(x); }
__device__ __forceinline__ void gelu_pair(float& x0, float& x1) {
  x0 = gelu_poly(x0);
  x1 = gelu_poly(x1);
}
#else
__device__ __forceinline__ float gelu_bf16(float x) {
  const float u = x * x;
  const float t = x * fmaf(u, 0.0356774081f, 0.7978845608f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
// Two elements at a time on packed fp32 pairs (common.cuh F2): 3 packed ops for the tanh argument, one f16x2 convert, one
// MUFU for both tanh values, 2 packed ops for 0.5 x (1 + tanh).
__device__ __forceinline__ F2 gelu_pair_f2(F2 x) {
  const F2 u = mul2(x, x);
  const F2 t = mul2(x, fma2(u, pack2(0.0356774081f, 0.0356774081f), pack2(0.7978845608f, 0.7978845608f)));
  float t0, t1;
  unpack2(t, t0, t1);
  uint32_t tp, th;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(tp) : "f"(t1), "f"(t0));  // {hi, lo} = {t1, t0}
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(tp));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&th));
  const F2 hx = mul2(x, pack2(0.5f, 0.5f));
  return fma2(hx, pack2(f.x, f.y), hx);
}
__device__ __forceinline__ void gelu_pair(float& x0, float& x1) {
  const F2 o = gelu_pair_f2(pack2(x0, x1));
  unpack2(o, x0, x1);
}
// Same tanh-form GELU on a packed pair with fp32 tanh (one MUFU per element, no f16 round trip): 5 packed FP32-pipe ops +
// 2 MUFU per TWO elements instead of 6 + 1 per element; the two tanh results are written into the halves of a register pair.
__device__ __forceinline__ F2 gelu_pair_f32(F2 x) {
  const F2 u = mul2(x, x);
  const F2 t = mul2(x, fma2(u, pack2(0.0356774081f, 0.0356774081f), pack2(0.7978845608f, 0.7978845608f)));
  float t0, t1, h0, h1;
  unpack2(t, t0, t1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(h0) : "f"(t0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(h1) : "f"(t1));
  const F2 hx = mul2(x, pack2(0.5f, 0.5f));
  return fma2(hx, pack2(h0, h1), hx);
}
#endif

// Epilogue activation.
// GELU, fp32 outputs: Abramowitz-Stegun 7.1.26 for erf (|err| < 1.5e-7, branch free).
// GELU, bf16 outputs: gelu_bf16 above.
template <bool OUT_BF16>
__device__ __forceinline__ float epi_act(float x, int act) {
  if (act == MSPI_ACT_RELU) return fmaxf(x, 0.f);
  if (act == MSPI_ACT_GELU) {
    if constexpr (OUT_BF16) {
      return gelu_bf16(x);
    } else {
      const float z = fabsf(x) * 0.70710678118654752440f;
      const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
      float poly = fmaf(t, 1.061405429f, -1.453152027f);
      poly = fmaf(poly, t, 1.421413741f);
      poly = fmaf(poly, t, -0.284496736f);
      poly = fmaf(poly, t, 0.254829592f);
      const float e = fmaf(-poly * t, __expf(-z * z), 1.f);
      return 0.5f * x * (1.f + copysignf(e, x));
    }
  }
  if (act == MSPI_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  return x;
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
// row_bytes = 128 / 64 / 32 -> SWIZZLE_128B / 64B / 32B (layout codes 2 / 4 / 6), 8-row groups 8*row_bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                  // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>((8 * row_bytes) >> 4) << 32;  // stride byte offset: 8 rows
  d |= static_cast<uint64_t>(1) << 46;                  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;
  return d;
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace tc
}  // namespace mspi
