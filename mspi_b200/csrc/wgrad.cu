// Weight gradient of the implicit-GEMM convolution (mspi_conv_gemm) for sm_100a.
//
//   dW[co][tap][ci] += sum over output positions p of  dY[p][co] * X[p + tap_off[tap]][ci]
//
// This is a GEMM whose reduction dimension is the (huge) position axis and whose two operands are both stored with
// the NON-reduction dimension contiguous (NDHWC: channels innermost).  tcgen05 takes such "MN-major" operands directly
// (instruction-descriptor bits 15/16), so nothing is transposed: a K step is a box of up to 128 positions, fetched
// by the same 5-D TMA boxes the forward kernel uses — dY at the box origin, X at the origin shifted by the tap (TMA
// zero-fill = the convolution's padding) — one box per 128-byte channel chunk, 128B-swizzled.  In shared memory a chunk
// is [positions][64 bf16 | 32 tf32], which is exactly the canonical MN-major SWIZZLE_128B layout (8-position groups
// 1024 B apart = SBO, channel chunks one chunk buffer apart = LBO).
//
// Work item = (position split, tap, Cin tile, Cout tile of 128); accumulators (128 x bn fp32) live in TMEM, double
// buffered; the epilogue adds them into the fp32 gradient tensor with red.global.add.f32 (split-K over positions).
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issue   warps 2..5: epilogue (one TMEM lane quarter each)
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mspi {
namespace {

using namespace tc;

constexpr int kThreads = 64 + 128;
constexpr int kMaxStages = 6;
constexpr int kBarrierBytes = 1024;
constexpr int kSmemBudget = 216 * 1024;

struct WgradParams {
  int box[4];
  int tiles_d[4];
  int pos_tiles, rows;      // number of position boxes, positions per box (K of one pipeline stage)
  int ntaps;
  int tap_off[MSPI_MAX_TAPS][4];
  int cout, cin, bn, ch;    // ch: channels per 128-byte chunk (64 bf16, 32 tf32)
  int m_tiles, n_tiles, splits, tiles_per_split;
  int a_slots, b_chunks;    // chunk buffers per stage: A (Cout side, <= 128/ch), B (Cin side, ceil(bn/ch))
  int chunk_bytes, stage_bytes, num_stages, tmem_cols, kmma;
  uint32_t idesc;
  long long s_co, s_ci, s_tap;
  float* dw;
};

// MN-major swizzled operand: start address, LBO = distance between 128-byte channel chunks, SBO = distance between
// swizzle atoms along the position (K) axis.
//   16-bit operands: SWIZZLE_128B (layout 2), atom = 8 positions x 128 B, SBO = 1024
//   tf32 operands  : the only MN-major layout tcgen05 accepts for 32-bit types is SWIZZLE_128B_BASE32B (layout 1): 32-byte
//                    chunks swizzled within 128 B, atom = 4 positions x 128 B, SBO = 512 (TMA: SWIZZLE_128B_ATOM_32B)
template <int KIND>
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes & 0x3FFFFu) >> 4) << 16;
  d |= static_cast<uint64_t>((KIND == MSPI_BF16 ? 1024 : 512) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(KIND == MSPI_BF16 ? 2 : 1) << 61;
  return d;
}

template <int KIND>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_x,
             const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_base;
  const uint32_t bar_empty = smem_base + 8 * kMaxStages;
  const uint32_t bar_tfull = smem_base + 16 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t tiles_base = smem_base + kBarrierBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_x) : "memory");
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int items_per_split = p.ntaps * p.n_tiles * p.m_tiles;
  const int total_items = p.splits * items_per_split;

  // item -> (split, tap, nt, mt): CTAs running side by side share a position range (dY / X boxes hit in L2)
#define MSPI_WG_DECODE(item)                                                    \
  int q_ = (item);                                                               \
  const int mt = q_ % p.m_tiles; q_ /= p.m_tiles;                                \
  const int nt = q_ % p.n_tiles; q_ /= p.n_tiles;                                \
  const int tap = q_ % p.ntaps;  q_ /= p.ntaps;                                  \
  const int pt0 = q_ * p.tiles_per_split;                                        \
  const int pt1 = min(pt0 + p.tiles_per_split, p.pos_tiles);                     \
  (void)mt; (void)nt; (void)tap;

  if (warp == 0) {
    const bool issuer = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      MSPI_WG_DECODE(item)
      const int a_chunks = min(p.a_slots, (p.cout - mt * 128 + p.ch - 1) / p.ch);
      const uint32_t tx = static_cast<uint32_t>((a_chunks + p.b_chunks) * p.chunk_bytes);
      for (int pt = pt0; pt < pt1; ++pt) {
        int t = pt, org[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          org[j] = (t % p.tiles_d[j]) * p.box[j];
          t /= p.tiles_d[j];
        }
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if (issuer) {
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, tx);
          const uint32_t sa = tiles_base + stage * p.stage_bytes;
          for (int c = 0; c < a_chunks; ++c)
            tma_load_5d(sa + c * p.chunk_bytes, &tma_dy, full, mt * 128 + c * p.ch, org[0], org[1], org[2], org[3]);
          const uint32_t sb = sa + p.a_slots * p.chunk_bytes;
          for (int c = 0; c < p.b_chunks; ++c)
            tma_load_5d(sb + c * p.chunk_bytes, &tma_x, full, nt * p.bn + c * p.ch, org[0] + p.tap_off[tap][0],
                        org[1] + p.tap_off[tap][1], org[2] + p.tap_off[tap][2], org[3] + p.tap_off[tap][3]);
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const bool issuer = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    const int mmas = p.rows / p.kmma;
    const uint32_t kstep = static_cast<uint32_t>(p.kmma * 128) >> 4;  // descriptor units (16 B) per MMA along positions
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      MSPI_WG_DECODE(item)
      if (pt1 <= pt0) continue;  // empty split tail: nothing to accumulate, nothing to store
      mbar_wait(bar_tempty + 8 * as, aphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * p.bn);
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (issuer) {
          const uint32_t sa = tiles_base + stage * p.stage_bytes;
          const uint64_t adesc = make_mn_desc<KIND>(sa, p.chunk_bytes);
          const uint64_t bdesc = make_mn_desc<KIND>(sa + p.a_slots * p.chunk_bytes, p.chunk_bytes);
          for (int k = 0; k < mmas; ++k)
            tc_mma<KIND>(tmem_d, adesc + kstep * k, bdesc + kstep * k, p.idesc, (pt > pt0 || k > 0) ? 1u : 0u);
          tc_commit(bar_empty + 8 * stage);
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
      if (issuer) tc_commit(bar_tfull + 8 * as);
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else {
    const int quarter = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      MSPI_WG_DECODE(item)
      if (pt1 <= pt0) continue;
      mbar_wait(bar_tfull + 8 * as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * p.bn);
      const int co = mt * 128 + quarter * 32 + lane;
      float* row = p.dw + static_cast<long long>(co) * p.s_co + static_cast<long long>(tap) * p.s_tap;
      const int ci0 = nt * p.bn;
      for (int c = 0; c < p.bn; c += 16) {
        uint32_t acc[16];
        __syncwarp();
        tmem_ld16(taddr + c, acc);
        tmem_ld_wait();
        if (co < p.cout) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int ci = ci0 + c + j;
            if (ci < p.cin) atomicAdd(row + static_cast<long long>(ci) * p.s_ci, __uint_as_float(acc[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }
#undef MSPI_WG_DECODE

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

int encode_act_map(CUtensorMap* map, EncodeTiledFn encode, int dtype, const void* base, const int32_t dims[5],
                   const int64_t strides[5], const int32_t box[5], int ch, const char* what) {
  const int es = dtype == MSPI_BF16 ? 2 : 4;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t span = static_cast<cuuint64_t>((dims[0] * es + 15) / 16 * 16);
  gdim[0] = static_cast<cuuint64_t>(dims[0]);
  bdim[0] = static_cast<cuuint32_t>(ch);
  for (int j = 1; j < 5; ++j) {
    gdim[j] = static_cast<cuuint64_t>(dims[j]);
    bdim[j] = static_cast<cuuint32_t>(box[j]);
    cuuint64_t st = static_cast<cuuint64_t>(strides[j]) * es;
    if (dims[j] == 1 && (st == 0 || st % 16 != 0)) st = span;  // never dereferenced
    gstr[j - 1] = st;
    if (st * gdim[j] > span) span = st * gdim[j];
  }
  CUresult r = encode(map, dtype == MSPI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5,
                      const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      dtype == MSPI_BF16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(MSPI_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed: %d dims=[%d,%d,%d,%d,%d] box=[%d,%d,%d,%d,%d]", what,
                     (int)r, dims[0], dims[1], dims[2], dims[3], dims[4], ch, box[1], box[2], box[3], box[4]);
  return MSPI_OK;
}

}  // namespace
}  // namespace mspi

using namespace mspi;

extern "C" int mspi_conv_wgrad(const MspiConvDesc* d, const void* x, const void* dy, float* dw, int64_t s_co, int64_t s_ci,
                               int64_t s_tap, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && dy && dw, "mspi_conv_wgrad: null argument");
  MSPI_CHECK_ARG(d->a_dtype == MSPI_BF16 || d->a_dtype == MSPI_F32, "a_dtype %d", d->a_dtype);
  MSPI_CHECK_ARG(d->o_dtype == d->a_dtype, "dY must have the dtype of X (cast it first)");
  MSPI_CHECK_ARG(d->k_row_bytes == 0 || d->k_row_bytes == 128, "stem descriptors (k_row_bytes != 128) are not supported");
  MSPI_CHECK_ARG(d->w_batch_dims[0] == 0, "batched weights are not supported");
  MSPI_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= MSPI_MAX_TAPS, "ntaps %d", d->ntaps);
  const bool bf16 = d->a_dtype == MSPI_BF16;
  const int es = bf16 ? 2 : 4;
  const int ch = 128 / es;
  const int kmma = bf16 ? 16 : 8;
  long long rows = 1;
  for (int j = 1; j < 5; ++j) {
    MSPI_CHECK_ARG(d->box[j] >= 1 && d->box[j] <= 256, "box[%d]=%d", j, d->box[j]);
    MSPI_CHECK_ARG((d->a_strides[j] * es) % 16 == 0 || d->a_dims[j] == 1, "a_strides[%d] not 16B aligned", j);
    MSPI_CHECK_ARG((d->o_strides[j - 1] * es) % 16 == 0 || d->o_dims[j - 1] == 1, "o_strides[%d] not 16B aligned", j - 1);
    rows *= d->box[j];
  }
  MSPI_CHECK_ARG(rows <= 128 && rows % kmma == 0, "box has %lld positions: need a multiple of %d, <= 128", rows, kmma);
  MSPI_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0,
                 "operands must be 16-byte aligned");
  const int cin = d->a_dims[0], cout = d->cout;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return set_error(MSPI_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");

  CUtensorMap map_x, map_dy;
  int rc = encode_act_map(&map_x, encode, d->a_dtype, x, d->a_dims, d->a_strides, d->box, ch, "X");
  if (rc != MSPI_OK) return rc;
  {
    int32_t dims[5] = {cout, d->o_dims[0], d->o_dims[1], d->o_dims[2], d->o_dims[3]};
    int64_t strides[5] = {1, d->o_strides[0], d->o_strides[1], d->o_strides[2], d->o_strides[3]};
    rc = encode_act_map(&map_dy, encode, d->a_dtype, dy, dims, strides, d->box, ch, "dY");
    if (rc != MSPI_OK) return rc;
  }

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.pos_tiles = 1;
  for (int j = 0; j < 4; ++j) {
    p.box[j] = d->box[j + 1];
    p.tiles_d[j] = (d->o_dims[j] + p.box[j] - 1) / p.box[j];
    p.pos_tiles *= p.tiles_d[j];
  }
  p.rows = static_cast<int>(rows);
  p.ntaps = d->ntaps;
  memcpy(p.tap_off, d->tap_off, sizeof(p.tap_off));
  p.cout = cout;
  p.cin = cin;
  p.ch = ch;
  p.kmma = kmma;
  p.m_tiles = (cout + 127) / 128;
  // Cin tile: as wide as TMEM (2 x bn columns) and the smem ring (>= 2 stages) allow
  int bn = (cin + 15) / 16 * 16;
  const int bn_max = bf16 ? 256 : 128;
  if (bn > bn_max) {
    const int nt = (cin + bn_max - 1) / bn_max;
    bn = ((cin + nt - 1) / nt + 15) / 16 * 16;
  }
  p.bn = bn;
  p.n_tiles = (cin + bn - 1) / bn;
  p.a_slots = (cout >= 128 ? 128 : (cout + ch - 1) / ch * ch) / ch;
  if (p.a_slots > 128 / ch) p.a_slots = 128 / ch;
  p.b_chunks = (bn + ch - 1) / ch;
  p.chunk_bytes = p.rows * 128;
  p.stage_bytes = (p.a_slots + p.b_chunks) * p.chunk_bytes;
  // The MMA always reads M = 128 rows of A, i.e. 128/ch chunk slots LBO apart, whatever a_slots is: with Cout < 128 the
  // extra slots alias the B chunks / the next stage (their accumulator rows are never stored), and behind the LAST stage
  // they must still be inside the allocation.
  const int slack = (128 / ch - p.a_slots) * p.chunk_bytes;
  p.num_stages = (kSmemBudget - kBarrierBytes - 1024 - slack) / p.stage_bytes;
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  MSPI_CHECK_ARG(p.num_stages >= 2, "shared memory budget leaves %d pipeline stages", p.num_stages);
  int cols = 32;
  while (cols < 2 * bn) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t fmt = bf16 ? 1u : 2u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(bn >> 3) << 17) |
            (static_cast<uint32_t>(128 >> 4) << 24);
  p.s_co = s_co;
  p.s_ci = s_ci;
  p.s_tap = s_tap;
  p.dw = dw;
  // split the position axis so that every SM has work, but keep >= 4 boxes per item where there are enough
  const int sms = num_sms();
  if (sms <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const long long base_items = static_cast<long long>(p.ntaps) * p.n_tiles * p.m_tiles;
  long long splits = (2ll * sms + base_items - 1) / base_items;
  const long long max_splits = (p.pos_tiles + 3) / 4;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.tiles_per_split = static_cast<int>((p.pos_tiles + splits - 1) / splits);
  p.splits = (p.pos_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  const long long total = base_items * p.splits;
  MSPI_CHECK_ARG(total < (1ll << 31), "too many work items");
  const size_t smem = 1024 + kBarrierBytes + static_cast<size_t>(p.num_stages) * p.stage_bytes + slack;
  auto kern = bf16 ? wgrad_kernel<MSPI_BF16> : wgrad_kernel<MSPI_F32>;
  MSPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int grid = static_cast<int>(total < sms ? total : sms);
  kern<<<grid, kThreads, smem, stream>>>(map_dy, map_x, p);
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}
