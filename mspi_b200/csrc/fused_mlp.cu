// Fused ConvNeXt MLP for sm_100a:   y = r + gamma * ( W2 . gelu(W1 . x + b1) + b2 )
//
// Replaces timm ConvNeXt Mlp (fc1 -> GELU -> fc2) + layer scale + residual (model_utils.py:361; stages 0 and 1,
// C = 96 / 192) with ONE kernel.  Unfused, the 4C-wide hidden activation is written to and re-read from HBM (2.1 GB
// per layer at stage 0, four times the layer's input + output), which makes those layers HBM/epilogue bound; here the
// hidden tile never leaves the SM:
//
//   TMA:   x tile [128 x C] -> smem (double buffered across tiles); W1 / W2 pieces stream through a smem ring
//   MMA1:  H_j [128 x 128] = x . W1_j^T          tcgen05.mma kind::f16, fp32 accumulators in TMEM (2 buffers)
//   epi:   H_j -> +b1 -> GELU -> bf16 -> smem, written in the 128B-swizzled K-major layout MMA2 reads as its A operand
//   MMA2:  Y [128 x C] += gelu(H_j) . W2_j^T      accumulators in TMEM columns 256..
//   epi:   Y -> gamma * (. + b2) + residual -> bf16 -> global
//
// One persistent CTA per SM, warp specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..17 = epilogue
// (the kernel is bound by the GELU epilogue: 16 warps give the schedulers four independent instruction streams each).
// Weight pieces are identical for every M tile, so the CTAs of a thread-block cluster (4 by default) stream them
// TOGETHER: each CTA fetches 1/CL of every piece and TMA-multicasts it into the ring of all CL CTAs, and a ring slot
// is released by a multicast tcgen05.commit from every consumer.  Without this the kernel is bound by re-reading the
// 147 KB (C=96) / 590 KB (C=192) of weights per tile from L2 (measured 3.9 TB/s L2->SM, profiles/r01_fused_mlp*).
// The MMA warp runs one software-pipelined stream over all (tile, hidden chunk) pairs: MMA1(g) is issued before
// MMA2(g-1), so the tensor pipe always has a GEMM to run while the epilogue warps apply GELU to the previous chunk.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mspi {
namespace {

using namespace tc;

constexpr int kHC = 128;                       // hidden columns per chunk
constexpr int kEpiWarps = 16;                  // four per TMEM lane quarter: each takes 32 of a chunk's 128 columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kChunkBytes = 128 * 128;         // one K chunk of a 128-row operand tile: 16 KB
constexpr int kMaxRing = 8;
constexpr int kMaxHBuf = 3;
constexpr int kVecBytes = 12288;               // staged per-channel vectors (8 KB) + LayerNorm partial sums (4 KB)
// MSPI_MLP_DEBUG (study bits: skip GELU / TMEM loads / A2 stores / residual I/O) only exists in builds made with
// -DMSPI_MLP_STUDY: as a run-time flag its dead branches still cost the epilogue 32 integer instructions per hidden chunk
// (ptxas hoists the fake accumulator values above the branch; ncu source page, profiles/r02_fused_mlp.md).
#ifdef MSPI_MLP_STUDY
#define MLP_DBG(p, bit) ((p).debug & (bit))
#else
#define MLP_DBG(p, bit) 0
#endif

struct MlpParams {
  int m_rows, c, kc1, nh, m_tiles, x_bufs;
  int debug;            // tuning aid (MSPI_MLP_DEBUG bit mask): 1 skip GELU math, 2 skip TMEM loads, 4 skip A2 stores
  int hbufs;            // TMEM buffers for the hidden chunk accumulators (3 when 3*128 + C <= 512, else 2)
  int cl;               // cluster size (CTAs sharing the weight stream); cluster_tiles = ceil(m_tiles / cl)
  int ring_stages, ring_stage_bytes;
  const float* b1;      // [4C]
  const float* scale;   // [C]  layer scale gamma
  const float* shift;   // [C]  gamma * b2
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  long long res_stride, y_stride;
  uint32_t idesc1, idesc2;
  int wide_io;          // residual / y rows are 32-byte aligned: 256-bit global loads and stores in the Y epilogue
  int y_staged;         // Y epilogue through shared memory: the residual tile arrives by TMA, is updated in place and leaves by
                        // TMA (a warp's direct row accesses cost 32 L1 tag cycles per instruction: 6144 per 128 x 96 tile)
  int ystage_bytes;
  int packed_gelu;      // GELU on packed fp32 pairs (fma.rn.f32x2 / mul / add) with one fp32 tanh per element
  // LayerNorm over the C channels of every output row, applied to the bf16-rounded block output before it is stored (the
  // next stage's downsample.0 LayerNorm2d, model_utils.py:361 / timm ConvNeXtStage): y = LN(r + gamma * mlp(x)) * w + b.
  // The block output itself is then never written: the downsample conv is its only reader.
  const float* ln_w;    // [C] or null
  const float* ln_b;    // [C]
  float ln_eps;
};

__device__ __forceinline__ void ldg256(const __nv_bfloat16* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(__nv_bfloat16* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// ---- CTA pair (tcgen05 cta_group::2), see conv_gemm.cu: both CTAs load their own x tile and HALF of every weight piece, all
// loads complete on the leader's barriers, the leader issues the M = 256 MMAs, commits are multicast to both CTAs and the
// peer's epilogue warps arrive on the leader's barriers.
__device__ __forceinline__ uint32_t leader_smem(uint32_t local) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local));
  return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit2_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// PAIR: CTA-pair instance (cluster of 2 required); a kernel that contains cta_group::2 instructions cannot be launched plainly.
template <bool PAIR, bool LN>
__global__ void __launch_bounds__(kThreads, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ CUtensorMap map_y, const __grid_constant__ MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  // barrier block
  const uint32_t ring_full = base, ring_empty = base + 8 * kMaxRing;
  const uint32_t x_full = base + 16 * kMaxRing, x_empty = x_full + 16;
  const uint32_t h_full = x_empty + 16, h_empty = h_full + 8 * kMaxHBuf;
  const uint32_t a2_full = h_empty + 8 * kMaxHBuf, a2_empty = a2_full + 16;
  const uint32_t y_full = a2_empty + 16, y_empty = y_full + 8;
  const uint32_t tmem_slot = y_empty + 8;
  const uint32_t res_full = y_empty + 16;
  float* s_b1 = reinterpret_cast<float*>(base_ptr + 1024);       // [4C]
  float* s_scale = s_b1 + 4 * p.c;                               // [C]
  float* s_shift = s_scale + p.c;                                // [C]
  float* s_lnw = s_shift + p.c;                                  // [C]   (8 C floats <= 6 KB so far)
  float* s_lnb = s_lnw + p.c;                                    // [C]
  float2* s_part = reinterpret_cast<float2*>(base_ptr + 1024 + 8192);   // [4 column groups][128 rows] (sum, sum of squares)
  const uint32_t x_base = base + 1024 + kVecBytes;               // 2 x kc1 chunks
  const uint32_t a2_base = x_base + static_cast<uint32_t>(p.x_bufs) * p.kc1 * kChunkBytes;  // 2 buffers x 2 chunks
  const uint32_t ystage_base = a2_base + 4u * kChunkBytes;       // [c/32 boxes][128 rows][64 B], 64B-swizzled (y_staged)
  const uint32_t ring_base = ystage_base + static_cast<uint32_t>(p.ystage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    for (int s = 0; s < p.ring_stages; ++s) {
      mbar_init(ring_full + 8 * s, 1);
      mbar_init(ring_empty + 8 * s, PAIR ? 1u : static_cast<uint32_t>(p.cl));  // one multicast commit per issuing CTA
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(x_full + 8 * s, 1);
      mbar_init(x_empty + 8 * s, 1);
      mbar_init(a2_full + 8 * s, PAIR ? 2 * kEpiWarps : kEpiWarps);  // pair: the leader waits for both CTAs' epilogues
      mbar_init(a2_empty + 8 * s, 1);
    }
    for (int s = 0; s < kMaxHBuf; ++s) {
      mbar_init(h_full + 8 * s, 1);
      mbar_init(h_empty + 8 * s, PAIR ? 2 * kEpiWarps : kEpiWarps);
    }
    mbar_init(y_full, 1);
    mbar_init(y_empty, PAIR ? 2 * kEpiWarps : kEpiWarps);
    mbar_init(res_full, 1);
    if (p.y_staged) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  pdl_wait();   // barrier init / TMEM allocation above overlap the previous kernel's drain (common.cuh)
  for (int i = threadIdx.x; i < 4 * p.c; i += kThreads) s_b1[i] = __ldg(p.b1 + i);
  for (int i = threadIdx.x; i < p.c; i += kThreads) {
    s_scale[i] = __ldg(p.scale + i);
    s_shift[i] = __ldg(p.shift + i);
    if (LN) {
      s_lnw[i] = __ldg(p.ln_w + i);
      s_lnb[i] = __ldg(p.ln_b + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const uint32_t tmem_y = tmem_base + static_cast<uint32_t>(p.hbufs * kHC);  // Y accumulators follow the H buffers
  const int x_buf_bytes = p.kc1 * kChunkBytes;
  const int w2_bytes = p.c * 128;
  const uint32_t rank = p.cl > 1 ? cluster_ctarank() : 0u;
  const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cl) - 1u);
  // tile schedule: cluster `cid` takes tile groups cid, cid + ncl, ...; CTA `rank` owns tile group*cl + rank.  Every CTA of
  // a cluster runs the same number of iterations (a tile index past the end is a dummy: loads zero-fill, stores are
  // masked), because all of them must consume every multicast weight piece.
  const int cid = p.cl > 1 ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int ncl = p.cl > 1 ? static_cast<int>(num_clusters_x()) : static_cast<int>(gridDim.x);
  const int groups = (p.m_tiles + p.cl - 1) / p.cl;
  if (p.cl > 1) cluster_sync_all();  // barrier inits of every CTA are visible before anyone multicasts

  if (warp == 0) {
    // ================================================================== TMA producer (warp-uniform, one elected issuer)
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int xs = 0;
      uint32_t xph = 0;
      auto load_x = [&](int tile) {
        mbar_wait(x_empty + 8 * xs, xph ^ 1u);
        if constexpr (PAIR) {
          if (issuer) {
            if (rank == 0) mbar_expect_tx(x_full + 8 * xs, 2u * static_cast<uint32_t>(x_buf_bytes));
            const uint32_t full_l = leader_smem(x_full + 8 * xs);
            for (int kc = 0; kc < p.kc1; ++kc)
              tma_load_2d_2sm(x_base + xs * x_buf_bytes + kc * kChunkBytes, &map_x, full_l, kc * 64, tile * 128);
          }
        } else if (issuer) {
          mbar_expect_tx(x_full + 8 * xs, static_cast<uint32_t>(x_buf_bytes));
          for (int kc = 0; kc < p.kc1; ++kc)
            tma_load_2d(x_base + xs * x_buf_bytes + kc * kChunkBytes, &map_x, x_full + 8 * xs, kc * 64, tile * 128);
        }
        __syncwarp();
        if (++xs == p.x_bufs) { xs = 0; xph ^= 1u; }
      };
      auto load_w1 = [&](int j) {
        for (int kc = 0; kc < p.kc1; ++kc) {
          mbar_wait(ring_empty + 8 * stage, phase ^ 1u);
          if constexpr (PAIR) {
            if (issuer) {   // this CTA's 64 of the piece's 128 hidden rows, into its own ring slot
              if (rank == 0) mbar_expect_tx(ring_full + 8 * stage, kChunkBytes);
              tma_load_2d_2sm(ring_base + stage * p.ring_stage_bytes, &map_w1, leader_smem(ring_full + 8 * stage), kc * 64,
                              j * kHC + static_cast<int>(rank) * (kHC / 2));
            }
          } else if (issuer) {
            mbar_expect_tx(ring_full + 8 * stage, kChunkBytes);
            if (p.cl > 1) {
              const int share = kHC / p.cl;  // rows of the piece this CTA fetches for the whole cluster
              tma_load_2d_mc(ring_base + stage * p.ring_stage_bytes + rank * share * 128, &map_w1, ring_full + 8 * stage,
                             kc * 64, j * kHC + rank * share, mc_mask);
            } else {
              tma_load_2d(ring_base + stage * p.ring_stage_bytes, &map_w1, ring_full + 8 * stage, kc * 64, j * kHC);
            }
          }
          __syncwarp();
          if (++stage == p.ring_stages) { stage = 0; phase ^= 1u; }
        }
      };
      auto load_w2 = [&](int j) {
        for (int kc2 = 0; kc2 < 2; ++kc2) {
          mbar_wait(ring_empty + 8 * stage, phase ^ 1u);
          if constexpr (PAIR) {
            if (issuer) {   // this CTA's C/2 of the piece's C output rows
              if (rank == 0) mbar_expect_tx(ring_full + 8 * stage, static_cast<uint32_t>(w2_bytes));
              tma_load_2d_2sm(ring_base + stage * p.ring_stage_bytes, &map_w2, leader_smem(ring_full + 8 * stage),
                              j * kHC + kc2 * 64, static_cast<int>(rank) * (p.c / 2));
            }
          } else if (issuer) {
            mbar_expect_tx(ring_full + 8 * stage, static_cast<uint32_t>(w2_bytes));
            if (p.cl > 1) {
              const int share = p.c / p.cl;
              tma_load_2d_mc(ring_base + stage * p.ring_stage_bytes + rank * share * 128, &map_w2, ring_full + 8 * stage,
                             j * kHC + kc2 * 64, rank * share, mc_mask);
            } else {
              tma_load_2d(ring_base + stage * p.ring_stage_bytes, &map_w2, ring_full + 8 * stage, j * kHC + kc2 * 64, 0);
            }
          }
          __syncwarp();
          if (++stage == p.ring_stages) { stage = 0; phase ^= 1u; }
        }
      };
      // with two x buffers the tile after the current one is requested before the current tile's weights
      if (p.x_bufs == 2 && cid < groups) load_x(cid * p.cl + rank);
      int prev_j = -1;
      for (int grp = cid; grp < groups; grp += ncl) {
        if (p.x_bufs == 1) load_x(grp * p.cl + rank);
        else if (grp + ncl < groups) load_x((grp + ncl) * p.cl + rank);
        for (int j = 0; j < p.nh; ++j) {
          load_w1(j);
          if (prev_j >= 0) load_w2(prev_j);
          prev_j = j;
        }
      }
      if (prev_j >= 0) load_w2(prev_j);
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    // The whole warp runs the control flow (barrier waits are warp-uniform); one elected lane issues tcgen05 ops.
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int xs = 0;
      uint32_t xph = 0;
      uint32_t g = 0, ytile = 0;
      uint32_t mb = 0, mph = 0;   // hidden buffer of chunk g and its phase
      int prev_j = -1;
      uint32_t prev_g = 0;
      // pair mode: the leader issues for both CTAs and every commit is multicast to both; the peer's MMA warp idles
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) tc_mma2_bf16(d, ad, bd, idesc, acc); else tc_mma<MSPI_BF16>(d, ad, bd, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {
        if constexpr (PAIR) tc_commit2_mc(bar); else tc_commit(bar);
      };
      auto commit_ring = [&](uint32_t bar) {
        if constexpr (PAIR) tc_commit2_mc(bar);
        else if (p.cl > 1) tc_commit_mc(bar, mc_mask);
        else tc_commit(bar);
      };
      const bool mma_active = !PAIR || rank == 0;
      auto mma2 = [&](int pj, uint32_t pg) {
        const uint32_t pb = pg & 1u;
        mbar_wait(a2_full + 8 * pb, (pg >> 1) & 1u);
        if (pj == 0) mbar_wait(y_empty, (ytile & 1u) ^ 1u);  // the previous tile's Y has been read out
        tc_fence_after();
        for (int kc2 = 0; kc2 < 2; ++kc2) {
          mbar_wait(ring_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(a2_base + (pb * 2 + kc2) * kChunkBytes, 128);
          const uint64_t bdesc = make_smem_desc(ring_base + stage * p.ring_stage_bytes, 128);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma(tmem_y, adesc + 2u * k, bdesc + 2u * k, p.idesc2, (pj | kc2 | k) != 0 ? 1u : 0u);
            commit_ring(ring_empty + 8 * stage);
          }
          __syncwarp();
          if (++stage == p.ring_stages) { stage = 0; phase ^= 1u; }
        }
        if (issuer) {
          commit(a2_empty + 8 * pb);
          if (pj == p.nh - 1) commit(y_full);
        }
        __syncwarp();
        if (pj == p.nh - 1) ++ytile;
      };
      for (int grp = cid; mma_active && grp < groups; grp += ncl) {
        mbar_wait(x_full + 8 * xs, xph);
        tc_fence_after();
        for (int j = 0; j < p.nh; ++j) {
          const uint32_t b = mb;
          mbar_wait(h_empty + 8 * b, mph ^ 1u);  // the epilogue has read H[b] of chunk g - hbufs
          tc_fence_after();
          for (int kc = 0; kc < p.kc1; ++kc) {
            mbar_wait(ring_full + 8 * stage, phase);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(x_base + xs * x_buf_bytes + kc * kChunkBytes, 128);
            const uint64_t bdesc = make_smem_desc(ring_base + stage * p.ring_stage_bytes, 128);
            const int ksteps = min(4, (p.c - kc * 64 + 15) >> 4);  // skip the zero-padded K tail
            if (issuer) {
              for (int k = 0; k < ksteps; ++k)
                mma(tmem_base + b * kHC, adesc + 2u * k, bdesc + 2u * k, p.idesc1, (kc | k) != 0 ? 1u : 0u);
              commit_ring(ring_empty + 8 * stage);
            }
            __syncwarp();
            if (++stage == p.ring_stages) { stage = 0; phase ^= 1u; }
          }
          if (issuer) {
            commit(h_full + 8 * b);
            if (j == p.nh - 1) commit(x_empty + 8 * xs);  // x tile free once these MMAs have read it
          }
          __syncwarp();
          if (prev_j >= 0) mma2(prev_j, prev_g);
          prev_j = j;
          prev_g = g;
          ++g;
          if (++mb == static_cast<uint32_t>(p.hbufs)) { mb = 0; mph ^= 1u; }
        }
        if (++xs == p.x_bufs) { xs = 0; xph ^= 1u; }
      }
      if (mma_active && prev_j >= 0) mma2(prev_j, prev_g);
    }
  } else {
    // ================================================================== epilogue warps
    // pair mode: the barriers the MMA warp waits on live in the leader CTA
    auto arrive = [&](uint32_t bar) {
      if constexpr (PAIR) mbar_arrive_cluster(leader_smem(bar)); else mbar_arrive(bar);
    };
    const int quarter = warp & 3;            // TMEM lanes 32*quarter .. +31 = this warp's rows
    const int colgrp = (warp - 2) >> 2;      // which 32 of a chunk's 128 hidden columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t lane_y = tmem_y + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t g = 0, ytile = 0;
    uint32_t hb = 0, hph = 0;   // TMEM hidden buffer of chunk g and its barrier phase (counters: no division in the loop)
    // The Y tile of a finished M tile is written one hidden chunk LATE (after the next tile's first GELU chunk): MMA2 of the
    // last chunk then has time to finish, and the residual rows requested right after the last GELU chunk have arrived.
    bool pend = false;
    long long pgrow = 0;
    bool pvalid = false;
    uint4 resq[3][2];
    int prow0 = 0;                       // first row of the pending tile
    const bool io_thread = warp == 2 && lane == 0;
    // ---- fused LayerNorm of the output rows (p.ln_w): a row's C channels are spread over the four column groups (same lane
    // of four warps), so each thread sums its bf16-rounded outputs, the partial sums meet in shared memory, and every thread
    // normalises its own columns.  s_part is safe to re-use from tile to tile: a thread reaches the next tile's Y epilogue
    // nh >= 3 hidden chunks later, and the A2 staging ring (two chunks deep, every epilogue warp arrives per chunk) keeps the
    // epilogue warps within two chunks of each other.
    constexpr bool ln = LN;   // compile time: the plain instance keeps its register allocation
    const float ln_inv_c = 1.f / static_cast<float>(p.c);
    auto ln_add = [&](uint32_t pk, float& s1, float& s2) {
      float a, b;
      unpack_bf16x2(pk, a, b);
      s1 += a + b;
      s2 = fmaf(a, a, fmaf(b, b, s2));
    };
    auto ln_exchange = [&](float s1, float s2, float& mean, float& rstd) {
      s_part[colgrp * 128 + row] = make_float2(s1, s2);
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 v = s_part[q * 128 + row];
        t1 += v.x;
        t2 += v.y;
      }
      mean = t1 * ln_inv_c;
      rstd = rsqrtf(fmaxf(fmaf(-mean, mean, t2 * ln_inv_c), 0.f) + p.ln_eps);
    };
    auto ln_apply = [&](uint32_t pk, int col, float mean, float rstd) {
      float a, b;
      unpack_bf16x2(pk, a, b);
      a = fmaf((a - mean) * rstd, s_lnw[col], s_lnb[col]);
      b = fmaf((b - mean) * rstd, s_lnw[col + 1], s_lnb[col + 1]);
      return pack_bf16x2(a, b);
    };
    auto y_epilogue = [&]() {
      mbar_wait(y_full, ytile & 1u);
      tc_fence_after();
      if (LN && p.y_staged) {
        // staged path (C <= 128: at most two 16-column pieces per thread).  Two passes over the accumulator instead of a
        // register stash (a stash of the packed outputs spilled 264 bytes per thread through an L1 that the 227 KB of
        // shared memory leave no room in: +0.32 ms per launch): pass 1 sums the bf16-rounded outputs, pass 2 recomputes
        // them — same arithmetic, same rounding — normalises and overwrites the residual in the staging tile.
        mbar_wait(res_full, ytile & 1u);
        float s1 = 0.f, s2 = 0.f, mean = 0.f, rstd = 0.f;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c0 = 16 * colgrp + 64 * u;
            if (c0 >= p.c) break;
            uint32_t acc[16];
            __syncwarp();
            tmem_ld16(lane_y + c0, acc);
            tmem_ld_wait();
            const uint32_t rowb = ystage_base + static_cast<uint32_t>(c0 >> 5) * (128u * 64u) + static_cast<uint32_t>(row) * 64u;
            const uint32_t k0 = static_cast<uint32_t>(c0 & 31) >> 3, swz = (static_cast<uint32_t>(row) >> 1) & 3u;
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int col = c0 + 8 * h8;
              const uint32_t addr = rowb + (((k0 + h8) ^ swz) << 4);
              uint4 rr;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w) : "r"(addr) : "memory");
              float res[8];
              unpack_bf16x2(rr.x, res[0], res[1]); unpack_bf16x2(rr.y, res[2], res[3]);
              unpack_bf16x2(rr.z, res[4], res[5]); unpack_bf16x2(rr.w, res[6], res[7]);
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float v0 = fmaf(__uint_as_float(acc[8 * h8 + 2 * e]), s_scale[col + 2 * e], s_shift[col + 2 * e]) + res[2 * e];
                const float v1 = fmaf(__uint_as_float(acc[8 * h8 + 2 * e + 1]), s_scale[col + 2 * e + 1], s_shift[col + 2 * e + 1]) + res[2 * e + 1];
                o[e] = pack_bf16x2(v0, v1);
                if (pass == 0) ln_add(o[e], s1, s2); else o[e] = ln_apply(o[e], col + 2 * e, mean, rstd);
              }
              if (pass == 1)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
          }
          if (pass == 0) ln_exchange(s1, s2, mean, rstd);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(y_empty);     // TMEM has been read twice: MMA2 of the next tile may overwrite Y
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // every epilogue thread has written its pieces
        if (io_thread) {
          for (int b = 0; b < p.c / 32; ++b)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_y),
                         "r"(ystage_base + static_cast<uint32_t>(b) * (128u * 64u)), "r"(b * 32), "r"(prow0)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++ytile;
        return;
      }
      if constexpr (LN) {
        // direct path (C = 192: three 16-column pieces per thread, residual rows prefetched into registers): same two passes
        float s1 = 0.f, s2 = 0.f, mean = 0.f, rstd = 0.f;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const int c0 = 16 * colgrp + 64 * u;
            if (c0 >= p.c) break;
            uint32_t acc[16];
            __syncwarp();
            tmem_ld16(lane_y + c0, acc);
            tmem_ld_wait();
            uint4 o2[2];
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int col = c0 + 8 * h8;
              const uint4 rr = resq[u][h8];
              float res[8];
              unpack_bf16x2(rr.x, res[0], res[1]); unpack_bf16x2(rr.y, res[2], res[3]);
              unpack_bf16x2(rr.z, res[4], res[5]); unpack_bf16x2(rr.w, res[6], res[7]);
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float v0 = fmaf(__uint_as_float(acc[8 * h8 + 2 * e]), s_scale[col + 2 * e], s_shift[col + 2 * e]) + (pvalid ? res[2 * e] : 0.f);
                const float v1 = fmaf(__uint_as_float(acc[8 * h8 + 2 * e + 1]), s_scale[col + 2 * e + 1], s_shift[col + 2 * e + 1]) + (pvalid ? res[2 * e + 1] : 0.f);
                o[e] = pack_bf16x2(v0, v1);
                if (pass == 0) ln_add(o[e], s1, s2); else o[e] = ln_apply(o[e], col + 2 * e, mean, rstd);
              }
              o2[h8] = make_uint4(o[0], o[1], o[2], o[3]);
            }
            if (pass == 1 && pvalid) {
              __nv_bfloat16* yp = p.y + pgrow * p.y_stride + c0;
              if (p.wide_io) {
                stg256(yp, o2[0], o2[1]);
              } else {
                *reinterpret_cast<uint4*>(yp) = o2[0];
                *reinterpret_cast<uint4*>(yp + 8) = o2[1];
              }
            }
          }
          if (pass == 0) ln_exchange(s1, s2, mean, rstd);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(y_empty);
        ++ytile;
        return;
      }
      if (p.y_staged) {
        mbar_wait(res_full, ytile & 1u);   // the residual tile has landed in the staging buffer
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int c0 = 16 * colgrp + 64 * u;
          if (c0 >= p.c) break;
          uint32_t acc[16];
          __syncwarp();
          tmem_ld16(lane_y + c0, acc);
          tmem_ld_wait();
          const uint32_t rowb = ystage_base + static_cast<uint32_t>(c0 >> 5) * (128u * 64u) + static_cast<uint32_t>(row) * 64u;
          const uint32_t k0 = static_cast<uint32_t>(c0 & 31) >> 3, swz = (static_cast<uint32_t>(row) >> 1) & 3u;
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const int col = c0 + 8 * h8;
            const uint32_t addr = rowb + (((k0 + h8) ^ swz) << 4);
            uint4 rr;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w) : "r"(addr) : "memory");
            float res[8];
            unpack_bf16x2(rr.x, res[0], res[1]); unpack_bf16x2(rr.y, res[2], res[3]);
            unpack_bf16x2(rr.z, res[4], res[5]); unpack_bf16x2(rr.w, res[6], res[7]);
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[e] = fmaf(__uint_as_float(acc[8 * h8 + e]), s_scale[col + e], s_shift[col + e]) + res[e];
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[0], v[1])),
                         "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
                         : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(y_empty);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // every epilogue thread has written its pieces
        if (io_thread) {
          for (int b = 0; b < p.c / 32; ++b)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_y),
                         "r"(ystage_base + static_cast<uint32_t>(b) * (128u * 64u)), "r"(b * 32), "r"(prow0)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++ytile;
        return;
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int c0 = 16 * colgrp + 64 * u;
        if (c0 >= p.c) break;
        uint32_t acc[16];
        __syncwarp();
        tmem_ld16(lane_y + c0, acc);
        tmem_ld_wait();
        if (pvalid && !MLP_DBG(p, 8)) {
          uint4 o2[2];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const int col = c0 + 8 * h8;
            const uint4 rr = resq[u][h8];
            float res[8];
            unpack_bf16x2(rr.x, res[0], res[1]); unpack_bf16x2(rr.y, res[2], res[3]);
            unpack_bf16x2(rr.z, res[4], res[5]); unpack_bf16x2(rr.w, res[6], res[7]);
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[e] = fmaf(__uint_as_float(acc[8 * h8 + e]), s_scale[col + e], s_shift[col + e]) + res[e];
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            o2[h8] = o;
          }
          __nv_bfloat16* yp = p.y + pgrow * p.y_stride + c0;
          if (p.wide_io) {
            stg256(yp, o2[0], o2[1]);
          } else {
            *reinterpret_cast<uint4*>(yp) = o2[0];
            *reinterpret_cast<uint4*>(yp + 8) = o2[1];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(y_empty);
      ++ytile;
    };
    for (int grp = cid; grp < groups; grp += ncl) {
      const long long grow = (static_cast<long long>(grp) * p.cl + rank) * 128 + row;
      const bool valid = grow < p.m_rows;
      for (int j = 0; j < p.nh; ++j) {
        const uint32_t ab = g & 1u, aph = (g >> 1) & 1u;
        mbar_wait(h_full + 8 * hb, hph);
        mbar_wait(a2_empty + 8 * ab, aph ^ 1u);  // MMA2 of chunk g-2 has finished reading this staging buffer
        tc_fence_after();
        uint32_t packed[16];
        const float* b1 = s_b1 + j * kHC + colgrp * 32;
        {
          uint32_t acc[32];
          __syncwarp();
          if (!MLP_DBG(p, 2)) {
            tmem_ld32(lane_addr + hb * kHC + colgrp * 32, acc);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) acc[q] = 0x3f800000u + q + lane;
          }
          if (MLP_DBG(p, 1)) {
#pragma unroll
            for (int q = 0; q < 16; ++q) packed[q] = pack_bf16x2(__uint_as_float(acc[2 * q]), __uint_as_float(acc[2 * q + 1]));
          } else if (p.packed_gelu) {
            // packed fp32 pairs end to end: TMEM registers are consecutive, so (acc[2k], acc[2k+1]) already is a register pair
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = *reinterpret_cast<const float4*>(b1 + 4 * q);
              const F2 g01 = gelu_pair_f32(add2(pack2(__uint_as_float(acc[4 * q + 0]), __uint_as_float(acc[4 * q + 1])), pack2(bb.x, bb.y)));
              const F2 g23 = gelu_pair_f32(add2(pack2(__uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3])), pack2(bb.z, bb.w)));
              float v0, v1, v2, v3;
              unpack2(g01, v0, v1);
              unpack2(g23, v2, v3);
              packed[2 * q] = pack_bf16x2(v0, v1);
              packed[2 * q + 1] = pack_bf16x2(v2, v3);
            }
          } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = *reinterpret_cast<const float4*>(b1 + 4 * q);
            const float v0 = gelu_bf16(__uint_as_float(acc[4 * q + 0]) + bb.x);
            const float v1 = gelu_bf16(__uint_as_float(acc[4 * q + 1]) + bb.y);
            const float v2 = gelu_bf16(__uint_as_float(acc[4 * q + 2]) + bb.z);
            const float v3 = gelu_bf16(__uint_as_float(acc[4 * q + 3]) + bb.w);
            packed[2 * q] = pack_bf16x2(v0, v1);
            packed[2 * q + 1] = pack_bf16x2(v2, v3);
          }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(h_empty + 8 * hb);  // H[hb] may be overwritten by MMA1 of chunk g + hbufs
        // 32 columns = four 16-byte pieces of K chunk (colgrp >> 1), pieces 4*(colgrp & 1) ..
        const uint32_t dst_row = a2_base + (ab * 2 + (colgrp >> 1)) * kChunkBytes + row * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (MLP_DBG(p, 4)) break;
          const int piece = 4 * (colgrp & 1) + i;
          const uint32_t dst = dst_row + (static_cast<uint32_t>(piece ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * i]), "r"(packed[4 * i + 1]),
                       "r"(packed[4 * i + 2]), "r"(packed[4 * i + 3])
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) arrive(a2_full + 8 * ab);
        ++g;
        if (++hb == static_cast<uint32_t>(p.hbufs)) { hb = 0; hph ^= 1u; }
        if (j == 0 && pend) {
          y_epilogue();
          pend = false;
        }
      }
      // this tile's hidden chunks are done: request its residual rows now, write its Y tile after the next GELU chunk
      pend = true;
      pgrow = grow;
      pvalid = valid;
      prow0 = static_cast<int>((static_cast<long long>(grp) * p.cl + rank) * 128);
      if (p.y_staged) {
        if (io_thread) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's store has read the buffer
          mbar_expect_tx(res_full, static_cast<uint32_t>(p.ystage_bytes));
          for (int b = 0; b < p.c / 32; ++b)
            tma_load_2d(ystage_base + static_cast<uint32_t>(b) * (128u * 64u), &map_r, res_full, b * 32, prow0);
        }
      } else if (valid && !MLP_DBG(p, 8)) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int c0 = 16 * colgrp + 64 * u;
          if (c0 < p.c) {
            // one 256-bit load per 16 columns: the 32 lanes of a warp read 32 different rows, so every load instruction
            // costs 32 L1 tag cycles whatever its width
            if (p.wide_io) {
              ldg256(p.residual + grow * p.res_stride + c0, resq[u][0], resq[u][1]);
            } else {
              resq[u][0] = __ldg(reinterpret_cast<const uint4*>(p.residual + grow * p.res_stride + c0));
              resq[u][1] = __ldg(reinterpret_cast<const uint4*>(p.residual + grow * p.res_stride + c0 + 8));
            }
          }
        }
      }
    }
    if (pend) y_epilogue();
    if (p.y_staged && io_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (p.cl > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into its ring or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace
}  // namespace mspi

using namespace mspi;

static int mlp_fused_impl(const void* x, const void* w1, const float* b1, const void* w2, const float* scale,
                          const float* shift, const void* residual, void* y, int64_t m, int c, int c_pad,
                          int64_t res_stride, int64_t y_stride, const float* ln_w, const float* ln_b, float ln_eps,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(x && w1 && b1 && w2 && scale && shift && residual && y && m > 0, "mspi_mlp_fused: null argument");
  MSPI_CHECK_ARG((c == 96 || c == 192) && c_pad % 64 == 0 && c_pad >= c && c_pad < c + 64, "mspi_mlp_fused: C %d / pad %d unsupported", c, c_pad);
  MSPI_CHECK_ARG(res_stride % 8 == 0 && y_stride % 8 == 0, "row strides must be multiples of 8 elements");
  MSPI_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2) |
                   reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "16-byte alignment");
  tc::EncodeTiledFn encode = tc::get_encode_fn();
  if (!encode) return set_error(MSPI_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if (num_sms() <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  const int hidden = 4 * c;
  auto make2d = [&](CUtensorMap* map, const void* ptr, uint64_t k, uint64_t rows, uint64_t row_stride_elems, uint32_t box_rows) {
    cuuint64_t gdim[2] = {k, rows};
    cuuint64_t gstr[1] = {row_stride_elems * 2};
    cuuint32_t bdim[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap map_x, map_w1, map_w2, map_r, map_y;
  memset(&map_r, 0, sizeof(map_r));
  memset(&map_y, 0, sizeof(map_y));
  CUresult r1 = make2d(&map_x, x, c, static_cast<uint64_t>(m), c, 128);
  // MSPI_MLP_CLUSTER: 1 = one CTA per tile; 2 | 4 = cluster sharing the weight stream by TMA multicast (measured: no gain, an
  // SM still ingests every piece); -2 (default) = CTA pair (tcgen05 cta_group::2): each SM ingests HALF of every weight piece
  int cl = -2;
  if (const char* e = getenv("MSPI_MLP_CLUSTER")) cl = atoi(e);
  const bool pair = cl == -2 && m > 128;
  if (cl == -2) cl = pair ? 2 : 1;
  if (cl != 1 && cl != 2 && cl != 4) cl = 1;
  CUresult r2 = make2d(&map_w1, w1, c_pad, hidden, c_pad, 128 / cl);
  CUresult r3 = make2d(&map_w2, w2, hidden, c, hidden, c / cl);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS || r3 != CUDA_SUCCESS)
    return set_error(MSPI_ERR_CUDA, "mspi_mlp_fused: cuTensorMapEncodeTiled failed (%d, %d, %d)", (int)r1, (int)r2, (int)r3);

  MlpParams p;
  memset(&p, 0, sizeof(p));
  p.m_rows = static_cast<int>(m);
  p.c = c;
  p.kc1 = c_pad / 64;
  p.nh = hidden / kHC;
  p.m_tiles = static_cast<int>((m + 127) / 128);
  p.ring_stage_bytes = (c > 128 ? c : 128) * 128 / (pair ? 2 : 1);   // pair: a slot holds this CTA's half of a piece
  p.b1 = b1;
  p.scale = scale;
  p.shift = shift;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.res_stride = res_stride;
  p.y_stride = y_stride;
  p.ln_w = ln_w;
  p.ln_b = ln_b;
  p.ln_eps = ln_eps;
  static const bool wide_on = [] { const char* e = getenv("MSPI_MLP_WIDE_IO"); return !e || atoi(e) != 0; }();
  p.wide_io = wide_on && (reinterpret_cast<uintptr_t>(residual) & 31) == 0 && (reinterpret_cast<uintptr_t>(y) & 31) == 0 &&
              (res_stride * 2) % 32 == 0 && (y_stride * 2) % 32 == 0;
  // Y epilogue through a [128 x C] staging tile (C = 96: 24 KB taken from the weight ring, which is not the limiter)
  static const bool stage_on = [] { const char* e = getenv("MSPI_MLP_STAGE_Y"); return !e || atoi(e) != 0; }();
  if (stage_on && c <= 128 && c % 32 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (res_stride * 2) % 16 == 0 && (y_stride * 2) % 16 == 0) {
    auto make_io = [&](CUtensorMap* map, const void* ptr, int64_t stride) {
      cuuint64_t gdim[2] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(m)};
      cuuint64_t gstr[1] = {static_cast<cuuint64_t>(stride) * 2};
      cuuint32_t bdim[2] = {32, 128};
      cuuint32_t estr[2] = {1, 1};
      return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (make_io(&map_r, residual, res_stride) == CUDA_SUCCESS && make_io(&map_y, y, y_stride) == CUDA_SUCCESS) {
      p.y_staged = 1;
      p.ystage_bytes = 128 * c * 2;
    }
  }
  const uint32_t mdim = pair ? 256u : 128u;   // pair: one MMA covers the 128 rows of both CTAs
  p.idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kHC >> 3) << 17) | ((mdim >> 4) << 24);
  p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(c >> 3) << 17) | ((mdim >> 4) << 24);
  p.x_bufs = c <= 96 ? 2 : 1;  // C = 192: the weight ring needs the room (30 pieces per tile, each one L2 round trip)
  const int fixed = 1024 + 1024 + kVecBytes + p.x_bufs * p.kc1 * kChunkBytes + 4 * kChunkBytes + p.ystage_bytes;
  p.ring_stages = (225 * 1024 - fixed) / p.ring_stage_bytes;
  if (p.ring_stages > kMaxRing) p.ring_stages = kMaxRing;
  MSPI_CHECK_ARG(p.ring_stages >= 2, "mspi_mlp_fused: shared memory leaves %d ring stages", p.ring_stages);
  const size_t smem = static_cast<size_t>(fixed) + static_cast<size_t>(p.ring_stages) * p.ring_stage_bytes;
  const bool lnk = ln_w != nullptr;
  auto kern = pair ? (lnk ? fused_mlp_kernel<true, true> : fused_mlp_kernel<true, false>)
                   : (lnk ? fused_mlp_kernel<false, true> : fused_mlp_kernel<false, false>);
  MSPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  p.cl = cl;
  if (const char* e = getenv("MSPI_MLP_DEBUG")) p.debug = atoi(e);
  static const int packed_on = [] { const char* e = getenv("MSPI_MLP_PACKED_GELU"); return e ? atoi(e) : 1; }();
  p.packed_gelu = packed_on;
  p.hbufs = (3 * kHC + c <= 512) ? 3 : 2;
  const int groups = (p.m_tiles + cl - 1) / cl;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent grid: as many clusters as can be resident at once (1 CTA per SM), never more than there are tile groups
  static int max_clusters[4][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
  const int ki = (pair ? 1 : 0) + (lnk ? 2 : 0);
  if (max_clusters[ki][cl] == 0) {
    cfg.gridDim = dim3(num_sms() / cl * cl, 1, 1);
    int n = 0;
    MSPI_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    max_clusters[ki][cl] = n > 0 ? n : 1;
  }
  int nclusters = max_clusters[ki][cl];
  if (nclusters > num_sms() / cl) nclusters = num_sms() / cl;
  if (nclusters > groups) nclusters = groups;
  cfg.gridDim = dim3(nclusters * cl, 1, 1);
  cfg.numAttrs = 1 + pdl_attr(&attr[1]);   // (the occupancy query above ran without it)
  MSPI_CUDA(cudaLaunchKernelEx(&cfg, kern, map_x, map_w1, map_w2, map_r, map_y, p));
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_mlp_fused(const void* x, const void* w1, const float* b1, const void* w2, const float* scale,
                              const float* shift, const void* residual, void* y, int64_t m, int c, int c_pad,
                              int64_t res_stride, int64_t y_stride, void* stream) {
  return mlp_fused_impl(x, w1, b1, w2, scale, shift, residual, y, m, c, c_pad, res_stride, y_stride, nullptr, nullptr, 0.f,
                        stream);
}

extern "C" int mspi_mlp_fused_ln(const void* x, const void* w1, const float* b1, const void* w2, const float* scale,
                                 const float* shift, const void* residual, void* y, int64_t m, int c, int c_pad,
                                 int64_t res_stride, int64_t y_stride, const float* ln_weight, const float* ln_bias,
                                 float ln_eps, void* stream) {
  MSPI_CHECK_ARG(ln_weight && ln_bias && ln_eps > 0.f, "mspi_mlp_fused_ln: null LayerNorm vector");
  return mlp_fused_impl(x, w1, b1, w2, scale, shift, residual, y, m, c, c_pad, res_stride, y_stride, ln_weight, ln_bias, ln_eps,
                        stream);
}
