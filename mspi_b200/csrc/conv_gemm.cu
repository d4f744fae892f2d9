// Implicit-GEMM convolution / linear kernel for sm_100a.
//
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (UMMA 128 x BN x 16,
//   cta_group::1, accumulators in TMEM, double buffered) -> tcgen05.ld epilogue (scale/shift = folded
//   BatchNorm or bias, activation, residual) -> global stores into a channel slice of the output.
//
// One CTA per SM, persistent over output tiles, warp specialised:
//   warp 0     : TMA producer (one elected lane)
//   warp 1     : TMEM allocation + MMA issue (one elected lane)
//   warps 2..9 : epilogue, two groups of four warps; warp w owns TMEM lanes 32*(w%4) .. +31, each group
//                takes every second 128-byte column chunk of the tile, stages it (swizzled) in shared
//                memory and one thread TMA-stores it (cp.async.bulk.tensor ... bulk_group).  Per-thread
//                row stores would write half sectors from 32 different lines per instruction, which
//                the memory system sustains at < 0.5 TB/s; the bulk store writes whole lines.
//
// The convolution is never materialised as an im2col matrix: an M-tile is a box of output positions in
// the (d1,d2,d3,d4) view of the NDHWC activation, and for filter tap `tap` the A operand is the same box
// shifted by tap_off[tap]; TMA zero-fills whatever falls outside the tensor, which is exactly the
// convolution's zero padding.  See include/mspi_b200.h (MspiConvDesc) for the contract.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mspi {
namespace {

#ifndef MSPI_EPI_SETS
#define MSPI_EPI_SETS 2
#endif
#ifndef MSPI_EPI_STAGE_BUFS
#define MSPI_EPI_STAGE_BUFS 1
#endif
constexpr int kEpiSets = MSPI_EPI_SETS;       // independent groups of 8 epilogue warps working on alternate column chunks
// Output staging buffers per epilogue set: with one, a set's next chunk cannot be written to shared memory before the bulk
// store of its previous chunk has finished READING the buffer (cp.async.bulk.wait_group.read 0 on the chunk's critical path);
// with two, that store drains while the next chunk is computed and staged (wait_group.read 1).  Measured (round 2, same
// box): two buffers are SLOWER — stage-2 fc1+GELU 0.232 -> 0.240 ms, stage-0 fc1 0.833 -> 0.882 ms — because the 32 KB come
// out of the operand ring (one pipeline stage less); the store's read-back is not what bounds the chunk chain.  Default 1.
constexpr int kStageBufs = kEpiSets > 1 ? MSPI_EPI_STAGE_BUFS : 2;
constexpr int kEpiWarps = 8 * kEpiSets;       // per set: two warps per TMEM lane quarter, they split the chunk's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;  // producer warp + MMA warp + epilogue warps
constexpr int kTileM = 128;
constexpr int kRowBytes = 128;               // one K chunk of one row: 64 bf16 or 32 tf32
constexpr int kABytes = kTileM * kRowBytes;  // 16 KB
constexpr int kBarrierBytes = 1024;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 220 * 1024;

// Study build (-DMSPI_GEMM_STUDY, loaded through MSPI_LIB): cycles the epilogue warps spend per phase, summed over warps.
//   [0] waiting for the accumulator (tmem_full)   [1] tcgen05.ld + wait::ld   [2] scale/shift/activation/pack
//   [3] staging buffer free (bulk wait_group.read + barrier)   [4] st.shared + proxy fence + barrier + bulk store issue
//   [5] chunks   [6] tiles   [7] whole epilogue loop
__device__ unsigned long long g_gemm_epi_cycles[8];
#ifdef MSPI_GEMM_STUDY
#define GEMM_CLK() clock64()
#else
#define GEMM_CLK() 0ll
#endif

struct GemmParams {
  int box[4];
  int tiles_d[4];
  int n_tiles, m_tiles;
  int ntaps, kchunks, cin_pad, bk_elems;
  int tap_off[MSPI_MAX_TAPS][4];
  int cout, bn;
  int o_dims[4];
  long long o_strides[4];
  long long r_strides[4];
  int o_dtype, r_dtype, act, has_res, res_after_act;
  const float* scale;
  const float* shift;
  const void* residual;
  void* y;
  uint32_t idesc;
  int num_stages, a_bytes, b_bytes, a_tx_bytes, tmem_cols;
  int row_bytes;  // K chunk of one row: 128 (default), 64 or 32 bytes; selects the swizzle mode of the operand tiles
  int ss_floats;  // staged scale/shift length (n_tiles * bn)
  int cl;         // CTAs per cluster that share every B (weight) tile by TMA multicast: 1 or 2
  int w_batched;  // 1: the B operand has its own matrix per (d3, d4) index of the M tile (batched GEMM, attention)
  int tma_store;  // 1: epilogue stages 128-byte output rows in shared memory and TMA-stores them
  int pair;       // 1: CTA pair (cl == 2, tcgen05 cta_group::2): b_bytes is this CTA's HALF of the B tile
  uint32_t idesc2;
  int res_tma;    // 1: the residual chunk arrives by TMA in the output staging buffer and is updated in place (fused_mlp.cu does
                  // the same): lane-per-row global loads of it cost 6500 cycles per chunk on the epilogue's critical path
  int y_box_bytes;  // bytes of one staged output / residual chunk: rows of the M box x 128
  int n_stage_bufs; // 16 KB output staging buffers behind the epilogue vectors
  // LayerNorm epilogue (LN instance): every GEMM row holds ln_groups groups of bn / ln_groups channels (the ConvNeXt stem
  // computes two output pixels per row); each group is normalised over its channels in fp32 and stored as bf16
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  int ln_groups;
  int ln_extra;     // bytes behind the staged scale / shift vectors: weight and bias per column, partial sums
};

using namespace tc;

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
// ---- CTA pair (tcgen05 cta_group::2): the two CTAs of a cluster compute one 256-row tile; each holds its own 128 rows of A
// and HALF of the B tile in shared memory, the leader (rank 0) issues the MMAs for both, accumulators land in each CTA's own
// TMEM.  Every operand load of either CTA signals the LEADER's full barrier.
__device__ __forceinline__ uint32_t leader_smem(uint32_t local) {   // same offset in CTA rank 0 of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local));
  return r;
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == MSPI_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tc_commit2_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------ kernel
// ACT / RES are compile-time for the layer types the model uses (ACT: MSPI_ACT_*, RES: 0 none, 1 added before
// the activation, 2 added after it); -1 selects the run-time value from the parameters (generic instance).
// PAIR: the CTA-pair (cta_group::2) variant.  It is a separate instantiation because a kernel that contains cta_group::2
// instructions can only be launched with an even cluster width (a plain launch fails with "cluster misconfiguration").
template <int KIND, int OUT, int ACT, int RES, bool PAIR = false, bool LN = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_y, const __grid_constant__ CUtensorMap tma_r,
                 const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr bool pair = PAIR;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // barrier block: full[8] empty[8] tmem_full[2] tmem_empty[2] tmem_ptr
  const uint32_t bar_full = smem_base;
  const uint32_t bar_empty = smem_base + 8 * kMaxStages;
  const uint32_t bar_tfull = smem_base + 16 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t bar_res = tmem_slot + 16;   // one per epilogue set: residual chunk landed in the staging buffer
  // staged epilogue vectors: scale[ss_floats], shift[ss_floats] (fp32); then, 1024 B aligned, the two output
  // staging buffers (TMA-store path) and the operand ring
  float* s_scale = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + kBarrierBytes);
  float* s_shift = s_scale + p.ss_floats;
  const uint32_t stage_base = (smem_base + kBarrierBytes + 8u * p.ss_floats + static_cast<uint32_t>(p.ln_extra) + 1023u) & ~1023u;
  const uint32_t tiles_base = stage_base + (p.tma_store ? static_cast<uint32_t>(p.n_stage_bufs) * kABytes : 0u);
  float* s_lnw = s_shift + p.ss_floats;     // LN instance: [bn] weight, [bn] bias (per GEMM column), then the partial sums
  float* s_lnb = s_lnw + p.bn;
  float* s_part = s_lnb + p.bn;             // [2][4 column parts][128 rows]
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();   // the next kernel of the stream may take this CTA's SM as soon as the CTA exits

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_y) : "memory");
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      // one (multicast) commit per CTA of the cluster; in pair mode only the leader's MMA warp commits
      mbar_init(bar_empty + 8 * s, pair ? 1u : static_cast<uint32_t>(p.cl));
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, pair ? 2 * kEpiWarps : kEpiWarps);  // pair: both CTAs' epilogues release the leader
    }
    for (int s = 0; s < kEpiSets; ++s) mbar_init(bar_res + 8 * s, 1);
    if (p.res_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_r) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (pair) {  // the same warp of both CTAs allocates the same columns in both TMEMs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"(static_cast<uint32_t>(p.tmem_cols))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"(static_cast<uint32_t>(p.tmem_cols))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // Everything above ran while the previous kernel of the stream was still finishing; from here on this kernel reads what
  // that kernel wrote (activations; in a training plan also the scale / shift vectors of the batch statistics).
  pdl_wait();
  for (int i = threadIdx.x; i < p.ss_floats; i += kThreads) {
    s_scale[i] = (p.scale != nullptr && i < p.cout) ? __ldg(p.scale + i) : 1.f;
    s_shift[i] = (p.shift != nullptr && i < p.cout) ? __ldg(p.shift + i) : 0.f;
  }
  if constexpr (LN) {
    const int gw = p.bn / p.ln_groups;
    for (int i = threadIdx.x; i < p.bn; i += kThreads) {
      s_lnw[i] = __ldg(p.ln_w + i % gw);
      s_lnb[i] = __ldg(p.ln_b + i % gw);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int k_iters = p.ntaps * p.kchunks;
  // Work items: with a cluster of `cl` CTAs, item q = (group of cl consecutive M tiles, N tile); CTA `rank` takes M tile
  // group*cl + rank.  All CTAs of a cluster walk the same items in lockstep because every B tile is fetched once per
  // cluster (each CTA loads 1/cl of its rows and multicasts them); an M tile past the end is a dummy (loads zero-fill,
  // nothing is stored).  cl == 1 degenerates to the plain persistent loop over tiles.
  uint32_t rank = 0, cid = blockIdx.x, ncl = gridDim.x;
  if (p.cl > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
    cluster_sync_all();  // every CTA's barriers are initialised before anyone multicasts into them
  }
  const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cl) - 1u);
  const int total_items = ((p.m_tiles + p.cl - 1) / p.cl) * p.n_tiles;
  constexpr int kFar = 0x3fffffff;  // a coordinate outside every tensor: dummy tiles read zeros and store nothing

  if (warp == 0) {
    // ================================================================== TMA producer
    // Warp-uniform control flow (all lanes wait on the barriers), one elected lane issues: with `if (lane == 0)` ptxas
    // wraps every uniform-datapath instruction (TMA, tcgen05) in a per-lane loop, and the single issuing thread's
    // instruction latency — not the tensor pipe — bounds the small-N / small-K layers (profiles/r01: fused MLP study).
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cid; item < total_items; item += ncl) {
        const int nt = item % p.n_tiles;
        int mt = (item / p.n_tiles) * p.cl + rank;
        const bool dummy = mt >= p.m_tiles;
        int org[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          org[j] = (mt % p.tiles_d[j]) * p.box[j];
          mt /= p.tiles_d[j];
        }
        if (dummy) org[3] = kFar;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int c1 = org[0] + p.tap_off[tap][0], c2 = org[1] + p.tap_off[tap][1];
          const int c3 = org[2] + p.tap_off[tap][2], c4 = org[3] + p.tap_off[tap][3];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if constexpr (pair) {
              // both CTAs load their A tile and their half of the B tile; everything lands on the leader's full barrier
              if (issuer) {
                const uint32_t full = bar_full + 8 * stage;
                if (rank == 0) mbar_expect_tx(full, 2u * static_cast<uint32_t>(p.a_tx_bytes + p.b_bytes));
                const uint32_t full_l = leader_smem(full);
                const uint32_t sa = tiles_base + stage * stage_bytes;
                tma_load_5d_2sm(sa, &tma_a, full_l, kc * p.bk_elems, c1, c2, c3, c4);
                tma_load_2d_2sm(sa + p.a_bytes, &tma_b, full_l, tap * p.cin_pad + kc * p.bk_elems,
                                nt * p.bn + static_cast<int>(rank) * (p.bn / 2));
              }
            } else if (issuer) {
              const uint32_t full = bar_full + 8 * stage;
              mbar_expect_tx(full, static_cast<uint32_t>(p.a_tx_bytes + p.b_bytes));
              const uint32_t sa = tiles_base + stage * stage_bytes;
              tma_load_5d(sa, &tma_a, full, kc * p.bk_elems, c1, c2, c3, c4);
              if (p.w_batched) {
                tma_load_4d(sa + p.a_bytes, &tma_b, full, tap * p.cin_pad + kc * p.bk_elems, nt * p.bn, org[2], org[3]);
              } else if (p.cl > 1) {
                const int share = p.bn / p.cl;  // rows of the B tile this CTA fetches for the whole cluster
                tma_load_2d_mc(sa + p.a_bytes + rank * share * p.row_bytes, &tma_b, full, tap * p.cin_pad + kc * p.bk_elems,
                               nt * p.bn + rank * share, mc_mask);
              } else {
                tma_load_2d(sa + p.a_bytes, &tma_b, full, tap * p.cin_pad + kc * p.bk_elems, nt * p.bn);
              }
            }
            __syncwarp();
            if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int mmas = p.row_bytes >> 5;  // one UMMA consumes 32 bytes of K per row
      if constexpr (pair) {
        // CTA pair: the leader alone issues (M = 256: its own 128 rows + the peer's), commits go to both CTAs
        if (rank == 0) {
          for (int item = cid; item < total_items; item += ncl) {
            mbar_wait(bar_tempty + 8 * as, aphase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * p.bn);
            for (int it = 0; it < k_iters; ++it) {
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              if (issuer) {
                const uint32_t sa = tiles_base + stage * stage_bytes;
                const uint64_t adesc = make_smem_desc(sa, p.row_bytes);
                const uint64_t bdesc = make_smem_desc(sa + p.a_bytes, p.row_bytes);
                for (int k = 0; k < mmas; ++k)
                  tc_mma2<KIND>(tmem_d, adesc + 2u * k, bdesc + 2u * k, p.idesc2, (it | k) != 0 ? 1u : 0u);
                tc_commit2_mc(bar_empty + 8 * stage, 3);
              }
              __syncwarp();
              if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
            }
            if (issuer) tc_commit2_mc(bar_tfull + 8 * as, 3);
            __syncwarp();
            if (++as == 2) { as = 0; aphase ^= 1u; }
          }
        }
      } else
      for (int item = cid; item < total_items; item += ncl) {
        mbar_wait(bar_tempty + 8 * as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * p.bn);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (issuer) {
            const uint32_t sa = tiles_base + stage * stage_bytes;
            const uint64_t adesc = make_smem_desc(sa, p.row_bytes);
            const uint64_t bdesc = make_smem_desc(sa + p.a_bytes, p.row_bytes);
            if (mmas == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma<KIND>(tmem_d, adesc + 2u * k, bdesc + 2u * k, p.idesc, (it | k) != 0 ? 1u : 0u);
            } else {
              for (int k = 0; k < mmas; ++k) tc_mma<KIND>(tmem_d, adesc + 2u * k, bdesc + 2u * k, p.idesc, (it | k) != 0 ? 1u : 0u);
            }
            // frees the smem slot (in every CTA of the cluster: the next B tile is multicast into all of them)
            if (p.cl > 1) tc_commit_mc(bar_empty + 8 * stage, mc_mask); else tc_commit(bar_empty + 8 * stage);
          }
          __syncwarp();
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
        if (issuer) tc_commit(bar_tfull + 8 * as);  // accumulator complete
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ================================================================== epilogue warps
    // The eight warps of a SET work on the same 128-byte column chunk of the tile (64 bf16 / 32 fp32 columns): warp w reads
    // TMEM lanes 32*(w%4).. (its rows) and the half of the chunk given by its group, HC columns.  A chunk is one serial
    // chain (tcgen05.ld -> math -> st.shared -> proxy fence -> barrier -> bulk store, ~2000 cycles: 4 elements/clk/SM, the
    // bound of every small-K layer), so kEpiSets sets run the chain on alternate chunks side by side, each with its own
    // staging buffer and named barrier.
    const int quarter = warp & 3;
    const int set = (warp - 2) >> 3;
    const int group = ((warp - 2) >> 2) & 1;
    const int row = quarter * 32 + lane;
    const bool leader = warp == 2 + 8 * set && lane == 0;
    int as = 0;
    uint32_t aphase = 0;
    constexpr bool out_bf16 = OUT == MSPI_BF16;
    constexpr int HC = out_bf16 ? 32 : 16;  // columns per thread per chunk (16 packed registers either way)
    constexpr int CH = 2 * HC;
    const int act = ACT >= 0 ? ACT : p.act;
    const bool has_res = RES >= 0 ? RES != 0 : p.has_res != 0;
    const bool res_after = RES >= 0 ? RES == 2 : p.res_after_act != 0;
    const int nchunks = (p.bn + CH - 1) / CH;
    // staging buffers: kEpiSets == 1 alternates two buffers (store_seq & 1); with two sets each owns one buffer — the other
    // set's chunk lies between two uses, so the previous bulk store has long finished reading it
    uint32_t store_seq = 0;
    const uint32_t my_stage = kEpiSets > 1 ? static_cast<uint32_t>(set) * kStageBufs * kABytes : 0u;
    const int bar_id = 1 + set;
    uint32_t chunk_seq = 0;  // column chunks processed by this CTA so far: chunk g of the kernel belongs to set g % kEpiSets
    long long cy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool res_tma = RES != 0 && p.res_tma != 0;   // residual through TMA + shared memory (needs the staged store path)
    uint32_t res_phase = 0;
    const uint32_t my_bar_res = bar_res + 8u * static_cast<uint32_t>(set);
    const long long t_loop0 = GEMM_CLK();
    for (int item = cid; item < total_items; item += ncl, chunk_seq += nchunks) {
      const int nt = item % p.n_tiles;
      int mt = (item / p.n_tiles) * p.cl + rank;
      const bool dummy = mt >= p.m_tiles;
      int r = row;
      bool valid = !dummy;
      long long yoff = 0, roff = 0;
      int org[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        org[j] = (mt % p.tiles_d[j]) * p.box[j];
        mt /= p.tiles_d[j];
        const int c = org[j] + (r % p.box[j]);
        r /= p.box[j];
        valid = valid && (c < p.o_dims[j]);
        yoff += static_cast<long long>(c) * p.o_strides[j];
        roff += static_cast<long long>(c) * p.r_strides[j];
      }
      valid = valid && (r == 0);
      if (dummy) org[3] = kFar;

      if constexpr (LN) {
        // ---- LayerNorm epilogue: the whole tile row (bn <= 192 columns) is in TMEM.  The 16 epilogue warps split it four
        // ways per TMEM lane quarter: thread (row, part) owns bn / 4 consecutive columns, adds the conv bias, and the
        // 4 / ln_groups threads that share a channel group exchange partial sums through shared memory — first the sums (mean),
        // then the squared deviations (two passes over registers: the stem's channels have |mean| >> sigma, so E[x^2] - mean^2
        // would cancel).  The normalised bf16 rows leave through the staging buffers and bulk tensor stores.
        mbar_wait(bar_tfull + 8 * as, aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * p.bn);
        const int part = (warp - 2) >> 2;
        const int cw = p.bn >> 2;                    // 16, 32 or 48 columns per thread
        const int c0 = part * cw;
        uint32_t acc[3][16];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (16 * u < cw) tmem_ld16(taddr + c0 + 16 * u, acc[u]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (pair) mbar_arrive_cluster(leader_smem(bar_tempty + 8 * as)); else mbar_arrive(bar_tempty + 8 * as);
        }
        float s1 = 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (16 * u < cw) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int c = c0 + 16 * u + e;
              const float v = fmaf(__uint_as_float(acc[u][e]), s_scale[c], s_shift[c]);
              acc[u][e] = __float_as_uint(v);
              s1 += v;
            }
          }
        const int ppg = 4 / p.ln_groups, g0 = (part / ppg) * ppg;
        const float inv_gw = static_cast<float>(p.ln_groups) / static_cast<float>(p.bn);
        s_part[part * 128 + row] = s1;
        asm volatile("bar.sync 3, %0;" ::"n"(kEpiWarps * 32) : "memory");
        float mean = 0.f;
        for (int q = 0; q < ppg; ++q) mean += s_part[(g0 + q) * 128 + row];
        mean *= inv_gw;
        float s2 = 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (16 * u < cw) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float dv = __uint_as_float(acc[u][e]) - mean;
              s2 = fmaf(dv, dv, s2);
            }
          }
        s_part[512 + part * 128 + row] = s2;
        // the bulk stores of the previous tile have finished reading the staging buffers before anyone writes them again
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 3, %0;" ::"n"(kEpiWarps * 32) : "memory");
        float var = 0.f;
        for (int q = 0; q < ppg; ++q) var += s_part[512 + (g0 + q) * 128 + row];
        const float rstd = rsqrtf(var * inv_gw + p.ln_eps);
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (16 * u < cw) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int c = c0 + 16 * u + 8 * h8;
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = fmaf((__uint_as_float(acc[u][8 * h8 + 2 * e]) - mean) * rstd, s_lnw[c + 2 * e], s_lnb[c + 2 * e]);
                const float b = fmaf((__uint_as_float(acc[u][8 * h8 + 2 * e + 1]) - mean) * rstd, s_lnw[c + 2 * e + 1], s_lnb[c + 2 * e + 1]);
                o[e] = pack_bf16x2(a, b);
              }
              const uint32_t piece = static_cast<uint32_t>(c) >> 3;       // 16-byte piece of the 128-byte chunk piece >> 3
              const uint32_t dst = stage_base + (piece >> 3) * kABytes + static_cast<uint32_t>(row) * kRowBytes +
                                   (((piece & 7u) ^ (static_cast<uint32_t>(row) & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
          }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 3, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (warp == 2 && lane == 0) {
          for (int cb = 0; cb * 64 < p.bn; ++cb)
            asm volatile(
                "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&tma_y),
                "r"(stage_base + static_cast<uint32_t>(cb) * kABytes), "r"(cb * 64), "r"(org[0]), "r"(org[1]), "r"(org[2]),
                "r"(org[3])
                : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (++as == 2) { as = 0; aphase ^= 1u; }
        continue;
      }

      // residual rows of this thread for one chunk (4 x 16 bytes), fetched early: issued just before use, every one of these
      // loads (32 different lines per warp instruction) exposes a full L2 / DRAM round trip to the chunk's critical path
      // (measured: 6500 cycles of "math" per chunk for ConvNeXt stage-2 fc2 + residual against 840 without the residual)
      const int n_base = nt * p.bn;
      const int first = kEpiSets > 1 ? static_cast<int>((static_cast<uint32_t>(set) + kEpiSets - chunk_seq % kEpiSets) % kEpiSets) : 0;
      uint4 rq[4];
#ifndef MSPI_EPI_RES_PREFETCH
#define MSPI_EPI_RES_PREFETCH 0   // measured (same box): stage-2 fc2 + residual 0.193 -> 0.206 ms WITH the register prefetch (16 more live
#endif                            // registers at the 96-register cap); the loads need a whole tile of lead, i.e. a TMA-staged tile
      const bool pre_ok = MSPI_EPI_RES_PREFETCH && has_res && (p.r_dtype == MSPI_BF16 || !out_bf16);
      auto fetch_res = [&](int ch_) {
        const int c0_ = ch_ * CH + group * HC;
        const int w_ = min(HC, p.bn - c0_);
#pragma unroll
        for (int g = 0; g < HC / 8; ++g) {
          const int ng = n_base + c0_ + 8 * g;
          if (valid && 8 * g < w_ && ng + 8 <= p.cout) {
            if (p.r_dtype == MSPI_BF16) {
              rq[g & 3] = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + roff + ng));
            } else {
              const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const float*>(p.residual) + roff + ng);
              rq[(2 * g) & 3] = __ldg(rp);
              rq[(2 * g + 1) & 3] = __ldg(rp + 1);
            }
          }
        }
      };
      if (pre_ok && first < nchunks) fetch_res(first);
      const long long t_w0 = GEMM_CLK();
      mbar_wait(bar_tfull + 8 * as, aphase);
      tc_fence_after();
      cy[0] += GEMM_CLK() - t_w0;
      cy[6] += 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(as * p.bn);
      // the sets alternate over the GLOBAL chunk sequence, so tiles with a single chunk (N <= 64 bf16 / 32 fp32: the stems,
      // most Inception branches) alternate between the sets as well and two tiles' chains overlap (`first`, above)
      if (first >= nchunks) {  // this set has no chunk in the tile: release the accumulator buffer (after the tfull wait above,
                               // so a set can never get ahead of the MMA warp and arrive twice in one phase)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (pair) mbar_arrive_cluster(leader_smem(bar_tempty + 8 * as)); else mbar_arrive(bar_tempty + 8 * as);
        }
      }
#pragma unroll 1
      for (int ch = first; ch < nchunks; ch += kEpiSets) {
        const int c0 = ch * CH + group * HC;
        const int width = min(HC, p.bn - c0);  // HC, 16 (bf16, odd multiple of 16), or <= 0 past the tile
        const int n0 = n_base + c0;
        if (res_tma && leader) {
          // the staging buffer is free once the bulk store of this set's previous chunk has finished READING it; the residual
          // chunk (same box, same swizzle as the store) then lands in it while the accumulator is read and scaled
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          const uint32_t sbuf = stage_base + (kEpiSets > 1 ? my_stage + (kStageBufs > 1 ? (store_seq & 1u) * kABytes : 0u)
                                                           : (store_seq & 1u) * kABytes);
          mbar_expect_tx(my_bar_res, static_cast<uint32_t>(p.y_box_bytes));
          tma_load_5d(sbuf, &tma_r, my_bar_res, n_base + ch * CH, org[0], org[1], org[2], org[3]);
        }
        uint32_t acc[HC];
        const long long t_c0 = GEMM_CLK();
        __syncwarp();  // tcgen05.ld is warp-collective
        if constexpr (out_bf16) {
          if (width >= 32) {
            tmem_ld32(taddr + c0, reinterpret_cast<uint32_t(&)[32]>(acc));
          } else if (width > 0) {
            uint32_t lo[16];
            tmem_ld16(taddr + c0, lo);
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc[j] = lo[j]; acc[HC - 16 + j] = 0u; }
          }
        } else {
          if (width > 0) tmem_ld16(taddr + c0, reinterpret_cast<uint32_t(&)[16]>(acc));
        }
        tmem_ld_wait();
        const long long t_c1 = GEMM_CLK();
        if (width <= 0) {
#pragma unroll
          for (int j = 0; j < HC; ++j) acc[j] = 0u;
        }
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale + n0);  // staged vectors are padded by 64
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift + n0);
        uint32_t packed[16];
#pragma unroll
        for (int g = 0; g < HC / 8; ++g) {  // groups of 8 columns
          const int ng = n0 + 8 * g;
          const bool live = (8 * g < width) && (ng < p.cout);
          const bool full8 = (ng + 8 <= p.cout);
          // Packed fp32 pairs end to end (fma.rn.f32x2 / add / mul: one issue slot per two elements; the epilogue warps are
          // bound by instruction issue — tools/prof_gemm_phases.py: 838 cycles of math per 32-element chunk without an
          // activation, 1680 with GELU, against 3072 cycles of MMA for a 128x256x384 tile).  TMEM registers are consecutive,
          // so (acc[2k], acc[2k+1]) already is a register pair.
          F2 v2[4];
          {
            const float4 s0 = sc4[2 * g], s1 = sc4[2 * g + 1], h0 = sh4[2 * g], h1 = sh4[2 * g + 1];
            v2[0] = fma2(pack2(__uint_as_float(acc[8 * g + 0]), __uint_as_float(acc[8 * g + 1])), pack2(s0.x, s0.y), pack2(h0.x, h0.y));
            v2[1] = fma2(pack2(__uint_as_float(acc[8 * g + 2]), __uint_as_float(acc[8 * g + 3])), pack2(s0.z, s0.w), pack2(h0.z, h0.w));
            v2[2] = fma2(pack2(__uint_as_float(acc[8 * g + 4]), __uint_as_float(acc[8 * g + 5])), pack2(s1.x, s1.y), pack2(h1.x, h1.y));
            v2[3] = fma2(pack2(__uint_as_float(acc[8 * g + 6]), __uint_as_float(acc[8 * g + 7])), pack2(s1.z, s1.w), pack2(h1.z, h1.w));
          }
          if (res_tma) {   // residual, activation and packing happen on the staged tile below
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float x0, x1;
              unpack2(v2[k], x0, x1);
              acc[8 * g + 2 * k] = __float_as_uint(x0);
              acc[8 * g + 2 * k + 1] = __float_as_uint(x1);
            }
            continue;
          }
          const bool have_res = has_res && valid && live;
          F2 r2[4];
          if (have_res) {
            if (pre_ok && full8) {   // prefetched before the accumulator wait / during the previous chunk's staging
              if (p.r_dtype == MSPI_BF16) {
                const uint4 a = rq[g];
                r2[0] = pack2(__uint_as_float(a.x << 16), __uint_as_float(a.x & 0xffff0000u));
                r2[1] = pack2(__uint_as_float(a.y << 16), __uint_as_float(a.y & 0xffff0000u));
                r2[2] = pack2(__uint_as_float(a.z << 16), __uint_as_float(a.z & 0xffff0000u));
                r2[3] = pack2(__uint_as_float(a.w << 16), __uint_as_float(a.w & 0xffff0000u));
              } else {
                const uint4 a = rq[(2 * g) & 3], b = rq[(2 * g + 1) & 3];
                r2[0] = pack2(__uint_as_float(a.x), __uint_as_float(a.y));
                r2[1] = pack2(__uint_as_float(a.z), __uint_as_float(a.w));
                r2[2] = pack2(__uint_as_float(b.x), __uint_as_float(b.y));
                r2[3] = pack2(__uint_as_float(b.z), __uint_as_float(b.w));
              }
            } else {
              float res[8];
              if (full8 && p.r_dtype == MSPI_BF16) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + roff + ng));
                unpack_bf16x2(a.x, res[0], res[1]); unpack_bf16x2(a.y, res[2], res[3]);
                unpack_bf16x2(a.z, res[4], res[5]); unpack_bf16x2(a.w, res[6], res[7]);
              } else if (full8) {
                const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + roff + ng);
                const float4 a = __ldg(rp), b = __ldg(rp + 1);
                res[0] = a.x; res[1] = a.y; res[2] = a.z; res[3] = a.w;
                res[4] = b.x; res[5] = b.y; res[6] = b.z; res[7] = b.w;
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int n = min(ng + j, p.cout - 1);
                  res[j] = p.r_dtype == MSPI_BF16 ? bf2f(static_cast<const __nv_bfloat16*>(p.residual)[roff + n])
                                                  : static_cast<const float*>(p.residual)[roff + n];
                }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) r2[k] = pack2(res[2 * k], res[2 * k + 1]);
            }
            if (!res_after) {
#pragma unroll
              for (int k = 0; k < 4; ++k) v2[k] = add2(v2[k], r2[k]);
            }
          }
          float v[8];
          if (out_bf16 && act == MSPI_ACT_GELU) {   // tanh form, one fp32 MUFU per element, everything else packed
#pragma unroll
            for (int k = 0; k < 4; ++k) v2[k] = gelu_pair_f32(v2[k]);
          } else if (act != MSPI_ACT_NONE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float x0, x1;
              unpack2(v2[k], x0, x1);
              v2[k] = pack2(epi_act<out_bf16>(x0, act), epi_act<out_bf16>(x1, act));
            }
          }
          if (have_res && res_after) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v2[k] = add2(v2[k], r2[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) unpack2(v2[k], v[2 * k], v[2 * k + 1]);
          if (p.tma_store) {
            if (out_bf16) {
              packed[4 * g + 0] = pack_bf16x2(v[0], v[1]);
              packed[4 * g + 1] = pack_bf16x2(v[2], v[3]);
              packed[4 * g + 2] = pack_bf16x2(v[4], v[5]);
              packed[4 * g + 3] = pack_bf16x2(v[6], v[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) packed[8 * g + j] = __float_as_uint(v[j]);
            }
          } else if (valid && live) {
            if (out_bf16) {
              __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(p.y) + yoff + ng;
              if (full8) {
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(yp) = o;
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (ng + j < p.cout) yp[j] = __float2bfloat16_rn(v[j]);
              }
            } else {
              float* yp = static_cast<float*>(p.y) + yoff + ng;
              if (full8 && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0)) {
                reinterpret_cast<float4*>(yp)[0] = make_float4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<float4*>(yp)[1] = make_float4(v[4], v[5], v[6], v[7]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (ng + j < p.cout) yp[j] = v[j];
              }
            }
          }
        }
        if (pre_ok && ch + kEpiSets < nchunks) fetch_res(ch + kEpiSets);   // lands while this chunk is staged and stored
        if (ch + kEpiSets >= nchunks) {
          // every TMEM read of this warp for this tile is done: hand the accumulator buffer back to the MMA warp now,
          // before the last chunk is staged and stored
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (pair) mbar_arrive_cluster(leader_smem(bar_tempty + 8 * as)); else mbar_arrive(bar_tempty + 8 * as);
          }
        }
        const long long t_c2 = GEMM_CLK();
        long long t_c3 = t_c2;
        if (p.tma_store) {
          // staging buffer (store_seq & 1) is free once the bulk store issued two chunks ago has finished READING it
          const uint32_t stage_buf = stage_base + (kEpiSets > 1 ? my_stage + (kStageBufs > 1 ? (store_seq & 1u) * kABytes : 0u)
                                                                : (store_seq & 1u) * kABytes);
          const uint32_t stage_row = stage_buf + row * kRowBytes;
          if (res_tma) {
            mbar_wait(my_bar_res, res_phase);   // the residual chunk has landed (the leader waited for the buffer before asking)
            res_phase ^= 1u;
            t_c3 = GEMM_CLK();
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // this thread's four 16-byte pieces: read the residual, combine, write back in place
              const uint32_t dst = stage_row + (static_cast<uint32_t>((4 * group + j) ^ (row & 7)) << 4);
              uint4 rr;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w) : "r"(dst) : "memory");
              uint4 o;
              if constexpr (out_bf16) {
                float rs[8], x[8];
                unpack_bf16x2(rr.x, rs[0], rs[1]); unpack_bf16x2(rr.y, rs[2], rs[3]);
                unpack_bf16x2(rr.z, rs[4], rs[5]); unpack_bf16x2(rr.w, rs[6], rs[7]);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float t = __uint_as_float(acc[8 * j + e]);
                  if (!res_after) t += rs[e];
                  t = epi_act<out_bf16>(t, act);
                  if (res_after) t += rs[e];
                  x[e] = t;
                }
                o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]);
                o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
              } else {
                const float rs[4] = {__uint_as_float(rr.x), __uint_as_float(rr.y), __uint_as_float(rr.z), __uint_as_float(rr.w)};
                float x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float t = __uint_as_float(acc[4 * j + e]);
                  if (!res_after) t += rs[e];
                  t = epi_act<out_bf16>(t, act);
                  if (res_after) t += rs[e];
                  x[e] = t;
                }
                o = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
            }
          } else {
          if (leader) {
            if (kEpiSets > 1 && kStageBufs == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
          t_c3 = GEMM_CLK();
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // this thread's four 16-byte pieces, 128B-swizzled like the tensor map expects
            const uint32_t dst = stage_row + (static_cast<uint32_t>((4 * group + j) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * j]), "r"(packed[4 * j + 1]),
                         "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3])
                         : "memory");
          }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
          if (leader) {
            asm volatile(
                "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(&tma_y),
                "r"(stage_buf), "r"(n_base + ch * CH), "r"(org[0]), "r"(org[1]), "r"(org[2]),
                "r"(org[3])
                : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++store_seq;
        }
        const long long t_c4 = GEMM_CLK();
        cy[1] += t_c1 - t_c0; cy[2] += t_c2 - t_c1; cy[3] += t_c3 - t_c2; cy[4] += t_c4 - t_c3; cy[5] += 1;
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
#ifdef MSPI_GEMM_STUDY
    cy[7] = GEMM_CLK() - t_loop0;
    if (lane == 0)
      for (int j = 0; j < 8; ++j) atomicAdd(&g_gemm_epi_cycles[j], static_cast<unsigned long long>(cy[j]));
#endif
    if (p.tma_store && leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (p.cl > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into its ring or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (pair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(static_cast<uint32_t>(p.tmem_cols))
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(static_cast<uint32_t>(p.tmem_cols))
                   : "memory");
  }
}

// -------------------------------------------------------------------------------- host side
}  // namespace
}  // namespace mspi

using namespace mspi;

extern "C" int mspi_debug_gemm_epilogue_cycles(uint64_t* out8, int reset) {
  unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  MSPI_CUDA(cudaDeviceSynchronize());
  if (out8 != nullptr) {
    MSPI_CUDA(cudaMemcpyFromSymbol(h, g_gemm_epi_cycles, sizeof(h)));
    for (int i = 0; i < 8; ++i) out8[i] = h[i];
  }
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    MSPI_CUDA(cudaMemcpyToSymbol(g_gemm_epi_cycles, z, sizeof(z)));
  }
  return MSPI_OK;
}

static int conv_gemm_impl(const MspiConvDesc* d, const void* x, const void* w, const float* scale, const float* shift,
                          const void* residual, void* y, const float* ln_w, const float* ln_b, float ln_eps, int ln_groups,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MSPI_CHECK_ARG(d && x && w && y, "mspi_conv_gemm: null argument");
  const bool ln = ln_w != nullptr;
  MSPI_CHECK_ARG(d->a_dtype == MSPI_BF16 || d->a_dtype == MSPI_F32, "a_dtype %d", d->a_dtype);
  const int elsize = d->a_dtype == MSPI_BF16 ? 2 : 4;
  const int row_bytes = d->k_row_bytes > 0 ? d->k_row_bytes : kRowBytes;
  MSPI_CHECK_ARG(row_bytes == 128 || row_bytes == 64 || row_bytes == 32, "k_row_bytes %d", row_bytes);
  const int bk = row_bytes / elsize;
  MSPI_CHECK_ARG(d->a_strides[0] == 1, "channel stride must be 1");
  MSPI_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= MSPI_MAX_TAPS, "ntaps %d", d->ntaps);
  MSPI_CHECK_ARG(d->bn >= 16 && d->bn <= 256 && d->bn % 16 == 0, "bn %d", d->bn);
  MSPI_CHECK_ARG(d->cin_pad > 0 && d->cin_pad % bk == 0, "cin_pad %d not a multiple of %d", d->cin_pad, bk);
  MSPI_CHECK_ARG(d->cout >= 1 && d->w_rows >= d->cout, "cout %d w_rows %d", d->cout, d->w_rows);
  MSPI_CHECK_ARG(!d->has_residual || residual, "residual requested but null");
  MSPI_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 "operands must be 16-byte aligned");
  long long rows = 1;
  for (int j = 1; j < 5; ++j) {
    MSPI_CHECK_ARG(d->box[j] >= 1 && d->box[j] <= 256, "box[%d]=%d", j, d->box[j]);
    MSPI_CHECK_ARG(d->a_dims[j] >= 1 && d->o_dims[j - 1] >= 1, "dims[%d]", j);
    MSPI_CHECK_ARG((d->a_strides[j] * elsize) % 16 == 0, "a_strides[%d] not 16B aligned", j);
    rows *= d->box[j];
  }
  MSPI_CHECK_ARG(rows <= kTileM, "box has %lld rows (> %d)", rows, kTileM);
  if (d->o_dtype == MSPI_BF16 && d->cout >= 8)
    for (int j = 0; j < 4; ++j)
      MSPI_CHECK_ARG(d->o_strides[j] % 8 == 0 || d->o_dims[j] == 1, "o_strides[%d] must be a multiple of 8", j);
  MSPI_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "output must be 16-byte aligned");

  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return set_error(MSPI_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");

  CUtensorMap map_a, map_b;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5] = {1, 1, 1, 1, 1};
    for (int j = 0; j < 5; ++j) gdim[j] = static_cast<cuuint64_t>(d->a_dims[j]);
    for (int j = 1; j < 5; ++j) gstr[j - 1] = static_cast<cuuint64_t>(d->a_strides[j]) * elsize;
    bdim[0] = bk;
    for (int j = 1; j < 5; ++j) bdim[j] = d->box[j];
    CUresult r = encode(&map_a, d->a_dtype == MSPI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32,
                        5, const_cast<void*>(x), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(MSPI_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d dims=[%d,%d,%d,%d,%d] box=[%d,%d,%d,%d,%d]",
                       (int)r, d->a_dims[0], d->a_dims[1], d->a_dims[2], d->a_dims[3], d->a_dims[4], bk,
                       d->box[1], d->box[2], d->box[3], d->box[4]);
  }
  const bool w_batched = d->w_batch_dims[0] > 0;
  // The layers are bound by what one SM can take in through TMA (~40 B/clk/SM, 11 TB/s over the chip: 40 KB per
  // 128x192x64 k-step = 940 clk against 384 clk of MMA).  A 2-CTA cluster that shares every B tile by TMA multicast
  // (MSPI_GEMM_CLUSTER=1) does not reduce the bytes an SM receives and measured no gain; the CTA pair (tcgen05
  // cta_group::2, mode 2, the default) does: the two SMs compute one 256-row tile and each ingests HALF of the B tile
  // (measured: readout.1 1.93 -> 1.67 ms, ConvNeXt stage-3 MLP GEMMs 0.18 -> 0.15 ms = 1.33 PF/s, step 39.9 -> 38.7 ms).
  int cl = 1, pair = 0;
  {
    long long m_tiles_est = 1;
    for (int j = 0; j < 4; ++j) m_tiles_est *= (d->o_dims[j] + d->box[j + 1] - 1) / d->box[j + 1];
    // MSPI_GEMM_CLUSTER: 0 = one CTA per tile, 1 = 2-CTA cluster with TMA-multicast B tiles, 2 (default) = CTA pair
    static const int mode = [] { const char* e = getenv("MSPI_GEMM_CLUSTER"); return e ? atoi(e) : 2; }();
    if (mode == 1 && !w_batched && d->bn % 16 == 0 && m_tiles_est >= 2) cl = 2;
    // mode 2: CTA pair (tcgen05 cta_group::2): each SM ingests HALF of every B tile
    if (mode == 2 && !w_batched && d->bn % 16 == 0 && row_bytes == 128 && m_tiles_est >= 2) { cl = 2; pair = 1; }
  }
  if (w_batched) {
    MSPI_CHECK_ARG(d->ntaps == 1 && d->w_batch_dims[0] == d->o_dims[2] && d->w_batch_dims[1] == d->o_dims[3] &&
                       d->box[3] == 1 && d->box[4] == 1,
                   "batched weights: one matrix per (d3, d4) index, boxes of 1 along d3/d4, a single tap");
    for (int j = 0; j < 3; ++j)
      MSPI_CHECK_ARG(d->w_strides[j] > 0 && (d->w_strides[j] * elsize) % 16 == 0, "w_strides[%d] not 16B aligned", j);
  }
  {
    cuuint64_t gdim[4] = {w_batched ? static_cast<cuuint64_t>(d->a_dims[0]) : static_cast<cuuint64_t>(d->ntaps) * d->cin_pad,
                          static_cast<cuuint64_t>(d->w_rows), 1, 1};
    cuuint64_t gstr[3] = {gdim[0] * elsize, 0, 0};
    cuuint32_t bdim[4] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(d->bn / cl), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (w_batched) {
      gdim[2] = static_cast<cuuint64_t>(d->w_batch_dims[0]);
      gdim[3] = static_cast<cuuint64_t>(d->w_batch_dims[1]);
      for (int j = 0; j < 3; ++j) gstr[j] = static_cast<cuuint64_t>(d->w_strides[j]) * elsize;
    }
    CUresult r = encode(&map_b, d->a_dtype == MSPI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32,
                        w_batched ? 4 : 2, const_cast<void*>(w), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(MSPI_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d K=%llu rows=%d bn=%d", (int)r,
                       (unsigned long long)gdim[0], d->w_rows, d->bn);
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = 1;
  for (int j = 0; j < 4; ++j) {
    p.box[j] = d->box[j + 1];
    p.o_dims[j] = d->o_dims[j];
    p.tiles_d[j] = (p.o_dims[j] + p.box[j] - 1) / p.box[j];
    p.m_tiles *= p.tiles_d[j];
    p.o_strides[j] = d->o_strides[j];
    p.r_strides[j] = d->r_strides[j];
  }
  p.n_tiles = (d->cout + d->bn - 1) / d->bn;
  p.ntaps = d->ntaps;
  p.bk_elems = bk;
  p.kchunks = (d->a_dims[0] + bk - 1) / bk;
  MSPI_CHECK_ARG(p.kchunks * bk <= d->cin_pad, "cin_pad %d smaller than padded channels %d", d->cin_pad, p.kchunks * bk);
  p.cin_pad = d->cin_pad;
  memcpy(p.tap_off, d->tap_off, sizeof(p.tap_off));
  p.cout = d->cout;
  p.bn = d->bn;
  p.o_dtype = d->o_dtype;
  p.r_dtype = d->r_dtype;
  p.act = d->act;
  p.has_res = d->has_residual;
  p.res_after_act = d->res_after_act;
  p.scale = scale;
  p.shift = shift;
  p.residual = residual;
  p.y = y;
  const uint32_t fmt = d->a_dtype == MSPI_BF16 ? 1u : 2u;  // UMMA F16F32Format: BF16 = 1, TF32 = 2
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(d->bn >> 3) << 17) |
            (static_cast<uint32_t>(kTileM >> 4) << 24);
  p.row_bytes = row_bytes;
  p.w_batched = w_batched ? 1 : 0;
  p.cl = cl;
  p.pair = pair;
  p.idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(d->bn >> 3) << 17) |
             (static_cast<uint32_t>((2 * kTileM) >> 4) << 24);
  p.a_bytes = kTileM * row_bytes;
  p.b_bytes = (pair ? d->bn / 2 : d->bn) * row_bytes;
  p.a_tx_bytes = static_cast<int>(rows) * row_bytes;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  p.ss_floats = p.n_tiles * d->bn + 64;  // padded: the last staged chunk may run past the tile
  if (ln) {
    MSPI_CHECK_ARG(ln_b && ln_eps > 0.f && (ln_groups == 1 || ln_groups == 2 || ln_groups == 4), "LayerNorm epilogue: bad argument");
    MSPI_CHECK_ARG(d->a_dtype == MSPI_BF16 && d->o_dtype == MSPI_BF16 && d->act == MSPI_ACT_NONE && !d->has_residual && !w_batched,
                   "LayerNorm epilogue: bf16 operands and output, no activation, no residual");
    MSPI_CHECK_ARG(d->cout == d->bn && d->bn % 64 == 0 && d->bn <= 192 && kEpiWarps == 16,
                   "LayerNorm epilogue: one N tile of 64, 128 or 192 columns (cout %d, bn %d)", d->cout, d->bn);
    p.ln_w = ln_w;
    p.ln_b = ln_b;
    p.ln_eps = ln_eps;
    p.ln_groups = ln_groups;
    p.ln_extra = 8 * d->bn + 2 * 4 * 128 * 4;
  }
  const int ss_bytes = (8 * p.ss_floats + p.ln_extra + 1023) & ~1023;
  MSPI_CHECK_ARG(ss_bytes <= 64 * 1024, "cout %d too large for the staged epilogue vectors", d->cout);

  // Output path: bulk tensor stores need 16-byte aligned row strides and N tiles that end on a 128-byte chunk.
  const int o_es = d->o_dtype == MSPI_BF16 ? 2 : 4;
  const int chunk_cols = kRowBytes / o_es;
  p.tma_store = (p.n_tiles == 1 || d->bn % chunk_cols == 0) ? 1 : 0;
  for (int j = 0; j < 4; ++j)
    if (p.o_dims[j] > 1 && (p.o_strides[j] * o_es) % 16 != 0) p.tma_store = 0;
  CUtensorMap map_y = map_a;
  if (p.tma_store) {
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5] = {1, 1, 1, 1, 1};
    gdim[0] = static_cast<cuuint64_t>(d->cout);
    bdim[0] = static_cast<cuuint32_t>(chunk_cols);
    cuuint64_t span = static_cast<cuuint64_t>((d->cout * o_es + 15) / 16 * 16);
    for (int j = 0; j < 4; ++j) {
      gdim[j + 1] = static_cast<cuuint64_t>(p.o_dims[j]);
      bdim[j + 1] = static_cast<cuuint32_t>(p.box[j]);
      cuuint64_t st = static_cast<cuuint64_t>(p.o_strides[j]) * o_es;
      if (p.o_dims[j] == 1 && (st == 0 || st % 16 != 0)) st = span;  // never dereferenced: any legal stride
      gstr[j] = st;
      if (st * gdim[j + 1] > span) span = st * gdim[j + 1];
    }
    CUresult r = encode(&map_y, d->o_dtype == MSPI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        5, y, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(MSPI_ERR_CUDA, "cuTensorMapEncodeTiled(Y) failed: %d cout=%d dims=[%d,%d,%d,%d] strides=[%lld,%lld,%lld,%lld]",
                       (int)r, d->cout, d->o_dims[0], d->o_dims[1], d->o_dims[2], d->o_dims[3], (long long)d->o_strides[0],
                       (long long)d->o_strides[1], (long long)d->o_strides[2], (long long)d->o_strides[3]);
  }
  // Residual through TMA: same box / swizzle as the staged store, so the chunk lands exactly where the epilogue writes its
  // result.  Needs the staged store path, a residual of the output's dtype and 16-byte aligned residual row strides.
  CUtensorMap map_r = map_y;
  p.y_box_bytes = static_cast<int>(rows) * kRowBytes;
  {
    static const bool res_tma_on = [] { const char* e = getenv("MSPI_GEMM_RES_TMA"); return !e || atoi(e) != 0; }();
    bool ok = res_tma_on && p.tma_store && d->has_residual && residual != nullptr && d->r_dtype == d->o_dtype &&
              (reinterpret_cast<uintptr_t>(residual) & 15) == 0;
    for (int j = 0; j < 4 && ok; ++j)
      if (p.o_dims[j] > 1 && (p.r_strides[j] * o_es) % 16 != 0) ok = false;
    if (ok) {
      cuuint64_t gdim[5], gstr[4];
      cuuint32_t bdim[5], estr[5] = {1, 1, 1, 1, 1};
      gdim[0] = static_cast<cuuint64_t>(d->cout);
      bdim[0] = static_cast<cuuint32_t>(chunk_cols);
      cuuint64_t span = static_cast<cuuint64_t>((d->cout * o_es + 15) / 16 * 16);
      for (int j = 0; j < 4; ++j) {
        gdim[j + 1] = static_cast<cuuint64_t>(p.o_dims[j]);
        bdim[j + 1] = static_cast<cuuint32_t>(p.box[j]);
        cuuint64_t st = static_cast<cuuint64_t>(p.r_strides[j]) * o_es;
        if (p.o_dims[j] == 1 && (st == 0 || st % 16 != 0)) st = span;  // never dereferenced: any legal stride
        gstr[j] = st;
        if (st * gdim[j + 1] > span) span = st * gdim[j + 1];
      }
      CUresult r = encode(&map_r, d->o_dtype == MSPI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                          5, const_cast<void*>(residual), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { ok = false; map_r = map_y; }
    }
    p.res_tma = ok ? 1 : 0;
  }
  if (ln) MSPI_CHECK_ARG(p.tma_store, "LayerNorm epilogue: output rows must allow bulk tensor stores (16-byte aligned strides)");
  p.n_stage_bufs = ln ? (d->bn + 63) / 64 : (kEpiSets > 1 ? kEpiSets * kStageBufs : 2);
  if (ln && p.n_stage_bufs < 2) p.n_stage_bufs = 2;
  const int out_stage_bytes = p.tma_store ? p.n_stage_bufs * kABytes : 0;
  p.num_stages = (kSmemBudget - kBarrierBytes - 1024 - ss_bytes - out_stage_bytes) / stage_bytes;
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  int cols = 32;
  while (cols < 2 * d->bn) cols <<= 1;
  p.tmem_cols = cols;
  MSPI_CHECK_ARG(p.num_stages >= 2, "shared memory budget leaves %d pipeline stages", p.num_stages);
  const size_t smem = 1024 + kBarrierBytes + ss_bytes + out_stage_bytes + static_cast<size_t>(p.num_stages) * stage_bytes;

  // Specialised instances for the (operand kind, output dtype, activation, residual) combinations the model
  // uses; anything else runs the generic instance that reads activation / residual mode at run time.
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);
  const int res_mode = d->has_residual ? (d->res_after_act ? 2 : 1) : 0;
  Kern kern = nullptr;
#define MSPI_PICK(K, O, A, R)                                                      \
  if (d->a_dtype == K && d->o_dtype == O && d->act == A && res_mode == R)          \
    kern = pair ? conv_gemm_kernel<K, O, A, R, true> : conv_gemm_kernel<K, O, A, R, false>;
  MSPI_PICK(MSPI_BF16, MSPI_BF16, MSPI_ACT_RELU, 0)
  MSPI_PICK(MSPI_BF16, MSPI_BF16, MSPI_ACT_NONE, 0)
  MSPI_PICK(MSPI_BF16, MSPI_BF16, MSPI_ACT_GELU, 0)
  MSPI_PICK(MSPI_BF16, MSPI_BF16, MSPI_ACT_NONE, 2)
  MSPI_PICK(MSPI_BF16, MSPI_BF16, MSPI_ACT_RELU, 1)
  MSPI_PICK(MSPI_BF16, MSPI_F32, MSPI_ACT_NONE, 0)
  MSPI_PICK(MSPI_BF16, MSPI_F32, MSPI_ACT_RELU, 0)
  MSPI_PICK(MSPI_F32, MSPI_F32, MSPI_ACT_NONE, 0)
  MSPI_PICK(MSPI_F32, MSPI_F32, MSPI_ACT_RELU, 0)
  MSPI_PICK(MSPI_F32, MSPI_F32, MSPI_ACT_GELU, 0)
  MSPI_PICK(MSPI_F32, MSPI_F32, MSPI_ACT_NONE, 2)
  MSPI_PICK(MSPI_F32, MSPI_BF16, MSPI_ACT_GELU, 0)
  MSPI_PICK(MSPI_BF16, MSPI_F32, MSPI_ACT_NONE, 2)
  MSPI_PICK(MSPI_BF16, MSPI_F32, MSPI_ACT_RELU, 1)
#undef MSPI_PICK
  if (kern == nullptr) {
    if (pair) {
      if (d->a_dtype == MSPI_BF16) kern = d->o_dtype == MSPI_BF16 ? conv_gemm_kernel<MSPI_BF16, MSPI_BF16, -1, -1, true>
                                                                  : conv_gemm_kernel<MSPI_BF16, MSPI_F32, -1, -1, true>;
      else kern = d->o_dtype == MSPI_BF16 ? conv_gemm_kernel<MSPI_F32, MSPI_BF16, -1, -1, true>
                                          : conv_gemm_kernel<MSPI_F32, MSPI_F32, -1, -1, true>;
    } else if (d->a_dtype == MSPI_BF16) kern = d->o_dtype == MSPI_BF16 ? conv_gemm_kernel<MSPI_BF16, MSPI_BF16, -1, -1>
                                                                : conv_gemm_kernel<MSPI_BF16, MSPI_F32, -1, -1>;
    else kern = d->o_dtype == MSPI_BF16 ? conv_gemm_kernel<MSPI_F32, MSPI_BF16, -1, -1>
                                        : conv_gemm_kernel<MSPI_F32, MSPI_F32, -1, -1>;
  }
  if (ln)
    kern = pair ? conv_gemm_kernel<MSPI_BF16, MSPI_BF16, MSPI_ACT_NONE, 0, true, true>
                : conv_gemm_kernel<MSPI_BF16, MSPI_BF16, MSPI_ACT_NONE, 0, false, true>;
  MSPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const long long total = static_cast<long long>(p.m_tiles) * p.n_tiles;
  MSPI_CHECK_ARG(total < (1ll << 31), "too many tiles");
  int grid = num_sms();
  if (grid <= 0) return set_error(MSPI_ERR_CUDA, "no CUDA device");
  if (cl == 1) {
    if (total < grid) grid = static_cast<int>(total);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(&attr[0]);
    MSPI_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_y, map_r, p));
  } else {
    const long long items = static_cast<long long>((p.m_tiles + cl - 1) / cl) * p.n_tiles;
    long long nclusters = grid / cl;
    if (items < nclusters) nclusters = items;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(nclusters * cl), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1 + pdl_attr(&attr[1]);
    MSPI_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_y, map_r, p));
  }
  MSPI_LAUNCH_CHECK();
  return MSPI_OK;
}

extern "C" int mspi_conv_gemm(const MspiConvDesc* d, const void* x, const void* w, const float* scale,
                              const float* shift, const void* residual, void* y, void* stream) {
  return conv_gemm_impl(d, x, w, scale, shift, residual, y, nullptr, nullptr, 0.f, 1, stream);
}

extern "C" int mspi_conv_gemm_ln(const MspiConvDesc* d, const void* x, const void* w, const float* scale, const float* shift,
                                 const float* ln_weight, const float* ln_bias, float ln_eps, int ln_groups, void* y,
                                 void* stream) {
  MSPI_CHECK_ARG(ln_weight && ln_bias, "mspi_conv_gemm_ln: null LayerNorm vector");
  return conv_gemm_impl(d, x, w, scale, shift, nullptr, y, ln_weight, ln_bias, ln_eps, ln_groups, stream);
}
