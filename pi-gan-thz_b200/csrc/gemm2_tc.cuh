// Two-CTA variant of the TN GEMM (cta_group::2): a pair of CTAs on one TPC computes a 256 x 256 tile, each CTA
// holding 128 rows of A, HALF of the B tile (128 of its 256 rows) and the 128 x 256 accumulator of its own rows in its
// tensor memory.  Why: with both operands in shared memory a single-CTA 128 x 256 x 64 k-block costs 48 KB of TMA
// writes plus 48 KB of UMMA operand reads per 512 MMA cycles — ~190 B/clk against 128 B/clk of shared-memory bandwidth
// (measured ceiling of the one-CTA mainloop: 1175 TFLOP/s at K=512, 890 at K=256).  Splitting B across
// the pair brings that to 32 + 32 KB per k-block, which the shared memory can sustain.
//
// Roles per CTA (320 threads, as in gemm_tc.cuh): warp 0 = TMA producer (own A rows + own half of B; the transaction
// bytes of BOTH CTAs land on the LEADER's full barrier), warp 1 = tcgen05 issuer (leader CTA only; its commits are
// multicast to both CTAs' barriers), warps 2..9 = two epilogue groups draining this CTA's accumulators with the same
// Epi interface as the one-CTA kernel (UnitInfo::m_tile = 2 * pair_tile + cluster rank).
#pragma once
#include "gemm_launch.cuh"

namespace pigan {

template <int STAGES_>
struct Gemm2Cfg {
  static constexpr int BLOCK_N = 256;       // per pair; each CTA stages BLOCK_N / 2 rows of B
  static constexpr int ACC_TILES = 1;
  static constexpr int STAGES = STAGES_;
  static constexpr bool MN_MAJOR = false;
  static constexpr int ACC_COLS = 256;
  static constexpr int ACC_BUFS = 2;
  static constexpr int B_HALF_BYTES = (BLOCK_N / 2) * kBlockK * 2;   // 16 KB
  static constexpr int STAGE_BYTES = kATileBytes + B_HALF_BYTES;     // 32 KB
  static constexpr int TX_BYTES_PAIR = 2 * STAGE_BYTES;              // both CTAs' loads of one stage
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES + 256;
  static constexpr int SMEM_BYTES = PIPE_BYTES + 1024;
};

__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in both CTAs once all prior MMAs of this thread completed
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (bit 24 of a shared::cluster address is
// the CTA's rank inside the pair)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}

template <class Cfg, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_x, const GemmShape g,
                const __grid_constant__ typename Epi::Params ep) {
  static_assert(Cfg::SMEM_BYTES + 2 * Epi::SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(!Epi::SPLIT && Epi::CLUSTER == 1 && !epi_early_release<Epi>::value,
                "two-CTA kernel: plain alternating epilogues only");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_smem = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = epi_smem + 2 * Epi::SMEM_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pairs = (int)gridDim.x / 2;
  const int pair_id = (int)blockIdx.x / 2;
  const int num_m_pairs = (g.num_m_tiles + 1) / 2;
  const int num_units = num_m_pairs * g.num_n_groups;   // k_splits == 1

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // the leader's producer; bytes arrive from both CTAs
      mbar_init(empty_bar(s), 1);   // one multicast commit per k-block
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 256);  // the owning epilogue group of BOTH CTAs (leader's copy is the one waited on)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();   // both CTAs' barriers and tensor memory exist before anything crosses the pair
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = pair_id; u < num_units; u += pairs) {
        const int m_pair = u / g.num_n_groups, n_group = u % g.num_n_groups;
        const int m_tile = 2 * m_pair + (int)rank;
        const int n0 = n_group * Cfg::BLOCK_N + (int)rank * (Cfg::BLOCK_N / 2);
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), Cfg::TX_BYTES_PAIR);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + kATileBytes;
          if (g.a_tail && kb == g.num_k_blocks - 1)
            tma_load_2d_2cta(sa, &tmap_x, full_bar(stage), 0, m_tile * kBlockM);
          else
            tma_load_2d_2cta(sa, &tmap_a, full_bar(stage), kb * kBlockK, m_tile * kBlockM);
          tma_load_2d_2cta(sb, &tmap_b, full_bar(stage), kb * kBlockK, n0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * kBlockM, Cfg::BLOCK_N, 0, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = pair_id; u < num_units; u += pairs, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait_cluster(tempty_bar(buf), (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * Cfg::ACC_COLS);
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + kATileBytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32u, 16u, 1024);
            const uint64_t db = umma_desc_sw128(sb + k * 32u, 16u, 1024);
            umma_f16_2cta(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2cta(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2cta(tfull_bar(buf));
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue groups (both CTAs)
    EpiCtx cx;
    cx.group = (warp - 2) >> 2;
    cx.smem = epi_smem + cx.group * Epi::SMEM_BYTES;
    cx.smem0 = epi_smem;
    cx.xbar = 0;
    cx.tempty = 0;
    cx.tid = threadIdx.x - 64 - 128 * cx.group;
    cx.q = warp & 3;
    cx.lane = lane;
    typename Epi::State st;
    Epi::init(ep, st, g, cx);
    int it = 0;
    for (int u = pair_id; u < num_units; u += pairs, ++it) {
      const int buf = it & 1;
      if (buf != cx.group) continue;
      UnitInfo w;
      w.m_tile = 2 * (u / g.num_n_groups) + (int)rank;
      w.n_group = u % g.num_n_groups;
      w.split = 0;
      w.kb_begin = 0;
      w.kb_end = g.num_k_blocks;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(tfull_bar(buf), use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(cx.q * 32) << 16) + (uint32_t)(buf * Cfg::ACC_COLS);
      Epi::unit(ep, st, g, w, tacc, cx);
      tc_fence_before();
      mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0));   // the leader's MMA thread waits for both CTAs
    }
    Epi::finish(ep, st, g, cx);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

template <class Cfg, class Epi>
int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb_half, const GemmShape& g,
                 const typename Epi::Params& ep, cudaStream_t st, const CUtensorMap* tx = nullptr) {
  auto kern = gemm2_tc_kernel<Cfg, Epi>;
  static bool attr_done = false;
  if (!attr_done) {
    PIGAN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES + 2 * Epi::SMEM_BYTES));
    attr_done = true;
  }
  if (g.k_splits != 1 || g.pair_mode) return fail(PIGAN_ERR_INVALID, "two-CTA GEMM: no split-K / pair mode");
  const int units = ((g.num_m_tiles + 1) / 2) * g.num_n_groups;
  if (units <= 0) return PIGAN_OK;
  int ctas = sm_count();
  if (ctas <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device");
  int grid = 2 * units < ctas ? 2 * units : ctas;
  grid -= grid % 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES + 2 * Epi::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  note_launch();
  PIGAN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb_half, tx ? *tx : tb_half, g, ep));
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

}  // namespace pigan
