// Persistent warp-specialised tcgen05 GEMM used by every hidden-layer contraction of the hot path
// (reference: nn.Linear forward/backward in core/models/{generator,discriminator,forward_model}.py).
//
//   TN mode  (MN_MAJOR = false):  D[M,N] = A[M,K] * B[N,K]^T      A,B fp16, K contiguous ("K-major")
//            forward layers (A = activations, B = weight [out,in]) and dX (B = weight^T copy).
//            Optional "A tail": the last 64-wide k-block of A comes from a second tensor (tmap_x) — the
//            spectrum tile whose spare columns carry the 4 structure parameters and a constant 1, so
//            torch.cat + bias of D's first layer ride inside the MMA.
//   NT mode  (MN_MAJOR = true):   D[M,N] = A[Kd,M]^T * B[Kd,N]    reduction over rows (the batch)
//            dW = dY^T * X with split-K over CTAs; operands are read as stored, no transposes.
//            Optional "B tail": for k-blocks >= b_tail_from_kb the last 64-column box of the last n-tile
//            comes from tmap_x (fake-row tail of the spectrum tensor).
//
// One CTA per SM, 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warps 2..5 and 6..9 = two epilogue groups (TMEM -> registers -> fused math -> smem -> TMA store).
// Operand tiles go through a STAGES-deep TMA/mbarrier ring in 128B-swizzled shared memory; fp32
// accumulators live in TMEM.  When two accumulator buffers fit (2*ACC_COLS <= 512) group e owns buffer e
// and handles every second unit, so two epilogues overlap each other and the MMAs of later units —
// measured on B200 the epilogue, not the tensor pipe, is what bounds these K<=1024 layers.
#pragma once
#include "ptx.cuh"

namespace pigan {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 halves = 128 B = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 320;   // 2 epilogue groups (the default); see gemm_threads<Epi>()
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB

struct GemmShape {
  int M;                // valid output rows
  int N;                // valid output columns
  int num_m_tiles;      // ceil(M / 128)
  int num_n_groups;     // ceil(N / (ACC_TILES * BLOCK_N))
  int num_k_blocks;     // ceil(K / 64)
  int k_splits;         // >= 1; units = m_tiles * n_groups * k_splits
  int b_wrap_k_blocks;  // NT mode: B operand row block index is taken modulo this (0 = off)
  int a_tail;           // TN mode: last k-block of A is read from tmap_x at column 0
  int a_split_kb;       // TN mode: k-blocks >= a_split_kb of A come from tmap_x, column (kb - a_split_kb) * 64
                        // (K-concatenation of two operands: [x | z] . [W | s W]^T); < 0: off
  int b_tail_from_kb;   // NT mode: k-blocks >= this read the last B box of the last n-tile from tmap_x (<0 off)
  int pair_mode;        // a CTA handles the two n-groups of an m-tile back to back (whole-row epilogues that split
                        // the row between the two epilogue groups / accumulator buffers); k_splits must be 1 and
                        // num_n_groups == 2 * cluster size
};

struct UnitInfo {
  int m_tile, n_group, kb_begin, kb_end, split;
};

__device__ __forceinline__ UnitInfo decode_unit(const GemmShape& g, int u) {
  UnitInfo w;
  const int tiles = g.num_m_tiles * g.num_n_groups;
  const int tile = u % tiles;
  w.split = u / tiles;
  w.m_tile = tile / g.num_n_groups;
  w.n_group = tile % g.num_n_groups;
  w.kb_begin = (int)(((long long)w.split * g.num_k_blocks) / g.k_splits);
  w.kb_end = (int)(((long long)(w.split + 1) * g.num_k_blocks) / g.k_splits);
  return w;
}

// it-th unit of this CTA (false: no more).  Every role of the kernel walks the same sequence.
template <int CLUSTER>
__device__ __forceinline__ bool get_unit(const GemmShape& g, int it, UnitInfo& w) {
  if (g.pair_mode) {
    const int m_tile = (int)blockIdx.x / CLUSTER + (it >> 1) * ((int)gridDim.x / CLUSTER);
    if (m_tile >= g.num_m_tiles) return false;
    w.m_tile = m_tile;
    w.n_group = 2 * ((int)blockIdx.x % CLUSTER) + (it & 1);
    w.split = 0;
    w.kb_begin = 0;
    w.kb_end = g.num_k_blocks;
    return true;
  }
  const int u = (int)blockIdx.x + it * (int)gridDim.x;
  if (u >= g.num_m_tiles * g.num_n_groups * g.k_splits) return false;
  w = decode_unit(g, u);
  return true;
}

// B_RES_KB > 0 ("resident B", TN mode): the CTA keeps the whole B operand of ITS n-group - B_RES_KB k-blocks, e.g.
// 4 x 32 KB for a 256-column group of a K = 256 layer - in shared memory for the lifetime of the kernel and only
// streams A.  Measured on B200 (profiles/r02_*): with both operands streamed a 128 x 256 x 256 unit moves 64 KB of A,
// 128 KB of B and 64 KB of output through the L2 <-> SM fabric (~42 B/clk per SM) for 2048 MMA cycles, i.e. the
// K = 256 layers were bound by re-reading the weights once per 128 rows; resident weights halve that traffic.
// Needs gridDim.x % num_n_groups == 0 (a CTA then sees one n-group only) and K <= 64 * B_RES_KB.
template <int BLOCK_N_, int ACC_TILES_, int STAGES_, bool MN_MAJOR_, int B_RES_KB_ = 0>
struct GemmCfg {
  static constexpr int BLOCK_N = BLOCK_N_;
  static constexpr int ACC_TILES = ACC_TILES_;
  static constexpr int STAGES = STAGES_;
  static constexpr bool MN_MAJOR = MN_MAJOR_;
  static constexpr int B_RES_KB = B_RES_KB_;
  static constexpr bool B_RESIDENT = B_RES_KB_ > 0;
  static constexpr int ACC_COLS = BLOCK_N * ACC_TILES;
  static constexpr int ACC_BUFS = (2 * ACC_COLS <= 512) ? 2 : 1;
  static constexpr int B_TILE_BYTES = BLOCK_N * kBlockK * 2;
  // keep every stage base 1024-B aligned (swizzle atom)
  static constexpr int B_TILE_ALLOC = (B_TILE_BYTES + 1023) / 1024 * 1024;
  static constexpr int RES_BYTES = B_RES_KB * B_TILE_ALLOC;
  static constexpr int STAGE_BYTES = kATileBytes + (B_RESIDENT ? 0 : B_TILE_ALLOC);
  static constexpr int TX_BYTES = kATileBytes + (B_RESIDENT ? 0 : B_TILE_BYTES);
  static constexpr int PIPE_BYTES = RES_BYTES + STAGES * STAGE_BYTES + 256 /*barriers*/;
  static constexpr int SMEM_BYTES = PIPE_BYTES + 1024 /*align slack*/;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
  static_assert(!MN_MAJOR || BLOCK_N % 64 == 0, "NT mode loads B in 64-wide boxes");
  static_assert(!B_RESIDENT || (!MN_MAJOR && ACC_TILES == 1), "resident B: TN mode, one accumulator tile per unit");
  static_assert(2 * STAGES + 10 <= 32, "barrier block");
};

// What an epilogue thread knows about itself.
struct EpiCtx {
  uint32_t smem;  // shared-memory address of this group's scratch region (Epi::SMEM_BYTES, 1024-aligned)
  int tid;        // 0..127 within the epilogue group
  int q;          // TMEM lane quarter of this warp (row in tile = q*32 + lane)
  int lane;
  int group;      // epilogue group 0/1 (0..3 with Epi::GROUPS = 4: accumulator buffer = group & 1, column half = group >> 1)
  uint32_t smem0; // scratch region of group 0 (data shared by both groups of a SPLIT epilogue lives there)
  uint32_t tempty; // "accumulator drained" mbarrier of the unit being processed (EARLY_RELEASE epilogues arrive on it)
  uint32_t xbar;  // this group's two cluster-exchange mbarriers (alternating between units)
};
// Epilogues that define `static constexpr bool EARLY_RELEASE = true` release the accumulator buffer themselves.
template <class E, class = void>
struct epi_early_release { static constexpr bool value = false; };
template <class E>
struct epi_early_release<E, decltype((void)E::EARLY_RELEASE)> { static constexpr bool value = E::EARLY_RELEASE; };
// Epilogues with `static constexpr bool FIXED_NGROUP = true` need every CTA to see one n-group only (gemm_launch.cuh)
template <class E, class = void>
struct epi_fixed_ngroup { static constexpr bool value = false; };
template <class E>
struct epi_fixed_ngroup<E, decltype((void)E::FIXED_NGROUP)> { static constexpr bool value = E::FIXED_NGROUP; };
// arrivals per phase on the cluster-exchange barriers: Epi::XBAR_COUNT when defined, else 128 per CTA of the cluster
template <class E, class = void>
struct epi_xbar_count { static constexpr int value = 128 * E::CLUSTER; };
template <class E>
struct epi_xbar_count<E, decltype((void)E::XBAR_COUNT)> { static constexpr int value = E::XBAR_COUNT; };
// Epilogue groups of 4 warps: 2 by default (group e owns accumulator buffer e), Epi::GROUPS = 4 puts four warps on
// every SM sub-partition - groups g and g + 2 then share the units of buffer g & 1 and split their columns
// (EpiCtx::group >> 1 = column half).  Epilogues that are bound by instruction latency, not by issue slots, need that.
template <class E, class = void>
struct epi_groups { static constexpr int value = 2; };
template <class E>
struct epi_groups<E, decltype((void)E::GROUPS)> { static constexpr int value = E::GROUPS; };
// shared memory of all epilogue groups together: GROUPS * Epi::SMEM_BYTES unless the epilogue names a total
template <class E, class = void>
struct epi_smem_total { static constexpr int value = epi_groups<E>::value * E::SMEM_BYTES; };
template <class E>
struct epi_smem_total<E, decltype((void)E::SMEM_TOTAL)> { static constexpr int value = E::SMEM_TOTAL; };
template <class E>
constexpr int gemm_threads() { return 64 + 128 * epi_groups<E>::value; }
__device__ __forceinline__ void epi_bar_sync(const EpiCtx& cx, int which) {  // named barrier over one group
  asm volatile("bar.sync %0, 128;" ::"r"(1 + 2 * cx.group + which) : "memory");
}

// Epilogue contract (all static, called by the 128 threads of an epilogue group):
//   struct Params;  struct State;                       // State lives in registers across units
//   static constexpr int SMEM_BYTES;                     // per-group scratch (multiple of 1024)
//   init(const Params&, State&, const GemmShape&, const EpiCtx&)   // before the unit loop
//   unit(const Params&, State&, const GemmShape&, const UnitInfo&, uint32_t tmem_acc, const EpiCtx&)
//        tmem_acc = TMEM address of this unit's first accumulator column, lane field already set
//        to this warp's 32-lane quarter; row handled by the thread = m_tile*128 + q*32 + lane.
//   finish(const Params&, State&, const GemmShape&, const EpiCtx&)  // after the loop
//   static constexpr bool SPLIT    // true: BOTH groups process every unit and split its columns between them
//                                  // (whole-row epilogues: LayerNorm, the forward-model output layer); false:
//                                  // group e owns accumulator buffer e and handles every second unit
//   static constexpr int CLUSTER   // CTAs per cluster (1 or 2).  2: the kernel is launched in clusters of two
//                                  // CTAs that work on the same m-tile with n_group = cluster rank and may
//                                  // exchange per-row partials through distributed shared memory
template <class Cfg, class Epi, int AB_FMT>
__global__ void __launch_bounds__(gemm_threads<Epi>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_x, const GemmShape g,
               const __grid_constant__ typename Epi::Params ep) {
  static_assert(Cfg::SMEM_BYTES + epi_smem_total<Epi>::value <= 232448, "shared memory budget");
  static_assert(epi_groups<Epi>::value == 2 || (epi_groups<Epi>::value == 4 && !Epi::SPLIT && Cfg::ACC_BUFS == 2),
                "4 epilogue groups: two per accumulator buffer");
  static_assert(Epi::SMEM_BYTES % 1024 == 0, "epilogue scratch must keep 1024-B alignment");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t res_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // resident B k-blocks (B_RESIDENT), else empty
  const uint32_t smem_base = res_base + Cfg::RES_BYTES;                 // operand ring
  const uint32_t epi_smem = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;  // 1024-aligned
  const uint32_t bar_base = epi_smem + epi_smem_total<Epi>::value;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  const uint32_t bres_bar = bar_base + 8u * (2 * Cfg::STAGES + 5);  // resident B has landed
  const uint32_t xbar = bar_base + 8u * (2 * Cfg::STAGES + 6);  // cluster exchange of row partials: [group][unit parity]
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), Epi::SPLIT ? 256 : 64 * epi_groups<Epi>::value);
    }
    for (int b = 0; b < 4; ++b) mbar_init(xbar + 8u * b, epi_xbar_count<Epi>::value);   // see EpiLnStore
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (Epi::CLUSTER > 1) cluster_sync_all();  // the peer's barriers exist before anyone arrives on them
  const uint32_t tmem_base = *tmem_slot_ptr;
  // launched as a programmatic dependent (gemm_launch.cuh): everything above overlapped the previous kernel's tail;
  // from here on its results are read and its inputs overwritten.  No-op for an ordinary launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      UnitInfo w;
      if constexpr (Cfg::B_RESIDENT) {
        // every unit of this CTA has the same n-group (host: gridDim.x % num_n_groups == 0): fetch its weights once
        if (get_unit<Epi::CLUSTER>(g, 0, w)) {
          mbar_arrive_expect_tx(bres_bar, (uint32_t)g.num_k_blocks * Cfg::B_TILE_BYTES);
          for (int kb = 0; kb < g.num_k_blocks; ++kb)
            tma_load_2d(res_base + kb * Cfg::B_TILE_ALLOC, &tmap_b, bres_bar, kb * kBlockK,
                        w.n_group * Cfg::BLOCK_N);
        }
      }
      for (int it = 0; get_unit<Epi::CLUSTER>(g, it, w); ++it) {
        for (int t = 0; t < Cfg::ACC_TILES; ++t) {
          const int nt = w.n_group * Cfg::ACC_TILES + t;
          const int n0 = nt * Cfg::BLOCK_N;
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), Cfg::TX_BYTES);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + kATileBytes;
            if constexpr (!Cfg::MN_MAJOR) {
              if (g.a_tail && kb == g.num_k_blocks - 1)
                tma_load_2d(sa, &tmap_x, full_bar(stage), 0, w.m_tile * kBlockM);
              else if (g.a_split_kb >= 0 && kb >= g.a_split_kb)
                tma_load_2d(sa, &tmap_x, full_bar(stage), (kb - g.a_split_kb) * kBlockK, w.m_tile * kBlockM);
              else
                tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kBlockK, w.m_tile * kBlockM);
              if constexpr (!Cfg::B_RESIDENT) tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kBlockK, n0);
            } else {
              const int kb_b = g.b_wrap_k_blocks ? (kb % g.b_wrap_k_blocks) : kb;
#pragma unroll
              for (int i = 0; i < kBlockM / 64; ++i)
                tma_load_2d(sa + i * 8192, &tmap_a, full_bar(stage), w.m_tile * kBlockM + i * 64,
                            kb * kBlockK);
#pragma unroll
              for (int j = 0; j < Cfg::BLOCK_N / 64; ++j) {
                const bool tail = (g.b_tail_from_kb >= 0) && (kb >= g.b_tail_from_kb) &&
                                  (j == Cfg::BLOCK_N / 64 - 1) && (nt == g.num_n_groups * Cfg::ACC_TILES - 1);
                if (tail)
                  tma_load_2d(sb + j * 8192, &tmap_x, full_bar(stage), 0, (kb - g.b_tail_from_kb) * kBlockK);
                else
                  tma_load_2d(sb + j * 8192, &tmap_b, full_bar(stage), n0 + j * 64, kb_b * kBlockK);
              }
            }
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc =
          umma_idesc_f16(kBlockM, Cfg::BLOCK_N, AB_FMT, Cfg::MN_MAJOR ? 1 : 0, Cfg::MN_MAJOR ? 1 : 0);
      constexpr uint32_t lbo = Cfg::MN_MAJOR ? 8192u : 16u;
      constexpr uint32_t kstep = Cfg::MN_MAJOR ? 2048u : 32u;  // bytes per UMMA_K step
      int stage = 0;
      uint32_t phase = 0;
      UnitInfo w;
      if constexpr (Cfg::B_RESIDENT) {
        if (get_unit<Epi::CLUSTER>(g, 0, w)) mbar_wait(bres_bar, 0);
      }
      for (int it = 0; get_unit<Epi::CLUSTER>(g, it, w); ++it) {
        const int buf = it % Cfg::ACC_BUFS;
        const uint32_t use = (uint32_t)(it / Cfg::ACC_BUFS);
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u);
        tc_fence_after();
        for (int t = 0; t < Cfg::ACC_TILES; ++t) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * Cfg::ACC_COLS + t * Cfg::BLOCK_N);
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = Cfg::B_RESIDENT ? res_base + (uint32_t)(kb * Cfg::B_TILE_ALLOC) : sa + kATileBytes;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t da = umma_desc_sw128(sa + k * kstep, lbo, 1024);
              const uint64_t db = umma_desc_sw128(sb + k * kstep, lbo, 1024);
              umma_f16(d_tmem, da, db, idesc, (kb > w.kb_begin || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(tfull_bar(buf));
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue groups
    EpiCtx cx;
    cx.group = (warp - 2) >> 2;
    cx.smem = epi_smem + cx.group * Epi::SMEM_BYTES;
    cx.smem0 = epi_smem;
    cx.xbar = xbar + 16u * (uint32_t)(cx.group & 1);
    cx.tid = threadIdx.x - 64 - 128 * cx.group;
    cx.q = warp & 3;  // TMEM lane quarter this warp may access
    cx.lane = lane;
    typename Epi::State st;
    Epi::init(ep, st, g, cx);
    UnitInfo w;
    for (int it = 0; get_unit<Epi::CLUSTER>(g, it, w); ++it) {
      const int buf = it % Cfg::ACC_BUFS;
      if (!Epi::SPLIT && (Cfg::ACC_BUFS == 2 ? buf : 0) != (cx.group & 1)) continue;  // groups e, e + 2 own buffer e
      const uint32_t use = (uint32_t)(it / Cfg::ACC_BUFS);
      mbar_wait(tfull_bar(buf), use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(cx.q * 32) << 16) + (uint32_t)(buf * Cfg::ACC_COLS);
      cx.tempty = tempty_bar(buf);
      Epi::unit(ep, st, g, w, tacc, cx);
      if constexpr (!epi_early_release<Epi>::value) {
        tc_fence_before();
        mbar_arrive(tempty_bar(buf));
      }
    }
    Epi::finish(ep, st, g, cx);
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (Epi::CLUSTER > 1) cluster_sync_all();  // nobody leaves while the peer may still write to it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pigan
