// Host-side launch helper + small epilogue utilities for gemm_tc_kernel.
#pragma once
#include "gemm_tc.cuh"
#include "host_util.h"

namespace pigan {

constexpr int kFmtF16 = 0;  // tcgen05 kind::f16 operand format codes
constexpr int kFmtBF16 = 1;

template <class Cfg, class Epi, int AB_FMT = kFmtF16>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& g,
                const typename Epi::Params& ep, cudaStream_t st, int max_ctas = 0,
                const CUtensorMap* tx = nullptr) {
  auto kern = gemm_tc_kernel<Cfg, Epi, AB_FMT>;
  static bool attr_done = false;
  if (!attr_done) {
    PIGAN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES + epi_smem_total<Epi>::value));
    attr_done = true;
  }
  // CTAs that can have work: one per unit, or one per (m-tile, cluster rank) in pair mode
  const int units = g.pair_mode ? g.num_m_tiles * Epi::CLUSTER : g.num_m_tiles * g.num_n_groups * g.k_splits;
  if (units <= 0) return PIGAN_OK;
  if (g.pair_mode && (g.k_splits != 1 || g.num_n_groups != 2 * Epi::CLUSTER || Cfg::ACC_BUFS != 2))
    return fail(PIGAN_ERR_INVALID, "pair mode needs num_n_groups == 2 * cluster size and two accumulator buffers");
  int ctas = sm_count();
  if (ctas <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device");
  if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
  int grid = units < ctas ? units : ctas;
  if constexpr (epi_fixed_ngroup<Epi>::value) {
    // the epilogue keeps per-column state for ONE n-group per CTA (u = block + it * grid, n_group = u % groups)
    if (g.k_splits != 1 || g.pair_mode) return fail(PIGAN_ERR_INVALID, "fixed-n-group epilogue: no split-K / pair mode");
    if (grid > g.num_n_groups) grid -= grid % g.num_n_groups;
  }
  if constexpr (Cfg::B_RESIDENT) {
    // a CTA keeps the weights of ONE n-group: its units must all share it (u = block + it * grid, n_group = u % groups)
    if (g.k_splits != 1 || g.pair_mode || g.num_k_blocks > Cfg::B_RES_KB)
      return fail(PIGAN_ERR_INVALID, "resident-B GEMM: K up to %d, no split-K / pair mode", Cfg::B_RES_KB * kBlockK);
    if (grid > g.num_n_groups) grid -= grid % g.num_n_groups;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(gemm_threads<Epi>());
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES + epi_smem_total<Epi>::value;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if constexpr (Epi::CLUSTER > 1) {
    // clusters of CLUSTER CTAs share an m-tile (n_group = cluster rank): the grid must be a multiple of it
    if (g.k_splits != 1 || g.num_n_groups != (g.pair_mode ? 2 : 1) * Epi::CLUSTER)
      return fail(PIGAN_ERR_INVALID, "cluster epilogue needs num_n_groups == cluster size (x2 in pair mode)");
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = Epi::CLUSTER;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    // clusters must be co-resident inside a GPC: on B200 only 33 clusters of 4 (132 of 148 SMs) fit at once
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cfg.gridDim = dim3(ctas - ctas % Epi::CLUSTER);
      int n = 0;
      PIGAN_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      max_clusters = n > 0 ? n : 1;
    }
    if (grid > max_clusters * Epi::CLUSTER) grid = max_clusters * Epi::CLUSTER;
    grid -= grid % Epi::CLUSTER;
    cfg.gridDim = dim3(grid);
  }
  if (pdl_enabled()) {
    // programmatic dependent launch: this kernel's CTAs may start (barrier init, tensor-memory allocation) while
    // the previous kernel on the stream drains; the kernel executes griddepcontrol.wait before it touches memory
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  note_launch();
  PIGAN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, tx ? *tx : tb, g, ep));
  PIGAN_CUDA_OK(cudaGetLastError());
  if (debug_sync_enabled()) {   // PIGAN_DEBUG_SYNC=1: name the kernel that faults (development aid; serialises everything)
    const cudaError_t err = cudaStreamSynchronize(st);
    if (err != cudaSuccess)
      return fail(PIGAN_ERR_CUDA, "%s while running %s (grid %d)", cudaGetErrorString(err), __PRETTY_FUNCTION__, grid);
  }
  return PIGAN_OK;
}

// TN operands: A [M,K] (box 128 x 64), B [N,K] (box BLOCK_N x 64).
template <class Cfg>
int make_tn_maps(CUtensorMap* ta, CUtensorMap* tb, const void* a, int m, int k, int lda, const void* b,
                 int n, int ldb) {
  PIGAN_TRY(make_tmap_f16_2d(ta, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, kBlockK, kBlockM));
  PIGAN_TRY(make_tmap_f16_2d(tb, b, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, kBlockK, Cfg::BLOCK_N));
  return PIGAN_OK;
}
// NT operands: A [Kd,M], B [Kd,N]; both loaded as boxes of 64 k-rows x 64 columns.
inline int make_nt_maps(CUtensorMap* ta, CUtensorMap* tb, const void* a, int kd_a, int m, int lda,
                        const void* b, int kd_b, int n, int ldb) {
  PIGAN_TRY(make_tmap_f16_2d(ta, a, (uint64_t)m, (uint64_t)kd_a, (uint64_t)lda, 64, kBlockK));
  PIGAN_TRY(make_tmap_f16_2d(tb, b, (uint64_t)n, (uint64_t)kd_b, (uint64_t)ldb, 64, kBlockK));
  return PIGAN_OK;
}

template <class Cfg>
GemmShape make_shape(int m, int n, int k, int k_splits = 1, int b_wrap_rows = 0) {
  GemmShape g;
  g.M = m;
  g.N = n;
  g.num_m_tiles = ceil_div(m, kBlockM);
  g.num_n_groups = ceil_div(n, Cfg::ACC_COLS);
  g.num_k_blocks = ceil_div(k, kBlockK);
  g.k_splits = k_splits < 1 ? 1 : (k_splits > g.num_k_blocks ? g.num_k_blocks : k_splits);
  g.b_wrap_k_blocks = b_wrap_rows / kBlockK;
  g.a_tail = 0;
  g.a_split_kb = -1;
  g.b_tail_from_kb = -1;
  g.pair_mode = 0;
  return g;
}

// ------------------------------------------------------------------ device-side epilogue helpers
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 32 consecutive fp32 values -> 32 fp16 at dst (64 B, 16-byte aligned)
__device__ __forceinline__ void store_f16x32(__half* dst, const float* v) {
  uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
    p[i] = u;
  }
}
__device__ __forceinline__ void store_f16x16(__half* dst, const float* v) {
  uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u;
    u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
    p[i] = u;
  }
}
// 32 consecutive fp16 values at src (64 B aligned to 16) -> fp32
__device__ __forceinline__ void load_f16x32(const __half* src, float* v) {
  const uint4* p = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = __ldg(p + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      v[8 * i + 2 * j] = f.x;
      v[8 * i + 2 * j + 1] = f.y;
    }
  }
}

// Butterfly "transpose-reduce": every lane holds 32 per-column values of its own row; afterwards lane l
// holds the sum over the warp's 32 rows of column l.  31 shuffles + adds instead of 32*5.
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      // keep the half of the columns this lane stays responsible for, send the other half
      const float keep = upper ? v[i + half] : v[i];
      const float send = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

}  // namespace pigan
