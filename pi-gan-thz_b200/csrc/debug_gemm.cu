// Test hooks: run the tcgen05 GEMM core with a trivial fp32 epilogue (every tile configuration the
// layers use) and with the production store / weight-gradient epilogues, so tests/ can compare them with a
// plain fp32 matmul of the same fp16 operands.
#include "../../include/pigan_b200_debug.h"
#include "epilogues.cuh"
#include "gemm2_tc.cuh"

namespace pigan {

// MODE 0: fp32 direct store; 1: no TMEM access (pipeline probe); 2: TMEM loads only (probe)
template <class Cfg, int MODE>
struct EpiDebug {
  struct Params {
    float* c;
    int ldc;
  };
  static constexpr int SMEM_BYTES = 0;
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  struct State {
    float acc;
  };
  __device__ static void init(const Params&, State& st, const GemmShape&, const EpiCtx&) { st.acc = 0.f; }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w, uint32_t tacc,
                              const EpiCtx& cx) {
    if constexpr (MODE == 1) return;
    const int row = w.m_tile * kBlockM + cx.q * 32 + cx.lane;
#pragma unroll 1
    for (int t = 0; t < Cfg::ACC_TILES; ++t) {
      const int n0 = (w.n_group * Cfg::ACC_TILES + t) * Cfg::BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < Cfg::BLOCK_N; c += 16) {
        float v[16];
        tmem_ld16(tacc + t * Cfg::BLOCK_N + c, v);
        tmem_ld_wait();
        if constexpr (MODE == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) st.acc += v[i];
        } else if (row < g.M) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + c + i < g.N) p.c[(size_t)row * p.ldc + n0 + c + i] = v[i];
        }
      }
    }
  }
  __device__ static void finish(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    if constexpr (MODE != 0) p.c[blockIdx.x * 256 + cx.group * 128 + cx.tid] = st.acc;
  }
};

template <class Cfg, int MODE>
static int run_tn(const void* a, const void* b, float* c, int m, int n, int k, cudaStream_t st) {
  CUtensorMap ta, tb;
  PIGAN_TRY(make_tn_maps<Cfg>(&ta, &tb, a, m, k, k, b, n, k));
  GemmShape g = make_shape<Cfg>(m, n, k);
  typename EpiDebug<Cfg, MODE>::Params ep{c, n};
  return launch_gemm<Cfg, EpiDebug<Cfg, MODE>>(ta, tb, g, ep, st);
}

template <bool BIAS, bool LRELU, bool RS, class Cfg>
static int run_linear_cfg(const void* a, const void* a_tail, const void* b, const float* bias, void* out,
                          float* rowstats, int m, int n, int k, cudaStream_t st) {
  using Epi = EpiStore<Cfg, BIAS, LRELU, RS>;
  CUtensorMap ta, tb, tx;
  PIGAN_TRY(make_tn_maps<Cfg>(&ta, &tb, a, m, k, k, b, n, k));
  GemmShape g = make_shape<Cfg>(m, n, k);
  if (a_tail) {
    PIGAN_TRY(make_tmap_f16_2d(&tx, a_tail, 64, (uint64_t)m, 64, kBlockK, kBlockM));
    g.a_tail = 1;
  }
  typename Epi::Params ep;
  ep.out.ptr = static_cast<__half*>(out); ep.out.ld = n; ep.out.rows = m; ep.out.cols = n;
  PIGAN_TRY(make_tmap_f16_store(&ep.out.map, out, (uint64_t)n, (uint64_t)m, (uint64_t)n));
  ep.bias = bias;
  ep.scale = nullptr;
  ep.rowstats = rowstats;
  ep.n_tiles = g.num_n_groups;
  ep.mask = nullptr;
  ep.mask_words = 0;
  return launch_gemm<Cfg, Epi>(ta, tb, g, ep, st, 0, a_tail ? &tx : nullptr);
}
// K <= 256: the production path keeps the weights resident in shared memory (engine.cu: CfgSR); g_force_streamed
// selects the streamed-operand kernel for the same shape (tools/gemm_perf.py compares the two)
static bool g_force_streamed = false;
template <bool BIAS, bool LRELU, bool RS>
static int run_linear(const void* a, const void* a_tail, const void* b, const float* bias, void* out,
                      float* rowstats, int m, int n, int k, cudaStream_t st) {
  if (k <= 256 && !g_force_streamed)
    return run_linear_cfg<BIAS, LRELU, RS, GemmCfg<256, 1, 4, false, 4>>(a, a_tail, b, bias, out, rowstats, m, n, k, st);
  return run_linear_cfg<BIAS, LRELU, RS, GemmCfg<256, 1, 4, false>>(a, a_tail, b, bias, out, rowstats, m, n, k, st);
}

template <bool BIAS, bool LRELU, bool RS>
static int run_linear2(const void* a, const void* a_tail, const void* b, const float* bias, void* out,
                       float* rowstats, int m, int n, int k, cudaStream_t st) {
  using Cfg = Gemm2Cfg<5>;
  using Epi = EpiStore<Cfg, BIAS, LRELU, RS>;
  CUtensorMap ta, tb, tx;
  PIGAN_TRY(make_tmap_f16_2d(&ta, a, (uint64_t)k, (uint64_t)m, (uint64_t)k, kBlockK, kBlockM));
  PIGAN_TRY(make_tmap_f16_2d(&tb, b, (uint64_t)k, (uint64_t)n, (uint64_t)k, kBlockK, Cfg::BLOCK_N / 2));
  GemmShape g = make_shape<Cfg>(m, n, k);
  if (a_tail) {
    PIGAN_TRY(make_tmap_f16_2d(&tx, a_tail, 64, (uint64_t)m, 64, kBlockK, kBlockM));
    g.a_tail = 1;
  }
  typename Epi::Params ep;
  ep.out.ptr = static_cast<__half*>(out); ep.out.ld = n; ep.out.rows = m; ep.out.cols = n;
  PIGAN_TRY(make_tmap_f16_store(&ep.out.map, out, (uint64_t)n, (uint64_t)m, (uint64_t)n));
  ep.bias = bias;
  ep.scale = nullptr;
  ep.rowstats = rowstats;
  ep.n_tiles = g.num_n_groups;
  ep.mask = nullptr;
  ep.mask_words = 0;
  return launch_gemm2<Cfg, Epi>(ta, tb, g, ep, st, a_tail ? &tx : nullptr);
}

}  // namespace pigan

using namespace pigan;

extern "C" int pigan_debug_gemm_tn(const void* a, const void* b, float* c, int32_t m, int32_t n, int32_t k,
                                   int32_t variant, void* stream) {
  PIGAN_CHECK_ARG(a && b && c && m > 0 && n > 0 && k > 0 && k % 8 == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (variant) {
    case 0: return run_tn<GemmCfg<256, 1, 4, false>, 0>(a, b, c, m, n, k, st);
    case 1: return run_tn<GemmCfg<256, 2, 4, false>, 0>(a, b, c, m, n, k, st);
    case 2: return run_tn<GemmCfg<144, 2, 4, false>, 0>(a, b, c, m, n, k, st);
    case 3: return run_tn<GemmCfg<128, 1, 4, false>, 0>(a, b, c, m, n, k, st);
    case 10: return run_tn<GemmCfg<256, 1, 4, false>, 1>(a, b, c, m, n, k, st);
    case 11: return run_tn<GemmCfg<256, 1, 4, false>, 2>(a, b, c, m, n, k, st);
    case 12: return run_tn<GemmCfg<256, 1, 3, false>, 1>(a, b, c, m, n, k, st);
    // resident-B probes (K <= 256): no-op epilogue with 3 / 4 / 6 A stages, TMEM loads only with 6
    case 20: return run_tn<GemmCfg<256, 1, 3, false, 4>, 1>(a, b, c, m, n, k, st);
    case 21: return run_tn<GemmCfg<256, 1, 4, false, 4>, 1>(a, b, c, m, n, k, st);
    case 22: return run_tn<GemmCfg<256, 1, 6, false, 4>, 1>(a, b, c, m, n, k, st);
    case 23: return run_tn<GemmCfg<256, 1, 6, false, 4>, 2>(a, b, c, m, n, k, st);
    case 24: return run_tn<GemmCfg<256, 1, 2, false>, 1>(a, b, c, m, n, k, st);
    case 25: return run_tn<GemmCfg<128, 1, 8, false, 4>, 1>(a, b, c, m, n, k, st);
    default: return fail(PIGAN_ERR_INVALID, "unknown gemm variant %d", variant);
  }
}

extern "C" int pigan_debug_force_streamed(int32_t on) {
  g_force_streamed = on != 0;
  return PIGAN_OK;
}

extern "C" int pigan_debug_linear(const void* a, const void* a_tail, const void* b, const float* bias,
                                  void* out_f16, float* rowstats, int32_t m, int32_t n, int32_t k,
                                  int32_t leaky, void* stream) {
  PIGAN_CHECK_ARG(a && b && out_f16 && m > 0 && n > 0 && k > 0 && k % 8 == 0 && n % 8 == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool rs = rowstats != nullptr;
  if (bias && leaky && rs) return run_linear<true, true, true>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (bias && leaky) return run_linear<true, true, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (bias && rs) return run_linear<true, false, true>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (bias) return run_linear<true, false, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (leaky) return run_linear<false, true, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  return run_linear<false, false, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
}

extern "C" int pigan_debug_linear2(const void* a, const void* a_tail, const void* b, const float* bias,
                                  void* out_f16, float* rowstats, int32_t m, int32_t n, int32_t k,
                                  int32_t leaky, void* stream) {
  PIGAN_CHECK_ARG(a && b && out_f16 && m > 0 && n > 0 && k > 0 && k % 8 == 0 && n % 8 == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool rs = rowstats != nullptr;
  if (bias && leaky && rs) return run_linear2<true, true, true>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (bias && leaky) return run_linear2<true, true, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (bias && rs) return run_linear2<true, false, true>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (bias) return run_linear2<true, false, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  if (leaky) return run_linear2<false, true, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
  return run_linear2<false, false, false>(a, a_tail, b, bias, out_f16, rowstats, m, n, k, st);
}

extern "C" int pigan_debug_gemm_nt(const void* a, const void* b, const void* b_tail, float* c, int32_t kd,
                                   int32_t m, int32_t n, int32_t k_splits, int32_t b_wrap_rows,
                                   int32_t tail_from_row, int32_t n_valid, int32_t bias_col, float* db,
                                   void* stream) {
  PIGAN_CHECK_ARG(a && b && c && m > 0 && n > 0 && kd > 0 && m % 8 == 0 && n % 8 == 0);
  PIGAN_CHECK_ARG(b_wrap_rows % kBlockK == 0 && tail_from_row % kBlockK == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  using Cfg = GemmCfg<256, 1, 4, true>;
  CUtensorMap ta, tb, tx;
  const int kd_b = b_wrap_rows > 0 ? b_wrap_rows : kd;
  PIGAN_TRY(make_nt_maps(&ta, &tb, a, kd, m, m, b, kd_b, n, n));
  GemmShape g = make_shape<Cfg>(m, n, kd, k_splits, b_wrap_rows);
  if (b_tail) {
    PIGAN_TRY(make_tmap_f16_2d(&tx, b_tail, 64, (uint64_t)(kd - tail_from_row), 64, 64, kBlockK));
    g.b_tail_from_kb = tail_from_row / kBlockK;
  }
  EpiWeightGrad<Cfg>::Params ep{c, n_valid > 0 ? n_valid : n, n_valid > 0 ? n_valid : n, 1.0f, bias_col, db};
  if (n_valid > 0) ep.ld = n_valid;
  return launch_gemm<Cfg, EpiWeightGrad<Cfg>>(ta, tb, g, ep, st, 0, b_tail ? &tx : nullptr);
}
