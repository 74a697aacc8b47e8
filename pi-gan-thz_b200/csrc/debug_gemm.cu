// Test hooks: run the tcgen05 GEMM core with trivial epilogues so tests/ can compare it with a plain
// fp32 matmul of the same fp16 operands.  Exercises every tile configuration the layers use.
#include "gemm_launch.cuh"

namespace pigan {

template <class Cfg>
struct EpiStoreF32 {
  struct Params {
    float* c;
    int ldc;
  };
  struct State {};
  __device__ static void init(State&) {}
  __device__ static void unit(const Params& p, State&, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, int q, int lane) {
    const int row = w.m_tile * kBlockM + q * 32 + lane;
#pragma unroll 1
    for (int t = 0; t < Cfg::ACC_TILES; ++t) {
      const int n0 = (w.n_group * Cfg::ACC_TILES + t) * Cfg::BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < Cfg::BLOCK_N; c += 16) {
        float v[16];
        tmem_ld16(tacc + t * Cfg::BLOCK_N + c, v);
        tmem_ld_wait();
        if (row < g.M) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + c + i < g.N) p.c[(size_t)row * p.ldc + n0 + c + i] = v[i];
        }
      }
    }
  }
  __device__ static void finish(const Params&, State&, int, int) {}
};

template <class Cfg>
struct EpiAtomicAddF32 {
  struct Params {
    float* c;
    int ldc;
  };
  struct State {};
  __device__ static void init(State&) {}
  __device__ static void unit(const Params& p, State&, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, int q, int lane) {
    const int row = w.m_tile * kBlockM + q * 32 + lane;
#pragma unroll 1
    for (int t = 0; t < Cfg::ACC_TILES; ++t) {
      const int n0 = (w.n_group * Cfg::ACC_TILES + t) * Cfg::BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < Cfg::BLOCK_N; c += 32) {
        float v[32];
        tmem_ld32(tacc + t * Cfg::BLOCK_N + c, v);
        tmem_ld_wait();
        if (row < g.M) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + c + i < g.N) atomicAdd(p.c + (size_t)row * p.ldc + n0 + c + i, v[i]);
        }
      }
    }
  }
  __device__ static void finish(const Params&, State&, int, int) {}
};

template <class Cfg>
static int run_tn(const void* a, const void* b, float* c, int m, int n, int k, cudaStream_t st) {
  CUtensorMap ta, tb;
  PIGAN_TRY(make_tn_maps<Cfg>(&ta, &tb, a, m, k, k, b, n, k));
  GemmShape g = make_shape<Cfg>(m, n, k);
  typename EpiStoreF32<Cfg>::Params ep{c, n};
  return launch_gemm<Cfg, EpiStoreF32<Cfg>>(ta, tb, g, ep, st);
}

}  // namespace pigan

using namespace pigan;

extern "C" int pigan_debug_gemm_tn(const void* a, const void* b, float* c, int32_t m, int32_t n, int32_t k,
                                   int32_t variant, void* stream) {
  PIGAN_CHECK_ARG(a && b && c && m > 0 && n > 0 && k > 0 && k % 8 == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (variant) {
    case 0: return run_tn<GemmCfg<256, 1, 4, false>>(a, b, c, m, n, k, st);
    case 1: return run_tn<GemmCfg<256, 2, 4, false>>(a, b, c, m, n, k, st);
    case 2: return run_tn<GemmCfg<144, 2, 4, false>>(a, b, c, m, n, k, st);
    case 3: return run_tn<GemmCfg<128, 1, 4, false>>(a, b, c, m, n, k, st);
    default: return fail(PIGAN_ERR_INVALID, "unknown gemm variant %d", variant);
  }
}

extern "C" int pigan_debug_gemm_nt(const void* a, const void* b, float* c, int32_t kd, int32_t m, int32_t n,
                                   int32_t k_splits, int32_t b_wrap_rows, void* stream) {
  PIGAN_CHECK_ARG(a && b && c && m > 0 && n > 0 && kd > 0 && m % 8 == 0 && n % 8 == 0);
  PIGAN_CHECK_ARG(b_wrap_rows % kBlockK == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  using Cfg = GemmCfg<256, 1, 4, true>;
  CUtensorMap ta, tb;
  const int kd_b = b_wrap_rows > 0 ? b_wrap_rows : kd;
  PIGAN_TRY(make_nt_maps(&ta, &tb, a, kd, m, m, b, kd_b, n, n));
  GemmShape g = make_shape<Cfg>(m, n, kd, k_splits, b_wrap_rows);
  EpiAtomicAddF32<Cfg>::Params ep{c, n};
  return launch_gemm<Cfg, EpiAtomicAddF32<Cfg>>(ta, tb, g, ep, st);
}
