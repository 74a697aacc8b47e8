// Host orchestration of the hot path: workspace carving, weight packing, and the kernel sequences behind
// the C ABI (Generator/Discriminator/ForwardModel forward, the PI-GAN train step of
// core/train/train_pigan.py:114-187 in seven phases, candidate scoring of
// core/evaluate/unified_evaluator.py:376-392).  No host<->device synchronisation anywhere: every call only
// enqueues work on the caller's stream, so a step can be captured in a CUDA graph.
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "elementwise.cuh"
#include "epilogues.cuh"
#include "layout.h"

namespace pigan {

namespace {

constexpr int kDwPartSlabs = 160;  // >= SM count: one slab per (output tile, k-split) unit
constexpr int kKp = 256;  // spectrum operand width (S + P + 2 spare columns <= 256)
using CfgS = GemmCfg<256, 1, 4, false>;      // store epilogues, both operands streamed (K > 256)
using CfgSR = GemmCfg<256, 1, 4, false, 4>;  // store epilogues, K <= 256: the n-group's weights stay in shared memory
using CfgP = GemmCfg<256, 1, 4, false>;      // no staging
using CfgO = GemmCfg<144, 2, 2, false>;      // forward-model output layer, generic epilogue (fp32 dump)
using CfgO4 = GemmCfg<144, 2, 4, false>;     // forward-model output layer, loss / scoring epilogues
using CfgW = GemmCfg<256, 1, 4, true>;       // weight gradients
using CfgL1 = GemmCfg<256, 1, 3, false>;     // Linear+LayerNorm, 256 columns per CTA
using CfgH = GemmCfg<256, 1, 3, false>;      // scoring: generator layer 2 + head + surrogate layer 1 (24 KB epilogue scratch)

// loss_sums indices (fp64)
enum { kSumD = 0, kSumAdv = 1, kSumRec = 2, kSumMet = 3, kSumMaxwell = 4, kSumLc1 = 5, kSumLc2 = 6, kSumRange = 7,
       kSumGradD = 8, kSumGradG = 9, kNumSums = 16 };

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// Widened dims on which the PI-GAN step runs (BASELINE config 5: hidden 2048, 2048-point spectra): the surrogate's
// conditions (surrogate_dims_ok) plus a spectrum length that is a multiple of 64 - the parameter / bias columns then
// fill the operand's last 64-column k-block, which is what the discriminator's fake rows swap - and generator /
// discriminator hidden widths of 256 / 512 / 1024 / 2048.
inline bool wide_gan_dims_ok(const PiganDims& d) {
  if (d.param_dim != 4 || d.metrics_dim < 2 || d.metrics_dim % 2 || d.spectrum_dim < 64 || d.spectrum_dim % 64) return false;
  if ((d.spectrum_dim + d.metrics_dim + 63) / 64 * 64 > 2560 || d.spectrum_dim > 2048) return false;
  for (int i = 0; i < 5; ++i) {
    const int h = d.f_hidden[i];
    if (h != 256 && h != 512 && h != 1024 && h != 2048) return false;
  }
  // (the streaming kernels give a thread 8 or 4 consecutive columns and a block whole rows: widths whose column
  // chunks divide the 256 threads of a block)
  for (int i = 0; i < 2; ++i) {
    const int g = d.g_hidden[i], h = d.d_hidden[i];
    if (g != 256 && g != 512 && g != 1024 && g != 2048) return false;
    if (h != 256 && h != 512 && h != 1024 && h != 2048) return false;
  }
  return true;
}

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

}  // namespace

}  // namespace pigan

using namespace pigan;

// Optional CUDA-event instrumentation of the kernel sequence (pigan_engine_profile_begin/_end): one event per
// section boundary on the caller's stream; off by default, so the hot path records nothing.
struct Prof {
  bool on = false;
  bool all = true;
  std::vector<std::string> wanted;
  std::vector<const char*> names;
  struct Rec { int slot; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  int open_slot = -1;
  cudaEvent_t open_ev = nullptr;
  cudaEvent_t next_event() {
    if (used == pool.size()) {
      cudaEvent_t ev;
      cudaEventCreate(&ev);
      pool.push_back(ev);
    }
    return pool[used++];
  }
  int slot_of(const char* name) {
    for (size_t i = 0; i < names.size(); ++i)
      if (names[i] == name || strcmp(names[i], name) == 0) return (int)i;
    names.push_back(name);
    return (int)names.size() - 1;
  }
  bool is_wanted(const char* name) const {
    if (all) return true;
    for (const auto& w : wanted)
      if (w == name) return true;
    return false;
  }
};

struct PiganEngine {
  Prof prof;
  PiganDims d;
  GenLayout gl;
  DiscLayout dl;
  FwdLayout fl;
  int64_t max_batch, bp;  // bp = max_batch rounded up to 128: row offset of the fake half in stacked tensors
  size_t ws_bytes;
  // full = the reference widths: every entry point, fused epilogues.  Otherwise (widened dims, BASELINE config 5 -
  // see check_dims) the engine serves the surrogate's paths (forward, VJP, training step) and, when `gan`, the PI-GAN
  // train step, all through generic-width pieces
  bool full = true;
  // gan = the PI-GAN train step runs on this engine: the reference widths (fused epilogues), or widened dims the
  // generic-width pieces cover (wide_gan_dims_ok) - then pigan_train_step[_phase] work, the module / scoring entry
  // points still need `full`
  bool gan = true;
  int kp = 256;           // spectrum operand width: S + P + 2 spare columns rounded up to 64 (256 at the reference dims)
  int dout_ld = 320;      // output-layer gradient operand: S + Mt columns padded to a multiple of 64
  int out_groups = 2;     // 256-column groups of the output layer (widened path: fp32 accumulator slabs)
  float* f_slab = nullptr;   // widened path: [bp / 128][out_groups][128][256] fp32 output-layer accumulators
  float* f_l1c = nullptr;    // widened path: closed-form LayerNorm constants of the first layer (launch_f_l1_wide)
  bool f_loaded = false;
  const float* f_params = nullptr;
  const float* center = nullptr;  // caller-provided spectrum centring row (pigan_engine_set_spectrum_center)
  // The frozen surrogate's forward pass of the G-step depends only on the generator output, not on the D-step: it
  // runs on a second, high-priority stream from the moment that output exists (phase 2) and is joined before the
  // generator-head backward (phase 3).  Its CTAs fill the SMs that the D-step's kernels leave idle in their last,
  // partial wave (512 row tiles on 148 SMs = 3.46 waves).  PIGAN_OVERLAP=0 keeps everything on the caller's stream.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool side_pending = false;

  // fp16 activations
  __half *xc_own;   // workspace copy of the spectrum operand; `xc` below points at it or at a caller-prepared operand
  __half *xc, *tail_f, *g_h1, *g_a1, *g_h2, *d_z1, *d_z2, *d_dh2, *d_dh1, *f_a1, *f_a2, *f_a3, *f_a4, *f_a5,
      *g_dy2, *g_da1;
  // fp16 weights
  __half *g_w1h, *g_w2h, *g_w2th, *d_w1h, *d_w2h, *d_w2th, *f_wh[6];
  // fp32 scratch
  uint32_t *d_mask1;                   // [2B][H1/32] sign bits of the discriminator's first activation
  float *dw_part;                      // [kDwPartSlabs][128][256] split-K weight-gradient slabs
  float *partials;                     // [kPartBlocks x kPartCols] two-stage batch reductions
  int bn_pending_blocks = 0;           // > 0: partial rows of BatchNorm sums wait in `partials` for the fused reduce+finalize
  bool whole_step = false;             // pigan_train_step (one GPU): nothing is exchanged between the phases
  float *dpre;                         // [B,4] generator head backward
  float *p, *pden, *dpden, *dp_lc, *dlogit, *prob, *row_err, *f_rowstats, *cvec, *g_beff, *d_beff, *d_wp;
  float *bn_sums, *bn_bwd_sums;        // [2*H1 + 2*H2], [2*H2 + 2*H1]
  float *mean1, *rstd1, *scale1, *bias1, *mean2, *rstd2, *scale2, *bias2;
  float *head_img;                     // [kHeadImgFloats] constants of EpiHeadF1
  float *zero_blk;                     // sums | bn_sums | bn_bwd_sums, cleared with one memset per step
  size_t zero_bytes;
  float *f_bias_out;                   // [288] zero padded
  double* sums;

  PiganEngine(const PiganDims& dims) : d(dims), gl(dims), dl(dims), fl(dims) {
    full = dims_are_default(dims);
    gan = full || wide_gan_dims_ok(dims);
    kp = full ? 256 : (int)((dims.spectrum_dim + dims.param_dim + 2 + 63) / 64 * 64);
    dout_ld = (int)((fl.OUT + 63) / 64 * 64);
    out_groups = (fl.OUT + 255) / 256;
  }

  size_t carve(void* ws) {
    Carver c(ws);
    const int64_t B = bp;
    const int64_t Bg = gan ? bp : 0;    // generator / discriminator activations: not on a surrogate-only engine
    const int H1 = gl.H1, H2 = gl.H2, D1 = dl.H1, D2 = dl.H2;
    // widened step: rows [bp, 2 bp) hold the fake rows' operand (the weight-gradient GEMM reads real and fake rows as
    // one matrix; at the reference widths a 64-column tail operand inside the GEMM does that)
    xc_own = c.take<__half>((full ? B : 2 * Bg) * kp);
    xc = xc_own;
    tail_f = c.take<__half>(B * 64);
    g_h1 = c.take<__half>(Bg * H1);
    g_a1 = c.take<__half>(Bg * H1);
    g_h2 = c.take<__half>(Bg * H2);
    d_z1 = c.take<__half>(2 * Bg * D1);
    d_z2 = c.take<__half>(2 * Bg * D2);
    d_dh2 = c.take<__half>(2 * Bg * D2);
    d_dh1 = c.take<__half>(2 * Bg * D1);
    f_a1 = c.take<__half>(B * fl.H[0]);
    f_a2 = c.take<__half>(B * fl.H[1]);
    f_a3 = c.take<__half>(B * fl.H[2]);
    f_a4 = c.take<__half>(B * fl.H[3]);
    f_a5 = c.take<__half>(B * fl.H[4]);
    g_dy2 = c.take<__half>(Bg * H2);
    g_da1 = c.take<__half>(Bg * H1);
    g_w1h = c.take<__half>((size_t)H1 * kp);
    g_w2h = c.take<__half>((size_t)H2 * H1);
    g_w2th = c.take<__half>((size_t)H1 * H2);
    d_w1h = c.take<__half>((size_t)D1 * kp);
    d_w2h = c.take<__half>((size_t)D2 * D1);
    d_w2th = c.take<__half>((size_t)D1 * D2);
    f_wh[0] = nullptr;
    int in = fl.H[0];
    for (int i = 1; i < 6; ++i) {
      const int out = i < 5 ? fl.H[i] : fl.OUT;
      f_wh[i] = c.take<__half>((size_t)out * in);
      in = out;
    }
    d_mask1 = c.take<uint32_t>((size_t)2 * Bg * (D1 / 32));
    dw_part = c.take<float>((size_t)kDwPartSlabs * kBlockM * 256);
    partials = c.take<float>((size_t)kPartBlocks * kPartCols);
    dpre = c.take<float>(B * 4);
    p = c.take<float>(B * 4);
    pden = c.take<float>(B * 4);
    dpden = c.take<float>(B * 4);
    dp_lc = c.take<float>(B * 4);
    dlogit = c.take<float>(2 * B);
    prob = c.take<float>(2 * B);
    row_err = c.take<float>(B);
    f_rowstats = c.take<float>(B * 8 * 2);
    cvec = c.take<float>(kp);
    g_beff = c.take<float>(H1);
    d_beff = c.take<float>(D1);
    d_wp = c.take<float>((size_t)D1 * 4);
    // one zero-filled block per step: loss sums (16 doubles) | BatchNorm forward sums | backward sums
    zero_blk = c.take<float>(32 + 2 * (2 * H1 + 2 * H2));
    zero_bytes = (32 + 2 * (2 * H1 + 2 * H2)) * sizeof(float);
    sums = reinterpret_cast<double*>(zero_blk);
    bn_sums = zero_blk ? zero_blk + 32 : nullptr;
    bn_bwd_sums = zero_blk ? zero_blk + 32 + 2 * H1 + 2 * H2 : nullptr;
    mean1 = c.take<float>(H1); rstd1 = c.take<float>(H1); scale1 = c.take<float>(H1); bias1 = c.take<float>(H1);
    mean2 = c.take<float>(H2); rstd2 = c.take<float>(H2); scale2 = c.take<float>(H2); bias2 = c.take<float>(H2);
    f_bias_out = c.take<float>(dout_ld > 288 ? dout_ld : 288);
    head_img = c.take<float>(kHeadImgFloats);
    f_slab = full ? nullptr : c.take<float>((size_t)(B / kBlockM) * out_groups * kBlockM * 256);
    f_l1c = c.take<float>(32);
    return (c.off + 255) & ~size_t(255);
  }
};

namespace pigan {
namespace {

// Closes the open profiling section and opens `name` (nullptr: just close).
void prof_mark(PiganEngine* e, const char* name, cudaStream_t st) {
  Prof& p = e->prof;
  if (!p.on) return;
  int slot = -1;
  if (name && p.is_wanted(name)) slot = p.slot_of(name);
  if (p.open_slot < 0 && slot < 0) return;
  cudaEvent_t ev = p.next_event();
  cudaEventRecord(ev, st);
  if (p.open_slot >= 0) p.recs.push_back({p.open_slot, p.open_ev, ev});
  p.open_slot = slot;
  p.open_ev = ev;
}
#define PM(name) prof_mark(e, name, st)

// Widened dims (BASELINE config 5: hidden 2048, 2048-point spectra; the widths of enhanced_forward_model.py:42-49):
// the surrogate's forward, VJP and training step run at any hidden width of 256/512/1024/2048 and any even S / Mt
// with round_up(S + Mt, 64) <= 2560, built from the plain store GEMMs + the streaming LayerNorm kernels (no fused
// LayerNorm / loss epilogues: at these widths the GEMMs are 126 MFLOP per sample and dominate).  The PI-GAN step
// runs on the subset wide_gan_dims_ok accepts; the stand-alone generator / discriminator module entry points, scoring
// and search stay at the reference widths (need_full).
bool surrogate_dims_ok(const PiganDims& d) {
  if (d.param_dim != 4 || d.spectrum_dim < 2 || d.metrics_dim < 2 || d.spectrum_dim % 2 || d.metrics_dim % 2) return false;
  if ((d.spectrum_dim + d.metrics_dim + 63) / 64 * 64 > 2560) return false;
  for (int i = 0; i < 5; ++i) {
    const int h = d.f_hidden[i];
    if (h != 256 && h != 512 && h != 1024 && h != 2048) return false;
  }
  return true;
}
int check_dims(const PiganDims& d) {
  if (!dims_are_default(d) && !surrogate_dims_ok(d))
    return fail(PIGAN_ERR_UNSUPPORTED,
                "dimensions: the full path implements the reference widths (S=250, P=4, Mt=8, G 512/256, D 512/256, "
                "F 256/512/1024/512/256); the widened paths take P=4, even S / Mt with S + Mt <= 2560 and hidden widths "
                "of 256/512/1024/2048 (include/pigan_b200.h)");
  return PIGAN_OK;
}
int need_full(const PiganEngine* e) {
  if (e && !e->full)
    return fail(PIGAN_ERR_UNSUPPORTED,
                "this engine was created with widened dimensions: the train step (pigan_train_step[_phase]) and the "
                "surrogate's entry points (pigan_forward_model_forward / _vjp / _input_grad, pigan_fwd_train_step) run "
                "at those; this entry point exists at the reference widths only");
  return PIGAN_OK;
}

// a_tail: [m, 64] replaces the last k-block of A.  a2 (k2 columns, pitch k2): A = [a | a2] along K, b has k + k2 columns.
template <class Cfg, class Epi>
int run_tn(typename Epi::Params& ep, const __half* a, int64_t m, int k, int lda, const __half* b, int n, int ldb,
           cudaStream_t st, const __half* a_tail = nullptr, const __half* a2 = nullptr, int k2 = 0) {
  CUtensorMap ta, tb, tx;
  PIGAN_TRY(make_tn_maps<Cfg>(&ta, &tb, a, (int)m, k, lda, b, n, ldb));
  GemmShape g = make_shape<Cfg>((int)m, n, k + k2);
  if (a_tail) {
    PIGAN_TRY(make_tmap_f16_2d(&tx, a_tail, 64, (uint64_t)m, 64, kBlockK, kBlockM));
    g.a_tail = 1;
  } else if (a2) {
    if (k % kBlockK != 0) return fail(PIGAN_ERR_INVALID, "K-concatenated operands: first part must be a multiple of 64");
    PIGAN_TRY(make_tmap_f16_2d(&tb, b, (uint64_t)(k + k2), (uint64_t)n, (uint64_t)ldb, kBlockK, Cfg::BLOCK_N));
    PIGAN_TRY(make_tmap_f16_2d(&tx, a2, (uint64_t)k2, (uint64_t)m, (uint64_t)k2, kBlockK, kBlockM));
    g.a_split_kb = k / kBlockK;
  }
  return launch_gemm<Cfg, Epi>(ta, tb, g, ep, st, 0, (a_tail || a2) ? &tx : nullptr);
}

int out_map(OutTile* m, __half* ptr, int64_t rows, int cols, int ld) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || ld % 8 != 0 || cols % 8 != 0)
    return fail(PIGAN_ERR_INVALID, "epilogue output must be 16-byte aligned with columns / pitch multiples of 8");
  m->ptr = ptr;
  m->ld = ld;
  m->rows = (int)rows;
  m->cols = cols;
  return make_tmap_f16_store(&m->map, ptr, (uint64_t)cols, (uint64_t)rows, (uint64_t)ld);
}

// PIGAN_STORE_GROUPS=2: two epilogue groups for the plain store epilogues (default four, see EpiStore)
bool store_groups4() {
  static const bool on = [] { const char* v = getenv("PIGAN_STORE_GROUPS"); return !(v && v[0] == '2'); }();
  return on;
}
template <class Cfg, class Epi>
int linear_store_run(const __half* a, int64_t rows, int k, const __half* w, int n, const float* bias, __half* out,
                     float* rowstats, cudaStream_t st, const __half* a_tail, uint32_t* mask) {
  typename Epi::Params ep;
  PIGAN_TRY(out_map(&ep.out, out, rows, n, n));
  ep.bias = bias;
  ep.scale = nullptr;
  ep.rowstats = rowstats;
  ep.n_tiles = ceil_div(n, 256);
  ep.mask = mask;
  ep.mask_words = n / 32;
  ep.colpart = nullptr;
  ep.col_groups = 0;
  return run_tn<Cfg, Epi>(ep, a, rows, k, k, w, n, k, st, a_tail);
}
// out[rows, n] = act(a . w^T + bias) (+ LayerNorm row partials)
template <bool BIAS, bool LRELU, bool RS, bool MASKOUT = false>
int linear_store(const __half* a, int64_t rows, int k, const __half* w, int n, const float* bias, __half* out,
                 float* rowstats, cudaStream_t st, const __half* a_tail = nullptr, uint32_t* mask = nullptr) {
  if constexpr (!RS) {
    if (store_groups4()) {
      if (k <= CfgSR::B_RES_KB * kBlockK)
        return linear_store_run<CfgSR, EpiStore<CfgSR, BIAS, LRELU, false, MASKOUT, false, false, 4>>(
            a, rows, k, w, n, bias, out, rowstats, st, a_tail, mask);
      return linear_store_run<CfgS, EpiStore<CfgS, BIAS, LRELU, false, MASKOUT, false, false, 4>>(
          a, rows, k, w, n, bias, out, rowstats, st, a_tail, mask);
    }
  }
  if (k <= CfgSR::B_RES_KB * kBlockK)
    return linear_store_run<CfgSR, EpiStore<CfgSR, BIAS, LRELU, RS, MASKOUT>>(a, rows, k, w, n, bias, out, rowstats, st,
                                                                             a_tail, mask);
  return linear_store_run<CfgS, EpiStore<CfgS, BIAS, LRELU, RS, MASKOUT>>(a, rows, k, w, n, bias, out, rowstats, st, a_tail,
                                                                         mask);
}

// out[rows, n] = fp16(a . w^T) and, from the same epilogue, the column sums / sums of squares of the stored values
// added into sum[n] / sumsq[n] (train-mode BatchNorm statistics of the generator's hidden layers); n = 256 or 512
int linear_store_colstats(const __half* a, int64_t rows, int k, const __half* w, int n, __half* out, float* sum,
                          float* sumsq, float* partials, cudaStream_t st, int* defer_blocks = nullptr) {
  using Epi = EpiStore<CfgL1, false, false, false, false, false, true>;
  if (n % 256 != 0) return fail(PIGAN_ERR_INVALID, "column statistics in the GEMM epilogue: n must be a multiple of 256");
  typename Epi::Params ep;
  PIGAN_TRY(out_map(&ep.out, out, rows, n, n));
  ep.bias = nullptr;
  ep.scale = nullptr;
  ep.rowstats = nullptr;
  ep.n_tiles = n / 256;
  ep.mask = nullptr;
  ep.mask_words = 0;
  ep.colpart = partials;
  ep.col_groups = n / 256;
  PIGAN_TRY((run_tn<CfgL1, Epi>(ep, a, rows, k, k, w, n, k, st)));
  // launch_gemm's grid: min(units, SMs) rounded down to a multiple of the n-groups; one partial row per CTA group
  const int groups = n / 256;
  const int units = ceil_div((int)rows, kBlockM) * groups;
  int grid = units < sm_count() ? units : sm_count();
  if (grid > groups) grid -= grid % groups;
  if (defer_blocks) {   // the caller reduces and finalises in one launch (launch_bn_reduce_finalize)
    *defer_blocks = grid / groups;
    return PIGAN_OK;
  }
  ReduceArgs r{partials, grid / groups, 2 * n, 2, {{sum, n, 1.f}, {sumsq, n, 1.f}}};
  launch_reduce_columns(r, st);
  return PIGAN_OK;
}

// out[rows, n] = fp16(LeakyReLU(LayerNorm(a . w^T + bias)))   n = 256, 512 (one CTA per row tile) or 1024 (cluster of 2)
long long* g_ln_trace = nullptr;  // pigan_engine_trace_layernorm
// bias / gamma / beta: device pointers (the layer's fp32 parameters); the epilogue stages its columns in shared memory
template <int CLUSTER, bool PAIR = false, class Cfg = CfgL1>
int linear_ln_c(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& g, int64_t rows, int k, int n,
                const float* bias, const float* gamma, const float* beta, __half* out, cudaStream_t st) {
  static const bool lsu_store = [] { const char* v = getenv("PIGAN_LN_TMA"); return v && v[0] == '0'; }();
  if (lsu_store) {   // comparison variant: read-back + global stores instead of TMA stores
    using Epi = EpiLnStore<Cfg, CLUSTER, PAIR, 4, false>;
    typename Epi::Params ep;
    PIGAN_TRY(out_map(&ep.out, out, rows, n, n));
    ep.bias = bias; ep.gamma = gamma; ep.beta = beta; ep.n_total = n;
    ep.trace = g_ln_trace ? g_ln_trace + (n == 1024 ? 1 : n == 256 ? 3 : (k == 256 ? 0 : 2)) * 320 : nullptr;
    return launch_gemm<Cfg, Epi>(ta, tb, g, ep, st);
  }
  using Epi = EpiLnStore<Cfg, CLUSTER, PAIR>;
  typename Epi::Params ep;
  PIGAN_TRY(out_map(&ep.out, out, rows, n, n));
  ep.bias = bias;
  ep.gamma = gamma;
  ep.beta = beta;
  ep.n_total = n;
  ep.trace = g_ln_trace ? g_ln_trace + (n == 1024 ? 1 : n == 256 ? 3 : (k == 256 ? 0 : 2)) * 320 : nullptr;
  return launch_gemm<Cfg, Epi>(ta, tb, g, ep, st);
}
int linear_ln(const __half* a, int64_t rows, int k, const __half* w, int n, const float* bias, const float* gamma,
              const float* beta, __half* out, cudaStream_t st) {
  CUtensorMap ta, tb;
  PIGAN_TRY(make_tn_maps<CfgL1>(&ta, &tb, a, (int)rows, k, k, w, n, k));
  const GemmShape g = make_shape<CfgL1>((int)rows, n, k);
  // (resident weights were measured for the K = 256 layer and make no difference here: 16 epilogue warps need the
  // shared memory more than the operand ring does)
  if (n == 256) return linear_ln_c<1>(ta, tb, g, rows, k, n, bias, gamma, beta, out, st);
  if (n == 512) return linear_ln_c<2>(ta, tb, g, rows, k, n, bias, gamma, beta, out, st);
  if (n == 1024) {
    // clusters of 4 (132 of 148 SMs) unless PIGAN_LN_PAIR=1: clusters of 2 whose CTAs walk two n-groups per row tile
    // (all SMs, but the row's statistics then wait for the second unit's MMAs and nothing overlaps)
    static const bool pair = [] { const char* v = getenv("PIGAN_LN_PAIR"); return v && v[0] == '1'; }();
    if (!pair) return linear_ln_c<4>(ta, tb, g, rows, k, n, bias, gamma, beta, out, st);
    GemmShape gp = g;
    gp.pair_mode = 1;
    return linear_ln_c<2, true>(ta, tb, gp, rows, k, n, bias, gamma, beta, out, st);
  }
  return fail(PIGAN_ERR_UNSUPPORTED, "LayerNorm width %d", n);
}

// dw[m_out, ld] += (1/gs) * a[kd, m_out]^T b[.., n]
int weight_grad(const __half* a, int64_t kd, int m_out, const __half* b, int64_t b_rows, int n, float* dw, int ld,
                int n_valid, float inv_gs, int bias_col, float* db, int64_t wrap_rows, const __half* b_tail,
                int64_t tail_from_row, float* part, cudaStream_t st, int lda = 0, const float* fix_cvec = nullptr,
                int fix_S = 0, int fix_P = 0) {
  CUtensorMap ta, tb, tx;
  PIGAN_TRY(make_nt_maps(&ta, &tb, a, (int)kd, m_out, lda > 0 ? lda : m_out, b, (int)b_rows, n, n));
  const int tiles_m = ceil_div(m_out, kBlockM), tiles_n = ceil_div(n, 256);
  const int tiles = tiles_m * tiles_n;
  int splits = sm_count() / (tiles > 0 ? tiles : 1);
  if (splits < 1) splits = 1;
  GemmShape g = make_shape<CfgW>(m_out, n, (int)kd, splits, (int)wrap_rows);
  if (tiles * g.k_splits > kDwPartSlabs) return fail(PIGAN_ERR_UNSUPPORTED, "weight-gradient scratch too small");
  if (b_tail) {
    PIGAN_TRY(make_tmap_f16_2d(&tx, b_tail, 64, (uint64_t)(kd - tail_from_row), 64, 64, kBlockK));
    g.b_tail_from_kb = (int)(tail_from_row / kBlockK);
  }
  using Epi = EpiWeightGradPartial<CfgW>;
  Epi::Params ep;
  PIGAN_TRY(make_tmap_f32_2d(&ep.part, part, 256, (uint64_t)tiles * g.k_splits * kBlockM, 256, kBlockM));
  PIGAN_TRY((launch_gemm<CfgW, Epi>(ta, tb, g, ep, st, 0, b_tail ? &tx : nullptr)));
  launch_dw_reduce(part, tiles_m, tiles_n, g.k_splits, dw, ld, m_out, n_valid, inv_gs, bias_col, db, st, fix_cvec, fix_S,
                   fix_P);
  return PIGAN_OK;
}

// ------------------------------------------------------------------------------------------ generator
// spectrum prep: centring vector, fp16 operand with [params | 1 1] in the spare columns
int prep_spectrum(PiganEngine* e, const float* x, const float* params, int64_t n, cudaStream_t st) {
  PM("prep_cast");
  e->xc = e->xc_own;
  if (e->center != nullptr) launch_copy_pad_f32(e->center, e->gl.S, e->cvec, e->kp, st);
  else launch_center_vec(x, n, e->gl.S, (int)(n < 512 ? n : 512), e->cvec, e->kp, st);
  launch_cast_center(x, e->cvec, params, e->xc, n, e->gl.S, e->gl.P, e->kp, st);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

int pack_generator(PiganEngine* e, const float* gp, bool need_backward, cudaStream_t st) {
  const GenLayout& L = e->gl;
  PM("pack_weights");
  launch_pack_net(gp + L.w1, L.S, L.S, L.P, 0, 0, gp + L.b1, e->cvec, e->g_w1h, e->kp, e->g_beff, L.H1, gp + L.w2, L.H2,
                  e->g_w2h, need_backward ? e->g_w2th : nullptr, nullptr, st);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

int pack_discriminator(PiganEngine* e, const float* dp, bool need_backward, cudaStream_t st) {
  const DiscLayout& L = e->dl;
  PM("pack_weights");
  launch_pack_net(dp + L.w1, L.IN, L.S, L.P, 1, 1, dp + L.b1, e->cvec, e->d_w1h, e->kp, e->d_beff, L.H1, dp + L.w2, L.H2,
                  e->d_w2h, need_backward ? e->d_w2th : nullptr, need_backward ? e->d_wp : nullptr, st);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

// G layer 1 -> h1 (pre-BN product WITHOUT its constant bias: see launch_pack_first_layer)
// PIGAN_GEMM_COLSTATS=0: statistics from a separate pass over the stored tensor (colstats_kernel), as in round 1
static bool gemm_colstats(const PiganEngine* e) {
  static const bool on = [] { const char* v = getenv("PIGAN_GEMM_COLSTATS"); return !(v && v[0] == '0'); }();
  return on && e->gl.H1 % 256 == 0 && e->gl.H2 % 256 == 0;
}
// train = also produce the layer's BatchNorm batch sums (bn_sums); false: g_bn_stats is not called (eval mode)
int g_layer1(PiganEngine* e, int64_t n, cudaStream_t st, bool train = false) {
  PM("g_l1_gemm");
  if (train && gemm_colstats(e))
    return linear_store_colstats(e->xc, n, e->kp, e->g_w1h, e->gl.H1, e->g_h1, e->bn_sums, e->bn_sums + e->gl.H1,
                                 e->partials, st, e->whole_step ? &e->bn_pending_blocks : nullptr);
  return linear_store<false, false, false>(e->xc, n, e->kp, e->g_w1h, e->gl.H1, nullptr, e->g_h1, nullptr, st);
}
int g_layer2(PiganEngine* e, const float* gp, int64_t n, cudaStream_t st, bool train = false) {
  PM("g_bn_relu_apply");
  launch_bn_relu_apply(e->g_h1, e->scale1, e->bias1, e->g_a1, n, e->gl.H1, st);
  (void)gp;
  PM("g_l2_gemm");
  if (train && gemm_colstats(e))
    return linear_store_colstats(e->g_a1, n, e->gl.H1, e->g_w2h, e->gl.H2, e->g_h2, e->bn_sums + 2 * e->gl.H1,
                                 e->bn_sums + 2 * e->gl.H1 + e->gl.H2, e->partials, st,
                                 e->whole_step ? &e->bn_pending_blocks : nullptr);
  return linear_store<false, false, false>(e->g_a1, n, e->gl.H1, e->g_w2h, e->gl.H2, nullptr, e->g_h2, nullptr, st);
}
void g_bn_stats(PiganEngine* e, int which, int64_t n, cudaStream_t st) {
  const int H1 = e->gl.H1, H2 = e->gl.H2;
  if (gemm_colstats(e)) return;   // the producing GEMM's epilogue already added them (g_layer1/2 with train = true)
  PM("g_bn_colstats");
  if (which == 1) launch_colstats(e->g_h1, n, H1, e->bn_sums, e->bn_sums + H1, e->partials, st);
  else launch_colstats(e->g_h2, n, H2, e->bn_sums + 2 * H1, e->bn_sums + 2 * H1 + H2, e->partials, st);
}
void g_bn_finalize(PiganEngine* e, int which, const float* gp, const float* offset, float* bn_buffers,
                   int64_t* nbt, double n_global, int num_updates, cudaStream_t st) {
  const GenLayout& L = e->gl;
  BnFinalizeArgs a;
  if (which == 1) {
    a.sum = e->bn_sums; a.sumsq = e->bn_sums + L.H1; a.offset = offset;
    a.gamma = gp + L.bn1_w; a.beta = gp + L.bn1_b;
    a.running_mean = bn_buffers ? bn_buffers + L.rm1 : nullptr;
    a.running_var = bn_buffers ? bn_buffers + L.rv1 : nullptr;
    a.num_batches_tracked = nbt ? reinterpret_cast<long long*>(nbt) : nullptr;
    a.mean = e->mean1; a.rstd = e->rstd1; a.scale = e->scale1; a.bias = e->bias1; a.C = L.H1;
  } else {
    a.sum = e->bn_sums + 2 * L.H1; a.sumsq = e->bn_sums + 2 * L.H1 + L.H2; a.offset = offset;
    a.gamma = gp + L.bn2_w; a.beta = gp + L.bn2_b;
    a.running_mean = bn_buffers ? bn_buffers + L.rm2 : nullptr;
    a.running_var = bn_buffers ? bn_buffers + L.rv2 : nullptr;
    a.num_batches_tracked = nbt ? reinterpret_cast<long long*>(nbt) + 1 : nullptr;
    a.mean = e->mean2; a.rstd = e->rstd2; a.scale = e->scale2; a.bias = e->bias2; a.C = L.H2;
  }
  a.n = n_global;
  a.num_updates = num_updates;
  PM("small");
  if (e->bn_pending_blocks > 0) {
    launch_bn_reduce_finalize(a, e->partials, e->bn_pending_blocks, st);
    e->bn_pending_blocks = 0;
    return;
  }
  launch_bn_finalize(a, st);
}

// the fused generator-head + surrogate-layer-1 epilogue (EpiHeadF1) is written for the reference widths
bool fused_head(const PiganEngine* e) {
  return e->f_loaded && e->gl.H2 == 256 && e->gl.P == 4 && e->fl.H[0] == 256;
}
// Generator in eval mode from the prepared operand e->xc: BatchNorm-1 (running statistics) + ReLU are folded
// into layer 1's epilogue, BatchNorm-2 + ReLU into the head kernel.  `packed`: weights / affines already set up.
int g_eval_setup(PiganEngine* e, const float* gp, const float* bn, cudaStream_t st) {
  const GenLayout& G = e->gl;
  PIGAN_TRY(pack_generator(e, gp, false, st));
  PM("small");
  launch_bn_eval_affine(bn + G.rm1, bn + G.rv1, gp + G.bn1_w, gp + G.bn1_b, e->g_beff, e->scale1, e->bias1, G.H1, st);
  launch_bn_eval_affine(bn + G.rm2, bn + G.rv2, gp + G.bn2_w, gp + G.bn2_b, gp + G.b2, e->scale2, e->bias2, G.H2, st);
  if (fused_head(e)) {
    // constants of the fused head + surrogate-layer-1 epilogue (EpiHeadF1)
    const FwdLayout& L = e->fl;
    const float* fp = e->f_params;
    launch_head_consts(e->scale2, e->bias2, gp + G.w3, gp + G.b3, fp + L.w[0], fp + L.b[0], fp + L.ln_w[0],
                       fp + L.ln_b[0], e->head_img, st);
  }
  return PIGAN_OK;
}
// with_f1: also emit the surrogate's first activation (e->f_a1) from the fused epilogue (needs the surrogate loaded)
int g_eval_forward(PiganEngine* e, const float* gp, int64_t n, float* p_out, cudaStream_t st, bool with_f1 = false) {
  const GenLayout& G = e->gl;
  {
    PM("g_l1_gemm");
    auto run = [&](auto epi_tag) -> int {
      using Epi = decltype(epi_tag);
      typename Epi::Params ep;
      PIGAN_TRY(out_map(&ep.out, e->g_a1, n, G.H1, G.H1));
      ep.bias = e->bias1;
      ep.scale = e->scale1;
      ep.rowstats = nullptr;
      ep.n_tiles = 0;
      ep.mask = nullptr;
      ep.mask_words = 0;
      ep.colpart = nullptr;
      ep.col_groups = 0;
      return run_tn<CfgSR, Epi>(ep, e->xc, n, kKp, kKp, e->g_w1h, G.H1, kKp, st);
    };
    if (store_groups4()) PIGAN_TRY(run(EpiStore<CfgSR, false, false, false, false, true, false, 4>{}));
    else PIGAN_TRY(run(EpiStore<CfgSR, false, false, false, false, true>{}));
  }
  if (with_f1 && fused_head(e)) {
    PM("g_l2_head_f1_gemm");
    using Epi = EpiHeadF1<CfgH>;
    Epi::Params ep;
    PIGAN_TRY(out_map(&ep.out, e->f_a1, n, e->fl.H[0], e->fl.H[0]));
    ep.p_out = p_out;
    ep.img = e->head_img;
    return run_tn<CfgH, Epi>(ep, e->g_a1, n, G.H1, G.H1, e->g_w2h, G.H2, G.H1, st);
  }
  PM("g_l2_gemm");
  PIGAN_TRY((linear_store<false, false, false>(e->g_a1, n, G.H1, e->g_w2h, G.H2, nullptr, e->g_h2, nullptr, st)));
  PM("g_head_fwd");
  launch_g_head_fwd(e->g_h2, e->scale2, e->bias2, gp + G.w3, gp + G.b3, p_out, nullptr, e->xc, nullptr, n, G.H2, kKp,
                    G.S, st);
  return PIGAN_OK;
}

// ------------------------------------------------------------------------------------------ discriminator
// z1 rows [row0, row0+n) = LeakyReLU([xc | tail] . w1h^T)   (bias and parameter columns inside the MMA)
int d_layer1(PiganEngine* e, int64_t n, int64_t row0, bool fake, cudaStream_t st) {
  PM("d_l1_gemm");
  return linear_store<false, true, false, true>(e->xc, n, e->kp, e->d_w1h, e->dl.H1, nullptr,
                                                e->d_z1 + row0 * e->dl.H1, nullptr, st, fake ? e->tail_f : nullptr,
                                                e->d_mask1 + row0 * (e->dl.H1 / 32));
}

struct DL2Opts {
  int64_t rows;          // rows of z1 / z2 processed (incl. the gap)
  int64_t rows_a;        // rows < rows_a carry label_a
  float label_a, label_b;
  int64_t gap_begin, gap_end;
  double global_batch;
  double* loss_sum;
  float* prob_out;
  bool store_z2, want_dlogit;
};
int d_layer2(PiganEngine* e, const float* dp, const DL2Opts& o, cudaStream_t st) {
  using Epi = EpiDiscL2<CfgS>;
  const DiscLayout& L = e->dl;
  if (L.H2 != 256) {
    // widened discriminator: the row's logit needs all H2 columns, which no longer sit in one accumulator tile -
    // layer 2 as a plain Linear + LeakyReLU store, layer 3 + Sigmoid + BCE as a streaming pass over z2
    PM("d_l2_gemm");
    PIGAN_TRY((linear_store<true, true, false>(e->d_z1, o.rows, L.H1, e->d_w2h, L.H2, dp + L.b2, e->d_z2, nullptr, st)));
    PM("d_logit_bce");
    launch_d_logit_bce(e->d_z2, dp + L.w3, dp + L.b3, o.rows, L.H2, o.rows_a, o.label_a, o.label_b, o.gap_begin,
                       o.gap_end, o.global_batch, o.loss_sum, o.want_dlogit ? e->dlogit : nullptr, o.prob_out, st);
    PIGAN_CUDA_OK(cudaGetLastError());
    return PIGAN_OK;
  }
  Epi::Params ep;
  PIGAN_TRY(out_map(&ep.z2, e->d_z2, o.rows, L.H2, L.H2));
  ep.b2 = dp + L.b2;
  ep.w3 = dp + L.w3;
  ep.b3 = dp + L.b3;
  ep.label_a = o.label_a;
  ep.label_b = o.label_b;
  ep.rows_a = (int)o.rows_a;
  ep.row_gap_begin = (int)o.gap_begin;
  ep.row_gap_end = (int)o.gap_end;
  ep.inv_batch = (float)(1.0 / o.global_batch);
  ep.grad_mult = 1.0f;  // GS == global batch
  ep.loss_sum = o.loss_sum;
  ep.dlogit = o.want_dlogit ? e->dlogit : nullptr;
  ep.prob_out = o.prob_out;
  ep.store_z2 = o.store_z2 ? 1 : 0;
  PM("d_l2_bce_gemm");
  return run_tn<CfgS, Epi>(ep, e->d_z1, o.rows, L.H1, L.H1, e->d_w2h, L.H2, L.H1, st);
}

// D layers 2+3 + BCE with the backward of layer 3 / layer 2's activation fused in (EpiDiscL2Bwd): writes dh2 instead
// of z2; param_grads (D-step): dw3 / db2 / db3 accumulate into d_grads.  PIGAN_FUSE_DL2=0: the two-kernel path.
bool fuse_dl2() {
  static const bool on = [] { const char* v = getenv("PIGAN_FUSE_DL2"); return !(v && v[0] == '0'); }();
  return on;
}
template <bool PG>
int d_layer2_bwd(PiganEngine* e, const float* dp, const DL2Opts& o, float* d_grads, float inv_gs, cudaStream_t st) {
  using Epi = EpiDiscL2Bwd<CfgL1, PG>;
  const DiscLayout& L = e->dl;
  typename Epi::Params ep;
  PIGAN_TRY(out_map(&ep.dh2, e->d_dh2, o.rows, L.H2, L.H2));
  ep.b2 = dp + L.b2;
  ep.w3 = dp + L.w3;
  ep.b3 = dp + L.b3;
  ep.label_a = o.label_a;
  ep.label_b = o.label_b;
  ep.rows_a = (int)o.rows_a;
  ep.row_gap_begin = (int)o.gap_begin;
  ep.row_gap_end = (int)o.gap_end;
  ep.inv_batch = (float)(1.0 / o.global_batch);
  ep.grad_mult = 1.0f;  // GS == global batch
  ep.loss_sum = o.loss_sum;
  ep.part = e->partials;
  PM("d_l2_bce_gemm");
  PIGAN_TRY((run_tn<CfgL1, Epi>(ep, e->d_z1, o.rows, L.H1, L.H1, e->d_w2h, L.H2, L.H1, st)));
  if constexpr (PG) {
    const int units = ceil_div((int)o.rows, kBlockM);
    const int grid = units < sm_count() ? units : sm_count();   // launch_gemm's grid: one partial row per CTA
    ReduceArgs r{e->partials, grid, 2 * L.H2 + 8, 3, {{d_grads + L.w3, L.H2, inv_gs}, {d_grads + L.b2, L.H2, inv_gs},
                                                        {d_grads + L.b3, 1, inv_gs}}};
    launch_reduce_columns(r, st);
  }
  return PIGAN_OK;
}

// ------------------------------------------------------------------------------------------ forward model
struct FOutOpts {
  int target_mode;  // 0 none, 1 one fp32 row (e->cvec holds it, padded), 2 the centred fp16 operand e->xc + e->cvec
  const float* target_metrics;
  const float* p_norm;
  double* sums;
  float* dp_lc;
  float lc_grad_mult;
  float* out_full;
  float* row_err;
  int f1_idx, f2_idx;
};
int f_out_layer(PiganEngine* e, const __half* a5, int64_t n, const FOutOpts& o, cudaStream_t st);
// Widened output layer (any S + Mt): the fp32 accumulators leave tensor memory as [128 x 256] slabs in e->f_slab
// (TMA stores of EpiWeightGradPartial, here on a TN product); the consumers add the bias in fp32
// (f_unslab_kernel / f_out_loss_slab_kernel).  Weights w6h: fp16 [OUT, H5] - rows beyond OUT are zero-filled by TMA.
int f_out_slab(PiganEngine* e, const __half* a5, int64_t n, const __half* w6h, cudaStream_t st) {
  const FwdLayout& L = e->fl;
  using Epi = EpiWeightGradPartial<CfgS>;
  Epi::Params ep;
  const int m_tiles = ceil_div((int)n, kBlockM);
  PIGAN_TRY(make_tmap_f32_2d(&ep.part, e->f_slab, 256, (uint64_t)m_tiles * e->out_groups * kBlockM, 256, kBlockM));
  PM("f_out_gemm");
  return run_tn<CfgS, Epi>(ep, a5, n, L.H[4], L.H[4], w6h, L.OUT, L.H[4], st);
}
// Eval-mode surrogate at widened dims: first layer (streaming), four Linear GEMMs with LayerNorm row partials from
// the epilogue + the in-place LayerNorm/LeakyReLU pass, output layer through the slabs.
int f_forward_wide(PiganEngine* e, const float* p_norm, int64_t n, float* out_full, cudaStream_t st) {
  const FwdLayout& L = e->fl;
  const float* fp = e->f_params;
  PM("f_l1");
  launch_f_l1_wide(p_norm, fp + L.w[0], fp + L.b[0], fp + L.ln_w[0], fp + L.ln_b[0], e->f_l1c, nullptr, e->f_a1,
                   nullptr, nullptr, nullptr, n, L.H[0], nullptr, st);
  __half* acts[5] = {e->f_a1, e->f_a2, e->f_a3, e->f_a4, e->f_a5};
  for (int i = 1; i < 5; ++i) {
    PM("f_hidden_gemm");
    PIGAN_TRY((linear_store<true, false, true>(acts[i - 1], n, L.H[i - 1], e->f_wh[i], L.H[i], fp + L.b[i], acts[i],
                                               e->f_rowstats, st)));
    PM("f_ln_apply");
    launch_ln_lrelu_apply(acts[i], e->f_rowstats, L.H[i] / 256, fp + L.ln_w[i], fp + L.ln_b[i], n, L.H[i], st);
  }
  PIGAN_TRY(f_out_slab(e, e->f_a5, n, e->f_wh[5], st));
  if (out_full != nullptr) {   // (the G-step's loss pass reads the slabs themselves)
    PM("f_out_unslab");
    launch_f_unslab(e->f_slab, e->out_groups, fp + L.b[5], out_full, n, L.OUT, st);
  }
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
int f_forward(PiganEngine* e, const float* p_norm, int64_t n, const FOutOpts& o, cudaStream_t st,
              bool have_a1 = false) {
  if (!e->f_loaded) return fail(PIGAN_ERR_INVALID, "forward model not loaded (pigan_engine_load_forward_model)");
  if (!e->full) {
    if (o.out_full == nullptr || o.sums != nullptr || o.row_err != nullptr || o.target_mode != 0 || have_a1)
      return need_full(e);
    return f_forward_wide(e, p_norm, n, o.out_full, st);
  }
  const FwdLayout& L = e->fl;
  const float* fp = e->f_params;
  PM("f_l1");
  if (!have_a1) launch_f_l1(p_norm, fp + L.w[0], fp + L.b[0], fp + L.ln_w[0], fp + L.ln_b[0], e->f_a1, n, L.H[0], st);
  __half* acts[5] = {e->f_a1, e->f_a2, e->f_a3, e->f_a4, e->f_a5};
  for (int i = 1; i < 5; ++i) {
    PM("f_hidden_gemm");
    PIGAN_TRY(linear_ln(acts[i - 1], n, L.H[i - 1], e->f_wh[i], L.H[i], fp + L.b[i], fp + L.ln_w[i], fp + L.ln_b[i],
                        acts[i], st));
  }
  return f_out_layer(e, e->f_a5, n, o, st);
}
// Output layer (forward_model.py:55) with the fused loss / error epilogue; weights from e->f_wh[5], e->f_bias_out
template <int TMODE, bool TRAIN>
int f_out_loss(PiganEngine* e, const __half* a5, int64_t n, const FOutOpts& o, cudaStream_t st) {
  const FwdLayout& L = e->fl;
  using Epi = EpiFwdLoss<CfgO4, TMODE, TRAIN>;
  typename Epi::Params ep;
  ep.bias = e->f_bias_out;
  ep.tcen = e->cvec;
  PIGAN_TRY(make_tmap_f16_2d(&ep.tgt, e->xc, kKp, (uint64_t)n, kKp, 64, kBlockM));
  ep.S = L.S;
  ep.Mt = L.Mt;
  ep.target_metrics = o.target_metrics;
  ep.p_norm = o.p_norm;
  ep.sums = o.sums;
  ep.dp_lc = o.dp_lc;
  ep.lc_grad_mult = o.lc_grad_mult;
  ep.row_err = o.row_err;
  ep.f1_idx = o.f1_idx;
  ep.f2_idx = o.f2_idx;
  PM("f_out_gemm");
  PIGAN_TRY((run_tn<CfgO4, Epi>(ep, a5, n, L.H[4], L.H[4], e->f_wh[5], L.OUT, L.H[4], st)));
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
// candidate search: one target row, spectrum columns only, double-buffered 128 x 256 tiles (EpiFwdScore);
// PIGAN_SCORE_TILE=288 keeps the 288-column loss epilogue for this mode too
int f_out_score(PiganEngine* e, const __half* a5, int64_t n, const FOutOpts& o, cudaStream_t st) {
  const FwdLayout& L = e->fl;
  using Epi = EpiFwdScore<CfgS>;
  Epi::Params ep;
  ep.bias = e->f_bias_out;
  ep.tcen = e->cvec;
  ep.S = L.S;
  ep.row_err = o.row_err;
  PM("f_out_gemm");
  // weight map over the S spectrum rows only: the tile's remaining columns are zero-filled by TMA
  PIGAN_TRY((run_tn<CfgS, Epi>(ep, a5, n, L.H[4], L.H[4], e->f_wh[5], L.S, L.H[4], st)));
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
int f_out_layer(PiganEngine* e, const __half* a5, int64_t n, const FOutOpts& o, cudaStream_t st) {
  const FwdLayout& L = e->fl;
  static const bool wide_score = [] { const char* v = getenv("PIGAN_SCORE_TILE"); return v && v[0] == '2' && v[1] == '8'; }();
  if (o.out_full == nullptr && o.sums == nullptr && o.row_err != nullptr && o.target_mode == 1 && L.S <= 256 &&
      L.H[4] % 64 == 0 && !wide_score)
    return f_out_score(e, a5, n, o, st);
  // the loss / scoring modes have their own lean epilogue; the generic one keeps the fp32 dump of the output
  if (o.out_full == nullptr && L.S <= 256 && L.OUT <= 288) {
    if (o.sums != nullptr && o.target_mode == 2 && o.row_err == nullptr) return f_out_loss<2, true>(e, a5, n, o, st);
    if (o.sums == nullptr && o.row_err != nullptr && o.target_mode == 2) return f_out_loss<2, false>(e, a5, n, o, st);
    if (o.sums == nullptr && o.row_err != nullptr && o.target_mode == 1) return f_out_loss<1, false>(e, a5, n, o, st);
  }
  using Epi = EpiFwdOut<CfgO>;
  Epi::Params ep;
  ep.bias = e->f_bias_out;
  ep.S = L.S;
  ep.Mt = L.Mt;
  ep.target_mode = o.target_mode;
  ep.target_row = e->cvec;
  ep.center = e->cvec;
  PIGAN_TRY(make_tmap_f16_2d(&ep.tgt, e->xc, kKp, (uint64_t)n, kKp, 64, kBlockM));
  ep.target_metrics = o.target_metrics;
  ep.p_norm = o.p_norm;
  ep.sums = o.sums;
  ep.dp_lc = o.dp_lc;
  ep.lc_grad_mult = o.lc_grad_mult;
  ep.out_full = o.out_full;
  ep.row_err = o.row_err;
  ep.f1_idx = o.f1_idx;
  ep.f2_idx = o.f2_idx;
  PM("f_out_gemm");
  PIGAN_TRY((run_tn<CfgO, Epi>(ep, a5, n, L.H[4], L.H[4], e->f_wh[5], L.OUT, L.H[4], st)));
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

// ------------------------------------------------------------------------------------------ train phases
// Generator head + BatchNorm-2 backward: same arguments for the statistics pass (phase 3) and the apply pass
// (phase 4, after the data-parallel reduction of the column sums).
GHeadBwdArgs head_bwd_args(PiganEngine* e, const PiganTrainArgs& a) {
  const GenLayout& G = e->gl;
  const float inv_gs = (float)(1.0 / (double)a.global_batch);
  float* gp = a.g_params;
  GHeadBwdArgs hb;
  hb.p = e->p; hb.dpden = e->dpden; hb.dp_lc = e->dp_lc;
  hb.dp_extra = a.dp_extra; hb.gs = (float)a.global_batch;
  hb.range_mult = a.lambda_param_range / (float)G.P;
  hb.h2 = e->g_h2; hb.scale = e->scale2; hb.bias = e->bias2; hb.mean = e->mean2; hb.rstd = e->rstd2;
  hb.w3 = gp + G.w3; hb.dy2 = e->g_dy2; hb.dw3 = a.g_grads + G.w3; hb.db3 = a.g_grads + G.b3;
  hb.sum_dy = e->bn_bwd_sums; hb.sum_dyx = e->bn_bwd_sums + G.H2; hb.range_sum = e->sums + kSumRange;
  hb.inv_gs = inv_gs; hb.rows = a.batch; hb.C = G.H2;
  hb.gamma = gp + G.bn2_w; hb.dbias = a.g_grads + G.b2; hb.dgamma = a.g_grads + G.bn2_w;
  hb.dbeta = a.g_grads + G.bn2_b; hb.inv_n = 1.0 / (double)a.global_batch; hb.part = e->partials; hb.dpre = e->dpre;
  hb.dpre_part = nullptr;   // placed inside the partial scratch by launch_g_head_bwd
  return hb;
}

bool overlap_enabled() {
  static const bool on = [] {
    const char* v = getenv("PIGAN_OVERLAP");
    return !(v && v[0] == '0');
  }();
  return on;
}
// surrogate forward of the G-step + its fused losses (train_pigan.py:156-172)
int g_step_surrogate(PiganEngine* e, const PiganTrainArgs& a, cudaStream_t st) {
  if (!e->full) {
    // widened: eval-mode chain into the fp32 slabs, then the loss pass (no gradient through F: train_pigan.py:156-157)
    if (!e->f_loaded) return fail(PIGAN_ERR_INVALID, "forward model not loaded (pigan_engine_load_forward_model)");
    PIGAN_TRY(f_forward_wide(e, e->p, a.batch, nullptr, st));
    PM("f_pigan_loss");
    const FwdLayout& L = e->fl;
    launch_f_pigan_loss_slab(e->f_slab, e->out_groups, e->f_params + L.b[5], a.spectrum, a.metrics_norm, e->p, a.batch,
                             L.S, L.Mt, a.f1_idx, a.f2_idx, a.lambda_lc, e->sums + kSumRec, e->dp_lc, st);
    PIGAN_CUDA_OK(cudaGetLastError());
    return PIGAN_OK;
  }
  FOutOpts fo{2, a.metrics_norm, e->p, e->sums + kSumRec, e->dp_lc, a.lambda_lc, nullptr, nullptr, a.f1_idx, a.f2_idx};
  return f_forward(e, e->p, a.batch, fo, st);
}
// phase 2, once e->p exists: start the surrogate chain on the side stream
int fork_surrogate(PiganEngine* e, const PiganTrainArgs& a, cudaStream_t st) {
  e->side_pending = false;
  if (!overlap_enabled() || e->prof.on || (a.flags & 1)) return PIGAN_OK;   // profiling keeps one stream so sections stay meaningful
  if (e->side == nullptr) {
    int lo = 0, hi = 0;
    PIGAN_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PIGAN_CUDA_OK(cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, hi));   // measured: priority is neutral
    PIGAN_CUDA_OK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    PIGAN_CUDA_OK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  }
  PIGAN_CUDA_OK(cudaEventRecord(e->ev_fork, st));
  PIGAN_CUDA_OK(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
  PIGAN_TRY(g_step_surrogate(e, a, e->side));
  PIGAN_CUDA_OK(cudaEventRecord(e->ev_join, e->side));
  e->side_pending = true;
  return PIGAN_OK;
}

int train_phase(PiganEngine* e, const PiganTrainArgs& a, int phase, cudaStream_t st) {
  const GenLayout& G = e->gl;
  const DiscLayout& D = e->dl;
  const int64_t B = a.batch;
  const int64_t BP = round_up(B, 128);  // fake half starts here in the stacked D tensors
  const double NG = (double)a.global_batch;
  const float inv_gs = (float)(1.0 / NG);  // gradient scale GS = global batch
  float* gp = a.g_params;
  float* dp = a.d_params;

  switch (phase) {
    case 0: {
      PM("memset");
      if (((reinterpret_cast<uintptr_t>(a.g_grads) | reinterpret_cast<uintptr_t>(a.d_grads)) & 15u) == 0) {
        // one kernel instead of four memset stream operations
        float* ptrs[4] = {a.g_grads, a.d_grads, reinterpret_cast<float*>(e->zero_blk), e->dpden};
        const int64_t nf[4] = {G.total, D.total, (int64_t)(e->zero_bytes / sizeof(float)), B * 4};
        launch_zero_buffers(ptrs, nf, 4, st);
      } else {
        PIGAN_CUDA_OK(cudaMemsetAsync(a.g_grads, 0, G.total * sizeof(float), st));
        PIGAN_CUDA_OK(cudaMemsetAsync(a.d_grads, 0, D.total * sizeof(float), st));
        PIGAN_CUDA_OK(cudaMemsetAsync(e->zero_blk, 0, e->zero_bytes, st));
        PIGAN_CUDA_OK(cudaMemsetAsync(e->dpden, 0, (size_t)B * 4 * sizeof(float), st));
      }
      if (a.spectrum_operand != nullptr) {
        // the caller prepared the fp16 operand (pigan_prepare_spectrum_operand, e.g. once per dataset) and names the
        // row it centred on: nothing to cast, half the bytes to move
        PM("prep_cast");
        e->xc = const_cast<__half*>(static_cast<const __half*>(a.spectrum_operand));
        launch_copy_pad_f32(a.spectrum_center, G.S, e->cvec, e->kp, st);
      } else {
        PIGAN_TRY(prep_spectrum(e, a.spectrum, a.params_denorm, B, st));
      }
      PIGAN_TRY(pack_generator(e, gp, true, st));
      PIGAN_TRY(g_layer1(e, B, st, true));
      g_bn_stats(e, 1, B, st);
      break;
    }
    case 1: {
      // the reference calls G.forward twice per step in train mode (train_pigan.py:131,148): same batch
      // statistics, running statistics updated twice
      g_bn_finalize(e, 1, gp, e->g_beff, a.g_bn_buffers, a.g_num_batches_tracked, NG, 2, st);
      PIGAN_TRY(g_layer2(e, gp, B, st, true));
      g_bn_stats(e, 2, B, st);
      break;
    }
    case 2: {
      g_bn_finalize(e, 2, gp, gp + G.b2, a.g_bn_buffers, a.g_num_batches_tracked, NG, 2, st);
      PM("g_head_fwd");
      launch_g_head_fwd(e->g_h2, e->scale2, e->bias2, gp + G.w3, gp + G.b3, e->p, e->pden, e->xc, e->tail_f, B, G.H2,
                        e->kp, G.S, st);
      PIGAN_TRY(fork_surrogate(e, a, st));
      if (!e->full) {
        // widened: the fake rows' operand = the real rows' spectrum columns + the head's 64-column tail (generated
        // parameters, the two constant-one columns), as rows [BP, BP + B) of the same matrix - read by the
        // discriminator's first-layer weight-gradient GEMM (the forward GEMM takes the tail as its last k-block)
        PM("d_fake_rows");
        const size_t pitch = (size_t)e->kp * sizeof(__half);
        PIGAN_CUDA_OK(cudaMemcpy2DAsync(e->xc + (size_t)BP * e->kp, pitch, e->xc, pitch, pitch - 128, (size_t)B,
                                        cudaMemcpyDeviceToDevice, st));
        PIGAN_CUDA_OK(cudaMemcpy2DAsync(e->xc + (size_t)BP * e->kp + (e->kp - 64), pitch, e->tail_f, 128, 128, (size_t)B,
                                        cudaMemcpyDeviceToDevice, st));
      }
      // ---- D-step (train_pigan.py:123-143)
      PIGAN_TRY(pack_discriminator(e, dp, true, st));
      PIGAN_TRY(d_layer1(e, B, 0, false, st));
      PIGAN_TRY(d_layer1(e, B, BP, true, st));
      DL2Opts o{BP + B, BP, 0.9f, 0.1f, B, BP, NG, e->sums + kSumD, nullptr, true, true};
      if (fuse_dl2() && D.H2 == 256) {
        PIGAN_TRY(d_layer2_bwd<true>(e, dp, o, a.d_grads, inv_gs, st));
      } else {
        PIGAN_TRY(d_layer2(e, dp, o, st));
        PM("d_l2_bwd");
        launch_d_l2_bwd(e->d_z2, e->dlogit, dp + D.w3, e->d_dh2, a.d_grads + D.w3, a.d_grads + D.b2, a.d_grads + D.b3,
                        BP + B, D.H2, inv_gs, e->partials, st);
      }
      {
        PM("d_dh1_gemm");
        auto run = [&](auto cfg_tag) -> int {
          using Cfg = decltype(cfg_tag);
          using Epi = EpiLeakyMaskStore<Cfg>;
          typename Epi::Params ep;
          PIGAN_TRY(out_map(&ep.out, e->d_dh1, BP + B, D.H1, D.H1));
          ep.mask = e->d_mask1;
          ep.mask_words = D.H1 / 32;
          return run_tn<Cfg, Epi>(ep, e->d_dh2, BP + B, D.H2, D.H2, e->d_w2th, D.H1, D.H2, st);
        };
        if (D.H2 <= CfgSR::B_RES_KB * kBlockK) PIGAN_TRY(run(CfgSR{}));
        else PIGAN_TRY(run(CfgS{}));   // widened: K = H2 streams
      }
      PM("d_dw2_gemm");
      PIGAN_TRY(weight_grad(e->d_dh2, BP + B, D.H2, e->d_z1, BP + B, D.H1, a.d_grads + D.w2, D.H1, D.H1, inv_gs, -1,
                            nullptr, 0, nullptr, 0, e->dw_part, st));
      PM("d_dw1_gemm");
      if (e->full)
        PIGAN_TRY(weight_grad(e->d_dh1, BP + B, D.H1, e->xc, BP, kKp, a.d_grads + D.w1, D.IN, D.IN, inv_gs, D.IN,
                              a.d_grads + D.b1, BP, e->tail_f, BP, e->dw_part, st, 0, e->cvec, D.S, D.P));
      else   // widened: rows [BP, BP + B) of the operand are the fake rows (built above), one plain product
        PIGAN_TRY(weight_grad(e->d_dh1, BP + B, D.H1, e->xc, BP + B, e->kp, a.d_grads + D.w1, D.IN, D.IN, inv_gs, D.IN,
                              a.d_grads + D.b1, 0, nullptr, 0, e->dw_part, st, 0, e->cvec, D.S, D.P));
      break;
    }
    case 3: {
      PM("clip_adam");
      double* sq = reinterpret_cast<double*>(e->partials);   // free here: block partials of the gradient norm
      const int nsq = launch_sumsq(a.d_grads, D.total, sq, st);
      AdamArgs ad{dp, a.d_grads, a.d_exp_avg, a.d_exp_avg_sq, D.total, a.lr_d, 0.5f, 0.999f, 1e-8f,
                  1.0 - pow(0.5, (double)a.step), 1.0 - pow(0.999, (double)a.step), sq, nsq, 1.0f};
      launch_clip_adam(ad, st);
      // ---- G-step (train_pigan.py:145-187) against the updated discriminator
      PIGAN_TRY(pack_discriminator(e, dp, true, st));
      PIGAN_TRY(d_layer1(e, B, 0, true, st));
      DL2Opts o{B, B, 1.0f, 1.0f, 0, 0, NG, e->sums + kSumAdv, nullptr, true, true};
      if (fuse_dl2() && D.H2 == 256) {
        PIGAN_TRY(d_layer2_bwd<false>(e, dp, o, nullptr, inv_gs, st));
      } else {
        PIGAN_TRY(d_layer2(e, dp, o, st));
        PM("d_l2_bwd");
        launch_d_l2_bwd(e->d_z2, e->dlogit, dp + D.w3, e->d_dh2, nullptr, nullptr, nullptr, B, D.H2, inv_gs, e->partials, st);
      }
      {
        PM("d_paramgrad_gemm");
        // streamed operands on purpose: the epilogue reads 16 bytes of first-layer weights per accumulator column
        // through L1, and the resident-weight configuration leaves no shared memory for an L1 cache (66 vs 37 us)
        using Epi = EpiDiscParamGrad<CfgP>;
        Epi::Params ep{e->d_mask1, D.H1 / 32, e->d_wp, e->dpden};
        PIGAN_TRY((run_tn<CfgP, Epi>(ep, e->d_dh2, B, D.H2, D.H2, e->d_w2th, D.H1, D.H2, st)));
      }
      if (e->side_pending) {
        PIGAN_CUDA_OK(cudaStreamWaitEvent(st, e->ev_join, 0));   // the surrogate chain started in phase 2
        e->side_pending = false;
      } else {
        PIGAN_TRY(g_step_surrogate(e, a, st));
      }
      PM("g_head_bwd");
      launch_g_head_bwd(head_bwd_args(e, a), false, st);
      break;
    }
    case 4: {
      PM("g_head_bwd");
      launch_g_head_bwd(head_bwd_args(e, a), true, st);
      PM("g_dw2_gemm");
      PIGAN_TRY(weight_grad(e->g_dy2, B, G.H2, e->g_a1, B, G.H1, a.g_grads + G.w2, G.H1, G.H1, inv_gs, -1, nullptr, 0,
                            nullptr, 0, e->dw_part, st));
      PM("g_da1_gemm");
      PIGAN_TRY((linear_store<false, false, false>(e->g_dy2, B, G.H2, e->g_w2th, G.H1, nullptr, e->g_da1, nullptr, st)));
      PM("g_bn_bwd_stats");
      launch_bn_bwd_stats(e->g_da1, e->g_h1, e->scale1, e->bias1, e->mean1, e->rstd1, e->bn_bwd_sums + 2 * G.H2,
                          e->bn_bwd_sums + 2 * G.H2 + G.H1, B, G.H1, e->partials, st);
      break;
    }
    case 5: {
      BnBwdArgs bb;
      bb.dy = e->g_da1; bb.h = e->g_h1; bb.relu_mask = 1;
      bb.scale = e->scale1; bb.bias = e->bias1; bb.mean = e->mean1; bb.rstd = e->rstd1; bb.gamma = gp + G.bn1_w;
      bb.sum_dy = e->bn_bwd_sums + 2 * G.H2; bb.sum_dyx = e->bn_bwd_sums + 2 * G.H2 + G.H1;
      bb.dh = e->g_da1; bb.dbias = nullptr; bb.dgamma = a.g_grads + G.bn1_w; bb.dbeta = a.g_grads + G.bn1_b;
      bb.inv_n = 1.0 / NG; bb.inv_gs = inv_gs; bb.rows = B; bb.C = G.H1; bb.part = e->partials;
      PM("g_bn_bwd_apply");
      launch_bn_bwd_apply(bb, st);
      PM("g_dw1_gemm");
      PIGAN_TRY(weight_grad(e->g_da1, B, G.H1, e->xc, B, e->kp, a.g_grads + G.w1, G.S, G.S, inv_gs, G.S + G.P,
                            a.g_grads + G.b1, 0, nullptr, 0, e->dw_part, st, 0, e->cvec, G.S, 0));
      break;
    }
    case 6: {
      PM("clip_adam");
      double* sq = reinterpret_cast<double*>(e->partials);
      const int nsq = launch_sumsq(a.g_grads, G.total, sq, st);
      AdamArgs ad{gp, a.g_grads, a.g_exp_avg, a.g_exp_avg_sq, G.total, a.lr_g, 0.5f, 0.999f, 1e-8f,
                  1.0 - pow(0.5, (double)a.step), 1.0 - pow(0.999, (double)a.step), sq, nsq, 1.0f};
      launch_clip_adam(ad, st);
      if (a.losses) {
        LossFinalizeArgs lf{e->sums, a.losses, NG, G.S, e->fl.Mt, G.P, a.lambda_recon, a.lambda_physics_spectrum,
                            a.lambda_physics_metrics, a.lambda_maxwell, a.lambda_lc, a.lambda_param_range,
                            a.lambda_bnn_kl};
        launch_loss_finalize(lf, st);
      }
      break;
    }
    default:
      return fail(PIGAN_ERR_INVALID, "train phase %d out of range", phase);
  }
  PM(nullptr);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

int check_train_args(PiganEngine* e, const PiganTrainArgs* a) {
  PIGAN_CHECK_ARG(e != nullptr && a != nullptr);
  if (!e->gan)
    return fail(PIGAN_ERR_UNSUPPORTED, "this engine's dimensions serve the surrogate only (PI-GAN step: the reference "
                                       "widths, or widened dims with spectrum_dim % 64 == 0 and generator / discriminator "
                                       "widths of 256 / 512 / 1024 / 2048)");
  if (!e->full && (a->spectrum_operand != nullptr || a->spectrum == nullptr || a->params_denorm == nullptr))
    return fail(PIGAN_ERR_UNSUPPORTED, "widened PI-GAN step: pass fp32 spectrum / params_denorm (no prepared operand)");
  PIGAN_CHECK_ARG(a->batch >= 2 && a->batch <= e->max_batch && a->global_batch >= a->batch);
  PIGAN_CHECK_ARG(a->metrics_norm != nullptr);
  PIGAN_CHECK_ARG(a->spectrum_operand ? (a->spectrum_center != nullptr &&
                                         (reinterpret_cast<uintptr_t>(a->spectrum_operand) & 15u) == 0)
                                      : (a->spectrum != nullptr && a->params_denorm != nullptr));
  // caller-prepared operand: the weight-gradient GEMM of D's first layer walks the operand in 128-row tiles (fake rows
  // wrap onto the same tiles), so a ragged slice would be read up to 127 rows past its end
  if (a->spectrum_operand != nullptr && a->batch % 128 != 0)
    return fail(PIGAN_ERR_INVALID, "spectrum_operand: batch %lld must be a multiple of 128 (pass spectrum / "
                "params_denorm for ragged batches)", (long long)a->batch);
  PIGAN_CHECK_ARG(a->g_params && a->g_grads && a->g_exp_avg && a->g_exp_avg_sq);
  PIGAN_CHECK_ARG(a->d_params && a->d_grads && a->d_exp_avg && a->d_exp_avg_sq);
  PIGAN_CHECK_ARG(a->step >= 1);
  PIGAN_CHECK_ARG(a->f1_idx >= 0 && a->f1_idx < e->fl.Mt && a->f2_idx >= 0 && a->f2_idx < e->fl.Mt);
  return PIGAN_OK;
}

}  // namespace
}  // namespace pigan

// =============================================================================================== C ABI
extern "C" size_t pigan_engine_workspace_bytes(const PiganDims* dims, int64_t max_batch) {
  PiganDims d;
  if (dims) d = *dims; else pigan_default_dims(&d);
  if ((!dims_are_default(d) && !surrogate_dims_ok(d)) || max_batch < 1) return 0;
  PiganEngine e(d);
  e.max_batch = max_batch;
  e.bp = round_up(max_batch, 128);
  return e.carve(nullptr);
}

extern "C" int pigan_engine_create(PiganEngine** out, const PiganDims* dims, int64_t max_batch, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  PIGAN_CHECK_ARG(out != nullptr && workspace != nullptr && max_batch >= 1);
  PiganDims d;
  if (dims) d = *dims; else pigan_default_dims(&d);
  PIGAN_TRY(check_dims(d));
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0)
    return fail(PIGAN_ERR_INVALID, "workspace must be 256-byte aligned");
  PiganEngine* e = new (std::nothrow) PiganEngine(d);
  if (!e) return fail(PIGAN_ERR_INVALID, "out of host memory");
  e->max_batch = max_batch;
  e->bp = round_up(max_batch, 128);
  const size_t need = e->carve(nullptr);
  if (workspace_bytes < need) {
    delete e;
    return fail(PIGAN_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  }
  e->ws_bytes = e->carve(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // zero once: padding rows/columns of the operand tensors are read (and multiplied by zeros) but never written
  cudaError_t err = cudaMemsetAsync(workspace, 0, need, st);
  if (err != cudaSuccess) {
    delete e;
    return fail(PIGAN_ERR_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(err));
  }
  *out = e;
  return PIGAN_OK;
}

extern "C" int pigan_engine_destroy(PiganEngine* e) {
  if (e) {
    for (cudaEvent_t ev : e->prof.pool) cudaEventDestroy(ev);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
  }
  delete e;
  return PIGAN_OK;
}

extern "C" int pigan_engine_load_forward_model(PiganEngine* e, const float* fp, void* stream) {
  PIGAN_CHECK_ARG(e != nullptr && fp != nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const FwdLayout& L = e->fl;
  int in = L.H[0];
  for (int i = 1; i < 6; ++i) {
    const int out = i < 5 ? L.H[i] : L.OUT;
    launch_cast_pad(fp + L.w[i], in, in, e->f_wh[i], in, out, st);
    in = out;
  }
  launch_copy_pad_f32(fp + L.b[5], L.OUT, e->f_bias_out, e->dout_ld > 288 ? e->dout_ld : 288, st);
  PIGAN_CUDA_OK(cudaGetLastError());
  e->f_params = fp;
  e->f_loaded = true;
  return PIGAN_OK;
}

extern "C" int pigan_prepare_spectrum_operand(const float* spectrum, const float* params_denorm, const float* center,
                                              int64_t n, int32_t spectrum_dim, int32_t param_dim, void* out_operand,
                                              void* stream) {
  PIGAN_CHECK_ARG(spectrum && center && out_operand && n >= 1);
  PIGAN_CHECK_ARG(spectrum_dim >= 1 && param_dim >= 0 && spectrum_dim + param_dim + 2 <= kKp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(center) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out_operand) & 15u) == 0);
  // the kernel reads center[j] for j < S only, so the row needs no padding
  launch_cast_center(spectrum, center, params_denorm, static_cast<__half*>(out_operand), n, spectrum_dim, param_dim, kKp,
                     st);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_engine_set_spectrum_center(PiganEngine* e, const float* center) {
  PIGAN_CHECK_ARG(e != nullptr);
  e->center = center;
  return PIGAN_OK;
}

extern "C" int pigan_generator_forward(PiganEngine* e, const float* gp, float* bn, int64_t* nbt, const float* x,
                                       int64_t n, int32_t training, float* out, void* stream) {
  PIGAN_CHECK_ARG(e && gp && x && out && n >= 1 && n <= e->max_batch);
  PIGAN_TRY(need_full(e));
  PIGAN_CHECK_ARG(training ? n >= 2 : bn != nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GenLayout& G = e->gl;
  PIGAN_TRY(prep_spectrum(e, x, nullptr, n, st));
  if (!training) {
    PIGAN_TRY(g_eval_setup(e, gp, bn, st));
    PIGAN_TRY(g_eval_forward(e, gp, n, out, st));
    PM(nullptr);
    PIGAN_CUDA_OK(cudaGetLastError());
    return PIGAN_OK;
  }
  PIGAN_TRY(pack_generator(e, gp, false, st));
  PIGAN_CUDA_OK(cudaMemsetAsync(e->bn_sums, 0, (2 * G.H1 + 2 * G.H2) * sizeof(float), st));
  PIGAN_TRY(g_layer1(e, n, st, true));
  g_bn_stats(e, 1, n, st);
  g_bn_finalize(e, 1, gp, e->g_beff, bn, nbt, (double)n, 1, st);
  PIGAN_TRY(g_layer2(e, gp, n, st, true));
  g_bn_stats(e, 2, n, st);
  g_bn_finalize(e, 2, gp, gp + G.b2, bn, nbt, (double)n, 1, st);
  PM("g_head_fwd");
  launch_g_head_fwd(e->g_h2, e->scale2, e->bias2, gp + G.w3, gp + G.b3, out, nullptr, e->xc, nullptr, n, G.H2, kKp,
                    G.S, st);
  PM(nullptr);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_discriminator_forward(PiganEngine* e, const float* dp, const float* x, const float* params,
                                           int64_t n, float* out_prob, void* stream) {
  PIGAN_CHECK_ARG(e && dp && x && params && out_prob && n >= 1 && n <= e->max_batch);
  PIGAN_TRY(need_full(e));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PIGAN_TRY(prep_spectrum(e, x, params, n, st));
  PIGAN_TRY(pack_discriminator(e, dp, false, st));
  PIGAN_TRY(d_layer1(e, n, 0, false, st));
  DL2Opts o{n, n, 1.0f, 1.0f, 0, 0, (double)n, nullptr, out_prob, false, false};
  PIGAN_TRY(d_layer2(e, dp, o, st));
  PM(nullptr);
  return PIGAN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Backward passes of the two trainable modules for callers that drive them through autograd (the drop-in modules'
// torch.autograd.Function wrappers, core/models/*.py) instead of the fused train step.  Both RECOMPUTE the forward
// pass from the saved inputs (the engine's activation buffers are shared by every module call, so they cannot be
// relied on between a forward and its backward), then run the same backward kernels as pigan_train_step.
// `grad_scale` s: the fp16 gradient tensors hold s x the true gradient (the train step uses s = global batch); the
// caller picks s so that s * max|upstream gradient| is O(1).  Outputs are unscaled.
// ---------------------------------------------------------------------------------------------------------------

// Generator (generator.py:28-33, train mode: batch statistics; the running buffers are NOT touched here):
// g_grads[G.total] = d/d(params) of sum(grad_p * G(x)); g_grads is overwritten.
extern "C" int pigan_generator_backward(PiganEngine* e, const float* gp, const float* x, int64_t n, const float* grad_p,
                                        float grad_scale, float* g_grads, void* stream) {
  PIGAN_CHECK_ARG(e && gp && x && grad_p && g_grads && n >= 2 && n <= e->max_batch && grad_scale > 0.f);
  PIGAN_TRY(need_full(e));
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(g_grads) & 15u) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GenLayout& G = e->gl;
  const float inv_gs = 1.0f / grad_scale;
  // ---- forward, activations into the engine's buffers
  PIGAN_TRY(prep_spectrum(e, x, nullptr, n, st));
  PIGAN_TRY(pack_generator(e, gp, true, st));
  PIGAN_CUDA_OK(cudaMemsetAsync(e->bn_sums, 0, (2 * G.H1 + 2 * G.H2) * sizeof(float), st));
  PIGAN_CUDA_OK(cudaMemsetAsync(e->bn_bwd_sums, 0, (2 * G.H1 + 2 * G.H2) * sizeof(float), st));
  PIGAN_CUDA_OK(cudaMemsetAsync(g_grads, 0, G.total * sizeof(float), st));
  PIGAN_TRY(g_layer1(e, n, st, true));
  g_bn_stats(e, 1, n, st);
  g_bn_finalize(e, 1, gp, e->g_beff, nullptr, nullptr, (double)n, 1, st);
  PIGAN_TRY(g_layer2(e, gp, n, st, true));
  g_bn_stats(e, 2, n, st);
  g_bn_finalize(e, 2, gp, gp + G.b2, nullptr, nullptr, (double)n, 1, st);
  launch_g_head_fwd(e->g_h2, e->scale2, e->bias2, gp + G.w3, gp + G.b3, e->p, nullptr, e->xc, nullptr, n, G.H2, kKp,
                    G.S, st);
  // ---- backward: head + BatchNorm-2, layer 2, BatchNorm-1, layer 1 (train_phase cases 3-5 without the loss terms)
  GHeadBwdArgs hb;
  hb.p = e->p; hb.dpden = nullptr; hb.dp_lc = nullptr; hb.dp_extra = grad_p; hb.gs = grad_scale; hb.range_mult = 0.f;
  hb.h2 = e->g_h2; hb.scale = e->scale2; hb.bias = e->bias2; hb.mean = e->mean2; hb.rstd = e->rstd2;
  hb.w3 = gp + G.w3; hb.dy2 = e->g_dy2; hb.dw3 = g_grads + G.w3; hb.db3 = g_grads + G.b3;
  hb.sum_dy = e->bn_bwd_sums; hb.sum_dyx = e->bn_bwd_sums + G.H2; hb.range_sum = nullptr;
  hb.inv_gs = inv_gs; hb.rows = n; hb.C = G.H2;
  hb.gamma = gp + G.bn2_w; hb.dbias = g_grads + G.b2; hb.dgamma = g_grads + G.bn2_w; hb.dbeta = g_grads + G.bn2_b;
  hb.inv_n = 1.0 / (double)n; hb.part = e->partials; hb.dpre = e->dpre; hb.dpre_part = nullptr;
  launch_g_head_bwd(hb, false, st);
  launch_g_head_bwd(hb, true, st);
  PIGAN_TRY(weight_grad(e->g_dy2, n, G.H2, e->g_a1, n, G.H1, g_grads + G.w2, G.H1, G.H1, inv_gs, -1, nullptr, 0, nullptr,
                        0, e->dw_part, st));
  PIGAN_TRY((linear_store<false, false, false>(e->g_dy2, n, G.H2, e->g_w2th, G.H1, nullptr, e->g_da1, nullptr, st)));
  launch_bn_bwd_stats(e->g_da1, e->g_h1, e->scale1, e->bias1, e->mean1, e->rstd1, e->bn_bwd_sums + 2 * G.H2,
                      e->bn_bwd_sums + 2 * G.H2 + G.H1, n, G.H1, e->partials, st);
  BnBwdArgs bb;
  bb.dy = e->g_da1; bb.h = e->g_h1; bb.relu_mask = 1;
  bb.scale = e->scale1; bb.bias = e->bias1; bb.mean = e->mean1; bb.rstd = e->rstd1; bb.gamma = gp + G.bn1_w;
  bb.sum_dy = e->bn_bwd_sums + 2 * G.H2; bb.sum_dyx = e->bn_bwd_sums + 2 * G.H2 + G.H1;
  bb.dh = e->g_da1; bb.dbias = nullptr; bb.dgamma = g_grads + G.bn1_w; bb.dbeta = g_grads + G.bn1_b;
  bb.inv_n = 1.0 / (double)n; bb.inv_gs = inv_gs; bb.rows = n; bb.C = G.H1; bb.part = e->partials;
  launch_bn_bwd_apply(bb, st);
  PIGAN_TRY(weight_grad(e->g_da1, n, G.H1, e->xc, n, kKp, g_grads + G.w1, G.S, G.S, inv_gs, G.S + G.P, g_grads + G.b1, 0,
                        nullptr, 0, e->dw_part, st, 0, e->cvec, G.S, 0));
  PM(nullptr);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

// Discriminator (discriminator.py:30-39): for out = D(x, params) [n] and an upstream gradient grad_out = dL/d(out),
// d_grads[D.total] = dL/d(D's parameters) (overwritten) and grad_params[n,4] = dL/d(params) (may be null).
// No gradient with respect to the spectrum x is produced (nothing in the reference asks for it).
extern "C" int pigan_discriminator_backward(PiganEngine* e, const float* dp, const float* x, const float* params,
                                            int64_t n, const float* grad_out, float grad_scale, float* d_grads,
                                            float* grad_params, void* stream) {
  PIGAN_CHECK_ARG(e && dp && x && params && grad_out && d_grads && n >= 1 && n <= e->max_batch && grad_scale > 0.f);
  PIGAN_TRY(need_full(e));
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_grads) & 15u) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DiscLayout& D = e->dl;
  const float inv_gs = 1.0f / grad_scale;
  // ---- forward: z1 + sign mask, z2, probabilities (into the dlogit buffer)
  PIGAN_TRY(prep_spectrum(e, x, params, n, st));
  PIGAN_TRY(pack_discriminator(e, dp, true, st));
  PIGAN_CUDA_OK(cudaMemsetAsync(d_grads, 0, D.total * sizeof(float), st));
  PIGAN_CUDA_OK(cudaMemsetAsync(e->dpden, 0, (size_t)n * 4 * sizeof(float), st));
  PIGAN_TRY(d_layer1(e, n, 0, false, st));
  DL2Opts o{n, n, 1.0f, 1.0f, 0, 0, (double)n, nullptr, e->dlogit, true, false};
  PIGAN_TRY(d_layer2(e, dp, o, st));
  // dlogit = s * grad_out * p * (1 - p)   (sigmoid backward)
  launch_sigmoid_bwd(e->dlogit, grad_out, grad_scale, n, st);
  // ---- backward (train_phase case 2 / 3 for one block of rows)
  launch_d_l2_bwd(e->d_z2, e->dlogit, dp + D.w3, e->d_dh2, d_grads + D.w3, d_grads + D.b2, d_grads + D.b3, n, D.H2,
                  inv_gs, e->partials, st);
  {
    using Epi = EpiLeakyMaskStore<CfgSR>;
    Epi::Params ep;
    PIGAN_TRY(out_map(&ep.out, e->d_dh1, n, D.H1, D.H1));
    ep.mask = e->d_mask1;
    ep.mask_words = D.H1 / 32;
    PIGAN_TRY((run_tn<CfgSR, Epi>(ep, e->d_dh2, n, D.H2, D.H2, e->d_w2th, D.H1, D.H2, st)));
  }
  PIGAN_TRY(weight_grad(e->d_dh2, n, D.H2, e->d_z1, n, D.H1, d_grads + D.w2, D.H1, D.H1, inv_gs, -1, nullptr, 0, nullptr,
                        0, e->dw_part, st));
  PIGAN_TRY(weight_grad(e->d_dh1, n, D.H1, e->xc, n, kKp, d_grads + D.w1, D.IN, D.IN, inv_gs, D.IN, d_grads + D.b1, 0,
                        nullptr, 0, e->dw_part, st, 0, e->cvec, D.S, D.P));
  if (grad_params != nullptr) {
    using Epi = EpiDiscParamGrad<CfgP>;
    Epi::Params ep{e->d_mask1, D.H1 / 32, e->d_wp, e->dpden};
    PIGAN_TRY((run_tn<CfgP, Epi>(ep, e->d_dh2, n, D.H2, D.H2, e->d_w2th, D.H1, D.H2, st)));
    launch_scale_copy(e->dpden, grad_params, inv_gs, n * 4, st);
  }
  PM(nullptr);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_forward_model_forward(PiganEngine* e, const float* p_norm, int64_t n, float* out,
                                           void* stream) {
  PIGAN_CHECK_ARG(e && p_norm && out && n >= 1 && n <= e->max_batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FOutOpts fo{0, nullptr, nullptr, nullptr, nullptr, 0.f, out, nullptr, 0, 1};
  PIGAN_TRY(f_forward(e, p_norm, n, fo, st));
  PM(nullptr);
  return PIGAN_OK;
}

extern "C" int pigan_train_step_phase(PiganEngine* e, const PiganTrainArgs* a, int32_t phase, void* stream) {
  PIGAN_TRY(check_train_args(e, a));
  return train_phase(e, *a, phase, static_cast<cudaStream_t>(stream));
}

extern "C" int pigan_train_step(PiganEngine* e, const PiganTrainArgs* a, void* stream) {
  PIGAN_TRY(check_train_args(e, a));
  static const bool fuse = [] { const char* v = getenv("PIGAN_FUSE_BN_FINALIZE"); return !(v && v[0] == '0'); }();
  e->whole_step = fuse;
  e->bn_pending_blocks = 0;
  int rc = PIGAN_OK;
  for (int ph = 0; ph <= 6 && rc == PIGAN_OK; ++ph) rc = train_phase(e, *a, ph, static_cast<cudaStream_t>(stream));
  e->whole_step = false;
  e->bn_pending_blocks = 0;
  return rc;
}

extern "C" int pigan_engine_profile_begin(PiganEngine* e, const char* sections_csv) {
  PIGAN_CHECK_ARG(e != nullptr);
  Prof& p = e->prof;
  p.on = true;
  p.recs.clear();
  p.used = 0;
  p.open_slot = -1;
  p.wanted.clear();
  p.all = (sections_csv == nullptr || sections_csv[0] == 0);
  if (!p.all) {
    std::string cur;
    for (const char* c = sections_csv;; ++c) {
      if (*c == ',' || *c == 0) {
        if (!cur.empty()) p.wanted.push_back(cur);
        cur.clear();
        if (*c == 0) break;
      } else {
        cur.push_back(*c);
      }
    }
  }
  return PIGAN_OK;
}

extern "C" int pigan_engine_profile_end(PiganEngine* e, char* report, size_t report_bytes) {
  PIGAN_CHECK_ARG(e != nullptr && report != nullptr && report_bytes > 0);
  Prof& p = e->prof;
  p.on = false;
  p.open_slot = -1;
  std::vector<double> ms(p.names.size(), 0.0);
  std::vector<long long> cnt(p.names.size(), 0);
  for (const auto& r : p.recs) {
    PIGAN_CUDA_OK(cudaEventSynchronize(r.b));
    float t = 0.f;
    PIGAN_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.slot] += t;
    cnt[r.slot] += 1;
  }
  p.recs.clear();
  p.used = 0;
  std::string out;
  char line[160];
  for (size_t i = 0; i < p.names.size(); ++i) {
    if (cnt[i] == 0) continue;
    snprintf(line, sizeof(line), "%s %lld %.6f\n", p.names[i], cnt[i], ms[i]);
    out += line;
  }
  if (out.size() + 1 > report_bytes) return fail(PIGAN_ERR_WORKSPACE, "profile report needs %zu bytes", out.size() + 1);
  memcpy(report, out.c_str(), out.size() + 1);
  return PIGAN_OK;
}

extern "C" float* pigan_engine_generator_output(PiganEngine* e) { return e ? e->p : nullptr; }
extern "C" float* pigan_engine_bn_sums(PiganEngine* e) { return e ? e->bn_sums : nullptr; }
extern "C" float* pigan_engine_bn_bwd_sums(PiganEngine* e) { return e ? e->bn_bwd_sums : nullptr; }
extern "C" double* pigan_engine_loss_sums(PiganEngine* e) { return e ? e->sums : nullptr; }

extern "C" int pigan_score_candidates(PiganEngine* e, const float* gp, const float* bn, const float* spectra,
                                      const float* target, const float* noise, float sigma, int64_t n,
                                      float* out_p, int32_t* out_viol, float* out_err, float* out_cons,
                                      void* stream) {
  PIGAN_CHECK_ARG(e && gp && bn && n >= 1 && n <= e->max_batch);
  PIGAN_TRY(need_full(e));
  PIGAN_CHECK_ARG((spectra != nullptr) != (target != nullptr && noise != nullptr));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GenLayout& G = e->gl;
  if (spectra) {
    PIGAN_TRY(prep_spectrum(e, spectra, nullptr, n, st));
  } else {
    // centre on the design target: the candidates differ from it by sigma * noise only
    PM("prep_cast");
    launch_copy_pad_f32(target, G.S, e->cvec, kKp, st);
    e->xc = e->xc_own;
    launch_cast_center_noise(target, noise, sigma, e->cvec, e->xc, nullptr, n, G.S, G.P, kKp, st);
  }
  PIGAN_TRY(g_eval_setup(e, gp, bn, st));
  float* p = out_p ? out_p : e->p;
  PIGAN_TRY(g_eval_forward(e, gp, n, p, st, true));
  float* err = out_err ? out_err : e->row_err;
  // per-candidate error against its own spectrum (given spectra) or against the design target (cvec = target)
  FOutOpts fo{spectra ? 2 : 1, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, err, 0, 1};
  PIGAN_TRY(f_forward(e, p, n, fo, st, fused_head(e)));
  PM("small");
  if (out_viol || out_cons) launch_score_finish(p, err, n, G.P, out_viol, out_cons, st);
  PM(nullptr);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

// Model-validation loop body (core/evaluate/unified_evaluator.py:439-468) in one call: cycle-consistency error
// mean((x - F(G(x)).spectrum)^2), prediction stability mean((G(x) - G(x + sigma * noise))^2) and the plausibility
// score mean(sigmoid(10 p - 5)) per row; the noise tensor is passed explicitly.
extern "C" int pigan_validate_model(PiganEngine* e, const float* gp, const float* bn, const float* spectra,
                                    const float* noise, float sigma, int64_t n, float* out_p, float* out_cycle_error,
                                    float* out_stability, float* out_plausibility, void* stream) {
  PIGAN_CHECK_ARG(e && gp && bn && spectra && noise && n >= 1 && n <= e->max_batch);
  PIGAN_TRY(need_full(e));
  PIGAN_CHECK_ARG(out_cycle_error && out_stability && out_plausibility);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GenLayout& G = e->gl;
  PIGAN_TRY(prep_spectrum(e, spectra, nullptr, n, st));
  PIGAN_TRY(g_eval_setup(e, gp, bn, st));
  float* p = out_p ? out_p : e->p;
  PIGAN_TRY(g_eval_forward(e, gp, n, p, st, true));
  FOutOpts fo{2, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, out_cycle_error, 0, 1};
  PIGAN_TRY(f_forward(e, p, n, fo, st, fused_head(e)));
  // the same generator on x + sigma * noise (same centring row, same packed weights).  Measured: feeding the noise as
  // a second operand against sigma * W1 (identical rounding of the x part in both passes) does not help - the score
  // is a difference of two outputs ~1e-3 apart and the fp16 rounding of the stored activations of the two passes
  // (2e-4 each) limits a row's value to ~20 %; means over rows agree to 1 % (tests/test_gpu_engine.py)
  PM("prep_cast");
  launch_cast_center_noise(spectra, noise, sigma, e->cvec, e->xc, nullptr, n, G.S, G.P, kKp, st, G.S);
  PIGAN_TRY(g_eval_forward(e, gp, n, e->pden, st, false));
  PM("small");
  launch_validation_scores(p, e->pden, n, G.P, out_stability, out_plausibility, st);
  PM(nullptr);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

// ------------------------------------------------------------------------------------------ surrogate training
// One optimiser step of pretrain_forward_model (core/train/pretrain_fwd_model.py:68-92): F in train
// mode (Dropout 0.2, counter based), MSE(spectrum) + MSE(metrics), backward through all six layers, clip 1.0, Adam.
//   phase 0: zero grads, pack weights, forward, loss, backward -> f_grads (local rows, already divided by the
//            global batch) and loss_sums (local sums of squares)          [data parallel: all-reduce both here]
//   phase 1: clip_grad_norm_ + Adam + the three reported losses
// Activations a_i reuse the engine's f_a1..f_a5 and packed-weight buffers, so the frozen-surrogate state of the
// engine is invalidated: call pigan_engine_load_forward_model again before PI-GAN steps or scoring.
namespace {
struct FTrainWs {
  __half* xhat[5];
  float* rstd[5];
  uint8_t* keepbits[5];   // Dropout keep-masks of the forward pass, 1 bit per element
  float* out32;
  __half* dout;
  __half* dbuf[2];
  __half* wth[6];     // transposed fp16 weights: [in_i, out_i] (layer 5: [256, dout_ld])
  float* grads_scratch;   // [L.total] parameter-gradient sink of pigan_forward_model_input_grad (weights frozen)
  float* dw1_tmp;     // [4][256]
  double* sumsq;
  float* loss_sums;   // [2] (used when the caller passes none)
  uint8_t* zero_from; // dw1_tmp .. loss_sums are cleared every step
  size_t zero_bytes;
  size_t carve(void* base, const FwdLayout& L, int64_t B, bool wide = false) {
    Carver c(base);
    const int64_t Bp = round_up(B, 128);
    const int dout_ld = (int)round_up(L.OUT, 64);   // 320 at the reference widths
    int wmax = 0;
    for (int i = 0; i < 5; ++i) {
      xhat[i] = c.take<__half>((size_t)Bp * L.H[i]);
      rstd[i] = c.take<float>(Bp);
      keepbits[i] = c.take<uint8_t>((size_t)Bp * (L.H[i] / 8));
      wmax = L.H[i] > wmax ? L.H[i] : wmax;
    }
    out32 = c.take<float>(wide ? 0 : (size_t)Bp * L.OUT);   // widened path: the engine's accumulator slabs instead
    dout = c.take<__half>((size_t)Bp * dout_ld);
    dbuf[0] = c.take<__half>((size_t)Bp * wmax);
    dbuf[1] = c.take<__half>((size_t)Bp * wmax);
    wth[0] = nullptr;
    for (int i = 1; i < 5; ++i) wth[i] = c.take<__half>((size_t)L.H[i - 1] * L.H[i]);
    wth[5] = c.take<__half>((size_t)L.H[4] * dout_ld);
    grads_scratch = c.take<float>((size_t)L.total);
    dw1_tmp = c.take<float>(4 * L.H[0]);
    zero_from = reinterpret_cast<uint8_t*>(dw1_tmp);
    sumsq = c.take<double>(1);
    loss_sums = c.take<float>(2);
    zero_bytes = base ? (size_t)(reinterpret_cast<uint8_t*>(loss_sums + 2) - zero_from) : 0;
    return (c.off + 255) & ~size_t(255);
  }
};

// input_grad: weights frozen (A19) — no weight-gradient GEMMs, the small parameter gradients go to a scratch buffer,
// the first layer keeps dh and the gradient with respect to params_norm is written to dp_out
struct FwdRunOpts {
  bool input_grad = false;
  float w_spec = 1.f, w_met = 1.f;
  float* dp_out = nullptr;
  const float* upstream = nullptr;   // input_grad: dL/d(output) [n, S+Mt] instead of the two MSE terms
};
int fwd_train_phase(PiganEngine* e, const PiganFwdTrainArgs& a, int phase, FTrainWs& w, cudaStream_t st,
                    const FwdRunOpts& opt = FwdRunOpts()) {
  const FwdLayout& L = e->fl;
  const int64_t n = a.batch;
  const int dout_ld = e->dout_ld;
  const bool wide = !e->full;   // widened dims: generic first-layer / loss kernels, output layer through the slabs
  float* fp = a.f_params;
  float* gr = opt.input_grad ? w.grads_scratch : a.f_grads;
  float* loss_sums = a.loss_sums ? a.loss_sums : w.loss_sums;
  const float inv_gs = (float)(1.0 / (double)a.global_batch);
  if (phase == 1) {
    PM("clip_adam");
    double* sq = reinterpret_cast<double*>(e->partials);
    const int nsq = launch_sumsq(gr, L.total, sq, st);
    AdamArgs ad{fp, gr, a.f_exp_avg, a.f_exp_avg_sq, L.total, a.lr, a.beta1, a.beta2, a.eps,
                1.0 - pow((double)a.beta1, (double)a.step), 1.0 - pow((double)a.beta2, (double)a.step), sq, nsq,
                a.max_norm};
    launch_clip_adam(ad, st);
    launch_f_train_losses(loss_sums, (double)a.global_batch * L.S, (double)a.global_batch * L.Mt, a.losses, st);
    PM(nullptr);
    PIGAN_CUDA_OK(cudaGetLastError());
    return PIGAN_OK;
  }
  DropoutArgs dr;
  dr.seed = a.dropout_seed;
  dr.first_row = a.first_row;
  dr.step = (unsigned int)a.step;
  dr.thresh16 = (unsigned int)lround((double)a.dropout_p * 65536.0);
  dr.keep_scale = 1.0f / (1.0f - a.dropout_p);
  if (fp != e->f_params) e->f_loaded = false;   // the packed weights below replace the frozen surrogate's
  PM("memset");
  if (((reinterpret_cast<uintptr_t>(gr) | reinterpret_cast<uintptr_t>(a.loss_sums)) & 15u) == 0 &&
      w.zero_bytes % sizeof(float) == 0) {
    // one kernel instead of four memset stream operations
    float* ptrs[4] = {gr, reinterpret_cast<float*>(w.zero_from), reinterpret_cast<float*>(w.wth[5]), a.loss_sums};
    const int64_t nf[4] = {L.total, (int64_t)(w.zero_bytes / sizeof(float)), (int64_t)L.H[4] * dout_ld / 2,
                           a.loss_sums ? 2 : 0};
    launch_zero_buffers(ptrs, nf, 4, st);
  } else {
    PIGAN_CUDA_OK(cudaMemsetAsync(gr, 0, (size_t)L.total * sizeof(float), st));
    PIGAN_CUDA_OK(cudaMemsetAsync(w.zero_from, 0, w.zero_bytes, st));
    if (a.loss_sums) PIGAN_CUDA_OK(cudaMemsetAsync(a.loss_sums, 0, 2 * sizeof(float), st));
    PIGAN_CUDA_OK(cudaMemsetAsync(w.wth[5], 0, (size_t)L.H[4] * dout_ld * sizeof(__half), st));
  }
  PM("pack_weights");
  for (int i = 1; i < 6; ++i) {
    const int in = L.H[i - 1], out = i < 5 ? L.H[i] : L.OUT;
    launch_cast_pad(fp + L.w[i], in, in, e->f_wh[i], in, out, st);
    launch_transpose_cast(fp + L.w[i], out, in, in, w.wth[i], i < 5 ? out : dout_ld, st);
  }
  launch_copy_pad_f32(fp + L.b[5], L.OUT, e->f_bias_out, dout_ld > 288 ? dout_ld : 288, st);
  // ---- forward (train mode)
  __half* act[5] = {e->f_a1, e->f_a2, e->f_a3, e->f_a4, e->f_a5};
  size_t moff[5];
  int hsum = 0;
  for (int i = 0; i < 5; ++i) {
    moff[i] = (size_t)hsum * (size_t)n;
    hsum += L.H[i];
  }
  auto mask = [&](int i) { return a.mask_dump ? a.mask_dump + moff[i] : nullptr; };
  PM("f_l1");
  if (wide)
    launch_f_l1_wide(a.params_norm, fp + L.w[0], fp + L.b[0], fp + L.ln_w[0], fp + L.ln_b[0], e->f_l1c, w.xhat[0],
                     act[0], w.rstd[0], mask(0), w.keepbits[0], n, L.H[0], &dr, st);
  else
    launch_f_l1_train(a.params_norm, fp + L.w[0], fp + L.b[0], fp + L.ln_w[0], fp + L.ln_b[0], w.xhat[0], act[0],
                      w.rstd[0], mask(0), w.keepbits[0], n, dr, st);
  for (int i = 1; i < 5; ++i) {
    PM("f_hidden_gemm");
    PIGAN_TRY((linear_store<true, false, true>(act[i - 1], n, L.H[i - 1], e->f_wh[i], L.H[i], fp + L.b[i], w.xhat[i],
                                               e->f_rowstats, st)));
    PM("f_ln_dropout");
    launch_ln_train(w.xhat[i], e->f_rowstats, fp + L.ln_w[i], fp + L.ln_b[i], act[i], w.rstd[i], mask(i), w.keepbits[i],
                    n, L.H[i], i, dr, st);
  }
  PM("f_out_gemm");
  if (wide) {
    if (opt.upstream == nullptr) PIGAN_TRY(f_out_slab(e, act[4], n, e->f_wh[5], st));   // a VJP does not need the output
  } else {
    FOutOpts fo{0, nullptr, nullptr, nullptr, nullptr, 0.f, w.out32, nullptr, 0, 1};
    PIGAN_TRY(f_out_layer(e, act[4], n, fo, st));
  }
  PM("f_out_loss");
  if (opt.upstream != nullptr)
    launch_f_upstream_cast(opt.upstream, L.OUT, w.dout, dout_ld, n, (float)a.global_batch, st);
  else if (wide)
    launch_f_out_loss_slab(e->f_slab, e->out_groups, fp + L.b[5], a.spectrum, a.metrics_norm, w.dout, dout_ld, n, L.S,
                           L.Mt, e->partials, gr + L.b[5], loss_sums, inv_gs, st, opt.w_spec, opt.w_met);
  else
    launch_f_out_loss(w.out32, a.spectrum, a.metrics_norm, w.dout, dout_ld, n, L.S, L.Mt, e->partials, gr + L.b[5],
                      loss_sums, inv_gs, st, opt.w_spec, opt.w_met);
  // ---- backward
  PM("f_wgrad_gemm");
  if (!opt.input_grad)
    PIGAN_TRY(weight_grad(w.dout, n, L.OUT, act[4], n, L.H[4], gr + L.w[5], L.H[4], L.H[4], inv_gs, -1, nullptr, 0,
                          nullptr, 0, e->dw_part, st, dout_ld));
  PM("f_dgrad_gemm");
  __half* d = w.dbuf[0];
  __half* d2 = w.dbuf[1];
  PIGAN_TRY((linear_store<false, false, false>(w.dout, n, dout_ld, w.wth[5], L.H[4], nullptr, d, nullptr, st)));
  for (int i = 4; i >= 0; --i) {
    PM("f_ln_bwd");
    // (widened dims: the first layer goes through the generic kernel too - dh stays in place - and its thin
    // weight / input gradients come from their own streaming kernels)
    launch_ln_bwd(d, w.xhat[i], w.rstd[i], fp + L.ln_w[i], fp + L.ln_b[i], (i == 0 && !wide) ? a.params_norm : nullptr,
                  w.keepbits[i], n, L.H[i], dr.keep_scale, e->partials, gr + L.ln_w[i], gr + L.ln_b[i], gr + L.b[i],
                  w.dw1_tmp, inv_gs, st, (i == 0 && opt.input_grad) ? 1 : 0);
    if (i == 0) {
      if (wide) {
        if (opt.input_grad) launch_f_dp_wide(d, fp + L.w[0], opt.dp_out, n, L.H[0], inv_gs, st);
        else launch_f_dw1_wide(d, a.params_norm, n, L.H[0], e->partials, w.dw1_tmp, gr + L.w[0], inv_gs, st);
      } else if (opt.input_grad) {
        launch_f_dp(d, fp + L.w[0], opt.dp_out, n, inv_gs, st);
      } else {
        launch_f_dw1_transpose(w.dw1_tmp, gr + L.w[0], st);
      }
      break;
    }
    PM("f_wgrad_gemm");
    if (!opt.input_grad)
      PIGAN_TRY(weight_grad(d, n, L.H[i], act[i - 1], n, L.H[i - 1], gr + L.w[i], L.H[i - 1], L.H[i - 1], inv_gs, -1,
                            nullptr, 0, nullptr, 0, e->dw_part, st));
    PM("f_dgrad_gemm");
    PIGAN_TRY((linear_store<false, false, false>(d, n, L.H[i], w.wth[i], L.H[i - 1], nullptr, d2, nullptr, st)));
    __half* t = d;
    d = d2;
    d2 = t;
  }
  PM(nullptr);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

int check_fwd_train_args(const PiganEngine* e, const PiganFwdTrainArgs* a, const void* ws, size_t ws_bytes) {
  PIGAN_CHECK_ARG(e != nullptr && a != nullptr && ws != nullptr);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255u) == 0);
  PIGAN_CHECK_ARG(a->batch >= 1 && a->batch <= e->max_batch && a->global_batch >= a->batch && a->first_row >= 0);
  PIGAN_CHECK_ARG(a->params_norm && a->spectrum && a->metrics_norm && a->losses);
  PIGAN_CHECK_ARG(a->f_params && a->f_grads && a->f_exp_avg && a->f_exp_avg_sq);
  PIGAN_CHECK_ARG(a->step >= 1 && a->dropout_p >= 0.f && a->dropout_p < 1.f);
  PIGAN_CHECK_ARG(a->beta1 >= 0.f && a->beta1 < 1.f && a->beta2 >= 0.f && a->beta2 < 1.f && a->max_norm > 0.f);
  FTrainWs w;
  const size_t need = w.carve(nullptr, e->fl, e->max_batch, !e->full);
  if (ws_bytes < need) return fail(PIGAN_ERR_WORKSPACE, "surrogate-training workspace too small: %zu < %zu", ws_bytes, need);
  return PIGAN_OK;
}
}  // namespace

extern "C" int pigan_forward_model_input_grad(PiganEngine* e, const float* f_params, const float* params_norm,
                                              const float* spectrum, const float* metrics_norm, int64_t n,
                                              float w_spectrum, float w_metrics, float* out_dp, float* out_losses,
                                              void* workspace, size_t workspace_bytes, void* stream) {
  PIGAN_CHECK_ARG(e && f_params && params_norm && spectrum && metrics_norm && out_dp && out_losses && workspace);
  PIGAN_CHECK_ARG(n >= 1 && n <= e->max_batch && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(out_dp) & 15u) == 0);
  FTrainWs w;
  const size_t need = w.carve(nullptr, e->fl, e->max_batch, !e->full);
  if (workspace_bytes < need)
    return fail(PIGAN_ERR_WORKSPACE, "surrogate-training workspace too small: %zu < %zu", workspace_bytes, need);
  w.carve(workspace, e->fl, e->max_batch, !e->full);
  PiganFwdTrainArgs a;
  memset(&a, 0, sizeof(a));
  a.params_norm = params_norm; a.spectrum = spectrum; a.metrics_norm = metrics_norm;
  a.batch = a.global_batch = n;
  a.f_params = const_cast<float*>(f_params);   // read only on this path
  a.step = 1;
  a.dropout_p = 0.f;                           // eval mode: Dropout is the identity
  a.losses = out_losses;
  FwdRunOpts opt;
  opt.input_grad = true;
  opt.w_spec = w_spectrum;
  opt.w_met = w_metrics;
  opt.dp_out = out_dp;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PIGAN_TRY(fwd_train_phase(e, a, 0, w, st, opt));
  // out_losses[0..1] = unweighted MSE(spectrum), MSE(metrics) of this batch
  launch_f_input_grad_losses(w.loss_sums, (double)n * e->fl.S, (double)n * e->fl.Mt, out_losses, st);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_forward_model_vjp(PiganEngine* e, const float* f_params, const float* params_norm,
                                       const float* grad_out, int64_t n, float* out_dp, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  PIGAN_CHECK_ARG(e && f_params && params_norm && grad_out && out_dp && workspace);
  PIGAN_CHECK_ARG(n >= 1 && n <= e->max_batch && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(out_dp) & 15u) == 0);
  FTrainWs w;
  const size_t need = w.carve(nullptr, e->fl, e->max_batch, !e->full);
  if (workspace_bytes < need)
    return fail(PIGAN_ERR_WORKSPACE, "surrogate-training workspace too small: %zu < %zu", workspace_bytes, need);
  w.carve(workspace, e->fl, e->max_batch, !e->full);
  PiganFwdTrainArgs a;
  memset(&a, 0, sizeof(a));
  a.params_norm = params_norm;
  a.batch = a.global_batch = n;
  a.f_params = const_cast<float*>(f_params);   // read only on this path
  a.step = 1;
  a.dropout_p = 0.f;                           // eval mode: Dropout is the identity
  FwdRunOpts opt;
  opt.input_grad = true;
  opt.dp_out = out_dp;
  opt.upstream = grad_out;
  PIGAN_TRY(fwd_train_phase(e, a, 0, w, static_cast<cudaStream_t>(stream), opt));
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" size_t pigan_fwd_train_workspace_bytes(const PiganEngine* e) {
  if (!e) return 0;
  FTrainWs w;
  return w.carve(nullptr, e->fl, e->max_batch, !e->full);
}

extern "C" int pigan_fwd_train_step_phase(PiganEngine* e, const PiganFwdTrainArgs* a, int32_t phase, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  PIGAN_TRY(check_fwd_train_args(e, a, workspace, workspace_bytes));
  if (phase < 0 || phase > 1) return fail(PIGAN_ERR_INVALID, "surrogate-training phase %d out of range", phase);
  FTrainWs w;
  w.carve(workspace, e->fl, e->max_batch, !e->full);
  return fwd_train_phase(e, *a, phase, w, static_cast<cudaStream_t>(stream));
}

extern "C" int pigan_fwd_train_step(PiganEngine* e, const PiganFwdTrainArgs* a, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  PIGAN_TRY(check_fwd_train_args(e, a, workspace, workspace_bytes));
  FTrainWs w;
  w.carve(workspace, e->fl, e->max_batch, !e->full);
  for (int ph = 0; ph <= 1; ++ph) PIGAN_TRY(fwd_train_phase(e, *a, ph, w, static_cast<cudaStream_t>(stream)));
  return PIGAN_OK;
}

// ------------------------------------------------------------------------------------------ inverse-design search
namespace {
constexpr int kSearchGroup = 16;  // chunks scored between two top-k merges
struct SearchWs {
  float* scores;      // [k + cap]   [0,k) best so far, [k, ...) this group's reconstruction errors
  float* params;      // [(k + cap) * 4]
  int64_t* iota;      // [k + cap]
  int64_t* best_idx;  // [k]
  float* sel_scores;  // [k]
  int64_t* sel_pos;   // [k]
  int64_t* tmp_idx;   // [k]
  float* tmp_params;  // [4k]
  void* topk_ws;
  size_t topk_bytes;
  size_t carve(void* base, int64_t cap, int k) {
    Carver c(base);
    scores = c.take<float>(k + cap);
    params = c.take<float>((size_t)(k + cap) * 4);
    iota = c.take<int64_t>(k + cap);
    best_idx = c.take<int64_t>(k);
    sel_scores = c.take<float>(k);
    sel_pos = c.take<int64_t>(k);
    tmp_idx = c.take<int64_t>(k);
    tmp_params = c.take<float>((size_t)k * 4);
    topk_bytes = pigan_topk_workspace_bytes(k + cap, k);
    topk_ws = c.take<uint8_t>(topk_bytes);
    return (c.off + 255) & ~size_t(255);
  }
};
}  // namespace

extern "C" size_t pigan_search_workspace_bytes(const PiganEngine* e, int32_t k) {
  if (!e || k < 1 || k > 4096) return 0;
  SearchWs w;
  return w.carve(nullptr, (int64_t)kSearchGroup * e->max_batch, k);
}

extern "C" int pigan_inverse_design_search(PiganEngine* e, const float* gp, const float* bn, const float* target,
                                           float sigma, uint64_t seed, int64_t first_candidate, int64_t count,
                                           int32_t k, float* out_scores, int64_t* out_indices, float* out_params,
                                           float* noise_dump, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  PIGAN_CHECK_ARG(e && gp && bn && target && out_scores && out_indices && out_params && workspace);
  PIGAN_TRY(need_full(e));
  PIGAN_CHECK_ARG(k >= 1 && k <= 4096 && count >= 0 && first_candidate >= 0);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0);
  if (!e->f_loaded) return fail(PIGAN_ERR_INVALID, "forward model not loaded (pigan_engine_load_forward_model)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GenLayout& G = e->gl;
  const int64_t chunk = e->max_batch;
  const int64_t cap = (int64_t)kSearchGroup * chunk;
  SearchWs w;
  const size_t need = w.carve(workspace, cap, k);
  if (workspace_bytes < need) return fail(PIGAN_ERR_WORKSPACE, "search workspace too small: %zu < %zu", workspace_bytes, need);
  e->xc = e->xc_own;
  launch_search_init(w.scores, w.iota, w.best_idx, w.params, k + cap, k, st);
  // weights and the target do not change during a search: pack / fold once
  launch_copy_pad_f32(target, G.S, e->cvec, kKp, st);   // centre on the design target
  PIGAN_TRY(g_eval_setup(e, gp, bn, st));
  for (int64_t g0 = 0; g0 < count; g0 += cap) {
    const int64_t in_group = count - g0 < cap ? count - g0 : cap;
    for (int64_t off = 0; off < in_group; off += chunk) {
      const int64_t n = in_group - off < chunk ? in_group - off : chunk;
      PM("prep_cast");
      launch_cast_center_philox(target, sigma, seed, first_candidate + g0 + off, e->cvec, e->xc,
                                noise_dump ? noise_dump + (size_t)(g0 + off) * G.S : nullptr, n, G.S, kKp, st);
      float* p = w.params + (size_t)(k + off) * 4;
      PIGAN_TRY(g_eval_forward(e, gp, n, p, st, true));
      FOutOpts fo{1, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, w.scores + k + off, 0, 1};
      PIGAN_TRY(f_forward(e, p, n, fo, st, fused_head(e)));
    }
    PM("topk");
    launch_fill_inf(w.scores + k + in_group, cap - in_group, st);
    PIGAN_TRY(pigan_topk_smallest(w.scores, w.iota, k + cap, k, 0, w.sel_scores, w.sel_pos, w.topk_ws, w.topk_bytes,
                                  stream));
    launch_search_gather(w.sel_pos, w.params, w.best_idx, first_candidate + g0, k, w.tmp_idx, w.tmp_params, st);
    launch_search_commit(w.sel_scores, w.tmp_idx, w.tmp_params, k, w.scores, w.best_idx, w.params, st);
  }
  PM(nullptr);
  PIGAN_CUDA_OK(cudaMemcpyAsync(out_scores, w.scores, (size_t)k * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PIGAN_CUDA_OK(cudaMemcpyAsync(out_indices, w.best_idx, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  PIGAN_CUDA_OK(cudaMemcpyAsync(out_params, w.params, (size_t)k * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_engine_trace_layernorm(void* device_buffer) {
  pigan::g_ln_trace = static_cast<long long*>(device_buffer);
  return PIGAN_OK;
}
