// Fused epilogues of the tcgen05 GEMM.  One thread owns one output row of the 128-row tile and walks
// its accumulator columns out of TMEM in blocks of 32 (the next block's load in flight); fp16 outputs are staged
// per warp through swizzled shared memory and written with TMA stores (coalesced, clipped at the tensor bounds).
//
// Measured on B200 these K<=1024 layers are bound by epilogue issue slots, not by the tensor pipe, so
// the epilogues are kept to a few instructions per element: per-column constants are fetched with
// uniform 128-bit loads from zero-padded arrays (no bounds checks), reductions over the batch (column
// sums) are left to the streaming kernels in elementwise.cu where a thread owns a column, and the first
// layers' bias / parameter columns ride inside the MMA (spare K columns of the spectrum tile).
//
// Gradient tensors stored in fp16 are pre-multiplied by the gradient scale GS (= global batch size) so
// 1/B-sized values stay in fp16's normal range; parameter gradients are un-scaled when accumulated.
#pragma once
#include "gemm_launch.cuh"

namespace pigan {

constexpr int kStageBytes = 16384;  // one [128 x 64] fp16 sub-tile (TMA-loaded target tiles of EpiFwdOut, fp32 dW boxes)
constexpr float kLeaky = 0.2f;

__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, kLeaky * x); }  // valid for slope < 1
__device__ __forceinline__ float lrelu_slope_from_out(float z) { return z > 0.f ? 1.f : kLeaky; }

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 32 consecutive per-column constants (uniform across the warp) as 8 float4 loads.
__device__ __forceinline__ void load_cols32(const float* __restrict__ base, float* out) {
  const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 t = __ldg(p + k);
    out[4 * k + 0] = t.x;
    out[4 * k + 1] = t.y;
    out[4 * k + 2] = t.z;
    out[4 * k + 3] = t.w;
  }
}

// Output tensor of a store epilogue: row-major fp16 [rows, cols] with row pitch ld (elements; ld % 8 == 0, 16-byte
// aligned base)
struct OutTile {
  CUtensorMap map;   // box = 32 rows x 32 columns, 64B swizzle (host_util: make_tmap_f16_store): TMA-store variant
  __half* ptr;       // LSU-store variant
  int ld, rows, cols;
};

// Per-warp staging of fp16 output blocks.  A warp owns 32 rows of the tile (its TMEM lane quarter); it packs 32
// columns of them into a [32 x 32] fp16 block in shared memory (64-byte rows, 16-byte chunks XOR-swizzled:
// conflict-free stores with a row per thread) and the block leaves either as a TMA store issued by lane 0 or through
// a read-back with four lanes per row and 16-byte global stores.  Nothing is shared between warps: one __syncwarp
// per block, no named barriers.  (Round 1 staged [128 x 64] blocks per epilogue GROUP behind two bar.sync each.)
// Measured (profiles/r02_time_step_*.txt): TMA stores win for every epilogue, by 3-8 % of the kernel - the read-back
// costs shared-memory cycles the operand ring needs; the LSU variant is kept for comparison (PIGAN_LN_TMA=0).
constexpr int kWarpBlockBytes = 2048;
constexpr int kEpiStagingBytes = 4 * 2 * kWarpBlockBytes;   // per group of 4 warps, two blocks per warp

// TMA = true : lane 0 issues a TMA store of the block; the buffer is reused once the TMA unit has read it (NBUF blocks
//              alternate).  Best for epilogues with slack (plain Linear / dX stores): nothing but the pack and four
//              16-byte shared-memory stores per block on the warp's critical path.
// TMA = false: the warp reads the block back (four lanes per row) and writes it with 16-byte global stores.  For
//              epilogues that must not wait for the TMA unit (see above).
template <int NBUF, bool TMA = true>
struct WarpStagerT {
  static_assert(NBUF == 1 || NBUF == 2, "blocks per warp");
  uint32_t cnt;
  uint32_t base;   // this warp's blocks
  __device__ __forceinline__ void init(const EpiCtx& cx) {
    cnt = 0;
    base = cx.smem + (uint32_t)(cx.tid >> 5) * (uint32_t)(NBUF * kWarpBlockBytes);
  }
  __device__ __forceinline__ void init_at(uint32_t warp_base) {
    cnt = 0;
    base = warp_base;
  }
  __device__ __forceinline__ uint32_t acquire(const EpiCtx& cx) {
    if constexpr (TMA) {
      if (cx.lane == 0) tma_store_wait_read<NBUF - 1>();  // the store that used this block NBUF commits ago has read it
      __syncwarp();
    } else {
      if (NBUF == 1) __syncwarp();   // the read-back of the previous block is done (two blocks: the sync of the block
                                     // in between already separates them)
    }
    return base + (NBUF == 2 ? (cnt & 1u) * kWarpBlockBytes : 0u);
  }
  // 32 values of this thread's row (lane = row within the warp's 32 rows)
  __device__ __forceinline__ static void put32(uint32_t buf, int lane, const float* v) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t addr = buf + (uint32_t)lane * 64u + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                   "r"(pack_half2(v[8 * j + 0], v[8 * j + 1])), "r"(pack_half2(v[8 * j + 2], v[8 * j + 3])),
                   "r"(pack_half2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_half2(v[8 * j + 6], v[8 * j + 7]))
                   : "memory");
    }
  }
  // already packed: 16 words = 32 halves
  __device__ __forceinline__ static void put32_packed(uint32_t buf, int lane, const uint32_t* h) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t addr = buf + (uint32_t)lane * 64u + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[4 * j]), "r"(h[4 * j + 1]),
                   "r"(h[4 * j + 2]), "r"(h[4 * j + 3])
                   : "memory");
    }
  }
  // col0: first of the 32 columns; row0: first row of the 128-row tile
  __device__ __forceinline__ void commit(const EpiCtx& cx, uint32_t buf, const OutTile& o, int col0, int row0) {
    if constexpr (TMA) {
      fence_proxy_async_smem();
      __syncwarp();
      if (cx.lane == 0) {
        tma_store_2d(&o.map, buf, col0, row0 + cx.q * 32);
        tma_store_commit();
      }
    } else {
      __syncwarp();
      const int chunk = cx.lane & 3;
      const int col = col0 + chunk * 8;
      const int rbase = row0 + cx.q * 32 + (cx.lane >> 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = (cx.lane >> 2) + 8 * i;   // row within the warp's 32
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                     : "r"(buf + (uint32_t)rl * 64u + (uint32_t)((chunk ^ ((rl >> 1) & 3)) << 4))
                     : "memory");
        const int row = rbase + 8 * i;
        if (row < o.rows && col < o.cols)
          *reinterpret_cast<uint4*>(o.ptr + (size_t)row * o.ld + col) = make_uint4(w0, w1, w2, w3);
      }
    }
    ++cnt;
  }
  __device__ __forceinline__ static void drain(const EpiCtx& cx) {
    if constexpr (TMA) {
      if (cx.lane == 0) tma_store_wait<0>();
    }
  }
};
using WarpStager = WarpStagerT<2>;

// Walks NCOLS accumulator columns of this thread's row: one tcgen05.ld.x64 (64 columns, 8 KB per warp) at a time,
// f(c, v) is called for each 32-column block.  Loads of several warps of an SM sub-partition overlap each other and
// the other warps' arithmetic: 8 epilogue warps drain and pack a 128 x 256 tile in ~1100 cycles this way against 2650
// with .x32 loads that are waited for one by one (tools/micro/ldtm_bench2.cu, profiles/r02_ldtm_bench2.txt).  Keeping a
// second register set in flight inside one warp was tried: ptxas places both 32-register destinations on the same
// range and spills the live set (840 bytes of spill stores per thread).
// GROUP = columns per trip of the (not unrolled) loop: 64, or 128 when f collects something per 128 columns in
// registers (mask words); the loop is deliberately NOT unrolled across trips - ptxas would hoist the next trip's
// load above this trip's arithmetic and run out of registers.
// X32 = true: one tcgen05.ld.x32 per block (32 live registers instead of 64) for the 16-warp epilogues, whose threads
// have 96 registers each.
template <int NCOLS, int GROUP = 64, bool X32 = false, class F>
__device__ __forceinline__ void drain_blocks32(uint32_t tacc, F&& f) {
  static_assert(NCOLS % GROUP == 0 && GROUP % 64 == 0, "pairs of 32-column blocks");
#pragma unroll 1
  for (int c0 = 0; c0 < NCOLS; c0 += GROUP) {
    if constexpr (X32) {
#pragma unroll
      for (int c = 0; c < GROUP; c += 32) {
        float v[32];
        tmem_ld32(tacc + c0 + c, v);
        tmem_ld_wait();
        f(c0 + c, c / 32, v);
      }
    } else {
#pragma unroll
      for (int c = 0; c < GROUP; c += 64) {
        float v[64];
        tmem_ld64(tacc + c0 + c, v);
        tmem_ld_wait();
        f(c0 + c, c / 32, v);
        f(c0 + c + 32, c / 32 + 1, v + 32);
      }
    }
  }
}
// .x32 loads, fully unrolled (callers that keep per-column state in registers: the LayerNorm stash)
template <int NCOLS, class F>
__device__ __forceinline__ void drain_blocks32_unrolled(uint32_t tacc, F&& f) {
#pragma unroll
  for (int c = 0; c < NCOLS; c += 32) {
    float v[32];
    tmem_ld32(tacc + c, v);
    tmem_ld_wait();
    f(c, v);
  }
}

// =====================================================================================================
// out = fp16(act(acc + bias)).  Linear layers of G (generator.py:18,21), F (forward_model.py:30-56) and
// the plain dX products of the backward pass.  ROWSTATS also emits per-row (sum, sum of squares) of the
// stored value over this tile's columns — the LayerNorm partials the consumer merges
// (forward_model.py:31-53).  `bias` must be zero-padded to a multiple of BLOCK_N.
// =====================================================================================================
// COLSTATS (train-mode BatchNorm, generator.py:19,22): the epilogue also produces the per-column sum and sum of squares
// of the STORED fp16 values over the rows this CTA handled - every warp reads its staged [32 x 32] block back column-
// wise (lane = column pair, half-warp = row parity: 16 conflict-free LDS.32 per block) while the TMA store of the same
// block is in flight, and keeps (sum, sumsq) x 2 columns per lane in its own 4 KB shared-memory slab; finish() writes
// one partial row per warp for reduce_partials_kernel.  Replaces colstats_kernel's second pass over the tensor (67 MB
// re-read per layer at B = 65 536).  A CTA must see ONE n-group for its whole life (FIXED_NGROUP: the launcher makes
// the grid a multiple of the number of n-groups, so n_group = blockIdx.x % groups); partial matrix: row
// blockIdx.x / groups, columns [sums of N | sums of squares of N].
// GROUPS_ = 4: four epilogue groups (16 warps); groups g and g + 2 share the units of accumulator buffer g & 1 and take
// 128 of the 256 columns each - twice the warps to hide the latency of the TMEM loads and the constant loads behind
// (no row or column statistics in this variant: those need the halves to meet).
template <class Cfg, bool BIAS, bool LRELU, bool ROWSTATS, bool MASKOUT = false, bool AFFINE_RELU = false,
          bool COLSTATS = false, int GROUPS_ = 2>
struct EpiStore {
  static_assert(GROUPS_ == 2 || (GROUPS_ == 4 && !ROWSTATS && !COLSTATS && Cfg::BLOCK_N == 256 && Cfg::ACC_BUFS == 2),
                "four groups: plain store variants on 256-column tiles");
  static constexpr int GROUPS = GROUPS_;
  static constexpr int NT = Cfg::BLOCK_N / (GROUPS_ / 2);   // columns per thread
  static constexpr int NBUF = GROUPS_ == 2 ? 2 : 1;         // staging blocks per warp
  static_assert(Cfg::BLOCK_N % 64 == 0 && Cfg::ACC_TILES == 1 && (!MASKOUT || Cfg::BLOCK_N % 128 == 0), "EpiStore tile shape");
  static_assert(!COLSTATS || (Cfg::BLOCK_N == 256 && !BIAS && !AFFINE_RELU && !LRELU),
                "column statistics: 256-column n-groups, rows beyond M must store zeros");
  static constexpr bool FIXED_NGROUP = COLSTATS;
  struct Params {
    OutTile out;
    const float* bias;
    const float* scale;  // AFFINE_RELU: out = relu(scale[c] * acc + bias[c]) — eval-mode BatchNorm + ReLU folded
                         // into the producing layer (generator.py:19-20 with running statistics)
    float* rowstats;  // [M][n_tiles][2]
    int n_tiles;
    uint32_t* mask;   // MASKOUT: [M][N/32] sign bits of the stored value (bit i of word c: column 32c+i > 0) —
                      // all the LeakyReLU backward needs, at 1/16 of the activation's bytes
    int mask_words;   // words per row (N / 32)
    float* colpart;   // COLSTATS: [gridDim.x / col_groups][2 * N] partial rows
    int col_groups;   // n-groups of the layer (N / 256)
  };
  static constexpr int kSlabBytes = 8 * 32 * 16;   // per warp: 8 column blocks x 32 lanes x float4
  static constexpr int SMEM_BYTES = 4 * NBUF * kWarpBlockBytes + (COLSTATS ? 4 * kSlabBytes : 0);
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  using Stager = WarpStagerT<NBUF>;
  struct State {
    Stager stg;
    uint32_t slab;
    int ng;
  };
  __device__ static void init(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    st.stg.init(cx);
    if constexpr (COLSTATS) {
      st.ng = (int)blockIdx.x % p.col_groups;
      st.slab = cx.smem + kEpiStagingBytes + (uint32_t)(cx.tid >> 5) * kSlabBytes + (uint32_t)cx.lane * 16u;
#pragma unroll
      for (int b = 0; b < 8; ++b)
        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(st.slab + b * 512u), "f"(0.f) : "memory");
    }
  }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, const EpiCtx& cx) {
    const int r = cx.q * 32 + cx.lane;
    const int half = GROUPS_ == 4 ? (cx.group >> 1) : 0;
    const int n0 = w.n_group * Cfg::BLOCK_N + half * NT;
    float s1 = 0.f, s2 = 0.f;
    uint32_t mw[4] = {0u, 0u, 0u, 0u};
    const int row = w.m_tile * kBlockM + r;
    drain_blocks32<NT, 64, GROUPS_ == 4>(tacc + (uint32_t)(half * NT), [&](int c, int, float* v) {
      if constexpr (AFFINE_RELU) {
        float b[32], sc[32];
        load_cols32(p.bias + n0 + c, b);
        load_cols32(p.scale + n0 + c, sc);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(sc[i], v[i], b[i]), 0.f);
      } else if constexpr (BIAS) {
        float b[32];
        load_cols32(p.bias + n0 + c, b);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += b[i];
      }
      if constexpr (LRELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lrelu(v[i]);
      }
      if constexpr (ROWSTATS) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s1 += v[i];
          s2 = fmaf(v[i], v[i], s2);
        }
      }
      if constexpr (MASKOUT) {
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
        // the last four words travel in a shift register (no indexing by a loop counter: the loop is rolled)
        mw[0] = mw[1]; mw[1] = mw[2]; mw[2] = mw[3]; mw[3] = bits;
        if ((c & 127) == 96 && row < g.M)   // 128 columns done: one 16-byte store
          *reinterpret_cast<uint4*>(p.mask + (size_t)row * p.mask_words + (n0 + c - 96) / 32) =
              make_uint4(mw[0], mw[1], mw[2], mw[3]);
      }
      const uint32_t buf = st.stg.acquire(cx);
      Stager::put32(buf, cx.lane, v);
      st.stg.commit(cx, buf, p.out, n0 + c, w.m_tile * kBlockM);
      if constexpr (COLSTATS) {
        // after commit's __syncwarp the whole block is visible: this lane sums columns 2j, 2j+1 over the rows of
        // its parity (the TMA store only reads the block, so both proceed together)
        const int j = cx.lane & 15, par = cx.lane >> 4;
        float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint32_t hw;
          asm volatile("ld.shared.b32 %0, [%1];"
                       : "=r"(hw)
                       : "r"(buf + (uint32_t)(2 * i + par) * 64u + (uint32_t)(((j >> 2) ^ (i & 3)) << 4) + (uint32_t)(j & 3) * 4u)
                       : "memory");
          const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&hw));
          a1 += x.x; a2 = fmaf(x.x, x.x, a2);
          b1 += x.y; b2 = fmaf(x.y, x.y, b2);
        }
        const uint32_t sl = st.slab + (uint32_t)(c >> 5) * 512u;
        float o0, o1, o2, o3;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o0), "=f"(o1), "=f"(o2), "=f"(o3) : "r"(sl) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sl), "f"(o0 + a1), "f"(o1 + a2), "f"(o2 + b1), "f"(o3 + b2)
                     : "memory");
      }
    });
    if constexpr (ROWSTATS) {
      if (row < g.M)
        *reinterpret_cast<float2*>(p.rowstats + ((size_t)row * p.n_tiles + w.n_group) * 2) = make_float2(s1, s2);
    }
  }
  __device__ static void finish(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    if constexpr (COLSTATS) {
      // one partial row per CTA: the eight warps' slabs (and the two row parities inside each) are added here, in a
      // fixed order, so that the reduction kernel reads 148 rows instead of 1184
      asm volatile("bar.sync 5, 256;" ::: "memory");   // both epilogue groups have finished their units
      const int t = cx.group * 128 + cx.tid;
      if (t < 128) {
        const int blk = t >> 4, j = t & 15;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const uint32_t slab = cx.smem0 + (uint32_t)(w >> 2) * SMEM_BYTES + kEpiStagingBytes + (uint32_t)(w & 3) * kSlabBytes +
                                (uint32_t)blk * 512u;
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            float o0, o1, o2, o3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(o0), "=f"(o1), "=f"(o2), "=f"(o3) : "r"(slab + (uint32_t)(par * 16 + j) * 16u) : "memory");
            s0 += o0; s1 += o1; s2 += o2; s3 += o3;
          }
        }
        const int N = p.col_groups * 256;
        float* row = p.colpart + (size_t)(blockIdx.x / p.col_groups) * (2 * N) + st.ng * 256;
        *reinterpret_cast<float2*>(row + blk * 32 + 2 * j) = make_float2(s0, s2);
        *reinterpret_cast<float2*>(row + N + blk * 32 + 2 * j) = make_float2(s1, s3);
      }
    }
    Stager::drain(cx);
  }
};

// =====================================================================================================
// Linear + LayerNorm + LeakyReLU in one epilogue (forward_model.py:35-53): out = fp16(lrelu(LN(acc + bias))).
// A thread owns one row (its TMEM lane) and walks the tile's 256 accumulator columns twice while they stay in
// tensor memory - pass 1: sum / sum of squares, pass 2: normalise, activate, store - so the pre-normalisation
// values never leave the SM and the normalisation works on the fp32 accumulator itself.  A row wider than 256 is
// split over a cluster of CTAs: CTA rank r owns columns [256 r, 256 r + 256) of the same 128 rows; the row partials
// travel through distributed shared memory as st.async stores that complete a transaction count on the receiving
// CTA's mbarrier (no fences, no arrive instructions).  Tensor memory holds two accumulator buffers, so the MMAs of
// the next row tile overlap both passes.
//
// Round 1 kept the row as packed fp16 in 128 registers between the passes ("stash") to release the accumulator early,
// which forces both passes to be fully unrolled: 45 KB of code per kernel, and the ncu source view showed the
// epilogue warps issuing one instruction every ~4 cycles (18 % of their stall samples "no instruction", 14 % the
// membar of the release-arrive on the peer's barrier).  tcgen05.ld delivers ~460 B/clk per SM (profiles/
// r02_ldtm_bench2.txt), so re-reading the accumulator is cheap; both passes are now rolled loops over 32-column
// chunks.  What the epilogue is bound by after that (profiles/r02_ln_trace_*.txt, tools/ln_trace.py): the three
// per-column constants of pass 2 - every constant a thread needs costs a cycle of the shared-memory / L1 data
// pipe per warp whether it comes from shared memory, L1 or the constant bank, because a thread owns a ROW.
// =====================================================================================================
//   PAIR = false: CTA rank r owns n-group r (N = 256 * CLUSTER); its two epilogue groups alternate row tiles.
//   PAIR = true : a CTA walks both n-groups 2r, 2r+1 of a row tile back to back (GemmShape::pair_mode, N = 512 *
//                 CLUSTER): group e owns n-group 2r+e / accumulator buffer e, all 2 * CLUSTER groups exchange their
//                 partials every row tile.  Used for N = 1024, where clusters of 4 would only cover 132 of 148 SMs.
//   GROUPS = 4  : four epilogue groups (16 warps, four per SM sub-partition): groups g and g + 2 share the units of
//                 accumulator buffer g & 1 and take 128 columns each.  The ncu source view of the two-group version
//                 showed the epilogue warps stalled 77 % of the time on fixed-latency dependencies and constant
//                 loads with only two warps per scheduler to hide them.
template <class Cfg, int CLUSTER_, bool PAIR = false, int GROUPS_ = 4, bool TMA_STORE = true>
struct EpiLnStore {
  static_assert(Cfg::BLOCK_N == 256 && Cfg::ACC_TILES == 1 && Cfg::ACC_BUFS == 2, "EpiLnStore tile shape");
  static_assert(GROUPS_ == 2 || GROUPS_ == 4, "epilogue groups");
  static constexpr bool SPLIT = false;
  static constexpr int GROUPS = GROUPS_;
  static constexpr int HALVES = GROUPS_ / 2;                 // column slices of a unit (one per group that shares it)
  static constexpr int PARTS = (PAIR ? 2 : 1) * CLUSTER_ * HALVES;   // row partials per row
  // cluster: one arrive.expect_tx per phase, the partials arrive as transaction bytes; single CTA: plain arrivals of
  // the threads that share the row
  static constexpr int XBAR_COUNT = CLUSTER_ > 1 ? 1 : 128 * PARTS;
  static constexpr int CLUSTER = CLUSTER_;
  static constexpr int NG = 256;           // columns of the row this CTA handles per unit
  static constexpr int NT = NG / HALVES;   // columns per thread
  struct Params {
    OutTile out;
    int n_total;         // LayerNorm width N = 256 * CLUSTER
    long long* trace;    // optional [units][5] clock64 stamps of CTA 0 / thread 0 of a group (tools/ln_trace.py)
    const float* bias;   // [N] device pointers into the layer's fp32 parameters (Linear bias, LayerNorm weight / bias)
    const float* gamma;
    const float* beta;
  };
  // shared memory of all groups together: per-warp store staging (two blocks per warp with 2 groups, one with 4) |
  // row-partial slots: [owner group 0/1][unit parity] buffers of [PARTS][128] float2
  static constexpr int NBUF = GROUPS_ == 2 ? 2 : 1;
  static constexpr int kStagingTotal = GROUPS_ * 4 * NBUF * kWarpBlockBytes;   // 32 KB either way
  static constexpr int kSlotBytes = PARTS * 128 * 8 > 4096 ? 8192 : 4096;
  static_assert(PARTS * 128 * 8 <= kSlotBytes, "slot buffer");
  // | bias, gamma, beta of the CTA's columns (W floats each).  Rounds 1 / early 2 read them from the kernel-parameter
  // constant bank: three streams 4 KB apart thrash the small constant cache - pass 2 (three constants per column)
  // took 9-32k cycles per unit against 1.7k for pass 1 (one constant per column), profiles/r02_ln_trace_*.txt.
  static constexpr int W = PAIR ? 512 : 256;
  static constexpr int kConstOff = kStagingTotal + 4 * kSlotBytes;
  static constexpr int SMEM_TOTAL = kConstOff + 8192;
  static_assert(3 * W * 4 <= 8192, "constant block");
  static constexpr int SMEM_BYTES = SMEM_TOTAL / GROUPS_;
  static_assert(SMEM_TOTAL % (1024 * GROUPS_) == 0, "alignment of the barrier block behind the epilogue scratch");
  struct State {
    WarpStagerT<NBUF, TMA_STORE> stg;
    uint32_t it;
    uint32_t rank;
  };
  __device__ static void init(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    st.stg.init_at(cx.smem0 + (uint32_t)((cx.group * 4 + (cx.tid >> 5)) * NBUF * kWarpBlockBytes));
    st.it = 0;
    st.rank = CLUSTER > 1 ? cluster_ctarank() : 0u;
    // this CTA's columns [rank * W, rank * W + W) of the three constant rows -> shared memory
    const int first = (int)st.rank * W;
    for (int i = cx.group * 128 + cx.tid; i < 3 * W; i += 128 * GROUPS_) {
      const int a = i / W, c = i - a * W;
      const float* src = a == 0 ? p.bias : a == 1 ? p.gamma : p.beta;
      sts_f32(cx.smem0 + kConstOff + 4u * (uint32_t)i, __ldg(src + first + c));
    }
    asm volatile("bar.sync 9, %0;" ::"n"(128 * GROUPS_) : "memory");   // all epilogue groups
  }
  // 16 consecutive per-column constants from shared memory (warp-uniform address: one broadcast per 16 bytes)
  __device__ static void ldc16(uint32_t addr, float* out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 t;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr + 16u * k));
      out[4 * k + 0] = t.x; out[4 * k + 1] = t.y; out[4 * k + 2] = t.z; out[4 * k + 3] = t.w;
    }
  }
  __device__ static void unit(const Params& p, State& st, const GemmShape&, const UnitInfo& w, uint32_t tacc,
                              const EpiCtx& cx) {
    const int r = cx.q * 32 + cx.lane;
    const int buf = cx.group & 1;            // accumulator buffer / (PAIR) n-group parity of this unit
    const int half = cx.group >> 1;          // column slice within the unit
    const uint32_t ngrp = PAIR ? st.rank * 2u + (uint32_t)buf : st.rank;   // which n-group of the row
    const uint32_t part = ngrp * HALVES + (uint32_t)half;                  // which partial of the row
    const int cofs = (int)ngrp * NG + half * NT;                           // first column of this thread
    const uint32_t cbias = cx.smem0 + kConstOff + 4u * (uint32_t)((PAIR ? buf * NG : 0) + half * NT);
    const uint32_t cgamma = cbias + 4u * W, cbeta = cbias + 8u * W;
    const uint32_t tcol = tacc + (uint32_t)(half * NT);
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && cx.tid == 0 && cx.group < 2 && st.it < 32;
    long long* trow = tr ? p.trace + (st.it * 2 + cx.group) * 5 : nullptr;
    if (tr) trow[0] = clock64();
    // ---- pass 1: row sum and sum of squares of acc + bias
    float s1, s2;
    {
      float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < NT; c += 32) {
        float v[32];
        tmem_ld32(tcol + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 32; h += 16) {
          float b[16];
          ldc16(cbias + 4u * (uint32_t)(c + h), b);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = v[h + i] + b[i];
            a1[i & 3] += x;
            a2[i & 3] = fmaf(x, x, a2[i & 3]);
          }
        }
      }
      s1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      s2 = (a2[0] + a2[1]) + (a2[2] + a2[3]);
    }
    if (tr) trow[1] = clock64();
    // ---- exchange the row partials (between the groups that share the unit and inside the cluster)
    if constexpr (PARTS > 1) {
      // The exchange of a unit lives in the region of ONE group: group 0 (PAIR: all groups of the CTA pair up on the
      // row) or the group that owns the accumulator buffer.  Barriers and slots alternate between units, so a peer
      // that is one unit ahead never touches the phase being waited for.
      const int owner = PAIR ? 0 : buf;
      const uint32_t par = st.it & 1u;
      const uint32_t slots = cx.smem0 + kStagingTotal + (uint32_t)(owner * 2 + (int)par) * kSlotBytes;
      const uint32_t xb = (PAIR ? cx.xbar - 16u * (uint32_t)buf : cx.xbar) + 8u * par;
      const uint32_t mine = slots + (part * 128u + (uint32_t)r) * 8u;
      if constexpr (CLUSTER > 1) {
        if (cx.tid == 0 && cx.group == owner) mbar_arrive_expect_tx(xb, PARTS * 128 * 8);
#pragma unroll
        for (uint32_t c = 0; c < (uint32_t)CLUSTER; ++c)
          st_async_f32x2(mapa_shared(mine, c), s1, s2, mapa_shared(xb, c));
      } else {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(mine), "f"(s1), "f"(s2) : "memory");
        mbar_arrive(xb);   // release (CTA scope) of the store above
      }
      mbar_wait(xb, (st.it >> 1) & 1u);
      s1 = 0.f;
      s2 = 0.f;
#pragma unroll
      for (int k = 0; k < PARTS; ++k) {
        float a, b;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                     : "=f"(a), "=f"(b) : "r"(slots + (uint32_t)(k * 128 + r) * 8u) : "memory");
        s1 += a;
        s2 += b;
      }
    }
    ++st.it;
    if (tr) trow[2] = clock64();
    const float inv_n = 1.0f / (float)p.n_total;
    const float mean = s1 * inv_n;
    const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    const float nmr = -mean * rstd;
    // ---- pass 2: normalise the fp32 accumulator, LeakyReLU, fp16, TMA store
#pragma unroll 1
    for (int c = 0; c < NT; c += 32) {
      float v[32];
      tmem_ld32(tcol + c, v);
      tmem_ld_wait();
      uint32_t hw[16];
#pragma unroll
      for (int h = 0; h < 32; h += 16) {
        float bs[16], gm[16], bt[16];
        ldc16(cbias + 4u * (uint32_t)(c + h), bs);
        ldc16(cgamma + 4u * (uint32_t)(c + h), gm);
        ldc16(cbeta + 4u * (uint32_t)(c + h), bt);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          // ((acc + b) - mean) * rstd = fma(acc, rstd, fma(b, rstd, -mean * rstd)); then the affine
          const float y0 = fmaf(fmaf(v[h + i], rstd, fmaf(bs[i], rstd, nmr)), gm[i], bt[i]);
          const float y1 = fmaf(fmaf(v[h + i + 1], rstd, fmaf(bs[i + 1], rstd, nmr)), gm[i + 1], bt[i + 1]);
          hw[(h + i) / 2] = pack_half2(lrelu(y0), lrelu(y1));
        }
      }
      const uint32_t sbuf = st.stg.acquire(cx);
      WarpStagerT<NBUF, TMA_STORE>::put32_packed(sbuf, cx.lane, hw);
      st.stg.commit(cx, sbuf, p.out, cofs + c, w.m_tile * kBlockM);
    }
    if (tr) trow[3] = clock64();
  }
  __device__ static void finish(const Params&, State&, const GemmShape&, const EpiCtx& cx) {
    WarpStagerT<NBUF, TMA_STORE>::drain(cx);
  }
};

// =====================================================================================================
// Discriminator layers 2+3 + BCE (discriminator.py:24-27, loss.py:17, train_pigan.py:127-137,152-154):
// z2 = LeakyReLU(acc + b2) is stored (fp16) for the backward pass; logit = z2.w3 + b3; p = sigmoid(logit);
// BCELoss against a per-row label (label_a for rows < rows_a, label_b after) summed into loss_sum; the
// gradient wrt the logit (autograd's BCE + sigmoid backward, scaled by GS) goes to dlogit[row].
// N must equal BLOCK_N = 256 (one tile holds the whole row).
// =====================================================================================================
template <class Cfg>
struct EpiDiscL2 {
  static_assert(Cfg::BLOCK_N == 256 && Cfg::ACC_TILES == 1, "EpiDiscL2 needs the whole row in one tile");
  struct Params {
    OutTile z2;           // [rows, 256] fp16 out (may be unused: store_z2 = 0)
    const float* b2;      // [256]
    const float* w3;      // [256]
    const float* b3;      // [1]
    float label_a, label_b;
    int rows_a;
    int row_gap_begin, row_gap_end;  // rows in [gap_begin, gap_end) are padding between the two halves
    float inv_batch;      // 1 / global batch (loss mean)
    float grad_mult;      // GS / global batch
    double* loss_sum;     // += sum over rows of BCE * inv_batch
    float* dlogit;        // [rows] out (scaled) or null
    float* prob_out;      // [rows] optional
    int store_z2;
  };
  static constexpr int SMEM_BYTES = kEpiStagingBytes;
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  struct State {
    WarpStager stg;
    float loss;
  };
  __device__ static void init(const Params&, State& st, const GemmShape&, const EpiCtx& cx) {
    st.stg.init(cx);
    st.loss = 0.f;
  }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, const EpiCtx& cx) {
    const int r = cx.q * 32 + cx.lane;
    const int row = w.m_tile * kBlockM + r;
    const bool valid = row < g.M && !(row >= p.row_gap_begin && row < p.row_gap_end);
    float lg[4] = {0.f, 0.f, 0.f, 0.f};  // independent chains (a single one is latency-bound)
    drain_blocks32<256>(tacc, [&](int c, int, float* v) {
      float b[32], w3[32];
      load_cols32(p.b2 + c, b);
      load_cols32(p.w3 + c, w3);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = lrelu(v[i] + b[i]);
        lg[i & 3] = fmaf(v[i], w3[i], lg[i & 3]);
      }
      if (p.store_z2) {
        const uint32_t buf = st.stg.acquire(cx);
        WarpStager::put32(buf, cx.lane, v);
        st.stg.commit(cx, buf, p.z2, c, w.m_tile * kBlockM);
      }
    });
    float logit = (lg[0] + lg[1]) + (lg[2] + lg[3]);
    if (valid) {
      logit += __ldg(p.b3);
      const float prob = 1.f / (1.f + expf(-logit));
      const float y = row < p.rows_a ? p.label_a : p.label_b;
      // nn.BCELoss forward (log clamped at -100) and autograd's backward (denominator clamped at 1e-12)
      const float lp = fmaxf(logf(prob), -100.f);
      const float l1p = fmaxf(log1pf(-prob), -100.f);
      st.loss += -(y * lp + (1.f - y) * l1p) * p.inv_batch;
      const float pq = prob * (1.f - prob);
      if (p.dlogit) p.dlogit[row] = (prob - y) / fmaxf(pq, 1e-12f) * pq * p.grad_mult;
      if (p.prob_out) p.prob_out[row] = prob;
    } else if (row < g.M && p.dlogit) {
      p.dlogit[row] = 0.f;
    }
  }
  __device__ static void finish(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    const float l = warp_sum(st.loss);
    if (cx.lane == 0 && l != 0.f && p.loss_sum) atomicAdd(p.loss_sum, (double)l);
    WarpStager::drain(cx);
  }
};

// =====================================================================================================
// Discriminator layers 2+3 + BCE + the BACKWARD of layers 3 and 2's activation in the same epilogue (train step):
// as EpiDiscL2 up to the logit gradient, then a second pass over the accumulator (still in tensor memory) writes
//   dh2[r, c] = dlogit_r * w3[c] * LeakyReLU'(x[r, c])      (fp16, scaled by GS like dlogit)
// instead of the activation - the layer's output never goes to HBM and the separate d_l2_bwd_kernel (a 134 MB round
// trip per 131 072 rows) disappears.  PG (D-step): the parameter gradients that kernel produced come from here too,
//   dw3[c] = sum_r dlogit_r * a[r, c],  db2[c] = sum_r dh2[r, c],  db3 = sum_r dlogit_r
// as per-CTA partial rows [dw3 (256) | db2 (256) | db3 | pad 7] for reduce_partials_kernel: every warp stages dh2
// (the block the TMA store takes) and dlogit * a as fp16 [32 x 32] blocks and reads them back column-wise (lane =
// column pair, half-warp = row parity), exactly like EpiStore's COLSTATS; the eight warps' sums meet in finish().
// =====================================================================================================
template <class Cfg, bool PG>
struct EpiDiscL2Bwd {
  static_assert(Cfg::BLOCK_N == 256 && Cfg::ACC_TILES == 1, "EpiDiscL2Bwd needs the whole row in one tile");
  struct Params {
    OutTile dh2;          // [rows, 256] fp16 out
    const float* b2;      // [256]
    const float* w3;      // [256]
    const float* b3;      // [1]
    float label_a, label_b;
    int rows_a;
    int row_gap_begin, row_gap_end;  // rows in [gap_begin, gap_end) are padding between the two halves
    float inv_batch;      // 1 / global batch (loss mean)
    float grad_mult;      // GS / global batch
    double* loss_sum;     // += sum over rows of BCE * inv_batch
    float* part;          // PG: [gridDim.x][2 * 256 + 8] partial rows
  };
  static constexpr int kSlabBytes = 8 * 32 * 16;    // per warp: 8 column blocks x 32 lanes x float4
  static constexpr int kAuxOff = kEpiStagingBytes;  // second staging block per warp (dlogit * a)
  static constexpr int kSlabOff = kAuxOff + (PG ? 4 * kWarpBlockBytes : 0);
  static constexpr int SMEM_BYTES = PG ? kSlabOff + 4 * kSlabBytes : kEpiStagingBytes;
  static_assert(SMEM_BYTES % 1024 == 0, "group regions stay 1024-aligned");
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  struct State {
    WarpStager stg;
    float loss, dl_sum;
    uint32_t aux, slab;
  };
  __device__ static void init(const Params&, State& st, const GemmShape&, const EpiCtx& cx) {
    st.stg.init(cx);
    st.loss = 0.f;
    st.dl_sum = 0.f;
    if constexpr (PG) {
      const uint32_t wq = (uint32_t)(cx.tid >> 5);
      st.aux = cx.smem + kAuxOff + wq * kWarpBlockBytes;
      st.slab = cx.smem + kSlabOff + wq * kSlabBytes + (uint32_t)cx.lane * 16u;
#pragma unroll
      for (int b = 0; b < 8; ++b)
        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(st.slab + b * 512u), "f"(0.f) : "memory");
    }
  }
  // column sums of a staged [32 x 32] fp16 block: this lane's two columns over the rows of its parity
  __device__ static void colsum2(uint32_t buf, int lane, float& a, float& b) {
    const int j = lane & 15, par = lane >> 4;
    a = 0.f;
    b = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t hw;
      asm volatile("ld.shared.b32 %0, [%1];"
                   : "=r"(hw)
                   : "r"(buf + (uint32_t)(2 * i + par) * 64u + (uint32_t)(((j >> 2) ^ (i & 3)) << 4) + (uint32_t)(j & 3) * 4u)
                   : "memory");
      const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&hw));
      a += x.x;
      b += x.y;
    }
  }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, const EpiCtx& cx) {
    const int r = cx.q * 32 + cx.lane;
    const int row = w.m_tile * kBlockM + r;
    const bool valid = row < g.M && !(row >= p.row_gap_begin && row < p.row_gap_end);
    // ---- pass 1: logit
    float lg[4] = {0.f, 0.f, 0.f, 0.f};
    drain_blocks32<256>(tacc, [&](int c, int, float* v) {
      float b[32], w3[32];
      load_cols32(p.b2 + c, b);
      load_cols32(p.w3 + c, w3);
#pragma unroll
      for (int i = 0; i < 32; ++i) lg[i & 3] = fmaf(lrelu(v[i] + b[i]), w3[i], lg[i & 3]);
    });
    float dl = 0.f;
    if (valid) {
      const float logit = (lg[0] + lg[1]) + (lg[2] + lg[3]) + __ldg(p.b3);
      const float prob = 1.f / (1.f + expf(-logit));
      const float y = row < p.rows_a ? p.label_a : p.label_b;
      // nn.BCELoss forward (log clamped at -100) and autograd's backward (denominator clamped at 1e-12)
      const float lp = fmaxf(logf(prob), -100.f);
      const float l1p = fmaxf(log1pf(-prob), -100.f);
      st.loss += -(y * lp + (1.f - y) * l1p) * p.inv_batch;
      const float pq = prob * (1.f - prob);
      dl = (prob - y) / fmaxf(pq, 1e-12f) * pq * p.grad_mult;
    }
    st.dl_sum += dl;
    // ---- pass 2: gradient with respect to the layer's pre-activation, straight from the accumulator
    drain_blocks32<256>(tacc, [&](int c, int, float* v) {
      float b[32], w3[32], t[32];
      load_cols32(p.b2 + c, b);
      load_cols32(p.w3 + c, w3);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = v[i] + b[i];
        const float slope = x > 0.f ? 1.f : kLeaky;
        if constexpr (PG) t[i] = dl * (x * slope);
        v[i] = dl * w3[i] * slope;
      }
      const uint32_t buf = st.stg.acquire(cx);
      WarpStager::put32(buf, cx.lane, v);
      if constexpr (PG) {
        __syncwarp();   // the previous block's column reads of the aux block are done
        WarpStager::put32(st.aux, cx.lane, t);
      }
      st.stg.commit(cx, buf, p.dh2, c, w.m_tile * kBlockM);
      if constexpr (PG) {
        float d0, d1, w0, w1;
        colsum2(buf, cx.lane, d0, d1);      // db2: columns 2j, 2j+1 of dh2
        colsum2(st.aux, cx.lane, w0, w1);   // dw3: the same columns of dlogit * a
        const uint32_t sl = st.slab + (uint32_t)(c >> 5) * 512u;
        float o0, o1, o2, o3;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o0), "=f"(o1), "=f"(o2), "=f"(o3) : "r"(sl) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sl), "f"(o0 + w0), "f"(o1 + d0), "f"(o2 + w1), "f"(o3 + d1)
                     : "memory");
      }
    });
  }
  __device__ static void finish(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    const float l = warp_sum(st.loss);
    if (cx.lane == 0 && l != 0.f && p.loss_sum) atomicAdd(p.loss_sum, (double)l);
    if constexpr (PG) {
      const float ds = warp_sum(st.dl_sum);
      WarpStager::drain(cx);   // the warp's staging blocks are free after this: the first one carries its sum of dlogit
      __syncwarp();
      if (cx.lane == 0) sts_f32(cx.smem + (uint32_t)(cx.tid >> 5) * 2u * kWarpBlockBytes, ds);
      asm volatile("bar.sync 5, 256;" ::: "memory");   // both epilogue groups have finished their units
      const int t = cx.group * 128 + cx.tid;
      float* row = p.part + (size_t)blockIdx.x * (2 * 256 + 8);
      if (t < 128) {
        const int blk = t >> 4, j = t & 15;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const uint32_t slab = cx.smem0 + (uint32_t)(w >> 2) * SMEM_BYTES + kSlabOff + (uint32_t)(w & 3) * kSlabBytes +
                                (uint32_t)blk * 512u;
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            float o0, o1, o2, o3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(o0), "=f"(o1), "=f"(o2), "=f"(o3) : "r"(slab + (uint32_t)(par * 16 + j) * 16u) : "memory");
            s0 += o0; s1 += o1; s2 += o2; s3 += o3;
          }
        }
        *reinterpret_cast<float2*>(row + blk * 32 + 2 * j) = make_float2(s0, s2);         // dw3
        *reinterpret_cast<float2*>(row + 256 + blk * 32 + 2 * j) = make_float2(s1, s3);   // db2
      } else if (t == 128) {
        float d = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w)
          d += lds_f32(cx.smem0 + (uint32_t)(w >> 2) * SMEM_BYTES + (uint32_t)(w & 3) * 2u * kWarpBlockBytes);
        row[512] = d;                                                                      // db3
      }
    }
    WarpStager::drain(cx);
  }
};

// =====================================================================================================
// dX through an in-place LeakyReLU (autograd of discriminator.py:23): out = acc * slope(sign of the saved
// activation z).  Used for dH1 = (dH2.W2) * LeakyReLU'(z1) in the D-step.
// =====================================================================================================
template <class Cfg>
struct EpiLeakyMaskStore {
  static_assert(Cfg::BLOCK_N % 64 == 0 && Cfg::ACC_TILES == 1, "tile shape");
  struct Params {
    OutTile out;
    const uint32_t* mask;  // [M][mask_words] sign bits of the saved activations (EpiStore MASKOUT)
    int mask_words;
  };
  static constexpr int SMEM_BYTES = kEpiStagingBytes;
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  struct State {
    WarpStager stg;
  };
  __device__ static void init(const Params&, State& st, const GemmShape&, const EpiCtx& cx) { st.stg.init(cx); }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, const EpiCtx& cx) {
    const int r = cx.q * 32 + cx.lane;
    const int row = w.m_tile * kBlockM + r;
    const int n0 = w.n_group * Cfg::BLOCK_N;
    const bool valid = row < g.M;
    // the row's sign bits for this tile (two 16-byte loads issued before the first tensor-memory load is waited for)
    static_assert(Cfg::BLOCK_N == 256, "mask words of a 256-column tile");
    const uint4* msrc = reinterpret_cast<const uint4*>(p.mask + (size_t)(valid ? row : 0) * p.mask_words + n0 / 32);
    const uint4 mA = __ldg(msrc), mB = __ldg(msrc + 1);
    drain_blocks32<Cfg::BLOCK_N, 128>(tacc, [&](int c, int k, float* v) {
      const uint4 m4 = (c >> 7) ? mB : mA;
      const uint32_t bits = k == 0 ? m4.x : k == 1 ? m4.y : k == 2 ? m4.z : m4.w;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= ((bits >> i) & 1u) ? 1.f : kLeaky;
      const uint32_t buf = st.stg.acquire(cx);
      WarpStager::put32(buf, cx.lane, v);
      st.stg.commit(cx, buf, p.out, n0 + c, w.m_tile * kBlockM);
    });
  }
  __device__ static void finish(const Params&, State&, const GemmShape&, const EpiCtx& cx) { WarpStager::drain(cx); }
};

// =====================================================================================================
// G-step gradient into the 4 parameter inputs of D (train_pigan.py:183 through discriminator.py:22,38):
// dParams[r][j] += sum_c (dH2.W2)[r][c] * LeakyReLU'(z1[r][c]) * W1p[c][j].  dH1 itself is never stored.
// `wp` is [Npad][4] fp32, zero-padded.
// =====================================================================================================
template <class Cfg>
struct EpiDiscParamGrad {
  static_assert(Cfg::ACC_TILES == 1 && Cfg::BLOCK_N % 128 == 0, "tile shape");
  struct Params {
    const uint32_t* mask;  // [M][mask_words] sign bits of z1 (EpiStore MASKOUT)
    int mask_words;
    const float* wp;   // [Npad][4]
    float* dparams;    // [M,4] += (scaled by GS)
  };
  static constexpr int SMEM_BYTES = 0;
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  struct State {};
  __device__ static void init(const Params&, State&, const GemmShape&, const EpiCtx&) {}
  __device__ static void unit(const Params& p, State&, const GemmShape& g, const UnitInfo& w, uint32_t tacc,
                              const EpiCtx& cx) {
    const int row = w.m_tile * kBlockM + cx.q * 32 + cx.lane;
    const int n0 = w.n_group * Cfg::BLOCK_N;
    const bool valid = row < g.M;
    // (one .x32 load per trip and the loop left rolled: the 32 uniform 16-byte weight loads per trip dominate, and
    // larger trips only add register pressure - 37 us against 53 us with .x64 loads and 128-column trips)
    const uint32_t* mrow = p.mask + (size_t)(valid ? row : 0) * p.mask_words + n0 / 32;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
    for (int c = 0; c < Cfg::BLOCK_N; c += 32) {
      float v[32];
      tmem_ld32(tacc + c, v);
      const uint32_t bits = __ldg(mrow + c / 32);
      tmem_ld_wait();
      const float4* wq = reinterpret_cast<const float4*>(p.wp) + n0 + c;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float gz = v[i] * (((bits >> i) & 1u) ? 1.f : kLeaky);
        const float4 w4 = __ldg(wq + i);
        a0 = fmaf(gz, w4.x, a0);
        a1 = fmaf(gz, w4.y, a1);
        a2 = fmaf(gz, w4.z, a2);
        a3 = fmaf(gz, w4.w, a3);
      }
    }
    if (valid) {
      atomicAdd(p.dparams + (size_t)row * 4 + 0, a0);
      atomicAdd(p.dparams + (size_t)row * 4 + 1, a1);
      atomicAdd(p.dparams + (size_t)row * 4 + 2, a2);
      atomicAdd(p.dparams + (size_t)row * 4 + 3, a3);
    }
  }
  __device__ static void finish(const Params&, State&, const GemmShape&, const EpiCtx&) {}
};

// =====================================================================================================
// Scoring path, generator layer 2 onwards in ONE epilogue (eval mode: BatchNorm uses running statistics, so rows
// never interact): ReLU(BN2(acc)) -> Linear(256,4) + tanh = the 4 predicted parameters (generator.py:22-25), then the
// forward surrogate's first layer Linear(4,256) + LayerNorm + LeakyReLU (forward_model.py:30-33) straight into the
// fp16 operand of its second layer.  The pre-BatchNorm activations, the generator output round trip and two
// streaming kernels disappear.  The accumulator is read once and released before the surrogate part starts.
//
// The surrogate's first layer is affine in the 4 parameters, so its LayerNorm statistics are closed-form:
//   h_c - mean = bc_c + wc_c . p   (bc, wc = bias / weight columns centred over c),   var = u^T Q u,  u = (1, p),
//   Q = 1/256 sum_c (bc_c, wc_c)(bc_c, wc_c)^T  (5 x 5, evaluated in fp64 per row: 15 FMAs).
// Per-column constants (14 KB image built by head_consts_kernel, elementwise.cu) are staged once per CTA in shared
// memory and read as warp-broadcast 16-byte loads:
//   floats [0,1536)    per column pair t: s_2t b_2t s_2t+1 b_2t+1 | w3[0..3][2t] | w3[0..3][2t+1]   (BN2 folded, head)
//   floats [1536,1540) b3 ; [1540,1570) Q as 15 doubles (upper triangle, row-major)
//   floats [2048,3072) gamma_c * wc_c[0..3] ; [3072,3584) (gamma_c * bc_c, beta_c)
// =====================================================================================================
constexpr int kHeadImgFloats = 3584;

template <class Cfg>
struct EpiHeadF1 {
  static_assert(Cfg::BLOCK_N == 256 && Cfg::ACC_TILES == 1 && Cfg::ACC_BUFS == 2, "EpiHeadF1 tile shape");
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  static constexpr bool EARLY_RELEASE = true;
  struct Params {
    OutTile out;        // surrogate layer-1 activations [M,256] fp16
    float* p_out;       // [M,4] predicted parameters (tanh output)
    const float* img;   // kHeadImgFloats constants, layout above
  };
  // per group: [0,16K) per-warp store staging | [16K,24K) one half of the constant image (group 0: floats [0,2048),
  // group 1: floats [2048,3584)); both groups read both halves.
  static constexpr int SMEM_BYTES = kEpiStagingBytes + 8192;
  struct State {
    WarpStager stg;
  };
  __device__ static uint32_t part_a(const EpiCtx& cx) { return cx.smem0 + kEpiStagingBytes; }
  __device__ static uint32_t part_b(const EpiCtx& cx) { return cx.smem0 + SMEM_BYTES + kEpiStagingBytes; }
  __device__ static float4 lds4(uint32_t addr) {
    float4 t;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr));
    return t;
  }
  __device__ static double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
  }
  __device__ static void init(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    st.stg.init(cx);
    const int base = cx.group == 0 ? 0 : 2048, n4 = cx.group == 0 ? 512 : 384;
    const uint32_t dst = cx.group == 0 ? part_a(cx) : part_b(cx);
    const float4* src = reinterpret_cast<const float4*>(p.img + base);
    for (int i = cx.tid; i < n4; i += 128) {
      const float4 t = __ldg(src + i);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u * i), "f"(t.x), "f"(t.y), "f"(t.z),
                   "f"(t.w) : "memory");
    }
    asm volatile("bar.sync 5, 256;" ::: "memory");
  }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w, uint32_t tacc,
                              const EpiCtx& cx) {
    const int r = cx.q * 32 + cx.lane;
    const int row = w.m_tile * kBlockM + r;
    const uint32_t pa = part_a(cx), pb = part_b(cx);
    // ---- generator head: the only pass over tensor memory
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    drain_blocks32<256>(tacc, [&](int c, int, float* v) {
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const uint32_t ad = pa + (uint32_t)(c / 2 + t) * 48u;
        const float4 sb = lds4(ad), w0 = lds4(ad + 16), w1 = lds4(ad + 32);
        const float a0 = fmaxf(fmaf(sb.x, v[2 * t], sb.y), 0.f);
        const float a1 = fmaxf(fmaf(sb.z, v[2 * t + 1], sb.w), 0.f);
        d0 = fmaf(a0, w0.x, d0); d1 = fmaf(a0, w0.y, d1); d2 = fmaf(a0, w0.z, d2); d3 = fmaf(a0, w0.w, d3);
        d0 = fmaf(a1, w1.x, d0); d1 = fmaf(a1, w1.y, d1); d2 = fmaf(a1, w1.z, d2); d3 = fmaf(a1, w1.w, d3);
      }
    });
    tc_fence_before();
    mbar_arrive(cx.tempty);   // accumulator drained: the MMAs of the unit after next may start
    const float4 b3 = lds4(pa + 1536 * 4);
    const float p0 = tanhf(d0 + b3.x), p1 = tanhf(d1 + b3.y), p2 = tanhf(d2 + b3.z), p3 = tanhf(d3 + b3.w);
    if (row < g.M) *reinterpret_cast<float4*>(p.p_out + (size_t)row * 4) = make_float4(p0, p1, p2, p3);
    // ---- surrogate layer 1: closed-form LayerNorm statistics, then one pass over the 256 columns
    float rs;
    {
      const double u[5] = {1.0, (double)p0, (double)p1, (double)p2, (double)p3};
      const uint32_t qa = pa + 1540 * 4;
      double var = 0.0;
      int k = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = i; j < 5; ++j, ++k) acc = fma(lds_f64(qa + 8u * k) * (j == i ? 1.0 : 2.0), u[j], acc);
        var = fma(acc, u[i], var);
      }
      rs = 1.0f / sqrtf(fmaxf((float)var, 0.f) + 1e-5f);
    }
    const float q0 = p0 * rs, q1 = p1 * rs, q2 = p2 * rs, q3 = p3 * rs;
#pragma unroll 1
    for (int blk = 0; blk < 8; ++blk) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int c = blk * 32 + i;
        const float4 gb = lds4(pb + 4096u + (uint32_t)c * 8u);   // (g*bc, beta) of columns c, c+1
        const float4 g0 = lds4(pb + (uint32_t)c * 16u), g1 = lds4(pb + (uint32_t)c * 16u + 16u);
        v[i] = lrelu(fmaf(g0.w, q3, fmaf(g0.z, q2, fmaf(g0.y, q1, fmaf(g0.x, q0, fmaf(gb.x, rs, gb.y))))));
        v[i + 1] = lrelu(fmaf(g1.w, q3, fmaf(g1.z, q2, fmaf(g1.y, q1, fmaf(g1.x, q0, fmaf(gb.z, rs, gb.w))))));
      }
      const uint32_t buf = st.stg.acquire(cx);
      WarpStager::put32(buf, cx.lane, v);
      st.stg.commit(cx, buf, p.out, blk * 32, w.m_tile * kBlockM);
    }
  }
  __device__ static void finish(const Params&, State&, const GemmShape&, const EpiCtx& cx) { WarpStager::drain(cx); }
};

// =====================================================================================================
// Forward-model output layer fused with everything the step derives from it without ever writing the
// [B,258] output (forward_model.py:56,74-75; train_pigan.py:159-170; loss.py:51-56,82-101;
// unified_evaluator.py:387): reconstruction MSE vs the real spectrum, metric MSE, mean squared second
// difference (maxwell), the two LC terms and d(LC)/d(params), per-row reconstruction error for
// candidate scoring, optional fp32 dump of the output for ForwardModel.forward.
// Tile: 2 accumulators of 144 columns = 288 >= 258, one thread sees its whole output row.
// =====================================================================================================
template <class Cfg>
struct EpiFwdOut {
  static_assert(Cfg::BLOCK_N == 144 && Cfg::ACC_TILES == 2 && Cfg::ACC_BUFS == 1, "EpiFwdOut tile shape");
  static constexpr bool SPLIT = true;   // group g handles output columns [144 g, 144 g + 144)
  static constexpr int CLUSTER = 1;
  struct Params {
    const float* bias;            // [288] zero-padded
    int S, Mt;
    // reconstruction target, one of:
    //   target_mode 1: one fp32 row for all candidates (target_row, zero-padded to 256)
    //   target_mode 2: per-row spectra as the centred fp16 operand copy [M,256] (tgt map, 128 x 64 boxes)
    //                  + the centring row (center, fp32, zero-padded to 256): x = fp16 + center
    int target_mode;
    const float* target_row;
    CUtensorMap tgt;
    const float* center;
    const float* target_metrics;  // [M,Mt] or null
    const float* p_norm;          // [M,4] generator output (LC loss) or null
    double* sums;                 // [0]=sum (recon-x)^2 [1]=sum (pm-m)^2 [2]=sum d2^2 [3]=sum lc1 [4]=sum lc2
    float* dp_lc;                 // [M,4] d(lambda_lc * LC)/dp * GS  or null
    float lc_grad_mult;           // lambda_lc * GS / global batch
    float* out_full;              // [M,S+Mt] fp32 or null (target_mode must be 0 or 1: shares the staging area)
    float* row_err;               // [M] mean_j (x - recon)^2 or null
    int f1_idx, f2_idx;
  };
  // group 0's region: [0, 64 KB) target tile, 4 swizzled [128 x 64] fp16 boxes (TMA); [64 KB, +8) its mbarrier;
  // [+1 KB, +2 KB) group 1's per-row partial errors (two buffers).  Every group's region also serves as its
  // per-warp [32][17] fp32 transposition buffers on the out_full path (no target tile in that mode).
  static constexpr int kTileBytes = 4 * kStageBytes;
  static constexpr int kPartOff = kTileBytes + 1024;
  static constexpr int SMEM_BYTES = kTileBytes + 3072;
  struct State {
    float s_rec, s_met, s_mx, s_lc1, s_lc2;
    uint32_t tphase, it;
  };
  __device__ static void issue_target(const Params& p, const EpiCtx& cx, int m_tile) {
    const uint32_t bar = cx.smem0 + kTileBytes;
    mbar_arrive_expect_tx(bar, kTileBytes);
#pragma unroll
    for (int b = 0; b < 4; ++b) tma_load_2d(cx.smem0 + b * kStageBytes, &p.tgt, bar, b * 64, m_tile * kBlockM);
  }
  __device__ static void sync_groups() { asm volatile("bar.sync 5, 256;" ::: "memory"); }
  __device__ static void init(const Params& p, State& st, const GemmShape& g, const EpiCtx& cx) {
    st.s_rec = st.s_met = st.s_mx = st.s_lc1 = st.s_lc2 = 0.f;
    st.tphase = 0;
    st.it = 0;
    if (p.target_mode == 2 && cx.group == 0 && cx.tid == 0) {
      mbar_init(cx.smem0 + kTileBytes, 1);
      fence_barrier_init();
      if ((int)blockIdx.x < g.num_m_tiles) issue_target(p, cx, blockIdx.x);
    }
    sync_groups();
  }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, const EpiCtx& cx) {
    const int rt = cx.q * 32 + cx.lane;               // row within the tile
    const int row0 = w.m_tile * kBlockM + cx.q * 32;  // first row of this warp
    const int row = row0 + cx.lane;
    const bool valid = row < g.M;
    const int OUT = p.S + p.Mt;
    const int jbase = cx.group * 144;
    const uint32_t tb = cx.smem + (uint32_t)(cx.tid >> 5) * (32u * 17u * 4u);  // this warp's transposition buffer
    float prev1 = 0.f, prev2 = 0.f;  // recon[j-1], recon[j-2]
    float rec = 0.f, met = 0.f, mx = 0.f, f1 = 0.f, f2 = 0.f;
    float rec4[4] = {0.f, 0.f, 0.f, 0.f}, mx4[4] = {0.f, 0.f, 0.f, 0.f};  // independent accumulation chains
    const float* ms = p.target_metrics ? p.target_metrics + (size_t)(valid ? row : 0) * p.Mt : nullptr;
    if (cx.group == 1) {
      // the second differences at columns 144 and 145 reach back into group 0's last two columns
      float v[16], b[16];
      tmem_ld16(tacc + 128, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + 128) + k);
        b[4 * k] = t.x; b[4 * k + 1] = t.y; b[4 * k + 2] = t.z; b[4 * k + 3] = t.w;
      }
      tmem_ld_wait();
      prev2 = v[14] + b[14];
      prev1 = v[15] + b[15];
    }
    if (p.target_mode == 2) {
      mbar_wait(cx.smem0 + kTileBytes, st.tphase);
      st.tphase ^= 1u;
    }
#pragma unroll 1
    for (int k16 = 0; k16 < 9; ++k16) {
      const int j0 = jbase + 16 * k16;
      float v[16];
      tmem_ld16(tacc + j0, v);  // the two 144-column accumulators are adjacent in TMEM
      float t[16], b[16];
      const bool has_t = p.target_mode != 0 && j0 < p.S;
      if (has_t) {
        const float* src = (p.target_mode == 1 ? p.target_row : p.center) + j0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 q4 = __ldg(reinterpret_cast<const float4*>(src) + k);
          t[4 * k] = q4.x; t[4 * k + 1] = q4.y; t[4 * k + 2] = q4.z; t[4 * k + 3] = q4.w;
        }
        if (p.target_mode == 2) {
          const uint32_t box = cx.smem0 + (uint32_t)(j0 >> 6) * kStageBytes + (uint32_t)rt * 128u;
          const int c8 = (j0 & 63) >> 3;  // first 16-byte chunk of these 16 columns inside the 64-column box
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t u0, u1, u2, u3;
            const uint32_t addr = box + (uint32_t)(((c8 + i) ^ (rt & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(addr) : "memory");
            const uint32_t uu[4] = {u0, u1, u2, u3};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&uu[k]));
              t[8 * i + 2 * k] += f.x;
              t[8 * i + 2 * k + 1] += f.y;
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(p.bias + j0) + k);
        b[4 * k] = q4.x; b[4 * k + 1] = q4.y; b[4 * k + 2] = q4.z; b[4 * k + 3] = q4.w;
      }
      tmem_ld_wait();
      if (j0 >= OUT) continue;
      if (j0 >= 2 && j0 + 16 <= p.S) {
        // interior spectrum chunk (all but the first and the last two of the 18): branch-free
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float o = v[i] + b[i];
          v[i] = o;
          if (has_t) {
            const float d = o - t[i];
            rec4[i & 3] = fmaf(d, d, rec4[i & 3]);
          }
          const float d2 = (o - prev1) - (prev1 - prev2);  // loss.py:51-53 difference of differences
          mx4[i & 3] = fmaf(d2, d2, mx4[i & 3]);
          prev2 = prev1;
          prev1 = o;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = j0 + i;
          const float o = v[i] + b[i];
          v[i] = o;
          if (j < p.S) {
            if (has_t) {
              const float d = o - t[i];
              rec = fmaf(d, d, rec);
            }
            if (j >= 2) {
              const float d2 = (o - prev1) - (prev1 - prev2);
              mx = fmaf(d2, d2, mx);
            }
            prev2 = prev1;
            prev1 = o;
          } else if (j < OUT) {
            const int k = j - p.S;
            if (ms) {
              const float d = o - __ldg(ms + k);
              met = fmaf(d, d, met);
            }
            if (k == p.f1_idx) f1 = o;
            if (k == p.f2_idx) f2 = o;
          }
        }
      }
      if (p.out_full) {
        // lane = row while writing to shared memory; 2 rows x 16 columns per store instruction
#pragma unroll
        for (int i = 0; i < 16; ++i) sts_f32(tb + (uint32_t)(cx.lane * 17 + i) * 4u, v[i]);
        __syncwarp();
        const int col = j0 + (cx.lane & 15);
        if (col < OUT) {
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int rr = 2 * i + (cx.lane >> 4);
            if (row0 + rr < g.M)
              p.out_full[(size_t)(row0 + rr) * OUT + col] = lds_f32(tb + (uint32_t)(rr * 17 + (cx.lane & 15)) * 4u);
          }
        }
        __syncwarp();
      }
    }
    rec += (rec4[0] + rec4[1]) + (rec4[2] + rec4[3]);
    mx += (mx4[0] + mx4[1]) + (mx4[2] + mx4[3]);
    // group 1 hands its part of the row's squared error to group 0 (two buffers alternate between units)
    const uint32_t part = cx.smem0 + kPartOff + (st.it & 1u) * 512u + (uint32_t)rt * 4u;
    ++st.it;
    if (cx.group == 1) sts_f32(part, rec);
    sync_groups();  // also: everyone is done with the target tile
    if (p.target_mode == 2 && cx.group == 0 && cx.tid == 0) {
      const int next = w.m_tile + (int)gridDim.x;   // fetched while the MMAs of the next unit run
      if (next < g.num_m_tiles) issue_target(p, cx, next);
    }
    if (!valid) return;
    st.s_rec += rec;
    st.s_met += met;
    st.s_mx += mx;
    if (cx.group == 0) {
      if (p.row_err) p.row_err[row] = (rec + lds_f32(part)) / (float)p.S;
      return;
    }
    if (p.p_norm) {
      const float4 pn = __ldg(reinterpret_cast<const float4*>(p.p_norm) + row);
      const float e1 = f1 - (0.4f * pn.x + 0.6f * pn.z);
      const float e2 = f2 - (0.3f * pn.y + 0.7f * pn.w);
      st.s_lc1 = fmaf(e1, e1, st.s_lc1);
      st.s_lc2 = fmaf(e2, e2, st.s_lc2);
      if (p.dp_lc) {
        // d/dp of (f - th)^2 with f constant (no grad through F, train_pigan.py:156-157): -2 e dth/dp
        const float m = -2.f * p.lc_grad_mult;
        float4 o = make_float4(m * e1 * 0.4f, m * e2 * 0.3f, m * e1 * 0.6f, m * e2 * 0.7f);
        *reinterpret_cast<float4*>(p.dp_lc + (size_t)row * 4) = o;
      }
    }
  }
  __device__ static void finish(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    float s[5] = {st.s_rec, st.s_met, st.s_mx, st.s_lc1, st.s_lc2};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float t = warp_sum(s[k]);
      if (cx.lane == 0 && t != 0.f && p.sums) atomicAdd(p.sums + k, (double)t);
    }
  }
};

// =====================================================================================================
// Candidate search (unified_evaluator.py:387 with ONE target spectrum for all rows): per-row reconstruction error
// row_err[r] = mean_j (F(p_r)[j] - target[j])^2 from the surrogate's output layer, spectrum columns only.  The 8
// metric columns are not needed here, so the tile is a plain 128 x 256 one (the weight map is built over the S = 250
// spectrum rows: TMA zero-fills the other six) and the accumulator is DOUBLE-buffered - EpiFwdLoss' 288 columns fill
// tensor memory once, so its epilogue and the next unit's MMAs take turns (ncu source view of the scoring launch: the
// epilogue warps wait for the accumulator in 39 % of their samples).  One per-column constant (bias - target), staged
// per group in shared memory; no smoothness term (train-time only).
// =====================================================================================================
template <class Cfg>
struct EpiFwdScore {
  static_assert(Cfg::BLOCK_N == 256 && Cfg::ACC_TILES == 1 && Cfg::ACC_BUFS == 2, "EpiFwdScore tile shape");
  struct Params {
    const float* bias;   // [>= 256] output-layer bias, zero-padded
    const float* tcen;   // [256] the target row, zero-padded
    int S;
    float* row_err;      // [M]
  };
  static constexpr int SMEM_BYTES = 1024;   // per group: (bias - target)[256]
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  struct State {};
  __device__ static void init(const Params& p, State&, const GemmShape&, const EpiCtx& cx) {
    for (int j = cx.tid; j < 256; j += 128)
      sts_f32(cx.smem + 4u * (uint32_t)j, j < p.S ? __ldg(p.bias + j) - __ldg(p.tcen + j) : 0.f);
    epi_bar_sync(cx, 0);
  }
  __device__ static void unit(const Params& p, State&, const GemmShape& g, const UnitInfo& w, uint32_t tacc,
                              const EpiCtx& cx) {
    const int row = w.m_tile * kBlockM + cx.q * 32 + cx.lane;
    float rec4[4] = {0.f, 0.f, 0.f, 0.f};
    drain_blocks32<256>(tacc, [&](int c, int, float* v) {
#pragma unroll
      for (int i4 = 0; i4 < 32; i4 += 4) {
        float4 t;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(cx.smem + 4u * (uint32_t)(c + i4)));
        const float d0 = v[i4] + t.x, d1 = v[i4 + 1] + t.y, d2 = v[i4 + 2] + t.z, d3 = v[i4 + 3] + t.w;
        rec4[0] = fmaf(d0, d0, rec4[0]);
        rec4[1] = fmaf(d1, d1, rec4[1]);
        rec4[2] = fmaf(d2, d2, rec4[2]);
        rec4[3] = fmaf(d3, d3, rec4[3]);
      }
    });
    if (row < g.M) p.row_err[row] = ((rec4[0] + rec4[1]) + (rec4[2] + rec4[3])) / (float)p.S;
  }
  __device__ static void finish(const Params&, State&, const GemmShape&, const EpiCtx&) {}
};

// =====================================================================================================
// The loss / scoring variants of the forward-model output layer (same tile shape and column split as EpiFwdOut, which
// keeps the generic path with the fp32 dump).  Round 1 ran every mode through one epilogue with run-time switches:
// 34 instructions per accumulator element, 7.7 % tensor-pipe activity (profiles/r02_fwd_gemms_ncu_before.csv).  Here
// the mode is a template parameter, the per-column constants (bias, bias - target centre) are staged once per CTA in
// shared memory, a trip handles 48 columns branch-free (one .x32 + one .x16 tensor-memory load in flight), and the
// second differences of the maxwell term come from the neighbouring registers instead of a serial chain.
//   TMODE 1: one target row for all rows (candidate search; tcen = the target)         d = acc + (bias - target)
//   TMODE 2: per-row targets as the centred fp16 operand copy (TMA-staged) + centre     d = acc + (bias - centre) - x16
//   TRAIN  : also metric MSE, LC terms and d(LC)/dp, sums for the loss scalars (train_pigan.py:159-170); else the
//            per-row reconstruction error (unified_evaluator.py:387)
// =====================================================================================================
template <class Cfg, int TMODE, bool TRAIN>
struct EpiFwdLoss {
  static_assert(Cfg::BLOCK_N == 144 && Cfg::ACC_TILES == 2 && Cfg::ACC_BUFS == 1, "EpiFwdLoss tile shape");
  static_assert(TMODE == 1 || TMODE == 2, "target mode");
  static constexpr bool SPLIT = true;   // group g handles output columns [144 g, 144 g + 144)
  static constexpr int CLUSTER = 1;
  struct Params {
    const float* bias;            // [288] zero-padded
    const float* tcen;            // [256] zero-padded: the target row (TMODE 1) or the operand's centring row (TMODE 2)
    CUtensorMap tgt;              // TMODE 2: centred fp16 operand copy [M,256], 128 x 64 boxes
    int S, Mt;
    const float* target_metrics;  // TRAIN: [M,Mt]
    const float* p_norm;          // TRAIN: [M,4] generator output (LC loss)
    double* sums;                 // TRAIN: [0]=sum (recon-x)^2 [1]=sum (pm-m)^2 [2]=sum d2^2 [3]=sum lc1 [4]=sum lc2
    float* dp_lc;                 // TRAIN: [M,4] d(lambda_lc * LC)/dp * GS  or null
    float lc_grad_mult;           // lambda_lc * GS / global batch
    float* row_err;               // !TRAIN: [M] mean_j (x - recon)^2
    int f1_idx, f2_idx;
  };
  // group 0's region: [0, 64 KB) target tile, 4 swizzled [128 x 64] fp16 boxes (TMA); [64 KB, +8) its mbarrier;
  // [+1 KB, +2 KB) group 1's per-row partial errors (two buffers); [+2 KB, +4.25 KB) bias[288] | bias - tcen [288]
  static constexpr int kTileBytes = 4 * kStageBytes;
  static constexpr int kPartOff = kTileBytes + 1024;
  static constexpr int kConstOff = kTileBytes + 2048;
  static constexpr int SMEM_BYTES = kTileBytes + 5120;
  static constexpr int SMEM_TOTAL = SMEM_BYTES;   // everything lives in group 0's region: the operand ring gets the rest
  struct State {
    float s_rec, s_met, s_mx, s_lc1, s_lc2;
    uint32_t tphase, it;
  };
  __device__ static void issue_target(const Params& p, const EpiCtx& cx, int m_tile) {
    const uint32_t bar = cx.smem0 + kTileBytes;
    mbar_arrive_expect_tx(bar, kTileBytes);
#pragma unroll
    for (int b = 0; b < 4; ++b) tma_load_2d(cx.smem0 + b * kStageBytes, &p.tgt, bar, b * 64, m_tile * kBlockM);
  }
  __device__ static void sync_groups() { asm volatile("bar.sync 5, 256;" ::: "memory"); }
  __device__ static float4 lds4(uint32_t addr) {
    float4 t;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr));
    return t;
  }
  __device__ static void init(const Params& p, State& st, const GemmShape& g, const EpiCtx& cx) {
    st.s_rec = st.s_met = st.s_mx = st.s_lc1 = st.s_lc2 = 0.f;
    st.tphase = 0;
    st.it = 0;
    if (TMODE == 2 && cx.group == 0 && cx.tid == 0) {
      mbar_init(cx.smem0 + kTileBytes, 1);
      fence_barrier_init();
      if ((int)blockIdx.x < g.num_m_tiles) issue_target(p, cx, blockIdx.x);
    }
    // per-column constants: bias | bias - tcen (columns >= S carry no target)
    for (int j = cx.tid + 128 * cx.group; j < 288; j += 256) {
      const float b = __ldg(p.bias + j);
      const float t = j < p.S ? __ldg(p.tcen + j) : 0.f;
      sts_f32(cx.smem0 + kConstOff + 4u * j, b);
      sts_f32(cx.smem0 + kConstOff + 4u * (288 + j), b - t);
    }
    sync_groups();
  }
  __device__ static void unit(const Params& p, State& st, const GemmShape& g, const UnitInfo& w,
                              uint32_t tacc, const EpiCtx& cx) {
    const int rt = cx.q * 32 + cx.lane;               // row within the tile
    const int row = w.m_tile * kBlockM + rt;
    const bool valid = row < g.M;
    const int OUT = p.S + p.Mt;
    const int jbase = cx.group * 144;
    const uint32_t cb = cx.smem0 + kConstOff;
    float rec4[4] = {0.f, 0.f, 0.f, 0.f}, mx4[4] = {0.f, 0.f, 0.f, 0.f};
    float met = 0.f, f1 = 0.f, f2 = 0.f;
    float p1 = 0.f, q = 0.f;   // o[j-1] and o[j-1] - o[j-2] entering the current trip
    if (cx.group == 1) {
      // the second differences at columns 144 and 145 reach back into group 0's last two columns
      float v[16];
      tmem_ld16(tacc + 128, v);
      const float4 b = lds4(cb + 4u * 140);
      tmem_ld_wait();
      const float o142 = v[14] + b.z, o143 = v[15] + b.w;
      p1 = o143;
      q = o143 - o142;
    }
    if constexpr (TMODE == 2) {
      mbar_wait(cx.smem0 + kTileBytes, st.tphase);
      st.tphase ^= 1u;
    }
#pragma unroll 1
    for (int trip = 0; trip < 3; ++trip) {
      const int j0 = jbase + 48 * trip;
      float v[48];
      tmem_ld32(tacc + j0, v);   // the two 144-column accumulators are adjacent in tensor memory
      tmem_ld16(tacc + j0 + 32, v + 32);
      // target values of these 48 columns (TMODE 2): six 16-byte chunks of this row in the swizzled boxes
      uint32_t xh[24];
      if constexpr (TMODE == 2) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int j = j0 + 8 * k < 256 ? j0 + 8 * k : 248;   // the tile has 256 columns; columns >= S are never used
          const uint32_t addr = cx.smem0 + (uint32_t)(j >> 6) * kStageBytes + (uint32_t)rt * 128u +
                                (uint32_t)(((((j & 63) >> 3)) ^ (rt & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(xh[4 * k]), "=r"(xh[4 * k + 1]), "=r"(xh[4 * k + 2]), "=r"(xh[4 * k + 3]) : "r"(addr) : "memory");
        }
      }
      tmem_ld_wait();
      if (j0 + 48 <= p.S) {
        // ---- all 48 columns are spectrum samples: branch-free
#pragma unroll
        for (int i4 = 0; i4 < 48; i4 += 4) {
          const float4 b4 = lds4(cb + 4u * (j0 + i4)), c4 = lds4(cb + 4u * (288 + j0 + i4));
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i4 + u;
            const float o = v[i] + bb[u];
            float d = v[i] + cc[u];
            if constexpr (TMODE == 2) {
              const float2 xf = __half22float2(*reinterpret_cast<const __half2*>(&xh[i >> 1]));
              d -= (i & 1) ? xf.y : xf.x;
            }
            rec4[u] = fmaf(d, d, rec4[u]);
            const float t1 = o - p1;           // loss.py:51-53 difference of differences
            float d2 = t1 - q;
            if (i < 2) d2 = (j0 + i >= 2) ? d2 : 0.f;   // the first two columns of the row have no second difference
            mx4[u] = fmaf(d2, d2, mx4[u]);
            q = t1;
            p1 = o;
          }
        }
      } else {
        // ---- the trip that holds the end of the spectrum, the metric columns and the padding
        const float* ms = (TRAIN && p.target_metrics) ? p.target_metrics + (size_t)(valid ? row : 0) * p.Mt : nullptr;
#pragma unroll
        for (int i = 0; i < 48; ++i) {
          const int j = j0 + i;
          const float o = v[i] + lds_f32(cb + 4u * j);
          if (j < p.S) {
            float d = v[i] + lds_f32(cb + 4u * (288 + j));
            if constexpr (TMODE == 2) {
              const float2 xf = __half22float2(*reinterpret_cast<const __half2*>(&xh[i >> 1]));
              d -= (i & 1) ? xf.y : xf.x;
            }
            rec4[i & 3] = fmaf(d, d, rec4[i & 3]);
            const float t1 = o - p1;
            const float d2 = t1 - q;
            if (j >= 2) mx4[i & 3] = fmaf(d2, d2, mx4[i & 3]);
            q = t1;
            p1 = o;
          } else if (TRAIN && j < OUT) {
            const int k = j - p.S;
            if (ms) {
              const float d = o - __ldg(ms + k);
              met = fmaf(d, d, met);
            }
            if (k == p.f1_idx) f1 = o;
            if (k == p.f2_idx) f2 = o;
          }
        }
      }
    }
    const float rec = (rec4[0] + rec4[1]) + (rec4[2] + rec4[3]);
    const float mx = (mx4[0] + mx4[1]) + (mx4[2] + mx4[3]);
    // group 1 hands its part of the row's squared error to group 0 (two buffers alternate between units)
    const uint32_t part = cx.smem0 + kPartOff + (st.it & 1u) * 512u + (uint32_t)rt * 4u;
    ++st.it;
    if (!TRAIN && cx.group == 1) sts_f32(part, rec);
    sync_groups();  // also: everyone is done with the target tile
    if (TMODE == 2 && cx.group == 0 && cx.tid == 0) {
      const int next = w.m_tile + (int)gridDim.x;   // fetched while the MMAs of the next unit run
      if (next < g.num_m_tiles) issue_target(p, cx, next);
    }
    if (!valid) return;
    if constexpr (!TRAIN) {
      if (cx.group == 0) p.row_err[row] = (rec + lds_f32(part)) / (float)p.S;
    } else {
      st.s_rec += rec;
      st.s_met += met;
      st.s_mx += mx;
      if (cx.group == 1 && p.p_norm) {
        const float4 pn = __ldg(reinterpret_cast<const float4*>(p.p_norm) + row);
        const float e1 = f1 - (0.4f * pn.x + 0.6f * pn.z);
        const float e2 = f2 - (0.3f * pn.y + 0.7f * pn.w);
        st.s_lc1 = fmaf(e1, e1, st.s_lc1);
        st.s_lc2 = fmaf(e2, e2, st.s_lc2);
        if (p.dp_lc) {
          // d/dp of (f - th)^2 with f constant (no grad through F, train_pigan.py:156-157): -2 e dth/dp
          const float m = -2.f * p.lc_grad_mult;
          *reinterpret_cast<float4*>(p.dp_lc + (size_t)row * 4) =
              make_float4(m * e1 * 0.4f, m * e2 * 0.3f, m * e1 * 0.6f, m * e2 * 0.7f);
        }
      }
    }
  }
  __device__ static void finish(const Params& p, State& st, const GemmShape&, const EpiCtx& cx) {
    if constexpr (TRAIN) {
      float s[5] = {st.s_rec, st.s_met, st.s_mx, st.s_lc1, st.s_lc2};
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float t = warp_sum(s[k]);
        if (cx.lane == 0 && t != 0.f && p.sums) atomicAdd(p.sums + k, (double)t);
      }
    }
  }
};

// =====================================================================================================
// Split-K weight gradients (autograd dW = dY^T X of every nn.Linear), atomic-free: every unit (output tile x
// k-split) stores its fp32 accumulator tile to its own [128 x 256] slab of a scratch buffer with TMA stores;
// dw_reduce_kernel (elementwise.cu) adds the slabs of a tile in a fixed order, un-scales by 1/GS, clips to the
// real [out,in] extent and routes the bias column.  (The first version used fp32 atomics: 37 splits x 512 KB of
// same-address reductions cost ~23 of a 60 us kernel and made the gradients run-to-run non-deterministic.)
// =====================================================================================================
template <class Cfg>
struct EpiWeightGradPartial {
  static_assert(Cfg::ACC_TILES == 1 && Cfg::BLOCK_N == 256, "tile shape");
  static constexpr bool SPLIT = true;   // group g drains columns [128 g, 128 g + 128)
  static constexpr int CLUSTER = 1;
  struct Params {
    CUtensorMap part;  // fp32 [units * 128, 256], box [128 x 32]
  };
  static constexpr int SMEM_BYTES = kStageBytes;  // one [128 x 32] fp32 box: the 4-stage operand ring needs the rest
  struct State {};
  __device__ static void init(const Params&, State&, const GemmShape&, const EpiCtx&) {}
  __device__ static void unit(const Params& p, State&, const GemmShape& g, const UnitInfo& w, uint32_t tacc,
                              const EpiCtx& cx) {
    const int r = cx.q * 32 + cx.lane;
    const int u = w.split * (g.num_m_tiles * g.num_n_groups) + w.m_tile * g.num_n_groups + w.n_group;
#pragma unroll 1
    for (int c = cx.group * 128; c < cx.group * 128 + 128; c += 32) {
      float v[32];
      tmem_ld32(tacc + c, v);
      if (cx.tid == 0) tma_store_wait_read<0>();  // the previous box has left shared memory
      epi_bar_sync(cx, 0);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t addr = cx.smem + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                     "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      epi_bar_sync(cx, 1);
      if (cx.tid == 0) {
        tma_store_2d(&p.part, cx.smem, c, u * kBlockM);
        tma_store_commit();
      }
    }
  }
  __device__ static void finish(const Params&, State&, const GemmShape&, const EpiCtx& cx) {
    if (cx.tid == 0) tma_store_wait<0>();
  }
};

// =====================================================================================================
// Split-K weight-gradient accumulation (autograd dW = dY^T X of every nn.Linear): fp32 atomics into the
// flat gradient buffer, un-scaling by 1/GS, clipping to the real [out,in] extent.  `bias_col` (>= 0)
// names the operand column that holds the constant 1 of the first layers: its product is the bias grad.
// =====================================================================================================
template <class Cfg>
struct EpiWeightGrad {
  static_assert(Cfg::ACC_TILES == 1 && Cfg::BLOCK_N % 32 == 0, "tile shape");
  struct Params {
    float* dw;       // [M, ld]
    int ld;
    int n_valid;
    float scale;     // 1 / GS
    int bias_col;    // -1: none
    float* db;       // [M]
  };
  static constexpr int SMEM_BYTES = 0;
  static constexpr bool SPLIT = false;
  static constexpr int CLUSTER = 1;
  struct State {};
  __device__ static void init(const Params&, State&, const GemmShape&, const EpiCtx&) {}
  __device__ static void unit(const Params& p, State&, const GemmShape& g, const UnitInfo& w, uint32_t tacc,
                              const EpiCtx& cx) {
    const int row = w.m_tile * kBlockM + cx.q * 32 + cx.lane;
    const int n0 = w.n_group * Cfg::BLOCK_N;
    const bool valid = row < g.M;
    float* dst = p.dw + (size_t)(valid ? row : 0) * p.ld;
#pragma unroll 1
    for (int c = 0; c < Cfg::BLOCK_N; c += 32) {
      float v[32];
      tmem_ld32(tacc + c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = n0 + c + i;
          if (col < p.n_valid) atomicAdd(dst + col, v[i] * p.scale);
          else if (col == p.bias_col) atomicAdd(p.db + row, v[i] * p.scale);
        }
      }
    }
  }
  __device__ static void finish(const Params&, State&, const GemmShape&, const EpiCtx&) {}
};

}  // namespace pigan
