#include "elementwise.cuh"

#include "host_util.h"

#include <math.h>

namespace pigan {

namespace {

constexpr int kThreads = 256;
constexpr float kBnEps = 1e-5f;
constexpr float kLnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;
constexpr float kSlope = 0.2f;

__device__ __forceinline__ void ld_h8(const __half* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void st_h8(__half* p, const float* v) {
  uint4 u;
  __half2 h;
  h = __floats2half2_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2half2_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2half2_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2half2_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void ld_f8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void zero8(float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
}
constexpr int kUnroll = 2;  // rows a thread keeps in flight per loop trip in the streaming kernels
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lrelu_f(float x) { return fmaxf(x, kSlope * x); }

// Column-chunk mapping shared by the batch-reduction kernels: a thread owns 8 consecutive columns
// (chunk `ch`) of every row it visits; `cpr` threads cover a row, a block covers 256/cpr rows at a time.
struct ColMap {
  int cpr, rpb, ch, rg;
  __device__ ColMap(int C) {
    cpr = C >> 3;
    rpb = kThreads / cpr;
    ch = threadIdx.x % cpr;
    rg = threadIdx.x / cpr;
  }
};
// Sum acc8 over the block's row groups and store it to this block's row of the partial-sum scratch
// (part = scratch + blockIdx.x * ld + array offset).  No atomics: thousands of same-line float atomics serialise
// in L2 (measured ~3.7 ns each on B200, 70 us for one BatchNorm statistic); reduce_partials_kernel finishes the
// sum in a fixed order, which also makes the batch statistics and gradients run-to-run deterministic.
__device__ __forceinline__ void block_colsum_partial(const float* acc8, float* part, const ColMap& m, float* sm) {
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[(m.rg * m.cpr + m.ch) * 8 + i] = acc8[i];
  __syncthreads();
  if (m.rg == 0) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t[i] = 0.f;
      for (int gq = 0; gq < m.rpb; ++gq) t[i] += sm[(gq * m.cpr + m.ch) * 8 + i];
    }
    float4* o = reinterpret_cast<float4*>(part + m.ch * 8);
    o[0] = make_float4(t[0], t[1], t[2], t[3]);
    o[1] = make_float4(t[4], t[5], t[6], t[7]);
  }
  __syncthreads();
}
__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum_f(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    t = warp_sum_f(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

inline int grid_for_rows(int64_t rows, int rows_per_block, int max_blocks = 148 * 8) {
  int64_t b = (rows + rows_per_block - 1) / rows_per_block;
  if (b < 1) b = 1;
  return (int)(b < max_blocks ? b : max_blocks);
}

// dst[c] += mult * sum_b part[b * ld + off + c] for up to 8 column segments (one block per 32 columns)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(ReduceArgs a) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (col < a.ld)
    for (int b = ty; b < a.nblocks; b += 32) s += a.part[(size_t)b * a.ld + col];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < a.ld) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][tx];
    int off = 0;
    for (int g = 0; g < a.nseg; ++g) {
      if (col < off + a.seg[g].ncols) {
        if (a.seg[g].dst) a.seg[g].dst[col - off] += a.seg[g].mult * t;
        break;
      }
      off += a.seg[g].ncols;
    }
  }
}
void launch_reduce_partials(const ReduceArgs& a, cudaStream_t st) {
  note_launch(), reduce_partials_kernel<<<(a.ld + 31) / 32, 1024, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------ spectrum prep
__global__ void center_vec_kernel(const float* __restrict__ x, int S, int rows_used, float* __restrict__ cvec,
                                  int Kp) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 32
  const int col = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (col < S)
    for (int r = ty; r < rows_used; r += 32) s += x[(size_t)r * S + col];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < Kp) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][tx];
    cvec[col] = col < S ? t / (float)rows_used : 0.f;
  }
}

template <bool NOISE>
__global__ void cast_center_kernel(const float* __restrict__ x, const float* __restrict__ noise, float sigma,
                                   const float* __restrict__ cvec, const float* __restrict__ params,
                                   __half* __restrict__ xc, long long rows, int S, int P, int Kp, int vec2) {
  const int cpr = Kp >> 3;
  const long long total = rows * cpr;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / cpr;
    const int j0 = (int)(idx % cpr) * 8;
    float v[8];
    if (vec2 && j0 + 8 <= S) {
      // 8-byte vector loads: row pitch S*4 and j0*4 are multiples of 8 (vec2 = even S, 8-byte aligned bases)
      float c[8];
      ld_f8(cvec + j0, c);
      const float2* src = reinterpret_cast<const float2*>((NOISE ? noise : x) + row * S + j0);
      float2 t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = __ldg(src + i);
      if (NOISE) {
        const float2* tg = reinterpret_cast<const float2*>(x + j0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 g = __ldg(tg + i);
          v[2 * i] = (g.x + sigma * t[i].x) - c[2 * i];
          v[2 * i + 1] = (g.y + sigma * t[i].y) - c[2 * i + 1];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[2 * i] = t[i].x - c[2 * i];
          v[2 * i + 1] = t[i].y - c[2 * i + 1];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + i;
        float t = 0.f;
        if (j < S) {
          if (NOISE) t = (x[j] + sigma * noise[row * S + j]) - cvec[j];
          else t = x[row * S + j] - cvec[j];
        } else if (j < S + P) {
          t = params ? params[row * P + (j - S)] - kParamCenter : 0.f;
        } else if (j < S + P + 2) {
          t = 1.f;
        }
        v[i] = t;
      }
    }
    st_h8(xc + row * Kp + j0, v);
  }
}

// ------------------------------------------------------------------------------------------ weight packing
__global__ void pack_first_layer_kernel(const float* __restrict__ w, int ld_src, int S, int P, int wp_cols,
                                        int bias_cols, const float* __restrict__ b, const float* __restrict__ cvec,
                                        __half* __restrict__ out, int Kp, float* __restrict__ b_eff_out, int rows) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= rows) return;
  const float* wr = w + (size_t)i * ld_src;
  float dot = 0.f;
  for (int j = lane; j < S; j += 32) dot = fmaf(cvec[j], wr[j], dot);
  if (wp_cols)
    for (int e = lane; e < P; e += 32) dot = fmaf(kParamCenter, wr[S + e], dot);
  dot = warp_sum_f(dot);
  const float be = b[i] + dot;
  const __half hi = __float2half_rn(be);
  const __half lo = __float2half_rn(be - __half2float(hi));
  if (lane == 0 && b_eff_out) b_eff_out[i] = be;
  for (int j = lane; j < Kp; j += 32) {
    __half o = __float2half_rn(0.f);
    if (j < S) o = __float2half_rn(wr[j]);
    else if (j < S + P) o = wp_cols ? __float2half_rn(wr[j]) : o;
    else if (j == S + P) o = bias_cols ? hi : o;
    else if (j == S + P + 1) o = bias_cols ? lo : o;
    out[(size_t)i * Kp + j] = o;
  }
}

__global__ void cast_pad_kernel(const float* __restrict__ src, int ld_src, int ncols, __half* __restrict__ dst,
                                int ld_dst, int rows) {
  const long long total = (long long)rows * ld_dst;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / ld_dst), c = (int)(idx % ld_dst);
    dst[idx] = __float2half_rn(c < ncols ? src[(size_t)r * ld_src + c] : 0.f);
  }
}

__global__ void transpose_cast_kernel(const float* __restrict__ src, int rows, int cols, int ld_src,
                                      __half* __restrict__ dst, int ld_dst) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = by + k, c = bx + tx;
    tile[k][tx] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = bx + k, r = by + tx;  // dst[c][r]
    if (c < cols && r < rows) dst[(size_t)c * ld_dst + r] = __float2half_rn(tile[tx][k]);
  }
}

__global__ void extract_wp_kernel(const float* __restrict__ w, int ld_src, int S, int P, float* __restrict__ wp,
                                  int rows, int rows_pad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_pad * 4) return;
  const int i = idx >> 2, e = idx & 3;
  wp[idx] = (i < rows && e < P) ? w[(size_t)i * ld_src + S + e] : 0.f;
}

__global__ void copy_pad_f32_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_pad) dst[idx] = idx < n ? src[idx] : 0.f;
}

// ------------------------------------------------------------------------------------------ BatchNorm
__global__ void __launch_bounds__(kThreads) colstats_kernel(const __half* __restrict__ h, long long rows, int C,
                                                            float* __restrict__ part) {
  __shared__ float sm[kThreads * 8];
  const ColMap m(C);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < rows; r0 += kUnroll * stride) {
    float v[kUnroll][8];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {  // independent loads in flight before any use
      const long long r = r0 + u * stride;
      if (r < rows) ld_h8(h + r * C + m.ch * 8, v[u]);
      else zero8(v[u]);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += v[u][i];
        q[i] = fmaf(v[u][i], v[u][i], q[i]);
      }
  }
  part += (size_t)blockIdx.x * 2 * C;
  block_colsum_partial(s, part, m, sm);
  block_colsum_partial(q, part + C, m, sm);
}

__global__ void bn_finalize_kernel(BnFinalizeArgs a) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.C; c += gridDim.x * blockDim.x) {
    const double ms = (double)a.sum[c] / a.n;
    double var = (double)a.sumsq[c] / a.n - ms * ms;
    if (var < 0.0) var = 0.0;
    const float mean = (float)ms;
    const float mean_full = (float)(ms + (a.offset ? (double)a.offset[c] : 0.0));
    const float varf = (float)var;
    const float rstd = 1.0f / sqrtf(varf + kBnEps);
    const float sc = a.gamma[c] * rstd;
    a.mean[c] = mean;
    a.rstd[c] = rstd;
    a.scale[c] = sc;
    a.bias[c] = a.beta[c] - mean * sc;
    if (a.running_mean != nullptr) {
      const float unbiased = (float)(var * (a.n / (a.n > 1.0 ? a.n - 1.0 : 1.0)));
      float rm = a.running_mean[c], rv = a.running_var[c];
      for (int u = 0; u < a.num_updates; ++u) {
        rm = (1.f - kBnMomentum) * rm + kBnMomentum * mean_full;
        rv = (1.f - kBnMomentum) * rv + kBnMomentum * unbiased;
      }
      a.running_mean[c] = rm;
      a.running_var[c] = rv;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.num_batches_tracked != nullptr)
    *a.num_batches_tracked += a.num_updates;
}

__global__ void bn_eval_affine_kernel(const float* rm, const float* rv, const float* gamma, const float* beta,
                                      const float* offset, float* scale, float* bias, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + kBnEps);
  scale[c] = sc;
  bias[c] = beta[c] + ((offset ? offset[c] : 0.f) - rm[c]) * sc;
}

__global__ void __launch_bounds__(kThreads) bn_relu_apply_kernel(const __half* __restrict__ h,
                                                                 const float* __restrict__ scale,
                                                                 const float* __restrict__ bias,
                                                                 __half* __restrict__ out, long long rows, int C) {
  const ColMap m(C);
  float sc[8], bi[8];
  ld_f8(scale + m.ch * 8, sc);
  ld_f8(bias + m.ch * 8, bi);
  for (long long r = (long long)blockIdx.x * m.rpb + m.rg; r < rows; r += (long long)gridDim.x * m.rpb) {
    float v[8];
    ld_h8(h + r * C + m.ch * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(sc[i], v[i], bi[i]), 0.f);
    st_h8(out + r * C + m.ch * 8, v);
  }
}

// warp per row
__global__ void __launch_bounds__(kThreads) g_head_fwd_kernel(
    const __half* __restrict__ h2, const float* __restrict__ scale, const float* __restrict__ bias,
    const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ p_out,
    float* __restrict__ pden_out, const __half* __restrict__ xc, __half* __restrict__ tail_fake, long long rows,
    int C, int Kp, int S) {
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  // per-lane constants of the first 256-column slab stay in registers across rows
  float sc0[8], bi0[8], wa0[8], wb0[8], wc0[8], wd0[8];
  {
    const int c = lane * 8;
    ld_f8(scale + c, sc0);
    ld_f8(bias + c, bi0);
    ld_f8(w3 + c, wa0);
    ld_f8(w3 + C + c, wb0);
    ld_f8(w3 + 2 * C + c, wc0);
    ld_f8(w3 + 3 * C + c, wd0);
  }
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    {
      float v[8];
      ld_h8(h2 + row * C + lane * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = fmaxf(fmaf(sc0[i], v[i], bi0[i]), 0.f);
        d0 = fmaf(a, wa0[i], d0);
        d1 = fmaf(a, wb0[i], d1);
        d2 = fmaf(a, wc0[i], d2);
        d3 = fmaf(a, wd0[i], d3);
      }
    }
    for (int c = 256 + lane * 8; c < C; c += 256) {
      float v[8], sc[8], bi[8], wa[8], wb[8], wc[8], wd[8];
      ld_h8(h2 + row * C + c, v);
      ld_f8(scale + c, sc);
      ld_f8(bias + c, bi);
      ld_f8(w3 + c, wa);
      ld_f8(w3 + C + c, wb);
      ld_f8(w3 + 2 * C + c, wc);
      ld_f8(w3 + 3 * C + c, wd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = fmaxf(fmaf(sc[i], v[i], bi[i]), 0.f);
        d0 = fmaf(a, wa[i], d0);
        d1 = fmaf(a, wb[i], d1);
        d2 = fmaf(a, wc[i], d2);
        d3 = fmaf(a, wd[i], d3);
      }
    }
    d0 = warp_sum_f(d0); d1 = warp_sum_f(d1); d2 = warp_sum_f(d2); d3 = warp_sum_f(d3);
    float p[4] = {tanhf(d0 + __ldg(b3 + 0)), tanhf(d1 + __ldg(b3 + 1)), tanhf(d2 + __ldg(b3 + 2)),
                  tanhf(d3 + __ldg(b3 + 3))};
    float pd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) pd[e] = (p[e] + 1.0f) / 2.0f * 0.6f + 2.2f;  // data_loader.py:238-252
    if (lane == 0) {
      *reinterpret_cast<float4*>(p_out + row * 4) = make_float4(p[0], p[1], p[2], p[3]);
      if (pden_out) *reinterpret_cast<float4*>(pden_out + row * 4) = make_float4(pd[0], pd[1], pd[2], pd[3]);
    }
    if (tail_fake != nullptr && lane < 8) {
      const int t0 = Kp - 64;  // first spectrum-operand column held by the tail
      float v[8];
      ld_h8(xc + row * Kp + t0 + lane * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = t0 + lane * 8 + i - S;
        if (e >= 0 && e < 4) v[i] = pd[e] - kParamCenter;
      }
      st_h8(tail_fake + row * 64 + lane * 8, v);
    }
  }
}

// Generator head backward (tanh, Linear(256,4), ReLU) fused with the BatchNorm-2 backward.  Two passes over h2
// instead of storing the pre-projection gradient dy in fp16: BatchNorm's backward removes the batch-mean and
// x-hat components of dy, which here carry most of its norm (the adversarial gradient pushes every sample the
// same way), so rounding dy before the projection would be amplified in what survives it.
//   APPLY = false: column sums of dy and dy*xhat, dW3, db3, range-loss sum
//   APPLY = true : dh = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)) -> fp16, db2, dgamma, dbeta
template <bool APPLY>
__global__ void __launch_bounds__(kThreads) g_head_bwd_kernel(GHeadBwdArgs a) {
  __shared__ float sm[kThreads * 8];
  const ColMap m(a.C);
  float sc[8], bi[8], mu[8], rs[8], w[4][8];
  ld_f8(a.scale + m.ch * 8, sc);
  ld_f8(a.bias + m.ch * 8, bi);
  ld_f8(a.mean + m.ch * 8, mu);
  ld_f8(a.rstd + m.ch * 8, rs);
#pragma unroll
  for (int j = 0; j < 4; ++j) ld_f8(a.w3 + j * a.C + m.ch * 8, w[j]);
  float dw[4][8], sdy[8], sdyx[8];  // APPLY: sdy accumulates dh (-> db2), sdyx/dw unused
  float gr[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sdy[i] = sdyx[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) dw[j][i] = 0.f;
  }
  if (APPLY) {
    ld_f8(a.gamma + m.ch * 8, gr);
    ld_f8(a.sum_dy + m.ch * 8, m1);
    ld_f8(a.sum_dyx + m.ch * 8, m2);
    if (blockIdx.x == 0 && m.rg == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (a.dgamma) a.dgamma[m.ch * 8 + i] += m2[i] * a.inv_gs;
        if (a.dbeta) a.dbeta[m.ch * 8 + i] += m1[i] * a.inv_gs;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      gr[i] *= rs[i];
      m1[i] = (float)((double)m1[i] * a.inv_n);
      m2[i] = (float)((double)m2[i] * a.inv_n);
    }
  }
  float db[4] = {0.f, 0.f, 0.f, 0.f};
  float range_acc = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long stride = (long long)gridDim.x * m.rpb;
  constexpr int U = 2;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < a.rows; r0 += U * stride) {
    float h[U][8];
    float4 p4[U], d4[U], l4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.rows) {
        ld_h8(a.h2 + r * a.C + m.ch * 8, h[u]);
        p4[u] = __ldg(reinterpret_cast<const float4*>(a.p) + r);
        d4[u] = a.dpden ? __ldg(reinterpret_cast<const float4*>(a.dpden) + r) : z4;
        l4[u] = a.dp_lc ? __ldg(reinterpret_cast<const float4*>(a.dp_lc) + r) : z4;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r >= a.rows) break;
      const float p[4] = {p4[u].x, p4[u].y, p4[u].z, p4[u].w};
      const float dd[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
      const float ll[4] = {l4[u].x, l4[u].y, l4[u].z, l4[u].w};
      float dpre[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = fmaxf(-p[j], 0.f), hi = fmaxf(p[j] - 1.f, 0.f);        // loss.py:121-123
        const float dp = (0.5f * 0.6f) * dd[j] + ll[j] + (2.f * hi - 2.f * lo) * a.range_mult;
        dpre[j] = dp * (1.f - p[j] * p[j]);                                      // tanh backward
        if (!APPLY && m.ch == 0) {
          db[j] += dpre[j];
          range_acc += lo * lo + hi * hi;
        }
      }
      float out[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float act = fmaxf(fmaf(sc[i], h[u][i], bi[i]), 0.f);
        float da = dpre[0] * w[0][i];
        da = fmaf(dpre[1], w[1][i], da);
        da = fmaf(dpre[2], w[2][i], da);
        da = fmaf(dpre[3], w[3][i], da);
        const float g = act > 0.f ? da : 0.f;
        const float xh = (h[u][i] - mu[i]) * rs[i];
        if (APPLY) {
          const float dh = gr[i] * (g - m1[i] - xh * m2[i]);
          out[i] = dh;
          sdy[i] += dh;
        } else {
          sdy[i] += g;
          sdyx[i] = fmaf(g, xh, sdyx[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) dw[j][i] = fmaf(dpre[j], act, dw[j][i]);
        }
      }
      if (APPLY) st_h8(a.dy2 + r * a.C + m.ch * 8, out);
    }
  }
  if (APPLY) {
    block_colsum_partial(sdy, a.part + (size_t)blockIdx.x * a.C, m, sm);
    return;
  }
  float* part = a.part + (size_t)blockIdx.x * 6 * a.C;
  block_colsum_partial(sdy, part, m, sm);
  block_colsum_partial(sdyx, part + a.C, m, sm);
#pragma unroll
  for (int j = 0; j < 4; ++j) block_colsum_partial(dw[j], part + (2 + j) * a.C, m, sm);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float t = block_sum(db[j], sm);
    if (threadIdx.x == 0) atomicAdd(a.db3 + j, t * a.inv_gs);
  }
  const float t = block_sum(range_acc, sm);
  if (threadIdx.x == 0 && a.range_sum) atomicAdd(a.range_sum, (double)t);
}

__global__ void __launch_bounds__(kThreads) bn_bwd_stats_kernel(
    const __half* __restrict__ da, const __half* __restrict__ h, const float* __restrict__ scale,
    const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ rstd,
    float* __restrict__ part, long long rows, int C) {
  __shared__ float sm[kThreads * 8];
  const ColMap m(C);
  float sc[8], bi[8], mu[8], rs[8], s1[8], s2[8];
  ld_f8(scale + m.ch * 8, sc);
  ld_f8(bias + m.ch * 8, bi);
  ld_f8(mean + m.ch * 8, mu);
  ld_f8(rstd + m.ch * 8, rs);
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < rows; r0 += kUnroll * stride) {
    float g[kUnroll][8], x[kUnroll][8];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long r = r0 + u * stride;
      if (r < rows) {
        ld_h8(da + r * C + m.ch * 8, g[u]);
        ld_h8(h + r * C + m.ch * 8, x[u]);
      } else {
        zero8(g[u]);
        zero8(x[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dy = fmaf(sc[i], x[u][i], bi[i]) > 0.f ? g[u][i] : 0.f;
        s1[i] += dy;
        s2[i] = fmaf(dy, (x[u][i] - mu[i]) * rs[i], s2[i]);
      }
  }
  part += (size_t)blockIdx.x * 2 * C;
  block_colsum_partial(s1, part, m, sm);
  block_colsum_partial(s2, part + C, m, sm);
}

__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(BnBwdArgs a) {
  __shared__ float sm[kThreads * 8];
  const ColMap m(a.C);
  float sc[8], bi[8], mu[8], rs[8], gr[8], m1[8], m2[8], sdh[8];
  ld_f8(a.scale + m.ch * 8, sc);
  ld_f8(a.bias + m.ch * 8, bi);
  ld_f8(a.mean + m.ch * 8, mu);
  ld_f8(a.rstd + m.ch * 8, rs);
  ld_f8(a.gamma + m.ch * 8, gr);
  ld_f8(a.sum_dy + m.ch * 8, m1);
  ld_f8(a.sum_dyx + m.ch * 8, m2);
  if (blockIdx.x == 0 && m.rg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (a.dgamma) a.dgamma[m.ch * 8 + i] += m2[i] * a.inv_gs;
      if (a.dbeta) a.dbeta[m.ch * 8 + i] += m1[i] * a.inv_gs;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    gr[i] *= rs[i];
    m1[i] = (float)((double)m1[i] * a.inv_n);
    m2[i] = (float)((double)m2[i] * a.inv_n);
    sdh[i] = 0.f;
  }
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < a.rows; r0 += kUnroll * stride) {
    float g[kUnroll][8], x[kUnroll][8];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long r = r0 + u * stride;
      if (r < a.rows) {
        ld_h8(a.dy + r * a.C + m.ch * 8, g[u]);
        ld_h8(a.h + r * a.C + m.ch * 8, x[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long r = r0 + u * stride;
      if (r >= a.rows) break;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float dy = g[u][i];
        if (a.relu_mask) dy = fmaf(sc[i], x[u][i], bi[i]) > 0.f ? dy : 0.f;
        const float xh = (x[u][i] - mu[i]) * rs[i];
        const float dh = gr[i] * (dy - m1[i] - xh * m2[i]);
        g[u][i] = dh;
        sdh[i] += dh;
      }
      st_h8(a.dh + r * a.C + m.ch * 8, g[u]);
    }
  }
  if (a.dbias) block_colsum_partial(sdh, a.part + (size_t)blockIdx.x * a.C, m, sm);
}

// ------------------------------------------------------------------------------------------ discriminator
__global__ void __launch_bounds__(kThreads) d_l2_bwd_kernel(const __half* __restrict__ z2,
                                                            const float* __restrict__ dlogit,
                                                            const float* __restrict__ w3, __half* __restrict__ dh2,
                                                            float* __restrict__ dw3, float* __restrict__ db2,
                                                            float* __restrict__ db3, long long rows, int C,
                                                            float inv_gs, float* __restrict__ part) {
  __shared__ float sm[kThreads * 8];
  const ColMap m(C);
  float w[8], sw[8], sb[8];
  ld_f8(w3 + m.ch * 8, w);
#pragma unroll
  for (int i = 0; i < 8; ++i) sw[i] = sb[i] = 0.f;
  float s3 = 0.f;
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < rows; r0 += kUnroll * stride) {
    float z[kUnroll][8], dl[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long r = r0 + u * stride;
      if (r < rows) {
        dl[u] = __ldg(dlogit + r);
        ld_h8(z2 + r * C + m.ch * 8, z[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long r = r0 + u * stride;
      if (r >= rows) break;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dh = dl[u] * w[i] * (z[u][i] > 0.f ? 1.f : kSlope);
        o[i] = dh;
        sw[i] = fmaf(dl[u], z[u][i], sw[i]);
        sb[i] += dh;
      }
      st_h8(dh2 + r * C + m.ch * 8, o);
      if (m.ch == 0) s3 += dl[u];
    }
  }
  if (dw3 != nullptr) {
    part += (size_t)blockIdx.x * 2 * C;
    block_colsum_partial(sw, part, m, sm);
    block_colsum_partial(sb, part + C, m, sm);
    const float t = block_sum(s3, sm);
    if (threadIdx.x == 0) atomicAdd(db3, t * inv_gs);
  }
}

// ------------------------------------------------------------------------------------------ forward model
// first layer (K = 4, forward_model.py:30-33): warp per row, lane owns 8 of the 256 columns
__global__ void __launch_bounds__(kThreads) f_l1_kernel(const float* __restrict__ p, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ lnw,
                                                        const float* __restrict__ lnb, __half* __restrict__ out,
                                                        long long rows, int C) {
  const int lane = threadIdx.x & 31;
  float4 wr[8];
  float b[8], gm[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wr[i] = __ldg(reinterpret_cast<const float4*>(w1) + lane * 8 + i);
  ld_f8(b1 + lane * 8, b);
  ld_f8(lnw + lane * 8, gm);
  ld_f8(lnb + lane * 8, bt);
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p) + row);
    float h[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = b[i];
      t = fmaf(q.x, wr[i].x, t);
      t = fmaf(q.y, wr[i].y, t);
      t = fmaf(q.z, wr[i].z, t);
      t = fmaf(q.w, wr[i].w, t);
      h[i] = t;
      s += t;
    }
    const float mean = warp_sum_f(s) / (float)C;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = h[i] - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = 1.0f / sqrtf(warp_sum_f(v) / (float)C + kLnEps);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = lrelu_f(fmaf((h[i] - mean) * rstd, gm[i], bt[i]));
    st_h8(out + row * C + lane * 8, h);
  }
}

__global__ void __launch_bounds__(kThreads) ln_lrelu_apply_kernel(__half* __restrict__ h,
                                                                  const float* __restrict__ rowstats, int n_tiles,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, long long rows,
                                                                  int N) {
  const ColMap m(N);
  float gm[8], bt[8];
  ld_f8(gamma + m.ch * 8, gm);
  ld_f8(beta + m.ch * 8, bt);
  const float inv_n = 1.0f / (float)N;
  for (long long r = (long long)blockIdx.x * m.rpb + m.rg; r < rows; r += (long long)gridDim.x * m.rpb) {
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(rowstats) + r * n_tiles + t);
      s1 += st.x;
      s2 += st.y;
    }
    const float mean = s1 * inv_n;
    const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + kLnEps);
    float v[8];
    ld_h8(h + r * N + m.ch * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = lrelu_f(fmaf((v[i] - mean) * rstd, gm[i], bt[i]));
    st_h8(h + r * N + m.ch * 8, v);
  }
}

// ------------------------------------------------------------------------------------------ optimiser
__global__ void __launch_bounds__(kThreads) sumsq_kernel(const float* __restrict__ g, long long n, double* out) {
  __shared__ float sm[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s = fmaf(g[i], g[i], s);
  const float t = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(out, (double)t);
}

__global__ void __launch_bounds__(kThreads) clip_adam_kernel(AdamArgs a) {
  // torch.nn.utils.clip_grad_norm_(max_norm) then optim.Adam.step() (train_pigan.py:142-143,186-187)
  const float total = (float)sqrt(*a.total_sq);
  float coef = a.max_norm / (total + 1e-6f);
  coef = coef > 1.f ? 1.f : coef;
  const float step = (float)((double)a.lr / a.bias_c1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(a.bias_c2));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
       i += (long long)gridDim.x * blockDim.x) {
    const float g = a.g[i] * coef;
    const float mm = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
    const float vv = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    a.g[i] = g;
    a.m[i] = mm;
    a.v[i] = vv;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + a.eps;
    a.p[i] = a.p[i] - step * (mm / denom);
  }
}

__global__ void dw_fixup_kernel(float* dw, int ld, int S, int P, const float* db, const float* cvec, int rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = S + P;
  if (idx >= rows * per) return;
  const int i = idx / per, j = idx % per;
  dw[(size_t)i * ld + j] += db[i] * (j < S ? cvec[j] : kParamCenter);
}

__global__ void loss_finalize_kernel(LossFinalizeArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double* s = a.sums;
  const float d_loss = (float)s[0];
  const float adv = (float)s[1];
  const float rec = (float)(s[2] / (a.batch * a.S));
  const float met = (float)(s[3] / (a.batch * a.Mt));
  const float mxw = a.S >= 3 ? (float)(s[4] / (a.batch * (a.S - 2))) : 0.f;
  const float lc = (float)(s[5] / a.batch) + (float)(s[6] / a.batch);
  const float rng = (float)(s[7] / (a.batch * a.P));
  const float kl = 0.f;
  const float g_loss = adv + a.lam_recon * rec + a.lam_phys_spec * rec + a.lam_phys_metrics * met +
                       a.lam_maxwell * mxw + a.lam_lc * lc + a.lam_range * rng + a.lam_kl * kl;
  a.out9[0] = d_loss;
  a.out9[1] = g_loss;
  a.out9[2] = adv;
  a.out9[3] = rec;
  a.out9[4] = met;
  a.out9[5] = mxw;
  a.out9[6] = lc;
  a.out9[7] = rng;
  a.out9[8] = kl;
}

__global__ void score_finish_kernel(const float* __restrict__ p, const float* __restrict__ err, long long rows,
                                    int P, int* __restrict__ viol, float* __restrict__ cons) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int v = 0;
  for (int j = 0; j < P; ++j) {
    const float x = p[r * P + j];
    v += (x < 0.f) | (x > 1.f);
  }
  if (viol) viol[r] = v;
  if (cons) cons[r] = 1.0f / (1.0f + err[r]);
}

}  // namespace

// =========================================================================================== launchers
void launch_center_vec(const float* x, int64_t rows, int S, int rows_used, float* cvec, int Kp, cudaStream_t st) {
  if (rows_used > rows) rows_used = (int)rows;
  note_launch(), center_vec_kernel<<<(Kp + 31) / 32, 1024, 0, st>>>(x, S, rows_used, cvec, Kp);
}
void launch_cast_center(const float* x, const float* cvec, const float* params, __half* xc, int64_t rows, int S,
                        int P, int Kp, cudaStream_t st) {
  const int64_t total = rows * (Kp / 8);
  const int grid = (int)((total + kThreads - 1) / kThreads < 148 * 16 ? (total + kThreads - 1) / kThreads : 148 * 16);
  note_launch(), cast_center_kernel<false><<<grid > 0 ? grid : 1, kThreads, 0, st>>>(x, nullptr, 0.f, cvec, params, xc, rows, S, P, Kp,
                                                                             (S % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0) ? 1 : 0);
}
void launch_cast_center_noise(const float* target, const float* noise, float sigma, const float* cvec, __half* xc,
                              float*, int64_t rows, int S, int P, int Kp, cudaStream_t st) {
  const int64_t total = rows * (Kp / 8);
  const int grid = (int)((total + kThreads - 1) / kThreads < 148 * 16 ? (total + kThreads - 1) / kThreads : 148 * 16);
  note_launch(), cast_center_kernel<true><<<grid > 0 ? grid : 1, kThreads, 0, st>>>(target, noise, sigma, cvec, nullptr, xc, rows, S, P, Kp,
                                                                            (S % 2 == 0 && (reinterpret_cast<uintptr_t>(target) & 7) == 0 &&
                                                                             (reinterpret_cast<uintptr_t>(noise) & 7) == 0) ? 1 : 0);
}
void launch_pack_first_layer(const float* w, int ld_src, int S, int P, int wp_cols, int bias_cols, const float* b,
                             const float* cvec, __half* out, int Kp, float* b_eff_out, int rows, cudaStream_t st) {
  note_launch(), pack_first_layer_kernel<<<(rows + 7) / 8, 256, 0, st>>>(w, ld_src, S, P, wp_cols, bias_cols, b, cvec, out, Kp,
                                                          b_eff_out, rows);
}
void launch_cast_pad(const float* src, int ld_src, int ncols, __half* dst, int ld_dst, int rows, cudaStream_t st) {
  const long long total = (long long)rows * ld_dst;
  int grid = (int)((total + kThreads - 1) / kThreads);
  if (grid > 148 * 8) grid = 148 * 8;
  note_launch(), cast_pad_kernel<<<grid, kThreads, 0, st>>>(src, ld_src, ncols, dst, ld_dst, rows);
}
void launch_transpose_cast(const float* src, int rows, int cols, int ld_src, __half* dst, int ld_dst,
                           cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  note_launch(), transpose_cast_kernel<<<grid, 256, 0, st>>>(src, rows, cols, ld_src, dst, ld_dst);
}
void launch_extract_wp(const float* w, int ld_src, int S, int P, float* wp, int rows, int rows_pad, cudaStream_t st) {
  note_launch(), extract_wp_kernel<<<(rows_pad * 4 + 255) / 256, 256, 0, st>>>(w, ld_src, S, P, wp, rows, rows_pad);
}
void launch_copy_pad_f32(const float* src, int n, float* dst, int n_pad, cudaStream_t st) {
  note_launch(), copy_pad_f32_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(src, n, dst, n_pad);
}
void launch_colstats(const __half* h, int64_t rows, int C, float* sum, float* sumsq, float* part, cudaStream_t st) {
  const int rpb = kThreads / (C / 8);
  const int grid = grid_for_rows(rows, rpb * 8, kPartBlocks);
  note_launch(), colstats_kernel<<<grid, kThreads, 0, st>>>(h, rows, C, part);
  ReduceArgs r{part, grid, 2 * C, 2, {{sum, C, 1.f}, {sumsq, C, 1.f}}};
  launch_reduce_partials(r, st);
}
void launch_bn_finalize(const BnFinalizeArgs& a, cudaStream_t st) {
  note_launch(), bn_finalize_kernel<<<(a.C + 255) / 256, 256, 0, st>>>(a);
}
void launch_bn_eval_affine(const float* rm, const float* rv, const float* gamma, const float* beta,
                           const float* offset, float* scale, float* bias, int C, cudaStream_t st) {
  note_launch(), bn_eval_affine_kernel<<<(C + 255) / 256, 256, 0, st>>>(rm, rv, gamma, beta, offset, scale, bias, C);
}
void launch_bn_relu_apply(const __half* h, const float* scale, const float* bias, __half* a, int64_t rows, int C,
                          cudaStream_t st) {
  const int rpb = kThreads / (C / 8);
  note_launch(), bn_relu_apply_kernel<<<grid_for_rows(rows, rpb * 4), kThreads, 0, st>>>(h, scale, bias, a, rows, C);
}
void launch_g_head_fwd(const __half* h2, const float* scale, const float* bias, const float* w3, const float* b3,
                       float* p_out, float* pden_out, const __half* xc, __half* tail_fake, int64_t rows, int C,
                       int Kp, int S, cudaStream_t st) {
  note_launch(), g_head_fwd_kernel<<<grid_for_rows(rows, 8 * 4), kThreads, 0, st>>>(h2, scale, bias, w3, b3, p_out, pden_out, xc,
                                                                     tail_fake, rows, C, Kp, S);
}
void launch_g_head_bwd(const GHeadBwdArgs& a, bool apply, cudaStream_t st) {
  const int rpb = kThreads / (a.C / 8);
  const int grid = grid_for_rows(a.rows, rpb * 8, kPartBlocks);
  if (apply) {
    note_launch(), g_head_bwd_kernel<true><<<grid, kThreads, 0, st>>>(a);
    ReduceArgs r{a.part, grid, a.C, 1, {{a.dbias, a.C, a.inv_gs}}};
    launch_reduce_partials(r, st);
  } else {
    note_launch(), g_head_bwd_kernel<false><<<grid, kThreads, 0, st>>>(a);
    ReduceArgs r{a.part, grid, 6 * a.C, 6,
                 {{a.sum_dy, a.C, 1.f}, {a.sum_dyx, a.C, 1.f}, {a.dw3, a.C, a.inv_gs}, {a.dw3 + a.C, a.C, a.inv_gs},
                  {a.dw3 + 2 * a.C, a.C, a.inv_gs}, {a.dw3 + 3 * a.C, a.C, a.inv_gs}}};
    launch_reduce_partials(r, st);
  }
}
void launch_bn_bwd_stats(const __half* da, const __half* h, const float* scale, const float* bias,
                         const float* mean, const float* rstd, float* sum_dy, float* sum_dyx, int64_t rows, int C,
                         float* part, cudaStream_t st) {
  const int rpb = kThreads / (C / 8);
  const int grid = grid_for_rows(rows, rpb * 8, kPartBlocks);
  note_launch(), bn_bwd_stats_kernel<<<grid, kThreads, 0, st>>>(da, h, scale, bias, mean, rstd, part, rows, C);
  ReduceArgs r{part, grid, 2 * C, 2, {{sum_dy, C, 1.f}, {sum_dyx, C, 1.f}}};
  launch_reduce_partials(r, st);
}
void launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t st) {
  const int rpb = kThreads / (a.C / 8);
  const int grid = grid_for_rows(a.rows, rpb * 8, kPartBlocks);
  note_launch(), bn_bwd_apply_kernel<<<grid, kThreads, 0, st>>>(a);
  if (a.dbias) {
    ReduceArgs r{a.part, grid, a.C, 1, {{a.dbias, a.C, a.inv_gs}}};
    launch_reduce_partials(r, st);
  }
}
void launch_d_l2_bwd(const __half* z2, const float* dlogit, const float* w3, __half* dh2, float* dw3, float* db2,
                     float* db3, int64_t rows, int C, float inv_gs, float* part, cudaStream_t st) {
  const int rpb = kThreads / (C / 8);
  const int grid = grid_for_rows(rows, rpb * 8, kPartBlocks);
  note_launch(), d_l2_bwd_kernel<<<grid, kThreads, 0, st>>>(z2, dlogit, w3, dh2, dw3, db2, db3, rows, C, inv_gs, part);
  if (dw3 != nullptr) {
    ReduceArgs r{part, grid, 2 * C, 2, {{dw3, C, inv_gs}, {db2, C, inv_gs}}};
    launch_reduce_partials(r, st);
  }
}
void launch_f_l1(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb, __half* out,
                 int64_t rows, int C, cudaStream_t st) {
  note_launch(), f_l1_kernel<<<grid_for_rows(rows, 8 * 4), kThreads, 0, st>>>(p, w1, b1, lnw, lnb, out, rows, C);
}
void launch_ln_lrelu_apply(__half* h, const float* rowstats, int n_tiles, const float* gamma, const float* beta,
                           int64_t rows, int N, cudaStream_t st) {
  const int rpb = kThreads / (N / 8);
  note_launch(), ln_lrelu_apply_kernel<<<grid_for_rows(rows, rpb * 4), kThreads, 0, st>>>(h, rowstats, n_tiles, gamma, beta, rows, N);
}
void launch_sumsq(const float* g, int64_t n, double* out, cudaStream_t st) {
  int grid = (int)((n + kThreads * 4 - 1) / (kThreads * 4));
  if (grid > 148 * 2) grid = 148 * 2;
  if (grid < 1) grid = 1;
  note_launch(), sumsq_kernel<<<grid, kThreads, 0, st>>>(g, n, out);
}
void launch_clip_adam(const AdamArgs& a, cudaStream_t st) {
  int grid = (int)((a.n + kThreads - 1) / kThreads);
  if (grid > 148 * 8) grid = 148 * 8;
  note_launch(), clip_adam_kernel<<<grid, kThreads, 0, st>>>(a);
}
void launch_dw_fixup(float* dw, int ld, int S, int P, const float* db, const float* cvec, int rows,
                     cudaStream_t st) {
  const int total = rows * (S + P);
  note_launch(), dw_fixup_kernel<<<(total + 255) / 256, 256, 0, st>>>(dw, ld, S, P, db, cvec, rows);
}
void launch_loss_finalize(const LossFinalizeArgs& a, cudaStream_t st) { note_launch(), loss_finalize_kernel<<<1, 32, 0, st>>>(a); }
void launch_score_finish(const float* p, const float* err, int64_t rows, int P, int32_t* violations,
                         float* consistency, cudaStream_t st) {
  note_launch(), score_finish_kernel<<<(int)((rows + 255) / 256), 256, 0, st>>>(p, err, rows, P, violations, consistency);
}

}  // namespace pigan
