#include "elementwise.cuh"

#include "host_util.h"

#include <math.h>

namespace pigan {

namespace {
// first statement of every kernel here: wait for the kernel this one was launched behind (host_util.h: launch_k)
// (An explicit early griddepcontrol.launch_dependents was measured and is slower: parked blocks of the successor
// take slots from this kernel's later waves; the implicit trigger at block exit is what is used.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// p[i] (sigmoid output) -> scale * grad_out[i] * p (1 - p): the upstream gradient carried through the sigmoid
__global__ void sigmoid_bwd_kernel(float* __restrict__ p, const float* __restrict__ g, float scale, long long n) {
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = p[i];
    p[i] = scale * g[i] * v * (1.0f - v);
  }
}
__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, float mult, long long n) {
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = mult * src[i];
}

// dst[c] += mult * sum_b part[b * ld + off + c] for up to 8 column segments (one block per 32 columns)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(ReduceArgs a) {
  pdl_wait();
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
  launch_k(reduce_partials_kernel, (a.ld + 31) / 32, 1024, 0, st, a);
}

// ------------------------------------------------------------------------------------------ spectrum prep
__global__ void center_vec_kernel(const float* __restrict__ x, int S, int rows_used, float* __restrict__ cvec,
                                  int Kp) {
  pdl_wait();
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
                                   __half* __restrict__ xc, long long rows, int S, int P, int Kp, int vec2,
                                   long long x_stride) {
  // NOISE: row r of the operand is x[r * x_stride + j] + sigma * noise[r, j] (x_stride = 0: one target row for all)
  pdl_wait();
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
        const float2* tg = reinterpret_cast<const float2*>(x + row * x_stride + j0);
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
          if (NOISE) t = (x[row * x_stride + j] + sigma * noise[row * S + j]) - cvec[j];
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

// Philox4x32-10 (Salmon et al., SC'11) keyed by the search seed, counter = (global candidate index, column block):
// the noise of candidate i does not depend on how candidates are cut into chunks, shards or ranks.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// 4 uniform words -> 4 standard normals (Box-Muller)
__device__ __forceinline__ void normal4(uint4 u, float* z) {
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)u.x + 0.5f) * k, u1 = (float)u.y * k;
  const float u2 = ((float)u.z + 0.5f) * k, u3 = (float)u.w * k;
  const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}
// xc[r, j] = fp16((target[j] + sigma * z(first + r, j)) - cvec[j]) (unified_evaluator.py:453-455 with in-kernel noise);
// noise_out (optional, [rows, S] fp32) receives z itself — tests replay it through the explicit-noise path
__global__ void cast_center_philox_kernel(const float* __restrict__ target, float sigma, unsigned long long seed,
                                          long long first, const float* __restrict__ cvec, __half* __restrict__ xc,
                                          float* __restrict__ noise_out, long long rows, int S, int Kp) {
  pdl_wait();
  const int cpr = Kp >> 3;
  const long long total = rows * cpr;
  const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / cpr;
    const int j0 = (int)(idx % cpr) * 8;
    const unsigned long long gi = (unsigned long long)(first + row);
    float z[8], v[8];
    normal4(philox4x32_10(make_uint4((unsigned int)gi, (unsigned int)(gi >> 32), (unsigned int)(j0 >> 2), 0u), key), z);
    normal4(philox4x32_10(make_uint4((unsigned int)gi, (unsigned int)(gi >> 32), (unsigned int)(j0 >> 2) + 1u, 0u), key),
            z + 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = j0 + i;
      float t = 0.f;
      if (j < S) {
        t = (target[j] + sigma * z[i]) - cvec[j];
        if (noise_out) noise_out[row * S + j] = z[i];
      } else if (j >= S + 4 && j < S + 6) {
        t = 1.f;
      }
      v[i] = t;
    }
    st_h8(xc + row * Kp + j0, v);
  }
}

// ---- running top-k of the inverse-design search (engine.cu: pigan_inverse_design_search)
__global__ void search_init_kernel(float* __restrict__ scores, long long* __restrict__ iota, long long* __restrict__ best_idx,
                                   float* __restrict__ params, long long total, int k) {
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    iota[i] = i;
    if (i < k) {
      scores[i] = __int_as_float(0x7f800000);
      best_idx[i] = -1;
      params[4 * i] = params[4 * i + 1] = params[4 * i + 2] = params[4 * i + 3] = 0.f;
    }
  }
}
__global__ void fill_inf_kernel(float* __restrict__ s, long long n) {
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s[i] = __int_as_float(0x7f800000);
}
// winners (positions in [0, k + cap)) -> temporaries
__global__ void search_gather_kernel(const long long* __restrict__ pos, const float* __restrict__ params,
                                     const long long* __restrict__ best_idx, long long group_base, int k,
                                     long long* __restrict__ tmp_idx, float* __restrict__ tmp_params) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const long long q = pos[i];
  tmp_idx[i] = q < k ? best_idx[q] : group_base + (q - k);
  *reinterpret_cast<float4*>(tmp_params + 4 * i) = *reinterpret_cast<const float4*>(params + 4 * q);
}
__global__ void search_commit_kernel(const float* __restrict__ sel_scores, const long long* __restrict__ tmp_idx,
                                     const float* __restrict__ tmp_params, int k, float* __restrict__ scores,
                                     long long* __restrict__ best_idx, float* __restrict__ params) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  scores[i] = sel_scores[i];
  best_idx[i] = tmp_idx[i];
  *reinterpret_cast<float4*>(params + 4 * i) = *reinterpret_cast<const float4*>(tmp_params + 4 * i);
}

// ------------------------------------------------------------------------------------------ weight packing
__global__ void pack_first_layer_kernel(const float* __restrict__ w, int ld_src, int S, int P, int wp_cols,
                                        int bias_cols, const float* __restrict__ b, const float* __restrict__ cvec,
                                        __half* __restrict__ out, int Kp, float* __restrict__ b_eff_out, int rows) {
  pdl_wait();
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
  pdl_wait();
  const long long total = (long long)rows * ld_dst;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / ld_dst), c = (int)(idx % ld_dst);
    dst[idx] = __float2half_rn(c < ncols ? src[(size_t)r * ld_src + c] : 0.f);
  }
}

__global__ void transpose_cast_kernel(const float* __restrict__ src, int rows, int cols, int ld_src,
                                      __half* __restrict__ dst, int ld_dst) {
  pdl_wait();
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
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_pad * 4) return;
  const int i = idx >> 2, e = idx & 3;
  wp[idx] = (i < rows && e < P) ? w[(size_t)i * ld_src + S + e] : 0.f;
}

// One launch per network and step for all operand copies of its weights (first layer packed with the effective
// bias, second layer cast, its transpose for dX, the parameter columns for the G-step gradient): block ranges
// [0,nA) first layer (8 rows per block), [nA,nA+nB) cast, then 32x32 transpose tiles, then the parameter columns.
struct PackNetArgs {
  // first layer
  const float* w1; int ld1, S, P, wp_cols, bias_cols; const float* b1; const float* cvec; __half* w1h; int Kp;
  float* b_eff; int H1;
  // second layer [H2, H1]
  const float* w2; int H2; __half* w2h; __half* w2th;  // w2th may be null
  float* wp;                                           // [H1][4] or null
  int nA, nB, nC;
};
__global__ void __launch_bounds__(256) pack_net_kernel(PackNetArgs a) {
  pdl_wait();
  __shared__ float tile[32][33];
  int b = blockIdx.x;
  if (b < a.nA) {
    const int lane = threadIdx.x & 31;
    const int i = b * 8 + (threadIdx.x >> 5);
    if (i >= a.H1) return;
    const float* wr = a.w1 + (size_t)i * a.ld1;
    float dot = 0.f;
    for (int j = lane; j < a.S; j += 32) dot = fmaf(a.cvec[j], wr[j], dot);
    if (a.wp_cols)
      for (int e = lane; e < a.P; e += 32) dot = fmaf(kParamCenter, wr[a.S + e], dot);
    dot = warp_sum_f(dot);
    const float be = a.b1[i] + dot;
    const __half hi = __float2half_rn(be);
    const __half lo = __float2half_rn(be - __half2float(hi));
    if (lane == 0 && a.b_eff) a.b_eff[i] = be;
    for (int j = lane; j < a.Kp; j += 32) {
      __half o = __float2half_rn(0.f);
      if (j < a.S) o = __float2half_rn(wr[j]);
      else if (j < a.S + a.P) o = a.wp_cols ? __float2half_rn(wr[j]) : o;
      else if (j == a.S + a.P) o = a.bias_cols ? hi : o;
      else if (j == a.S + a.P + 1) o = a.bias_cols ? lo : o;
      a.w1h[(size_t)i * a.Kp + j] = o;
    }
    return;
  }
  b -= a.nA;
  if (b < a.nB) {
    const int total = a.H2 * a.H1;
    for (int idx = (b * 256 + threadIdx.x) * 4; idx < total; idx += a.nB * 256 * 4) {
      const float4 v = *reinterpret_cast<const float4*>(a.w2 + idx);
      __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(a.w2h + idx) = u;
    }
    return;
  }
  b -= a.nB;
  if (b < a.nC) {
    // w2 [H2 rows, H1 cols] -> w2th [H1, H2]
    const int tiles_x = a.H1 / 32;
    const int bx = (b % tiles_x) * 32, by = (b / tiles_x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) tile[k][tx] = a.w2[(size_t)(by + k) * a.H1 + bx + tx];
    __syncthreads();
    for (int k = ty; k < 32; k += 8) a.w2th[(size_t)(bx + k) * a.H2 + by + tx] = __float2half_rn(tile[tx][k]);
    return;
  }
  b -= a.nC;
  const int idx = b * 256 + threadIdx.x;
  if (a.wp != nullptr && idx < a.H1 * 4) {
    const int i = idx >> 2, e = idx & 3;
    a.wp[idx] = e < a.P ? a.w1[(size_t)i * a.ld1 + a.S + e] : 0.f;
  }
}

// Constant image of the fused generator-head + surrogate-layer-1 epilogue (layout: EpiHeadF1, epilogues.cuh).
// One block of 256 threads, thread c = column c of both the generator's second layer and the surrogate's first.
__global__ void head_consts_kernel(const float* __restrict__ scale2, const float* __restrict__ bias2,
                                   const float* __restrict__ w3, const float* __restrict__ b3,
                                   const float* __restrict__ fw1, const float* __restrict__ fb1,
                                   const float* __restrict__ flnw, const float* __restrict__ flnb,
                                   float* __restrict__ out) {
  pdl_wait();
  __shared__ double red[256];
  __shared__ double mean[5];
  const int c = threadIdx.x;
  // head part, column pairs
  {
    float* o = out + (c >> 1) * 12;
    o[(c & 1) * 2 + 0] = scale2[c];
    o[(c & 1) * 2 + 1] = bias2[c];
    for (int j = 0; j < 4; ++j) o[4 + (c & 1) * 4 + j] = w3[j * 256 + c];
  }
  if (c < 4) out[1536 + c] = b3[c];
  for (int i = 1570 + c; i < 2048; i += 256) out[i] = 0.f;
  // centred surrogate layer: u_c = (bias_c, w_c[0..3]) minus its mean over the columns
  double u[5] = {(double)fb1[c], (double)fw1[c * 4 + 0], (double)fw1[c * 4 + 1], (double)fw1[c * 4 + 2],
                 (double)fw1[c * 4 + 3]};
  for (int j = 0; j < 5; ++j) {
    red[c] = u[j];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (c < s) red[c] += red[c + s];
      __syncthreads();
    }
    if (c == 0) mean[j] = red[0] / 256.0;
    __syncthreads();
  }
  for (int j = 0; j < 5; ++j) u[j] -= mean[j];
  double* q = reinterpret_cast<double*>(out + 1540);   // 8-byte aligned: 1540 * 4 = 6160
  int k = 0;
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j, ++k) {
      red[c] = u[i] * u[j];
      __syncthreads();
      for (int s = 128; s > 0; s >>= 1) {
        if (c < s) red[c] += red[c + s];
        __syncthreads();
      }
      if (c == 0) q[k] = red[0] / 256.0;
      __syncthreads();
    }
  const float gam = flnw[c];
  for (int j = 0; j < 4; ++j) out[2048 + c * 4 + j] = gam * (float)u[1 + j];
  out[3072 + c * 2 + 0] = gam * (float)u[0];
  out[3072 + c * 2 + 1] = flnb[c];
}

__global__ void copy_pad_f32_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad) {
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_pad) dst[idx] = idx < n ? src[idx] : 0.f;
}

// ---- 4-columns-per-thread layout of the batch-coupled streaming kernels --------------------------------------
// A thread owns 4 consecutive columns (8-byte fp16 loads) of every row it visits and keeps kU rows in flight;
// per-column constants are folded algebraically so a kernel needs <= 64 registers and 4+ blocks fit per SM —
// these kernels are latency-bound, measured 12-24 % active warps and 5-38 % DRAM throughput at 124-187 registers.
constexpr int kU = 4;
struct ColMap4 {
  int cpr, rpb, ch, rg;
  __device__ ColMap4(int C) {
    cpr = C >> 2;
    rpb = kThreads / cpr;
    ch = threadIdx.x % cpr;
    rg = threadIdx.x / cpr;
  }
};
__device__ __forceinline__ void ld_h4(const __half* p, float* v) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st_h4(__half* p, const float* v) {
  uint2 u;
  __half2 h;
  h = __floats2half2_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2half2_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void ld_f4(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
// block partial of 4 column sums -> this block's scratch row (see block_colsum_partial)
__device__ __forceinline__ void block_colsum_partial4(const float* acc4, float* part, const ColMap4& m, float* sm) {
  *reinterpret_cast<float4*>(sm + (m.rg * m.cpr + m.ch) * 4) = make_float4(acc4[0], acc4[1], acc4[2], acc4[3]);
  __syncthreads();
  if (m.rg == 0) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int gq = 0; gq < m.rpb; ++gq) {
      const float4 x = *reinterpret_cast<const float4*>(sm + (gq * m.cpr + m.ch) * 4);
      t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
    }
    *reinterpret_cast<float4*>(part + m.ch * 4) = t;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ BatchNorm
__global__ void __launch_bounds__(kThreads, 4) colstats_kernel(const __half* __restrict__ h, long long rows, int C,
                                                               float* __restrict__ part) {
  pdl_wait();
  __shared__ float sm[kThreads * 4];
  const ColMap4 m(C);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < rows; r0 += kU * stride) {
    float v[kU][4];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * stride;
      if (r < rows) ld_h4(h + r * C + m.ch * 4, v[u]);
      else v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i] += v[u][i];
        q[i] = fmaf(v[u][i], v[u][i], q[i]);
      }
  }
  part += (size_t)blockIdx.x * 2 * C;
  block_colsum_partial4(s, part, m, sm);
  block_colsum_partial4(q, part + C, m, sm);
}

__device__ __forceinline__ void bn_finalize_column(const BnFinalizeArgs& a, int c, float sum, float sumsq) {
  const double ms = (double)sum / a.n;
  double var = (double)sumsq / a.n - ms * ms;
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
__global__ void bn_finalize_kernel(BnFinalizeArgs a) {
  pdl_wait();
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.C; c += gridDim.x * blockDim.x)
    bn_finalize_column(a, c, a.sum[c], a.sumsq[c]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.num_batches_tracked != nullptr)
    *a.num_batches_tracked += a.num_updates;
}
// reduce_partials_kernel + bn_finalize_kernel in one launch (single-GPU step: nothing to exchange between the two):
// part[nblocks][2 * C] = per-block [sums | sums of squares]; a.sum / a.sumsq (the step's zeroed accumulators) receive
// the totals as before.  One block per 32 columns.
__global__ void __launch_bounds__(1024) bn_reduce_finalize_kernel(BnFinalizeArgs a, const float* __restrict__ part,
                                                                  int nblocks) {
  pdl_wait();
  __shared__ float sm[2][32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  float s = 0.f, q = 0.f;
  if (col < a.C)
    for (int b = ty; b < nblocks; b += 32) {
      s += part[(size_t)b * 2 * a.C + col];
      q += part[(size_t)b * 2 * a.C + a.C + col];
    }
  sm[0][ty][tx] = s;
  sm[1][ty][tx] = q;
  __syncthreads();
  if (ty == 0 && col < a.C) {
    float t = 0.f, u = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      t += sm[0][k][tx];
      u += sm[1][k][tx];
    }
    t += a.sum[col];      // same association as the two-kernel path: accumulator (zero) + partial total
    u += a.sumsq[col];
    const_cast<float*>(a.sum)[col] = t;
    const_cast<float*>(a.sumsq)[col] = u;
    bn_finalize_column(a, col, t, u);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.num_batches_tracked != nullptr)
    *a.num_batches_tracked += a.num_updates;
}

__global__ void bn_eval_affine_kernel(const float* rm, const float* rv, const float* gamma, const float* beta,
                                      const float* offset, float* scale, float* bias, int C) {
  pdl_wait();
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
  pdl_wait();
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

// ------------------------------------------------------------------------------------------ generator head
// p = tanh(relu(bn2(h2)) W3^T + b3) (generator.py:22-25), denormalisation (data_loader.py:238-252) and the fake-row
// tail of the spectrum operand.  A warp takes 8 rows per trip with coalesced 16-byte loads; a lane owns 8 of the 256
// columns (its BatchNorm affine and W3 slices stay in registers) and forms the 8 x 4 partial dot products, which a
// 31-shuffle transposing reduction turns into one finished output per lane (lane L: row L/4, parameter L%4), so tanh
// runs once per output and the [rows,4] stores are coalesced.  C must be 256.
constexpr int kHeadRows = 8;   // rows per warp and trip
// v[0..31] per lane -> returns sum over the lanes of v[lane]
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}
__global__ void __launch_bounds__(kThreads, 2) g_head_fwd_kernel(
    const __half* __restrict__ h2, const float* __restrict__ scale, const float* __restrict__ bias,
    const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ p_out,
    float* __restrict__ pden_out, const __half* __restrict__ xc, __half* __restrict__ tail_fake, long long rows,
    int C, int Kp, int S) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  float sc[8], bi[8], w[4][8];
  ld_f8(scale + lane * 8, sc);
  ld_f8(bias + lane * 8, bi);
#pragma unroll
  for (int e = 0; e < 4; ++e) ld_f8(w3 + e * C + lane * 8, w[e]);
  const float b3l = __ldg(b3 + (lane & 3));
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * kHeadRows;
  for (long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kHeadRows; r0 < rows;
       r0 += wstride) {
    uint4 u[kHeadRows];
#pragma unroll
    for (int q = 0; q < kHeadRows; ++q) {
      const long long row = r0 + q < rows ? r0 + q : rows - 1;
      u[q] = *reinterpret_cast<const uint4*>(h2 + row * C + lane * 8);
    }
    float part[32];
#pragma unroll
    for (int q = 0; q < kHeadRows; ++q) {
      const uint32_t hw[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
      float a[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
        a[2 * j] = fmaxf(fmaf(sc[2 * j], f.x, bi[2 * j]), 0.f);
        a[2 * j + 1] = fmaxf(fmaf(sc[2 * j + 1], f.y, bi[2 * j + 1]), 0.f);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) d = fmaf(a[i], w[e][i], d);
        part[q * 4 + e] = d;
      }
    }
    const float pv = tanhf(warp_transpose_sum32(part) + b3l);       // lane L: row r0 + L/4, parameter L%4
    const float pd = (pv + 1.0f) / 2.0f * 0.6f + 2.2f;               // data_loader.py:238-252
    const long long row = r0 + (lane >> 2);
    if (row < rows) {
      p_out[r0 * 4 + lane] = pv;
      if (pden_out) pden_out[r0 * 4 + lane] = pd;
    }
    if (tail_fake != nullptr) {
      // last 64 operand columns of the 8 rows (8 chunks of 16 bytes each), parameter columns replaced
      const int t0 = Kp - 64;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int t = j * 32 + lane, r = t >> 3, ch = t & 7;
        float pr[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pr[e] = __shfl_sync(0xffffffffu, pd, r * 4 + e);
        if (r0 + r < rows) {
          float v[8];
          ld_h8(xc + (r0 + r) * Kp + t0 + ch * 8, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = t0 + ch * 8 + k - S;
            if (e >= 0 && e < 4) v[k] = pr[e] - kParamCenter;
          }
          st_h8(tail_fake + (r0 + r) * 64 + ch * 8, v);
        }
      }
    }
  }
}

// Generator head backward, fused with the BatchNorm-2 backward, in three kernels that never store the
// pre-projection gradient dy = relu'(.) * (dpre W3):
//   g_head_dpre   dpre[r, 0:4] = dL/d(pre-tanh) from the adversarial, LC and range terms; db3, range-loss sum
//   g_head_bwd<0> batch moments S0[j,c] = sum_r m dpre_j, S1[j,c] = sum_r m h dpre_j (m = ReLU mask, h = stored
//                 pre-BN value); g_head_moments_reduce turns them into sum dy, sum dy*xhat and dW3
//   g_head_bwd<1> dh = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)) as A*dy - B*h + C0 -> fp16, db2
// Recomputing dy costs 4 FMAs per element; storing it in fp16 before BatchNorm's projection would amplify its
// rounding (the adversarial gradient pushes every sample the same way, so most of dy is projected out).
__global__ void __launch_bounds__(kThreads) g_head_dpre_kernel(GHeadBwdArgs a) {
  pdl_wait();
  __shared__ float sm[32];
  float db[4] = {0.f, 0.f, 0.f, 0.f};
  float range_acc = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.rows;
       r += (long long)gridDim.x * blockDim.x) {
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(a.p) + r);
    const float4 d4 = a.dpden ? __ldg(reinterpret_cast<const float4*>(a.dpden) + r) : z4;
    const float4 l4 = a.dp_lc ? __ldg(reinterpret_cast<const float4*>(a.dp_lc) + r) : z4;
    const float4 e4 = a.dp_extra ? __ldg(reinterpret_cast<const float4*>(a.dp_extra) + r) : z4;
    const float p[4] = {p4.x, p4.y, p4.z, p4.w};
    const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
    // the caller's extra term is a true gradient; the others carry the gradient scale GS
    const float ll[4] = {fmaf(e4.x, a.gs, l4.x), fmaf(e4.y, a.gs, l4.y), fmaf(e4.z, a.gs, l4.z), fmaf(e4.w, a.gs, l4.w)};
    float dpre[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = fmaxf(-p[j], 0.f), hi = fmaxf(p[j] - 1.f, 0.f);        // loss.py:121-123
      const float dp = (0.5f * 0.6f) * dd[j] + ll[j] + (2.f * hi - 2.f * lo) * a.range_mult;
      dpre[j] = dp * (1.f - p[j] * p[j]);                                      // tanh backward
      db[j] += dpre[j];
      range_acc += lo * lo + hi * hi;
    }
    *reinterpret_cast<float4*>(a.dpre + r * 4) = make_float4(dpre[0], dpre[1], dpre[2], dpre[3]);
  }
  // per-block partials (db3[0..3], range sum) -> a.dpre_part[block][8]; g_head_moments_finish_kernel adds them in
  // block order (float atomics here made the bias gradient, hence the whole run, irreproducible in the last bit)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float t = block_sum(db[j], sm);
    if (threadIdx.x == 0) a.dpre_part[blockIdx.x * 8 + j] = t;
  }
  const float t = block_sum(range_acc, sm);
  if (threadIdx.x == 0) a.dpre_part[blockIdx.x * 8 + 4] = t;
}

template <bool APPLY>
__global__ void __launch_bounds__(kThreads, 2) g_head_bwd_kernel(GHeadBwdArgs a) {
  pdl_wait();
  __shared__ float sm[kThreads * 4];
  const ColMap4 m(a.C);
  const int c0 = m.ch * 4;
  float sc[4], bi[4];
  ld_f4(a.scale + c0, sc);
  ld_f4(a.bias + c0, bi);
  const long long stride = (long long)gridDim.x * m.rpb;
  constexpr int U = APPLY ? 4 : 6;   // rows in flight per thread: these kernels are bound by memory-level parallelism
  if (!APPLY) {
    float s0[4][4], s1[4][4];  // [j][column]
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s0[j][i] = s1[j][i] = 0.f;
    for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < a.rows; r0 += U * stride) {
      float h[U][4];
      float4 dq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * stride;
        if (r < a.rows) {
          ld_h4(a.h2 + r * a.ld + c0, h[u]);
          dq[u] = __ldg(reinterpret_cast<const float4*>(a.dpre) + r);
        } else {
          h[u][0] = h[u][1] = h[u][2] = h[u][3] = 0.f;
          dq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float dp[4] = {dq[u].x, dq[u].y, dq[u].z, dq[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float mf = fmaf(sc[i], h[u][i], bi[i]) > 0.f ? 1.f : 0.f;
          const float mh = mf * h[u][i];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s0[j][i] = fmaf(mf, dp[j], s0[j][i]);
            s1[j][i] = fmaf(mh, dp[j], s1[j][i]);
          }
        }
      }
    }
    float* part = a.part + (size_t)blockIdx.x * 8 * a.C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      block_colsum_partial4(s0[j], part + j * a.C, m, sm);
      block_colsum_partial4(s1[j], part + (4 + j) * a.C, m, sm);
    }
  } else {
    float w[4][4], A[4], Bc[4], C0[4], sdh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ld_f4(a.w3 + j * a.ld + c0, w[j]);
    {
      float gm[4], rs[4], mu[4], m1[4], m2[4];
      ld_f4(a.gamma + c0, gm);
      ld_f4(a.rstd + c0, rs);
      ld_f4(a.mean + c0, mu);
      ld_f4(a.sum_dy + c0, m1);
      ld_f4(a.sum_dyx + c0, m2);
      if (blockIdx.x == 0 && m.rg == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (a.dgamma) a.dgamma[c0 + i] += m2[i] * a.inv_gs;
          if (a.dbeta) a.dbeta[c0 + i] += m1[i] * a.inv_gs;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float mean_dy = (float)((double)m1[i] * a.inv_n);
        const float mean_dyx = (float)((double)m2[i] * a.inv_n);
        A[i] = gm[i] * rs[i];
        Bc[i] = A[i] * mean_dyx * rs[i];
        C0[i] = A[i] * (mean_dyx * rs[i] * mu[i] - mean_dy);
        sdh[i] = 0.f;
      }
    }
    for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < a.rows; r0 += U * stride) {
      float h[U][4];
      float4 dq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * stride;
        if (r < a.rows) {
          ld_h4(a.h2 + r * a.ld + c0, h[u]);
          dq[u] = __ldg(reinterpret_cast<const float4*>(a.dpre) + r);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * stride;
        if (r >= a.rows) break;
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float dy = dq[u].x * w[0][i];
          dy = fmaf(dq[u].y, w[1][i], dy);
          dy = fmaf(dq[u].z, w[2][i], dy);
          dy = fmaf(dq[u].w, w[3][i], dy);
          dy = fmaf(sc[i], h[u][i], bi[i]) > 0.f ? dy : 0.f;
          const float dh = fmaf(A[i], dy, fmaf(-Bc[i], h[u][i], C0[i]));
          out[i] = dh;
          sdh[i] += dh;
        }
        st_h4(a.dy2 + r * a.ld + c0, out);
      }
    }
    block_colsum_partial4(sdh, a.part + (size_t)blockIdx.x * a.C, m, sm);
  }
}

// Sums the moment partials over blocks (part[nblocks][8][C]: four arrays of sum dpre_j * 1[relu], four of
// sum dpre_j * relu'd xhat terms) in a fixed order and finishes sum dy, sum dy*xhat and dW3 from them; one extra block
// adds g_head_dpre_kernel's partials (db3, range-loss sum).  One launch (was: a summing kernel and a finishing kernel):
// a block owns 8 columns; thread = (column, array, one of 16 row groups).
__global__ void __launch_bounds__(1024) g_head_moments_kernel(GHeadBwdArgs a, const float* __restrict__ part,
                                                              int nblocks, int dpre_blocks) {
  pdl_wait();
  if (blockIdx.x == gridDim.x - 1) {   // db3 and the range-loss sum from g_head_dpre_kernel's partials, fixed order
    __shared__ double dsm[128][8];
    const int q = threadIdx.x & 7, grp = threadIdx.x >> 3;   // 128 groups x 8 quantities (5 used)
    double t = 0.0;
    if (q < 5)
      for (int b = grp; b < dpre_blocks; b += 128) t += (double)a.dpre_part[b * 8 + q];
    dsm[grp][q] = t;
    __syncthreads();
    if (threadIdx.x < 5) {
      double tt = 0.0;
      for (int k = 0; k < 128; ++k) tt += dsm[k][threadIdx.x];
      if (threadIdx.x < 4) a.db3[threadIdx.x] += (float)tt * a.inv_gs;
      else if (a.range_sum) *a.range_sum += tt;
    }
    return;
  }
  __shared__ float sm[16][64];
  __shared__ float tot[8][8];
  const int cl = threadIdx.x & 7, k = (threadIdx.x >> 3) & 7, rg = threadIdx.x >> 6;
  const int c = blockIdx.x * 8 + cl;
  float s = 0.f;
  if (c < a.C) {
    const float* col = part + (size_t)k * a.C + c;
#pragma unroll 4
    for (int b = rg; b < nblocks; b += 16) s += col[(size_t)b * 8 * a.C];
  }
  sm[rg][threadIdx.x & 63] = s;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) t += sm[q][threadIdx.x];
    tot[k][cl] = t;
  }
  __syncthreads();
  if (threadIdx.x >= 8 || c >= a.C) return;
  const float sc = a.scale[c], bi = a.bias[c], mu = a.mean[c], rs = a.rstd[c];
  float sdy = 0.f, sdyh = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float w = a.w3[j * a.ld + c];
    const float t0 = tot[j][cl], t1 = tot[4 + j][cl];
    sdy = fmaf(w, t0, sdy);
    sdyh = fmaf(w, t1, sdyh);
    a.dw3[j * a.ld + c] += a.inv_gs * (sc * t1 + bi * t0);   // sum dpre_j * relu(sc*h+bi)
  }
  a.sum_dy[c] += sdy;
  a.sum_dyx[c] += rs * (sdyh - mu * sdy);
}

__global__ void __launch_bounds__(kThreads, 4) bn_bwd_stats_kernel(
    const __half* __restrict__ da, const __half* __restrict__ h, const float* __restrict__ scale,
    const float* __restrict__ bias, const float* __restrict__ mean, const float* __restrict__ rstd,
    float* __restrict__ part, long long rows, int C, int ld) {
  pdl_wait();
  __shared__ float sm[kThreads * 4];
  const ColMap4 m(C);
  const int c0 = m.ch * 4;
  float sc[4], bi[4], mu[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  ld_f4(scale + c0, sc);
  ld_f4(bias + c0, bi);
  ld_f4(mean + c0, mu);
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < rows; r0 += kU * stride) {
    float g[kU][4], x[kU][4];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      long long r = r0 + u * stride;
      const bool ok = r < rows;
      r = ok ? r : rows - 1;
      ld_h4(da + r * ld + c0, g[u]);
      ld_h4(h + r * ld + c0, x[u]);
      if (!ok) g[u][0] = g[u][1] = g[u][2] = g[u][3] = 0.f;   // dy = 0: contributes nothing
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dy = fmaf(sc[i], x[u][i], bi[i]) > 0.f ? g[u][i] : 0.f;
        s1[i] += dy;
        s2[i] = fmaf(dy, x[u][i] - mu[i], s2[i]);   // rstd applied once at the end
      }
  }
  float rs[4];
  ld_f4(rstd + c0, rs);
#pragma unroll
  for (int i = 0; i < 4; ++i) s2[i] *= rs[i];
  part += (size_t)blockIdx.x * 2 * C;
  block_colsum_partial4(s1, part, m, sm);
  block_colsum_partial4(s2, part + C, m, sm);
}

// dh = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)), dy = da * ReLU mask, written as A*dy - B*h + C0
__global__ void __launch_bounds__(kThreads, 3) bn_bwd_apply_kernel(BnBwdArgs a) {
  pdl_wait();
  __shared__ float sm[kThreads * 4];
  const ColMap4 m(a.C);
  const int c0 = m.ch * 4;
  float sc[4], bi[4], A[4], Bc[4], C0[4], sdh[4];
  ld_f4(a.scale + c0, sc);
  ld_f4(a.bias + c0, bi);
  {
    float gm[4], rs[4], mu[4], m1[4], m2[4];
    ld_f4(a.gamma + c0, gm);
    ld_f4(a.rstd + c0, rs);
    ld_f4(a.mean + c0, mu);
    ld_f4(a.sum_dy + c0, m1);
    ld_f4(a.sum_dyx + c0, m2);
    if (blockIdx.x == 0 && m.rg == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (a.dgamma) a.dgamma[c0 + i] += m2[i] * a.inv_gs;
        if (a.dbeta) a.dbeta[c0 + i] += m1[i] * a.inv_gs;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float mean_dy = (float)((double)m1[i] * a.inv_n);
      const float mean_dyx = (float)((double)m2[i] * a.inv_n);
      A[i] = gm[i] * rs[i];
      Bc[i] = A[i] * mean_dyx * rs[i];
      C0[i] = A[i] * (mean_dyx * rs[i] * mu[i] - mean_dy);
      sdh[i] = 0.f;
    }
  }
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < a.rows; r0 += kU * stride) {
    float g[kU][4], x[kU][4];
#pragma unroll
    for (int u = 0; u < kU; ++u) {  // clamped index: all loads issue back to back, no branches
      long long r = r0 + u * stride;
      r = r < a.rows ? r : a.rows - 1;
      ld_h4(a.dy + r * a.ld + c0, g[u]);
      ld_h4(a.h + r * a.ld + c0, x[u]);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * stride;
      if (r >= a.rows) break;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float dy = g[u][i];
        if (a.relu_mask) dy = fmaf(sc[i], x[u][i], bi[i]) > 0.f ? dy : 0.f;
        const float dh = fmaf(A[i], dy, fmaf(-Bc[i], x[u][i], C0[i]));
        g[u][i] = dh;
        sdh[i] += dh;
      }
      st_h4(a.dh + r * a.ld + c0, g[u]);
    }
  }
  if (a.dbias) block_colsum_partial4(sdh, a.part + (size_t)blockIdx.x * a.C, m, sm);
}

// ------------------------------------------------------------------------------------------ discriminator
__global__ void __launch_bounds__(kThreads, 3) d_l2_bwd_kernel(const __half* __restrict__ z2,
                                                               const float* __restrict__ dlogit,
                                                               const float* __restrict__ w3, __half* __restrict__ dh2,
                                                               float* __restrict__ dw3, float* __restrict__ db2,
                                                               float* __restrict__ db3, long long rows, int C,
                                                               float inv_gs, float* __restrict__ part) {
  pdl_wait();
  // 8 columns per thread (16-byte loads and stores), 4 rows in flight: bound by memory-level parallelism
  __shared__ float sm[kThreads * 8];
  const ColMap m(C);
  const int c0 = m.ch * 8;
  float w[8], sw[8], sb[8];
  ld_f8(w3 + c0, w);
  zero8(sw);
  zero8(sb);
  float s3 = 0.f;
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < rows; r0 += kU * stride) {
    uint4 zraw[kU];   // rows stay packed until they are used
    float dl[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * stride;
      if (r < rows) {
        dl[u] = __ldg(dlogit + r);
        zraw[u] = *reinterpret_cast<const uint4*>(z2 + r * C + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * stride;
      if (r >= rows) break;
      float o[8], z[8];
      const uint32_t zw[4] = {zraw[u].x, zraw[u].y, zraw[u].z, zraw[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&zw[j]));
        z[2 * j] = f.x;
        z[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dh = dl[u] * w[i] * (z[i] > 0.f ? 1.f : kSlope);
        o[i] = dh;
        sw[i] = fmaf(dl[u], z[i], sw[i]);
        sb[i] += dh;
      }
      st_h8(dh2 + r * C + c0, o);
      if (m.ch == 0) s3 += dl[u];
    }
  }
  if (dw3 != nullptr) {
    part += (size_t)blockIdx.x * (2 * C + 8);
    block_colsum_partial(sw, part, m, sm);
    block_colsum_partial(sb, part + C, m, sm);
    const float t = block_sum(s3, sm);
    if (threadIdx.x == 0) part[2 * C] = t;   // db3: column 2C of the partial row (no float atomics)
  }
}

// ------------------------------------------------------------------------------------------ forward model
// first layer (K = 4, forward_model.py:30-33): a warp handles 4 rows per trip (independent reduction chains), a
// lane owns 8 of the 256 columns
__global__ void __launch_bounds__(kThreads) f_l1_kernel(const float* __restrict__ p, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ lnw,
                                                        const float* __restrict__ lnb, __half* __restrict__ out,
                                                        long long rows, int C) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  float4 wr[8];
  float b[8], gm[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wr[i] = __ldg(reinterpret_cast<const float4*>(w1) + lane * 8 + i);
  ld_f8(b1 + lane * 8, b);
  ld_f8(lnw + lane * 8, gm);
  ld_f8(lnb + lane * 8, bt);
  constexpr int R = 4;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.0f / (float)C;
  for (long long g0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; g0 < rows;
       g0 += wstride * R) {
    float4 q[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = g0 + u < rows ? g0 + u : rows - 1;
      q[u] = __ldg(reinterpret_cast<const float4*>(p) + row);
    }
    float h[R][8], s[R], v[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      s[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = b[i];
        t = fmaf(q[u].x, wr[i].x, t);
        t = fmaf(q[u].y, wr[i].y, t);
        t = fmaf(q[u].z, wr[i].z, t);
        t = fmaf(q[u].w, wr[i].w, t);
        h[u][i] = t;
        s[u] += t;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      s[u] *= inv_c;  // mean
      v[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = h[u][i] - s[u];
        v[u] = fmaf(d, d, v[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (g0 + u >= rows) break;
      const float rstd = 1.0f / sqrtf(v[u] * inv_c + kLnEps);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[u][i] = lrelu_f(fmaf((h[u][i] - s[u]) * rstd, gm[i], bt[i]));
      st_h8(out + (g0 + u) * C + lane * 8, h[u]);
    }
  }
}

__global__ void __launch_bounds__(kThreads) ln_lrelu_apply_kernel(__half* __restrict__ h,
                                                                  const float* __restrict__ rowstats, int n_tiles,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, long long rows,
                                                                  int N) {
  pdl_wait();
  const ColMap m(N);
  float gm[8], bt[8];
  ld_f8(gamma + m.ch * 8, gm);
  ld_f8(beta + m.ch * 8, bt);
  const float inv_n = 1.0f / (float)N;
  constexpr int U = 4;   // rows in flight per thread (one was latency-bound: 40 % of the HBM peak at N = 2048)
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rg; r0 < rows; r0 += U * stride) {
    uint4 raw[U];
    float mean[U], rstd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride < rows ? r0 + u * stride : rows - 1;
      raw[u] = *reinterpret_cast<const uint4*>(h + r * N + m.ch * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride < rows ? r0 + u * stride : rows - 1;
      // up to 8 partials per row (N <= 2048), fully unrolled and predicated: the loads issue back to back (the rolled
      // loop paid one L2 latency per partial)
      float s1 = 0.f, s2 = 0.f;
      const float2* rs = reinterpret_cast<const float2*>(rowstats) + r * n_tiles;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (t < n_tiles) {
          const float2 st = __ldg(rs + t);
          s1 += st.x;
          s2 += st.y;
        }
      }
      mean[u] = s1 * inv_n;
      rstd[u] = 1.0f / sqrtf(fmaxf(s2 * inv_n - mean[u] * mean[u], 0.f) + kLnEps);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r >= rows) break;
      const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
      float v[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = lrelu_f(fmaf((v[i] - mean[u]) * rstd[u], gm[i], bt[i]));
      st_h8(h + r * N + m.ch * 8, v);
    }
  }
}

// ------------------------------------------------------------------------------------------ optimiser
// clears up to 4 fp32 buffers in one launch: replaces a row of cudaMemsetAsync calls, each of which is a separate
// stream operation that also breaks the chain of programmatic dependent launches
__global__ void __launch_bounds__(kThreads) zero_buffers_kernel(ZeroArgs a) {
  pdl_wait();
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4* p = reinterpret_cast<uint4*>(a.ptr[k]);
    const long long n16 = a.nfloat[k] >> 2;
    for (long long i = tid; i < n16; i += nth) p[i] = z;
    if (tid < (a.nfloat[k] & 3)) a.ptr[k][(n16 << 2) + tid] = 0.f;
  }
}

// squared gradient norm, stage 1: one fp64 partial per block (no atomics: clip_adam_kernel adds the partials in a
// fixed order, so the clip coefficient - and with it every weight - is run-to-run reproducible)
__global__ void __launch_bounds__(kThreads) sumsq_kernel(const float* __restrict__ g, long long n,
                                                         double* __restrict__ parts) {
  pdl_wait();
  __shared__ double smd[8];
  double s = 0.0;
  float f = 0.f;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    f = fmaf(g[i], g[i], f);
    if (++cnt == 64) {   // bounded fp32 runs, fp64 across them
      s += (double)f;
      f = 0.f;
      cnt = 0;
    }
  }
  s += (double)f;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) smd[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += smd[k];
    parts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kThreads) clip_adam_kernel(AdamArgs a) {
  pdl_wait();
  // torch.nn.utils.clip_grad_norm_(max_norm) then optim.Adam.step() (train_pigan.py:142-143,186-187)
  __shared__ double tot_sm;
  if (threadIdx.x < 32) {   // every block adds the partials of sumsq_kernel in the same order
    double t = 0.0;
    for (int k = threadIdx.x; k < a.n_parts; k += 32) t += a.sq_parts[k];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) tot_sm = t;
  }
  __syncthreads();
  const float total = (float)sqrt(tot_sm);
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

// dw[m, n] += scale * sum_s part[unit(s, tile(m,n))][m % 128][n % 256]; the bias column goes to db[m].
// fix_cvec != nullptr (first layers, whose operand is centred: x - c): the product with the centring row is folded in,
// dw[m, n] += db[m] * (n < fix_S ? c[n] : 2.5) for n < fix_S + fix_P, where db[m] = scale * (bias column's sum) - db
// must be zero on entry.  Every thread of a row sums the bias column itself (a warp-uniform address: one broadcast
// load per slab), which saves the separate fix-up launch.
template <bool FIX>
__global__ void __launch_bounds__(256) dw_reduce_kernel(const float* __restrict__ part, int tiles_m, int tiles_n,
                                                        int splits, float* __restrict__ dw, int ld, int m_valid,
                                                        int n_valid, float scale, int bias_col,
                                                        float* __restrict__ db, const float* __restrict__ fix_cvec,
                                                        int fix_S, int fix_P) {
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= tiles_n * 256 || m >= m_valid) return;
  const bool is_w = n < n_valid, is_b = n == bias_col;
  if (!is_w && !is_b) return;
  const int tiles = tiles_m * tiles_n;
  const size_t stride = (size_t)tiles * 128 * 256;
  const float* p = part + ((size_t)((m >> 7) * tiles_n + (n >> 8)) * 128 + (m & 127)) * 256 + (n & 255);
  // the row's bias column: the same address for every thread of the warp (one broadcast load per slab)
  const int bc = FIX ? bias_col : 0;
  const float* q = part + ((size_t)((m >> 7) * tiles_n + (bc >> 8)) * 128 + (m & 127)) * 256 + (bc & 255);
  float t = 0.f, tb = 0.f;
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    const float a0 = p[(size_t)s * stride], a1 = p[(size_t)(s + 1) * stride];
    const float a2 = p[(size_t)(s + 2) * stride], a3 = p[(size_t)(s + 3) * stride];
    if constexpr (FIX) {
      const float b0 = q[(size_t)s * stride], b1 = q[(size_t)(s + 1) * stride];
      const float b2 = q[(size_t)(s + 2) * stride], b3 = q[(size_t)(s + 3) * stride];
      tb += (b0 + b1) + (b2 + b3);
    }
    t += (a0 + a1) + (a2 + a3);
  }
  for (; s < splits; ++s) {
    t += p[(size_t)s * stride];
    if constexpr (FIX) tb += q[(size_t)s * stride];
  }
  if (!is_w) {
    db[m] += scale * t;
    return;
  }
  float v = dw[(size_t)m * ld + n] + scale * t;
  if constexpr (FIX) {
    if (n < fix_S + fix_P) v += (scale * tb) * (n < fix_S ? fix_cvec[n] : kParamCenter);
  }
  dw[(size_t)m * ld + n] = v;
}

__global__ void dw_fixup_kernel(float* dw, int ld, int S, int P, const float* db, const float* cvec, int rows) {
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = S + P;
  if (idx >= rows * per) return;
  const int i = idx / per, j = idx % per;
  dw[(size_t)i * ld + j] += db[i] * (j < S ? cvec[j] : kParamCenter);
}

__global__ void loss_finalize_kernel(LossFinalizeArgs a) {
  pdl_wait();
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
  pdl_wait();
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


// dout[r, 0:ld] = fp16(scale * g[r, 0:cols]) (zero beyond cols): an upstream gradient as the surrogate's backward
// operand (pigan_forward_model_vjp); scale = the gradient scale GS the backward chain carries
__global__ void f_upstream_cast_kernel(const float* __restrict__ g, int cols, __half* __restrict__ dout, int ld,
                                       long long rows, float scale) {
  pdl_wait();
  const long long total = rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld;
    const int c = (int)(i - r * ld);
    dout[i] = __float2half_rn(c < cols ? scale * g[r * cols + c] : 0.f);
  }
}

// Model-validation scores (unified_evaluator.py:453-468): stability[r] = mean_j (p[r,j] - p_noisy[r,j])^2,
// plausibility[r] = mean_j sigmoid(10 p[r,j] - 5)
__global__ void validation_scores_kernel(const float* __restrict__ p, const float* __restrict__ pn, long long rows,
                                         int P, float* __restrict__ stab, float* __restrict__ plaus) {
  pdl_wait();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f, q = 0.f;
  for (int j = 0; j < P; ++j) {
    const float a = p[r * P + j], d = a - pn[r * P + j];
    s = fmaf(d, d, s);
    q += 1.0f / (1.0f + expf(-(a * 10.f - 5.f)));
  }
  if (stab) stab[r] = s / (float)P;
  if (plaus) plaus[r] = q / (float)P;
}

// ------------------------------------------------------------------------------------------ surrogate training
// Training-mode forward model (pretrain_fwd_model.py:68-92, forward_model.py:28-52): Linear -> LayerNorm ->
// LeakyReLU -> Dropout(0.2).  The keep-mask is counter based: Philox4x32-10 keyed by the caller's seed, counter =
// (global row, layer * 4096 + column / 8, step): 8 columns per call, 16 bits each, dropped when bits < p * 65536.
// It depends only on (seed, step, global row, layer, column), so the backward pass regenerates it and the result
// does not depend on how rows are sharded over GPUs.
__device__ __forceinline__ unsigned int drop_keep8(const DropoutArgs& d, long long grow, int layer, int col8) {
  const uint4 u = philox4x32_10(make_uint4((unsigned int)grow, (unsigned int)(grow >> 32),
                                           (unsigned int)(layer * 4096 + col8), d.step),
                                make_uint2((unsigned int)d.seed, (unsigned int)(d.seed >> 32)));
  const unsigned int w[4] = {u.x, u.y, u.z, u.w};
  unsigned int keep = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) keep |= (((w[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu) >= d.thresh16 ? 1u : 0u) << i;
  return keep;
}

// layer 1 (K = 4): a warp per row, a lane owns 8 of the 256 columns.  Writes xhat (normalised, pre-affine), the
// post-dropout activation, 1/std per row and (tests) the keep-mask as bytes.
__global__ void __launch_bounds__(kThreads) f_l1_train_kernel(const float* __restrict__ p, const float* __restrict__ w1,
                                                              const float* __restrict__ b1,
                                                              const float* __restrict__ lnw,
                                                              const float* __restrict__ lnb, __half* __restrict__ xhat,
                                                              __half* __restrict__ act, float* __restrict__ rstd_out,
                                                              unsigned char* __restrict__ mask_out,
                                                              unsigned char* __restrict__ keepbits, long long rows,
                                                              DropoutArgs dr) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  float4 wr[8];
  float b[8], gm[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wr[i] = __ldg(reinterpret_cast<const float4*>(w1) + lane * 8 + i);
  ld_f8(b1 + lane * 8, b);
  ld_f8(lnw + lane * 8, gm);
  ld_f8(lnb + lane * 8, bt);
  constexpr int R = 4;   // rows per warp and trip: four independent reduction chains in flight
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * R;
  for (long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; r0 < rows; r0 += wstride) {
    float h[R][8], s[R], v[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = r0 + u < rows ? r0 + u : rows - 1;
      const float4 q = __ldg(reinterpret_cast<const float4*>(p) + row);
      s[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h[u][i] = fmaf(q.w, wr[i].w, fmaf(q.z, wr[i].z, fmaf(q.y, wr[i].y, fmaf(q.x, wr[i].x, b[i]))));
        s[u] += h[u][i];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const float mean = s[u] * (1.0f / 256.0f);
      v[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h[u][i] -= mean;
        v[u] = fmaf(h[u][i], h[u][i], v[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = r0 + u;
      if (row >= rows) break;
      const float rstd = 1.0f / sqrtf(v[u] * (1.0f / 256.0f) + kLnEps);
      const unsigned int keep = drop_keep8(dr, dr.first_row + row, 0, lane);
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h[u][i] *= rstd;
        a[i] = (keep >> i) & 1u ? lrelu_f(fmaf(h[u][i], gm[i], bt[i])) * dr.keep_scale : 0.f;
      }
      st_h8(xhat + row * 256 + lane * 8, h[u]);
      st_h8(act + row * 256 + lane * 8, a);
      keepbits[row * 32 + lane] = (unsigned char)keep;   // 1 bit per element: the backward pass reads it back
      if (lane == 0) rstd_out[row] = rstd;
      if (mask_out) {
        unsigned long long m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= (unsigned long long)((keep >> i) & 1u) << (8 * i);
        *reinterpret_cast<unsigned long long*>(mask_out + row * 256 + lane * 8) = m;
      }
    }
  }
}

// layers 2..5: h (fp16 Linear output incl. bias, in `xhat`) + fp32 row partials from the GEMM epilogue -> xhat in
// place, activation, 1/std.  A warp per row, NCH = N / 256 chunks of 8 columns per lane.
template <int NCH>
__global__ void __launch_bounds__(kThreads) ln_train_kernel(__half* __restrict__ xhat, const float* __restrict__ rowstats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __half* __restrict__ act,
                                                            float* __restrict__ rstd_out,
                                                            unsigned char* __restrict__ mask_out,
                                                            unsigned char* __restrict__ keepbits, long long rows,
                                                            int layer, DropoutArgs dr) {
  pdl_wait();
  constexpr int N = NCH * 256;
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int t = 0; t < NCH; ++t) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(rowstats) + row * NCH + t);
      s1 += st.x;
      s2 += st.y;
    }
    const float mean = s1 * (1.0f / N);
    const float rstd = 1.0f / sqrtf(fmaxf(s2 * (1.0f / N) - mean * mean, 0.f) + kLnEps);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int c0 = j * 256 + lane * 8;
      float h[8], gm[8], bt[8], a[8];
      ld_h8(xhat + row * N + c0, h);
      ld_f8(gamma + c0, gm);
      ld_f8(beta + c0, bt);
      const unsigned int keep = drop_keep8(dr, dr.first_row + row, layer, c0 >> 3);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h[i] = (h[i] - mean) * rstd;
        a[i] = (keep >> i) & 1u ? lrelu_f(fmaf(h[i], gm[i], bt[i])) * dr.keep_scale : 0.f;
      }
      st_h8(xhat + row * N + c0, h);
      st_h8(act + row * N + c0, a);
      keepbits[row * (N / 8) + (c0 >> 3)] = (unsigned char)keep;
      if (mask_out) {
        unsigned long long m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= (unsigned long long)((keep >> i) & 1u) << (8 * i);
        *reinterpret_cast<unsigned long long*>(mask_out + row * N + c0) = m;
      }
    }
    if (lane == 0) rstd_out[row] = rstd;
  }
}

// Output layer loss: out [rows, S + Mt] fp32 against the spectrum / normalised metrics (MSELoss means,
// pretrain_fwd_model.py:80-84).  dout (fp16, ld columns, zero padded) = d(loss_spec + loss_metrics)/d(out) * GS
// (GS = global batch: every gradient of the step carries that scale until dw_reduce / the partial reductions remove
// it, which keeps the fp16 operands in range).  Per block: column sums of dout (bias gradient) and the two sums
// of squares go to part[block][0 .. S+Mt) and part[block][ld_part - 2 .. ld_part).
__global__ void __launch_bounds__(kThreads) f_out_loss_kernel(const float* __restrict__ out,
                                                              const float* __restrict__ spectrum,
                                                              const float* __restrict__ metrics,
                                                              __half* __restrict__ dout, int ld, long long rows, int S,
                                                              int Mt, float* __restrict__ part, int ld_part,
                                                              float w_spec, float w_met) {
  pdl_wait();
  // a lane owns the column pairs (2 lane + 64 j, +1), j < 5: 8-byte loads, 4-byte stores; S, Mt and ld are even,
  // so a pair never straddles the spectrum / metrics / padding boundaries.  Two rows per warp and trip.
  __shared__ float sm[8][328];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int OUT = S + Mt;
  const float gs_spec = w_spec * 2.0f / (float)S, gs_met = w_met * 2.0f / (float)Mt;   // loss weights (1, 1 in training)
  constexpr int J = 5;    // covers ld <= 320
  constexpr int R = 2;
  float colsum[J][2];
#pragma unroll
  for (int j = 0; j < J; ++j) colsum[j][0] = colsum[j][1] = 0.f;
  float sq_spec = 0.f, sq_met = 0.f;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * R;
  for (long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * R; r0 < rows; r0 += wstride) {
    float2 o[R][J], t[R][J];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = r0 + u < rows ? r0 + u : rows - 1;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = j * 64 + 2 * lane;
        o[u][j] = t[u][j] = make_float2(0.f, 0.f);
        if (c < OUT) {
          o[u][j] = *reinterpret_cast<const float2*>(out + row * OUT + c);
          t[u][j] = c < S ? __ldg(reinterpret_cast<const float2*>(spectrum + row * S + c))
                          : __ldg(reinterpret_cast<const float2*>(metrics + row * Mt + (c - S)));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (r0 + u >= rows) break;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = j * 64 + 2 * lane;
        if (c >= ld) continue;
        const float d0 = o[u][j].x - t[u][j].x, d1 = o[u][j].y - t[u][j].y;   // zero in the padding columns
        const float sq = fmaf(d0, d0, d1 * d1);
        if (c < S) sq_spec += sq;
        else sq_met += sq;
        const float gsc = c < S ? gs_spec : gs_met;
        const float g0 = d0 * gsc, g1 = d1 * gsc;
        *reinterpret_cast<__half2*>(dout + (r0 + u) * ld + c) = __floats2half2_rn(g0, g1);
        colsum[j][0] += g0;
        colsum[j][1] += g1;
      }
    }
  }
  sq_spec = warp_sum_f(sq_spec);
  sq_met = warp_sum_f(sq_met);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    sm[warp][j * 64 + 2 * lane] = colsum[j][0];
    sm[warp][j * 64 + 2 * lane + 1] = colsum[j][1];
  }
  if (lane == 0) {
    sm[warp][320] = sq_spec;
    sm[warp][321] = sq_met;
  }
  __syncthreads();
  float* prow = part + (size_t)blockIdx.x * ld_part;
  for (int c = threadIdx.x; c < ld_part; c += blockDim.x) {
    const int src = c >= ld_part - 2 ? 320 + (c - (ld_part - 2)) : c;
    float tt = 0.f;
    if (src < 322 && (c < OUT || c >= ld_part - 2))
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) tt += sm[wq][src];
    prow[c] = tt;
  }
}

// Backward through Dropout, LeakyReLU and LayerNorm of one surrogate layer:
//   dact = da * keep / (1 - p);  dy = dact * (y > 0 ? 1 : 0.2), y = gamma * xhat + beta;  dxh = dy * gamma
//   dh = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat))                      (written over da)
// (keep = the forward pass's mask, 1 bit per element in keepbits: reading 320 B per row is cheaper than a second
// Philox evaluation, which was 38 % of this kernel's instructions)
// and per-column sums over the rows for dbeta (dy), dgamma (dy * xhat), dbias (dh) -- FIRST: also dW1 (dh * p_j) --
// as per-block partials part[block][k * N + c], k = 0..2 (3..6), finished by reduce_partials_kernel.
// Layout: a thread owns 8 consecutive columns for the whole kernel (its 24 / 56 column accumulators stay in
// registers) and R = 4 rows per trip; WPR = N / 256 warps share a row and meet through shared memory for the two row
// means, so a block of 8 warps works on 8 / WPR row groups x 4 rows at a time.  (The first version gave a warp the
// whole row: 96 accumulators per lane at N = 1024, 254 registers, 8 warps per SM, 2.7x the HBM time.)
template <int NCH, bool FIRST>
__global__ void __launch_bounds__(kThreads, 2) ln_bwd_kernel(__half* __restrict__ da, const __half* __restrict__ xhat,
                                                             const float* __restrict__ rstd_in,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ p_in,
                                                             const unsigned char* __restrict__ keepbits,
                                                             long long rows, float keep_scale,
                                                             float* __restrict__ part, int store_dh) {
  pdl_wait();
  constexpr int N = NCH * 256;
  constexpr int NQ = FIRST ? 7 : 3;
  constexpr int WPR = NCH;          // warps per row
  constexpr int RG = 8 / WPR;       // row groups per block
  constexpr int R = FIRST ? 2 : 4;  // rows per group and trip (the first layer carries 56 column accumulators)
  __shared__ float sm[NQ * N];
  __shared__ float xch[2][RG][R][WPR][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp % WPR, grp = warp / WPR;
  const int c0 = wr * 256 + lane * 8;
  float gm[8], bt[8], acc[NQ][8];
  ld_f8(gamma + c0, gm);
  ld_f8(beta + c0, bt);
#pragma unroll
  for (int k = 0; k < NQ; ++k) zero8(acc[k]);
  const long long rows_per_trip = (long long)gridDim.x * RG * R;
  const long long trips = (rows + rows_per_trip - 1) / rows_per_trip;
  auto unpack = [](const uint4& u, float* v) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  };
  for (long long t = 0; t < trips; ++t) {
    const long long rbase = t * rows_per_trip + ((long long)blockIdx.x * RG + grp) * R;
    // the rows stay packed (fp16) in registers between the two phases; dy is recomputed in phase 2
    uint4 graw[R], xraw[R];
    unsigned int keep[R];
    float m1[R], m2[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = rbase + u;
      if (row < rows) {
        graw[u] = *reinterpret_cast<const uint4*>(da + row * N + c0);
        xraw[u] = *reinterpret_cast<const uint4*>(xhat + row * N + c0);
      } else {
        graw[u] = make_uint4(0u, 0u, 0u, 0u);
        xraw[u] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = rbase + u;
      keep[u] = row < rows ? (unsigned int)__ldg(keepbits + row * (N / 8) + (c0 >> 3)) : 0u;
      float g[8], xh[8];
      unpack(graw[u], g);
      unpack(xraw[u], xh);
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = fmaf(xh[i], gm[i], bt[i]);
        const float dy = (keep[u] >> i) & 1u ? g[i] * keep_scale * (y > 0.f ? 1.f : kSlope) : 0.f;
        acc[0][i] += dy;
        acc[1][i] = fmaf(dy, xh[i], acc[1][i]);
        const float dx = dy * gm[i];
        a1 += dx;
        a2 = fmaf(dx, xh[i], a2);
      }
      m1[u] = warp_sum_f(a1);
      m2[u] = warp_sum_f(a2);
    }
    if constexpr (WPR > 1) {
      const int par = (int)(t & 1);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          xch[par][grp][u][wr][0] = m1[u];
          xch[par][grp][u][wr][1] = m2[u];
        }
      }
      __syncthreads();   // uniform trip count: every warp of the block gets here
#pragma unroll
      for (int u = 0; u < R; ++u) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) {
          s1 += xch[par][grp][u][w][0];
          s2 += xch[par][grp][u][w][1];
        }
        m1[u] = s1;
        m2[u] = s2;
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = rbase + u;
      if (row >= rows) continue;
      const float rstd = __ldg(rstd_in + row);
      const float mu1 = m1[u] * (1.0f / N), mu2 = m2[u] * (1.0f / N);
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (FIRST) q = __ldg(reinterpret_cast<const float4*>(p_in) + row);
      float g[8], xh[8], dh[8];
      unpack(graw[u], g);
      unpack(xraw[u], xh);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = fmaf(xh[i], gm[i], bt[i]);
        const float dy = (keep[u] >> i) & 1u ? g[i] * keep_scale * (y > 0.f ? 1.f : kSlope) : 0.f;
        dh[i] = rstd * (dy * gm[i] - mu1 - xh[i] * mu2);
        acc[2][i] += dh[i];
        if constexpr (FIRST) {
          acc[3][i] = fmaf(dh[i], q.x, acc[3][i]);
          acc[4][i] = fmaf(dh[i], q.y, acc[4][i]);
          acc[5][i] = fmaf(dh[i], q.z, acc[5][i]);
          acc[6][i] = fmaf(dh[i], q.w, acc[6][i]);
        }
      }
      if (!FIRST || store_dh) st_h8(da + row * N + c0, dh);   // the first layer keeps dh only for the input gradient
    }
  }
  // block combine over the row groups in a fixed order (deterministic), then this block's row of the partial scratch
  for (int i = threadIdx.x; i < NQ * N; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  for (int gq = 0; gq < RG; ++gq) {
    if (grp == gq) {
#pragma unroll
      for (int k = 0; k < NQ; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[k * N + c0 + i] += acc[k][i];
    }
    __syncthreads();
  }
  float* prow = part + (size_t)blockIdx.x * (NQ * N);
  for (int i = threadIdx.x; i < NQ * N; i += blockDim.x) prow[i] = sm[i];
}

// input gradient of the surrogate: dp[r, j] = scale * sum_c dh1[r, c] * W1[c, j]  (W1 [256, 4]); a warp per row
__global__ void __launch_bounds__(kThreads) f_dp_kernel(const __half* __restrict__ dh1, const float* __restrict__ w1,
                                                        float* __restrict__ dp, long long rows, float scale) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  float4 wr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wr[i] = __ldg(reinterpret_cast<const float4*>(w1) + lane * 8 + i);
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    float d[8];
    ld_h8(dh1 + row * 256 + lane * 8, d);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a0 = fmaf(d[i], wr[i].x, a0);
      a1 = fmaf(d[i], wr[i].y, a1);
      a2 = fmaf(d[i], wr[i].z, a2);
      a3 = fmaf(d[i], wr[i].w, a3);
    }
    a0 = warp_sum_f(a0); a1 = warp_sum_f(a1); a2 = warp_sum_f(a2); a3 = warp_sum_f(a3);
    if (lane == 0) *reinterpret_cast<float4*>(dp + row * 4) = make_float4(a0 * scale, a1 * scale, a2 * scale, a3 * scale);
  }
}

// dW1 arrives from the partial reduction as [4][256] (k-major); the state_dict layout is [256][4]
__global__ void f_dw1_transpose_kernel(const float* __restrict__ src, float* __restrict__ dw1) {
  pdl_wait();
  const int c = threadIdx.x;   // 256 threads
#pragma unroll
  for (int j = 0; j < 4; ++j) dw1[c * 4 + j] = src[j * 256 + c];
}

__global__ void f_input_grad_losses_kernel(const float* __restrict__ sums, double n_spec, double n_met, float* out) {
  pdl_wait();
  if (threadIdx.x != 0) return;
  out[0] = (float)((double)sums[0] / n_spec);
  out[1] = (float)((double)sums[1] / n_met);
}
__global__ void f_train_losses_kernel(const float* __restrict__ sums, double n_spec, double n_met, float* out) {
  pdl_wait();
  if (threadIdx.x != 0) return;
  const float ls = (float)((double)sums[0] / n_spec), lm = (float)((double)sums[1] / n_met);
  out[0] = ls + lm;
  out[1] = ls;
  out[2] = lm;
}


// ------------------------------------------------------------------------------------------ widened surrogate
// BASELINE config 5 (hidden 2048, 2048-point spectra): the surrogate's streaming kernels for any hidden width
// N = 256 * NCH (NCH = 1, 2, 4, 8).  Same arithmetic as the reference-width kernels above; a lane owns 8 columns of
// every 256-column chunk of its row.
// First layer (K = 4, forward_model.py:30-33).  With K = 4 the LayerNorm statistics of a row are a closed form in its
// four inputs: h_c - mean = U_c . p + v_c with the CENTRED weights U_c = W_c - mean_c(W), v_c = b_c - mean(b), so
//   var = p^T A p + 2 p^T B + C,   A = mean_c(U_c^T U_c), B = mean_c(U_c v_c), C = mean_c(v_c^2)
// (a positive semi-definite form: no cancellation).  f_l1_consts_kernel computes the 20 constants in fp64 once per
// call; the row kernel then needs no reduction over the row at all - a thread owns 8 columns (their centred weights
// stay in registers for the whole kernel) and walks the rows.  (The first version gave a warp a row and re-read the
// 32 KB weight matrix per row from L1: 655 us for a 540 MB output at B = 65 536, N = 2048.)
constexpr int kL1Consts = 20;   // [0:4] mean_c(W), [4] mean(b), [5:15] A (upper triangle, row-major), [15:19] B, [19] C
__global__ void __launch_bounds__(256) f_l1_consts_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                                          int N, float* __restrict__ out) {
  pdl_wait();
  __shared__ double sm[256][15];
  __shared__ double mean[5];
  const int t = threadIdx.x;
  double a[15];
  for (int k = 0; k < 5; ++k) a[k] = 0.0;
  for (int c = t; c < N; c += 256) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(w1) + c);
    a[0] += w.x; a[1] += w.y; a[2] += w.z; a[3] += w.w; a[4] += b1[c];
  }
  for (int k = 0; k < 5; ++k) sm[t][k] = a[k];
  __syncthreads();
  if (t < 5) {
    double s = 0.0;
    for (int i = 0; i < 256; ++i) s += sm[i][t];
    mean[t] = s / N;
  }
  __syncthreads();
  for (int k = 0; k < 15; ++k) a[k] = 0.0;
  for (int c = t; c < N; c += 256) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(w1) + c);
    const double u[4] = {w.x - mean[0], w.y - mean[1], w.z - mean[2], w.w - mean[3]};
    const double v = b1[c] - mean[4];
    int k = 0;
    for (int i = 0; i < 4; ++i)
      for (int j = i; j < 4; ++j) a[k++] += u[i] * u[j];
    for (int i = 0; i < 4; ++i) a[10 + i] += u[i] * v;
    a[14] += v * v;
  }
  for (int k = 0; k < 15; ++k) sm[t][k] = a[k];
  __syncthreads();
  if (t < 15) {
    double s = 0.0;
    for (int i = 0; i < 256; ++i) s += sm[i][t];
    out[5 + t] = (float)(s / N);
  }
  if (t < 5) out[t] = (float)mean[t];
}

template <bool TRAIN>
__global__ void __launch_bounds__(kThreads) f_l1_wide_kernel(const float* __restrict__ p, const float* __restrict__ w1,
                                                             const float* __restrict__ b1,
                                                             const float* __restrict__ lnw,
                                                             const float* __restrict__ lnb,
                                                             const float* __restrict__ consts,
                                                             __half* __restrict__ xhat, __half* __restrict__ act,
                                                             float* __restrict__ rstd_out,
                                                             unsigned char* __restrict__ mask_out,
                                                             unsigned char* __restrict__ keepbits, long long rows,
                                                             int N, DropoutArgs dr) {
  pdl_wait();
  const ColMap m(N);
  const int c0 = m.ch * 8;
  float k[kL1Consts];
#pragma unroll
  for (int i = 0; i < kL1Consts; ++i) k[i] = __ldg(consts + i);
  float4 u[8];
  float v[8], gm[8], bt[8];
  ld_f8(b1 + c0, v);
  ld_f8(lnw + c0, gm);
  ld_f8(lnb + c0, bt);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(w1) + c0 + i);
    u[i] = make_float4(w.x - k[0], w.y - k[1], w.z - k[2], w.w - k[3]);
    v[i] -= k[4];
  }
  for (long long row = (long long)blockIdx.x * m.rpb + m.rg; row < rows; row += (long long)gridDim.x * m.rpb) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p) + row);
    // var = q^T A q + 2 q^T B + C
    float var = k[19];
    var = fmaf(q.x, fmaf(q.x, k[5], 2.f * fmaf(q.y, k[6], fmaf(q.z, k[7], fmaf(q.w, k[8], k[15])))), var);
    var = fmaf(q.y, fmaf(q.y, k[9], 2.f * fmaf(q.z, k[10], fmaf(q.w, k[11], k[16]))), var);
    var = fmaf(q.z, fmaf(q.z, k[12], 2.f * fmaf(q.w, k[13], k[17])), var);
    var = fmaf(q.w, fmaf(q.w, k[14], 2.f * k[18]), var);
    const float rstd = 1.0f / sqrtf(fmaxf(var, 0.f) + kLnEps);
    float xh[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      xh[i] = fmaf(q.w, u[i].w, fmaf(q.z, u[i].z, fmaf(q.y, u[i].y, fmaf(q.x, u[i].x, v[i])))) * rstd;
    if constexpr (TRAIN) {
      const unsigned int keep = drop_keep8(dr, dr.first_row + row, 0, c0 >> 3);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = (keep >> i) & 1u ? lrelu_f(fmaf(xh[i], gm[i], bt[i])) * dr.keep_scale : 0.f;
      st_h8(xhat + row * N + c0, xh);
      keepbits[row * (N / 8) + (c0 >> 3)] = (unsigned char)keep;
      if (mask_out) {
        unsigned long long mk = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) mk |= (unsigned long long)((keep >> i) & 1u) << (8 * i);
        *reinterpret_cast<unsigned long long*>(mask_out + row * N + c0) = mk;
      }
      if (m.ch == 0) rstd_out[row] = rstd;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = lrelu_f(fmaf(xh[i], gm[i], bt[i]));
    }
    st_h8(act + row * N + c0, a);
  }
}

// dW1 (k-major partials [4][N]) = sum_r dh1[r, c] * p[r, j]: the first layer's weight gradient when its LayerNorm
// backward ran through the generic kernel (ln_bwd_kernel<NCH, false>, dh kept in place).  A thread owns 8 columns.
__global__ void __launch_bounds__(kThreads) f_dw1_wide_kernel(const __half* __restrict__ dh1, const float* __restrict__ p,
                                                              long long rows, int N, float* __restrict__ part) {
  pdl_wait();
  __shared__ float sm[kThreads * 8];
  const ColMap m(N);
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) zero8(acc[k]);
  for (long long r = (long long)blockIdx.x * m.rpb + m.rg; r < rows; r += (long long)gridDim.x * m.rpb) {
    float d[8];
    ld_h8(dh1 + r * N + m.ch * 8, d);
    const float4 q = __ldg(reinterpret_cast<const float4*>(p) + r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] = fmaf(d[i], q.x, acc[0][i]);
      acc[1][i] = fmaf(d[i], q.y, acc[1][i]);
      acc[2][i] = fmaf(d[i], q.z, acc[2][i]);
      acc[3][i] = fmaf(d[i], q.w, acc[3][i]);
    }
  }
  float* prow = part + (size_t)blockIdx.x * (4 * N);
#pragma unroll
  for (int k = 0; k < 4; ++k) block_colsum_partial(acc[k], prow + k * N, m, sm);
}

// dp[r, j] = scale * sum_c dh1[r, c] * W1[c, j] for any N = 256 * NCH; a warp per row
__global__ void __launch_bounds__(kThreads) f_dp_wide_kernel(const __half* __restrict__ dh1, const float* __restrict__ w1,
                                                             float* __restrict__ dp, long long rows, int nch,
                                                             float scale) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int N = nch * 256;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j = 0; j < nch; ++j) {
      const int c0 = j * 256 + lane * 8;
      float d[8];
      ld_h8(dh1 + row * N + c0, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w1) + c0 + i);
        a0 = fmaf(d[i], w.x, a0);
        a1 = fmaf(d[i], w.y, a1);
        a2 = fmaf(d[i], w.z, a2);
        a3 = fmaf(d[i], w.w, a3);
      }
    }
    a0 = warp_sum_f(a0); a1 = warp_sum_f(a1); a2 = warp_sum_f(a2); a3 = warp_sum_f(a3);
    if (lane == 0) *reinterpret_cast<float4*>(dp + row * 4) = make_float4(a0 * scale, a1 * scale, a2 * scale, a3 * scale);
  }
}

__global__ void f_dw1_transpose_wide_kernel(const float* __restrict__ src, float* __restrict__ dw1, int N) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) dw1[c * 4 + j] = src[j * N + c];
}

// The widened output layer leaves its fp32 accumulators as [128 x 256] slabs (EpiWeightGradPartial on a TN GEMM):
// element (r, c) of the [rows, OUT] product sits at slab[((r / 128) * ngroups + c / 256) * 32768 + (r % 128) * 256 + c % 256].
__device__ __forceinline__ size_t slab_index(long long r, int c, int ngroups) {
  return ((size_t)(r >> 7) * ngroups + (size_t)(c >> 8)) * 32768u + (size_t)(r & 127) * 256u + (size_t)(c & 255);
}
// out[r, c] = slab(r, c) + bias[c]: the row-major fp32 output of pigan_forward_model_forward
__global__ void __launch_bounds__(kThreads) f_unslab_kernel(const float* __restrict__ slab, int ngroups,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            long long rows, int OUT) {
  pdl_wait();
  const int pairs = OUT >> 1;   // OUT is even
  for (long long r = blockIdx.x; r < rows; r += gridDim.x)
    for (int q = threadIdx.x; q < pairs; q += blockDim.x) {
      const int c = 2 * q;
      const float2 v = *reinterpret_cast<const float2*>(slab + slab_index(r, c, ngroups));
      const float2 b = __ldg(reinterpret_cast<const float2*>(bias + c));
      *reinterpret_cast<float2*>(out + r * OUT + c) = make_float2(v.x + b.x, v.y + b.y);
    }
}
// f_out_loss_kernel for the slab layout and any width (ld <= 512 * kWideJ): a block walks rows, a thread owns the
// column pairs (2 t + 512 j, +1).  Same outputs: dout (fp16, GS-scaled, zero padded to ld), per-block column sums of
// dout in part[block][0 .. OUT) and the two sums of squares in part[block][ld_part - 2 .. ld_part).
constexpr int kWideJ = 5;
__global__ void __launch_bounds__(kThreads) f_out_loss_slab_kernel(const float* __restrict__ slab, int ngroups,
                                                                   const float* __restrict__ bias,
                                                                   const float* __restrict__ spectrum,
                                                                   const float* __restrict__ metrics,
                                                                   __half* __restrict__ dout, int ld, long long rows,
                                                                   int S, int Mt, float* __restrict__ part,
                                                                   int ld_part, float w_spec, float w_met) {
  pdl_wait();
  __shared__ float sm[8];
  const int OUT = S + Mt;
  const float gs_spec = w_spec * 2.0f / (float)S, gs_met = w_met * 2.0f / (float)Mt;
  float colsum[kWideJ][2];
  float2 bj[kWideJ];
#pragma unroll
  for (int j = 0; j < kWideJ; ++j) {
    colsum[j][0] = colsum[j][1] = 0.f;
    const int c = j * 512 + 2 * (int)threadIdx.x;
    bj[j] = c < OUT ? __ldg(reinterpret_cast<const float2*>(bias + c)) : make_float2(0.f, 0.f);
  }
  float sq_spec = 0.f, sq_met = 0.f;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
#pragma unroll
    for (int j = 0; j < kWideJ; ++j) {
      const int c = j * 512 + 2 * (int)threadIdx.x;
      if (c >= ld) continue;
      float g0 = 0.f, g1 = 0.f;
      if (c < OUT) {
        const float2 o = *reinterpret_cast<const float2*>(slab + slab_index(r, c, ngroups));
        const float2 t = c < S ? __ldg(reinterpret_cast<const float2*>(spectrum + r * S + c))
                               : __ldg(reinterpret_cast<const float2*>(metrics + r * Mt + (c - S)));
        const float d0 = o.x + bj[j].x - t.x, d1 = o.y + bj[j].y - t.y;
        const float sq = fmaf(d0, d0, d1 * d1);
        if (c < S) sq_spec += sq;
        else sq_met += sq;
        const float gsc = c < S ? gs_spec : gs_met;
        g0 = d0 * gsc;
        g1 = d1 * gsc;
        colsum[j][0] += g0;
        colsum[j][1] += g1;
      }
      *reinterpret_cast<__half2*>(dout + r * ld + c) = __floats2half2_rn(g0, g1);
    }
  }
  float* prow = part + (size_t)blockIdx.x * ld_part;
#pragma unroll
  for (int j = 0; j < kWideJ; ++j) {
    const int c = j * 512 + 2 * (int)threadIdx.x;
    if (c < OUT) *reinterpret_cast<float2*>(prow + c) = make_float2(colsum[j][0], colsum[j][1]);
  }
  for (int c = OUT + (int)threadIdx.x; c < ld_part - 2; c += blockDim.x) prow[c] = 0.f;
  const float t0 = block_sum(sq_spec, sm);
  const float t1 = block_sum(sq_met, sm);
  if (threadIdx.x == 0) {
    prow[ld_part - 2] = t0;
    prow[ld_part - 1] = t1;
  }
}


// ------------------------------------------------------------------------------------------ widened PI-GAN step
// Generic-width versions of the pieces the reference-width step fuses into GEMM epilogues (BASELINE config 5: hidden
// 2048, 2048-point spectra).  At those widths the GEMMs carry 193 MFLOP per sample and these passes a few percent.
// Generator head (generator.py:22-25) for any C = 256 k: a warp per row, a lane owns 8 columns of every 256-column
// chunk.  Same outputs as g_head_fwd_kernel (p, denormalised p, the fake-row tail of the spectrum operand).
__global__ void __launch_bounds__(kThreads) wide_head_fwd_kernel(
    const __half* __restrict__ h2, const float* __restrict__ scale, const float* __restrict__ bias,
    const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ p_out,
    float* __restrict__ pden_out, const __half* __restrict__ xc, __half* __restrict__ tail_fake, long long rows,
    int C, int Kp, int S) {
  pdl_wait();
  constexpr int R = 4;   // rows per warp and trip: the per-column constants (BatchNorm affine, 4 rows of W3) are
                         // fetched once per chunk for all of them (one row per trip was bound by those L1 reads)
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5) * R;
  for (long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; r0 < rows; r0 += wstride) {
    float acc[R][4];
#pragma unroll
    for (int u = 0; u < R; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float sc[8], bi[8], w[4][8], h[R][8];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const long long row = r0 + u < rows ? r0 + u : rows - 1;
        ld_h8(h2 + row * C + c0, h[u]);
      }
      ld_f8(scale + c0, sc);
      ld_f8(bias + c0, bi);
#pragma unroll
      for (int e = 0; e < 4; ++e) ld_f8(w3 + (size_t)e * C + c0, w[e]);
#pragma unroll
      for (int u = 0; u < R; ++u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h[u][i] = fmaxf(fmaf(sc[i], h[u][i], bi[i]), 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[u][e] = fmaf(h[u][i], w[e][i], acc[u][e]);
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[u][e] = warp_sum_f(acc[u][e]);
    const float b3l = __ldg(b3 + (lane & 3));
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = r0 + u;
      if (row >= rows) break;   // uniform across the warp
      const float mine = (lane & 3) == 0 ? acc[u][0] : (lane & 3) == 1 ? acc[u][1] : (lane & 3) == 2 ? acc[u][2] : acc[u][3];
      const float pv = tanhf(mine + b3l);                            // every lane: parameter lane % 4
      const float pd = (pv + 1.0f) / 2.0f * 0.6f + 2.2f;             // data_loader.py:238-252
      if (lane < 4) {
        p_out[row * 4 + lane] = pv;
        if (pden_out) pden_out[row * 4 + lane] = pd;
      }
      if (tail_fake != nullptr) {
        // last 64 operand columns of the row (8 chunks of 16 bytes), parameter columns replaced
        const int t0 = Kp - 64;
        float pr[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pr[e] = __shfl_sync(0xffffffffu, pd, e);
        if (lane < 8) {
          float v[8];
          ld_h8(xc + row * Kp + t0 + lane * 8, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = t0 + lane * 8 + k - S;
            if (e >= 0 && e < 4) v[k] = pr[e] - kParamCenter;
          }
          st_h8(tail_fake + row * 64 + lane * 8, v);
        }
      }
    }
  }
}

// Discriminator layer 3 + Sigmoid + BCE (discriminator.py:26-28, loss.py:8-17) on the stored layer-2 activation
// z2 [rows, C]: logit = z2 . w3 + b3; rows < rows_a carry label_a, the others label_b; rows in [gap_begin, gap_end) are
// padding.  Same arithmetic as EpiDiscL2: loss_sum += sum BCE / global batch, dlogit (scaled by GS = global batch).
__global__ void __launch_bounds__(kThreads) d_logit_bce_kernel(const __half* __restrict__ z2,
                                                               const float* __restrict__ w3,
                                                               const float* __restrict__ b3, long long rows, int C,
                                                               long long rows_a, float label_a, float label_b,
                                                               long long gap_begin, long long gap_end,
                                                               float inv_batch, double* __restrict__ loss_sum,
                                                               float* __restrict__ dlogit,
                                                               float* __restrict__ prob_out) {
  pdl_wait();
  __shared__ float sm[8];
  const int lane = threadIdx.x & 31;
  float loss = 0.f;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const bool valid = !(row >= gap_begin && row < gap_end);
    float lg = 0.f;
    if (valid) {
      for (int c0 = lane * 8; c0 < C; c0 += 256) {
        float z[8], w[8];
        ld_h8(z2 + row * C + c0, z);
        ld_f8(w3 + c0, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) lg = fmaf(z[i], w[i], lg);
      }
    }
    lg = warp_sum_f(lg);
    if (lane != 0) continue;
    if (valid) {
      const float logit = lg + __ldg(b3);
      const float prob = 1.f / (1.f + expf(-logit));
      const float y = row < rows_a ? label_a : label_b;
      const float lp = fmaxf(logf(prob), -100.f);
      const float l1p = fmaxf(log1pf(-prob), -100.f);
      loss += -(y * lp + (1.f - y) * l1p) * inv_batch;
      const float pq = prob * (1.f - prob);
      if (dlogit) dlogit[row] = (prob - y) / fmaxf(pq, 1e-12f) * pq;
      if (prob_out) prob_out[row] = prob;
    } else if (dlogit) {
      dlogit[row] = 0.f;
    }
  }
  const float t = block_sum(loss, sm);
  if (threadIdx.x == 0 && t != 0.f && loss_sum) atomicAdd(loss_sum, (double)t);
}

// Surrogate losses of the G-step (train_pigan.py:156-172) from the widened output layer's fp32 slabs: per row
// recon = slab + bias; sums[0] += sum (recon - x)^2, [1] += sum (pm - m)^2, [2] += sum of squared second differences
// (loss.py:29-64), [3], [4] += the two LC terms (loss.py:67-101); dp_lc [rows, 4] = lambda_lc * dLC/dp * GS (no gradient
// through F: the reference evaluates it under no_grad).
__global__ void __launch_bounds__(kThreads) f_pigan_loss_slab_kernel(const float* __restrict__ slab, int ngroups,
                                                                     const float* __restrict__ bias,
                                                                     const float* __restrict__ spectrum,
                                                                     const float* __restrict__ metrics,
                                                                     const float* __restrict__ p_norm, long long rows,
                                                                     int S, int Mt, int f1_idx, int f2_idx,
                                                                     float lc_grad_mult, double* __restrict__ sums,
                                                                     float* __restrict__ dp_lc) {
  pdl_wait();
  // A warp per row, a lane owns the column pairs (2 lane + 64 j, +1): S and Mt are even, so a pair never straddles the
  // spectrum / metrics boundary.  The second difference at a column needs its two left neighbours: they are the
  // previous lane's pair (a shuffle) or, for lane 0, the last pair of the previous 64-column chunk (carried along).
  __shared__ float red[8];
  const int lane = threadIdx.x & 31;
  const int OUT = S + Mt;
  float rec = 0.f, met = 0.f, mx = 0.f, lc1 = 0.f, lc2 = 0.f;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wstride) {
    float carry_x = 0.f, carry_y = 0.f;   // recon[c0 - 2], recon[c0 - 1] of the chunk being processed
    float f1 = 0.f, f2 = 0.f;
#pragma unroll 2
    for (int c0 = 0; c0 < OUT; c0 += 64) {
      const int c = c0 + 2 * lane;
      const bool in = c < OUT;
      float2 o = make_float2(0.f, 0.f), t = make_float2(0.f, 0.f);
      if (in) {
        const float2 v = *reinterpret_cast<const float2*>(slab + slab_index(r, c, ngroups));
        const float2 b = __ldg(reinterpret_cast<const float2*>(bias + c));
        o = make_float2(v.x + b.x, v.y + b.y);
        t = c < S ? __ldg(reinterpret_cast<const float2*>(spectrum + r * S + c))
                  : __ldg(reinterpret_cast<const float2*>(metrics + r * Mt + (c - S)));
      }
      float left_x = __shfl_up_sync(0xffffffffu, o.x, 1), left_y = __shfl_up_sync(0xffffffffu, o.y, 1);
      if (lane == 0) {
        left_x = carry_x;
        left_y = carry_y;
      }
      carry_x = __shfl_sync(0xffffffffu, o.x, 31);
      carry_y = __shfl_sync(0xffffffffu, o.y, 31);
      if (!in) continue;
      const float d0 = o.x - t.x, d1 = o.y - t.y;
      if (c < S) {
        rec = fmaf(d0, d0, fmaf(d1, d1, rec));
        if (c >= 2) {   // loss.py:51-53: difference of differences at columns c and c + 1
          const float da = (o.x - left_y) - (left_y - left_x);
          const float db = (o.y - o.x) - (o.x - left_y);
          mx = fmaf(da, da, fmaf(db, db, mx));
        }
      } else {
        met = fmaf(d0, d0, fmaf(d1, d1, met));
        const int k = c - S;
        if (k == f1_idx) f1 = o.x;
        if (k + 1 == f1_idx) f1 = o.y;
        if (k == f2_idx) f2 = o.x;
        if (k + 1 == f2_idx) f2 = o.y;
      }
    }
    if (p_norm != nullptr) {
      f1 = warp_sum_f(f1);   // one lane holds each value, the others zero
      f2 = warp_sum_f(f2);
      if (lane == 0) {
        const float4 pn = __ldg(reinterpret_cast<const float4*>(p_norm) + r);
        const float e1 = f1 - (0.4f * pn.x + 0.6f * pn.z);
        const float e2 = f2 - (0.3f * pn.y + 0.7f * pn.w);
        lc1 = fmaf(e1, e1, lc1);
        lc2 = fmaf(e2, e2, lc2);
        if (dp_lc) {
          const float m = -2.f * lc_grad_mult;
          *reinterpret_cast<float4*>(dp_lc + r * 4) = make_float4(m * e1 * 0.4f, m * e2 * 0.3f, m * e1 * 0.6f, m * e2 * 0.7f);
        }
      }
    }
  }
  const float v[5] = {rec, met, mx, lc1, lc2};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float t = block_sum(v[k], red);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(sums + k, (double)t);
  }
}

}  // namespace

// =========================================================================================== launchers
void launch_center_vec(const float* x, int64_t rows, int S, int rows_used, float* cvec, int Kp, cudaStream_t st) {
  if (rows_used > rows) rows_used = (int)rows;
  launch_k(center_vec_kernel, (Kp + 31) / 32, 1024, 0, st, x, S, rows_used, cvec, Kp);
}
void launch_cast_center(const float* x, const float* cvec, const float* params, __half* xc, int64_t rows, int S,
                        int P, int Kp, cudaStream_t st) {
  const int64_t total = rows * (Kp / 8);
  const int grid = (int)((total + kThreads - 1) / kThreads < 148 * 16 ? (total + kThreads - 1) / kThreads : 148 * 16);
  launch_k(cast_center_kernel<false>, grid > 0 ? grid : 1, kThreads, 0, st, x, nullptr, 0.f, cvec, params, xc, rows, S, P, Kp,
                                                                             (S % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0) ? 1 : 0,
                                                                             0ll);
}
void launch_cast_center_noise(const float* target, const float* noise, float sigma, const float* cvec, __half* xc,
                              float*, int64_t rows, int S, int P, int Kp, cudaStream_t st, int64_t target_stride) {
  const int64_t total = rows * (Kp / 8);
  const int grid = (int)((total + kThreads - 1) / kThreads < 148 * 16 ? (total + kThreads - 1) / kThreads : 148 * 16);
  launch_k(cast_center_kernel<true>, grid > 0 ? grid : 1, kThreads, 0, st, target, noise, sigma, cvec, nullptr, xc, rows, S, P, Kp,
                                                                            (S % 2 == 0 && (reinterpret_cast<uintptr_t>(target) & 7) == 0 &&
                                                                             (reinterpret_cast<uintptr_t>(noise) & 7) == 0) ? 1 : 0,
                                                                            (long long)target_stride);
}
void launch_cast_center_philox(const float* target, float sigma, uint64_t seed, int64_t first, const float* cvec,
                               __half* xc, float* noise_out, int64_t rows, int S, int Kp, cudaStream_t st) {
  const int64_t total = rows * (Kp / 8);
  const int grid = (int)((total + kThreads - 1) / kThreads < 148 * 16 ? (total + kThreads - 1) / kThreads : 148 * 16);
  launch_k(cast_center_philox_kernel, grid > 0 ? grid : 1, kThreads, 0, st, 
      target, sigma, (unsigned long long)seed, (long long)first, cvec, xc, noise_out, rows, S, Kp);
}
void launch_search_init(float* scores, int64_t* iota, int64_t* best_idx, float* params, int64_t total, int k,
                        cudaStream_t st) {
  launch_k(search_init_kernel, 148 * 4, 256, 0, st, scores, reinterpret_cast<long long*>(iota),
                                                       reinterpret_cast<long long*>(best_idx), params, total, k);
}
void launch_fill_inf(float* s, int64_t n, cudaStream_t st) {
  if (n <= 0) return;
  launch_k(fill_inf_kernel, 148 * 2, 256, 0, st, s, n);
}
void launch_search_gather(const int64_t* pos, const float* params, const int64_t* best_idx, int64_t group_base, int k,
                          int64_t* tmp_idx, float* tmp_params, cudaStream_t st) {
  launch_k(search_gather_kernel, (k + 255) / 256, 256, 0, st, 
      reinterpret_cast<const long long*>(pos), params, reinterpret_cast<const long long*>(best_idx), group_base, k,
      reinterpret_cast<long long*>(tmp_idx), tmp_params);
}
void launch_search_commit(const float* sel_scores, const int64_t* tmp_idx, const float* tmp_params, int k,
                          float* scores, int64_t* best_idx, float* params, cudaStream_t st) {
  launch_k(search_commit_kernel, (k + 255) / 256, 256, 0, st, sel_scores, reinterpret_cast<const long long*>(tmp_idx),
                                                                tmp_params, k, scores,
                                                                reinterpret_cast<long long*>(best_idx), params);
}
void launch_pack_net(const float* w1, int ld1, int S, int P, int wp_cols, int bias_cols, const float* b1,
                     const float* cvec, __half* w1h, int Kp, float* b_eff, int H1, const float* w2, int H2,
                     __half* w2h, __half* w2th, float* wp, cudaStream_t st) {
  PackNetArgs a;
  a.w1 = w1; a.ld1 = ld1; a.S = S; a.P = P; a.wp_cols = wp_cols; a.bias_cols = bias_cols; a.b1 = b1; a.cvec = cvec;
  a.w1h = w1h; a.Kp = Kp; a.b_eff = b_eff; a.H1 = H1; a.w2 = w2; a.H2 = H2; a.w2h = w2h; a.w2th = w2th; a.wp = wp;
  a.nA = (H1 + 7) / 8;
  a.nB = (H2 * H1 / 4 + 255) / 256;
  if (a.nB > 128) a.nB = 128;
  a.nC = w2th ? (H1 / 32) * (H2 / 32) : 0;
  const int nD = wp ? (H1 * 4 + 255) / 256 : 0;
  launch_k(pack_net_kernel, a.nA + a.nB + a.nC + nD, 256, 0, st, a);
}
void launch_head_consts(const float* scale2, const float* bias2, const float* w3, const float* b3, const float* fw1,
                        const float* fb1, const float* flnw, const float* flnb, float* out, cudaStream_t st) {
  launch_k(head_consts_kernel, 1, 256, 0, st, scale2, bias2, w3, b3, fw1, fb1, flnw, flnb, out);
}
void launch_pack_first_layer(const float* w, int ld_src, int S, int P, int wp_cols, int bias_cols, const float* b,
                             const float* cvec, __half* out, int Kp, float* b_eff_out, int rows, cudaStream_t st) {
  launch_k(pack_first_layer_kernel, (rows + 7) / 8, 256, 0, st, w, ld_src, S, P, wp_cols, bias_cols, b, cvec, out, Kp,
                                                          b_eff_out, rows);
}
void launch_cast_pad(const float* src, int ld_src, int ncols, __half* dst, int ld_dst, int rows, cudaStream_t st) {
  const long long total = (long long)rows * ld_dst;
  int grid = (int)((total + kThreads - 1) / kThreads);
  if (grid > 148 * 8) grid = 148 * 8;
  launch_k(cast_pad_kernel, grid, kThreads, 0, st, src, ld_src, ncols, dst, ld_dst, rows);
}
void launch_transpose_cast(const float* src, int rows, int cols, int ld_src, __half* dst, int ld_dst,
                           cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  launch_k(transpose_cast_kernel, grid, 256, 0, st, src, rows, cols, ld_src, dst, ld_dst);
}
void launch_extract_wp(const float* w, int ld_src, int S, int P, float* wp, int rows, int rows_pad, cudaStream_t st) {
  launch_k(extract_wp_kernel, (rows_pad * 4 + 255) / 256, 256, 0, st, w, ld_src, S, P, wp, rows, rows_pad);
}
void launch_copy_pad_f32(const float* src, int n, float* dst, int n_pad, cudaStream_t st) {
  launch_k(copy_pad_f32_kernel, (n_pad + 255) / 256, 256, 0, st, src, n, dst, n_pad);
}
void launch_reduce_columns(const ReduceArgs& a, cudaStream_t st) { launch_reduce_partials(a, st); }
void launch_sigmoid_bwd(float* p_inout, const float* grad_out, float scale, int64_t n, cudaStream_t st) {
  launch_k(sigmoid_bwd_kernel, (int)((n + 255) / 256), 256, 0, st, p_inout, grad_out, scale, (long long)n);
}
void launch_scale_copy(const float* src, float* dst, float mult, int64_t n, cudaStream_t st) {
  launch_k(scale_copy_kernel, (int)((n + 255) / 256), 256, 0, st, src, dst, mult, (long long)n);
}
void launch_colstats(const __half* h, int64_t rows, int C, float* sum, float* sumsq, float* part, cudaStream_t st) {
  const int rpb = kThreads / (C / 4);
  const int grid = grid_for_rows(rows, rpb * kU * 2, kPartBlocks);
  launch_k(colstats_kernel, grid, kThreads, 0, st, h, rows, C, part);
  ReduceArgs r{part, grid, 2 * C, 2, {{sum, C, 1.f}, {sumsq, C, 1.f}}};
  launch_reduce_partials(r, st);
}
void launch_bn_finalize(const BnFinalizeArgs& a, cudaStream_t st) {
  launch_k(bn_finalize_kernel, (a.C + 255) / 256, 256, 0, st, a);
}
void launch_bn_reduce_finalize(const BnFinalizeArgs& a, const float* part, int nblocks, cudaStream_t st) {
  launch_k(bn_reduce_finalize_kernel, (a.C + 31) / 32, 1024, 0, st, a, part, nblocks);
}
void launch_bn_eval_affine(const float* rm, const float* rv, const float* gamma, const float* beta,
                           const float* offset, float* scale, float* bias, int C, cudaStream_t st) {
  launch_k(bn_eval_affine_kernel, (C + 255) / 256, 256, 0, st, rm, rv, gamma, beta, offset, scale, bias, C);
}
void launch_bn_relu_apply(const __half* h, const float* scale, const float* bias, __half* a, int64_t rows, int C,
                          cudaStream_t st) {
  const int rpb = kThreads / (C / 8);
  launch_k(bn_relu_apply_kernel, grid_for_rows(rows, rpb * 4), kThreads, 0, st, h, scale, bias, a, rows, C);
}
void launch_g_head_fwd(const __half* h2, const float* scale, const float* bias, const float* w3, const float* b3,
                       float* p_out, float* pden_out, const __half* xc, __half* tail_fake, int64_t rows, int C,
                       int Kp, int S, cudaStream_t st) {
  if (C != 256) {   // the register layout of g_head_fwd_kernel is the reference width's
    launch_k(wide_head_fwd_kernel, grid_for_rows(rows, 8 * 4 * 2, 148 * 8), kThreads, 0, st, h2, scale, bias, w3, b3, p_out,
             pden_out, xc, tail_fake, (long long)rows, C, Kp, S);
    return;
  }
  launch_k(g_head_fwd_kernel, grid_for_rows(rows, 8 * kHeadRows * 2, 148 * 2), kThreads, 0, st, h2, scale, bias, w3, b3,
           p_out, pden_out, xc, tail_fake, (long long)rows, C, Kp, S);
}
void launch_d_logit_bce(const __half* z2, const float* w3, const float* b3, int64_t rows, int C, int64_t rows_a,
                        float label_a, float label_b, int64_t gap_begin, int64_t gap_end, double global_batch,
                        double* loss_sum, float* dlogit, float* prob_out, cudaStream_t st) {
  launch_k(d_logit_bce_kernel, grid_for_rows(rows, 8 * 4, 148 * 8), kThreads, 0, st, z2, w3, b3, (long long)rows, C,
           (long long)rows_a, label_a, label_b, (long long)gap_begin, (long long)gap_end,
           (float)(1.0 / global_batch), loss_sum, dlogit, prob_out);
}
void launch_f_pigan_loss_slab(const float* slab, int ngroups, const float* bias, const float* spectrum,
                              const float* metrics, const float* p_norm, int64_t rows, int S, int Mt, int f1_idx,
                              int f2_idx, float lc_grad_mult, double* sums, float* dp_lc, cudaStream_t st) {
  launch_k(f_pigan_loss_slab_kernel, grid_for_rows(rows, 8 * 4, 148 * 8), kThreads, 0, st,
           slab, ngroups, bias, spectrum, metrics, p_norm, (long long)rows, S, Mt, f1_idx, f2_idx, lc_grad_mult, sums,
           dp_lc);
}
// Column segment [off, off + w) of the head's arguments: the thread-owns-4-columns kernels cover at most 1024 columns
// per launch (256 threads a row), wider layers (the widened generator, C = 2048) run them once per segment with the
// row pitch ld = C.  One segment (C <= 1024, ld = C) is the reference-width path, unchanged.
static GHeadBwdArgs head_segment(const GHeadBwdArgs& a, int off, int w) {
  GHeadBwdArgs b = a;
  b.C = w;
  b.ld = a.C;
  b.h2 = a.h2 + off; b.dy2 = a.dy2 ? a.dy2 + off : nullptr;
  b.scale = a.scale + off; b.bias = a.bias + off; b.mean = a.mean + off; b.rstd = a.rstd + off;
  b.w3 = a.w3 + off; b.dw3 = a.dw3 + off; b.sum_dy = a.sum_dy + off; b.sum_dyx = a.sum_dyx + off;
  b.gamma = a.gamma ? a.gamma + off : nullptr; b.dbias = a.dbias ? a.dbias + off : nullptr;
  b.dgamma = a.dgamma ? a.dgamma + off : nullptr; b.dbeta = a.dbeta ? a.dbeta + off : nullptr;
  return b;
}
void launch_g_head_bwd(const GHeadBwdArgs& a, bool apply, cudaStream_t st) {
  const int seg = a.C > 1024 ? 1024 : a.C;
  const int rpb = kThreads / (seg / 4);
  // partial scratch: [grid][8 * seg] moment rows, then 8 * seg totals, then the dpre partials (<= 296 x 8)
  const size_t scratch = (size_t)kPartBlocks * kPartCols;
  int cap = (int)((scratch - 8 * (size_t)seg - 8 * 296) / (8 * (size_t)seg));
  if (cap > kPartBlocks) cap = kPartBlocks;
  const int grid = grid_for_rows(a.rows, rpb * (apply ? 4 : 6) * 2, cap);
  if (apply) {
    for (int off = 0; off < a.C; off += seg) {
      const GHeadBwdArgs b = head_segment(a, off, seg);
      launch_k(g_head_bwd_kernel<true>, grid, kThreads, 0, st, b);
      ReduceArgs r{a.part, grid, seg, 1, {{b.dbias, seg, a.inv_gs}}};
      launch_reduce_partials(r, st);
    }
  } else {
    float* tot = a.part + (size_t)cap * 8 * seg;   // after the partial rows
    float* dpre_part = tot + 8 * seg;              // [dpre blocks][8] after the moment totals
    const int dpre_blocks = grid_for_rows(a.rows, kThreads, 148 * 2);
    {
      GHeadBwdArgs b = a;
      b.ld = a.C;
      b.dpre_part = dpre_part;
      launch_k(g_head_dpre_kernel, dpre_blocks, kThreads, 0, st, b);
    }
    for (int off = 0; off < a.C; off += seg) {
      GHeadBwdArgs b = head_segment(a, off, seg);
      b.dpre_part = dpre_part;
      launch_k(g_head_bwd_kernel<false>, grid, kThreads, 0, st, b);
      // db3 and the range-loss sum (the dpre partials) are added by the first segment's launch only
      launch_k(g_head_moments_kernel, (seg + 7) / 8 + 1, 1024, 0, st, b, (const float*)a.part, grid,
               off == 0 ? dpre_blocks : 0);
    }
  }
}
void launch_bn_bwd_stats(const __half* da, const __half* h, const float* scale, const float* bias,
                         const float* mean, const float* rstd, float* sum_dy, float* sum_dyx, int64_t rows, int C,
                         float* part, cudaStream_t st) {
  const int seg = C > 1024 ? 1024 : C;   // column segments of <= 1024 (thread-owns-4-columns mapping), pitch C
  const int rpb = kThreads / (seg / 4);
  const int grid = grid_for_rows(rows, rpb * kU * 2, kPartBlocks);
  for (int off = 0; off < C; off += seg) {
    launch_k(bn_bwd_stats_kernel, grid, kThreads, 0, st, da + off, h + off, scale + off, bias + off, mean + off,
             rstd + off, part, rows, seg, C);
    ReduceArgs r{part, grid, 2 * seg, 2, {{sum_dy + off, seg, 1.f}, {sum_dyx + off, seg, 1.f}}};
    launch_reduce_partials(r, st);
  }
}
void launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t st) {
  const int seg = a.C > 1024 ? 1024 : a.C;   // see launch_bn_bwd_stats
  const int rpb = kThreads / (seg / 4);
  const int grid = grid_for_rows(a.rows, rpb * kU * 2, kPartBlocks);
  for (int off = 0; off < a.C; off += seg) {
    BnBwdArgs b = a;
    b.C = seg;
    b.ld = a.C;
    b.dy = a.dy + off; b.h = a.h + off; b.dh = a.dh + off;
    b.scale = a.scale + off; b.bias = a.bias + off; b.mean = a.mean + off; b.rstd = a.rstd + off;
    b.gamma = a.gamma + off; b.sum_dy = a.sum_dy + off; b.sum_dyx = a.sum_dyx + off;
    b.dbias = a.dbias ? a.dbias + off : nullptr;
    b.dgamma = a.dgamma ? a.dgamma + off : nullptr;
    b.dbeta = a.dbeta ? a.dbeta + off : nullptr;
    launch_k(bn_bwd_apply_kernel, grid, kThreads, 0, st, b);
    if (b.dbias) {
      ReduceArgs r{a.part, grid, seg, 1, {{b.dbias, seg, a.inv_gs}}};
      launch_reduce_partials(r, st);
    }
  }
}
void launch_d_l2_bwd(const __half* z2, const float* dlogit, const float* w3, __half* dh2, float* dw3, float* db2,
                     float* db3, int64_t rows, int C, float inv_gs, float* part, cudaStream_t st) {
  const int rpb = kThreads / (C / 8);
  int cap = (int)(((size_t)kPartBlocks * kPartCols) / (size_t)(2 * C + 8));   // partial rows inside the scratch
  if (cap > kPartBlocks) cap = kPartBlocks;
  const int grid = grid_for_rows(rows, rpb * kU * 2, cap);
  launch_k(d_l2_bwd_kernel, grid, kThreads, 0, st, z2, dlogit, w3, dh2, dw3, db2, db3, rows, C, inv_gs, part);
  if (dw3 != nullptr) {
    ReduceArgs r{part, grid, 2 * C + 8, 3, {{dw3, C, inv_gs}, {db2, C, inv_gs}, {db3, 1, inv_gs}}};
    launch_reduce_partials(r, st);
  }
}
void launch_f_l1(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb, __half* out,
                 int64_t rows, int C, cudaStream_t st) {
  launch_k(f_l1_kernel, grid_for_rows(rows, 8 * 4 * 2, 148 * 4), kThreads, 0, st, p, w1, b1, lnw, lnb, out, rows, C);
}
void launch_ln_lrelu_apply(__half* h, const float* rowstats, int n_tiles, const float* gamma, const float* beta,
                           int64_t rows, int N, cudaStream_t st) {
  const int rpb = kThreads / (N / 8);
  launch_k(ln_lrelu_apply_kernel, grid_for_rows(rows, rpb * 4 * 2), kThreads, 0, st, h, rowstats, n_tiles, gamma, beta, rows, N);
}
void launch_zero_buffers(float* const* ptrs, const int64_t* nfloat, int count, cudaStream_t st) {
  ZeroArgs a;
  long long most = 0;
  for (int k = 0; k < 4; ++k) {
    a.ptr[k] = k < count ? ptrs[k] : nullptr;
    a.nfloat[k] = k < count ? (long long)nfloat[k] : 0;
    most = a.nfloat[k] > most ? a.nfloat[k] : most;
  }
  int grid = (int)((most / 4 + kThreads - 1) / kThreads);
  if (grid > 148 * 4) grid = 148 * 4;
  if (grid < 1) grid = 1;
  launch_k(zero_buffers_kernel, grid, kThreads, 0, st, a);
}
int launch_sumsq(const float* g, int64_t n, double* parts, cudaStream_t st) {
  int grid = (int)((n + kThreads * 4 - 1) / (kThreads * 4));
  if (grid > 148 * 2) grid = 148 * 2;
  if (grid < 1) grid = 1;
  launch_k(sumsq_kernel, grid, kThreads, 0, st, g, (long long)n, parts);
  return grid;
}
void launch_clip_adam(const AdamArgs& a, cudaStream_t st) {
  int grid = (int)((a.n + kThreads - 1) / kThreads);
  if (grid > 148 * 8) grid = 148 * 8;
  launch_k(clip_adam_kernel, grid, kThreads, 0, st, a);
}
void launch_dw_reduce(const float* part, int tiles_m, int tiles_n, int splits, float* dw, int ld, int m_valid,
                      int n_valid, float scale, int bias_col, float* db, cudaStream_t st, const float* fix_cvec,
                      int fix_S, int fix_P) {
  if (fix_cvec != nullptr)
    launch_k(dw_reduce_kernel<true>, dim3(tiles_n, m_valid), 256, 0, st, part, tiles_m, tiles_n, splits, dw, ld, m_valid,
             n_valid, scale, bias_col, db, fix_cvec, fix_S, fix_P);
  else
    launch_k(dw_reduce_kernel<false>, dim3(tiles_n, m_valid), 256, 0, st, part, tiles_m, tiles_n, splits, dw, ld,
             m_valid, n_valid, scale, bias_col, db, fix_cvec, fix_S, fix_P);
}
void launch_dw_fixup(float* dw, int ld, int S, int P, const float* db, const float* cvec, int rows,
                     cudaStream_t st) {
  const int total = rows * (S + P);
  launch_k(dw_fixup_kernel, (total + 255) / 256, 256, 0, st, dw, ld, S, P, db, cvec, rows);
}
void launch_loss_finalize(const LossFinalizeArgs& a, cudaStream_t st) { launch_k(loss_finalize_kernel, 1, 32, 0, st, a); }
void launch_f_upstream_cast(const float* g, int cols, __half* dout, int ld, int64_t rows, float scale, cudaStream_t st) {
  launch_k(f_upstream_cast_kernel, 148 * 8, 256, 0, st, g, cols, dout, ld, (long long)rows, scale);
}
void launch_validation_scores(const float* p, const float* p_noisy, int64_t rows, int P, float* stability,
                              float* plausibility, cudaStream_t st) {
  launch_k(validation_scores_kernel, (int)((rows + 255) / 256), 256, 0, st, p, p_noisy, (long long)rows, P, stability,
           plausibility);
}
void launch_score_finish(const float* p, const float* err, int64_t rows, int P, int32_t* violations,
                         float* consistency, cudaStream_t st) {
  launch_k(score_finish_kernel, (int)((rows + 255) / 256), 256, 0, st, p, err, rows, P, violations, consistency);
}


void launch_f_l1_train(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb,
                       __half* xhat, __half* act, float* rstd, unsigned char* mask, unsigned char* keepbits,
                       int64_t rows, const DropoutArgs& dr, cudaStream_t st) {
  launch_k(f_l1_train_kernel, grid_for_rows(rows, 8 * 4 * 2, 148 * 4), kThreads, 0, st, p, w1, b1, lnw, lnb, xhat, act, rstd,
           mask, keepbits, (long long)rows, dr);
}
void launch_ln_train(__half* xhat, const float* rowstats, const float* gamma, const float* beta, __half* act,
                     float* rstd, unsigned char* mask, unsigned char* keepbits, int64_t rows, int N, int layer,
                     const DropoutArgs& dr, cudaStream_t st) {
  const int grid = grid_for_rows(rows, 8 * 4, 148 * 4);
  if (N == 256) launch_k(ln_train_kernel<1>, grid, kThreads, 0, st, xhat, rowstats, gamma, beta, act, rstd, mask, keepbits, (long long)rows, layer, dr);
  else if (N == 512) launch_k(ln_train_kernel<2>, grid, kThreads, 0, st, xhat, rowstats, gamma, beta, act, rstd, mask, keepbits, (long long)rows, layer, dr);
  else if (N == 1024) launch_k(ln_train_kernel<4>, grid, kThreads, 0, st, xhat, rowstats, gamma, beta, act, rstd, mask, keepbits, (long long)rows, layer, dr);
  else launch_k(ln_train_kernel<8>, grid, kThreads, 0, st, xhat, rowstats, gamma, beta, act, rstd, mask, keepbits, (long long)rows, layer, dr);
}
void launch_f_out_loss(const float* out, const float* spectrum, const float* metrics, __half* dout, int ld,
                       int64_t rows, int S, int Mt, float* part, float* db_out, float* loss_sums, float inv_gs,
                       cudaStream_t st, float w_spec, float w_met) {
  const int OUT = S + Mt, ld_part = (OUT + 2 + 7) / 8 * 8;
  const int grid = grid_for_rows(rows, 8 * 2 * 4, kPartBlocks);
  launch_k(f_out_loss_kernel, grid, kThreads, 0, st, out, spectrum, metrics, dout, ld, (long long)rows, S, Mt, part,
           ld_part, w_spec, w_met);
  ReduceArgs r;
  r.part = part; r.nblocks = grid; r.ld = ld_part; r.nseg = 3;
  r.seg[0] = {db_out, OUT, inv_gs};
  r.seg[1] = {nullptr, ld_part - 2 - OUT, 0.f};
  r.seg[2] = {loss_sums, 2, 1.f};
  launch_reduce_partials(r, st);
}
void launch_ln_bwd(__half* da, const __half* xhat, const float* rstd, const float* gamma, const float* beta,
                   const float* p_in, const unsigned char* keepbits, int64_t rows, int N, float keep_scale,
                   float* part, float* dgamma, float* dbeta, float* dbias, float* dw1_kmajor, float inv_gs,
                   cudaStream_t st, int store_dh) {
  // partial rows are NQ * N floats wide: keep nblocks * NQ * N inside the scratch (kPartBlocks * kPartCols floats)
  const int nq = p_in ? 7 : 3;
  int cap = (int)(((size_t)kPartBlocks * kPartCols) / ((size_t)nq * N));
  if (cap > 148 * 2) cap = 148 * 2;
  const int grid = grid_for_rows(rows, (8 / (N / 256)) * (p_in ? 2 : 4) * 4, cap);
  if (p_in) launch_k(ln_bwd_kernel<1, true>, grid, kThreads, 0, st, da, xhat, rstd, gamma, beta, p_in, keepbits, (long long)rows, keep_scale, part, store_dh);
  else if (N == 256) launch_k(ln_bwd_kernel<1, false>, grid, kThreads, 0, st, da, xhat, rstd, gamma, beta, p_in, keepbits, (long long)rows, keep_scale, part, store_dh);
  else if (N == 512) launch_k(ln_bwd_kernel<2, false>, grid, kThreads, 0, st, da, xhat, rstd, gamma, beta, p_in, keepbits, (long long)rows, keep_scale, part, store_dh);
  else if (N == 1024) launch_k(ln_bwd_kernel<4, false>, grid, kThreads, 0, st, da, xhat, rstd, gamma, beta, p_in, keepbits, (long long)rows, keep_scale, part, store_dh);
  else launch_k(ln_bwd_kernel<8, false>, grid, kThreads, 0, st, da, xhat, rstd, gamma, beta, p_in, keepbits, (long long)rows, keep_scale, part, store_dh);
  ReduceArgs r;
  r.part = part; r.nblocks = grid; r.ld = nq * N; r.nseg = p_in ? 4 : 3;
  r.seg[0] = {dbeta, N, inv_gs};
  r.seg[1] = {dgamma, N, inv_gs};
  r.seg[2] = {dbias, N, inv_gs};
  if (p_in) r.seg[3] = {dw1_kmajor, 4 * N, inv_gs};
  launch_reduce_partials(r, st);
}
void launch_f_dp(const __half* dh1, const float* w1, float* dp, int64_t rows, float scale, cudaStream_t st) {
  launch_k(f_dp_kernel, grid_for_rows(rows, 8 * 4, 148 * 4), kThreads, 0, st, dh1, w1, dp, (long long)rows, scale);
}
void launch_f_dw1_transpose(const float* src, float* dw1, cudaStream_t st) {
  launch_k(f_dw1_transpose_kernel, 1, 256, 0, st, src, dw1);
}
void launch_f_l1_wide(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb,
                      float* consts, __half* xhat, __half* act, float* rstd, unsigned char* mask,
                      unsigned char* keepbits, int64_t rows, int N, const DropoutArgs* dr, cudaStream_t st) {
  launch_k(f_l1_consts_kernel, 1, 256, 0, st, w1, b1, N, consts);
  const int rpb = kThreads / (N / 8);
  const int grid = grid_for_rows(rows, rpb * 16, 148 * 8);
  if (dr) {
    launch_k(f_l1_wide_kernel<true>, grid, kThreads, 0, st, p, w1, b1, lnw, lnb, (const float*)consts, xhat, act, rstd,
             mask, keepbits, (long long)rows, N, *dr);
  } else {
    DropoutArgs none = {};
    launch_k(f_l1_wide_kernel<false>, grid, kThreads, 0, st, p, w1, b1, lnw, lnb, (const float*)consts,
             (__half*)nullptr, act, (float*)nullptr, (unsigned char*)nullptr, (unsigned char*)nullptr,
             (long long)rows, N, none);
  }
}
void launch_f_dw1_wide(const __half* dh1, const float* p, int64_t rows, int N, float* part, float* dw1_kmajor,
                       float* dw1, float inv_gs, cudaStream_t st) {
  int cap = (int)(((size_t)kPartBlocks * kPartCols) / ((size_t)4 * N));
  if (cap > 148 * 2) cap = 148 * 2;
  const int grid = grid_for_rows(rows, (kThreads / (N / 8)) * 8, cap);
  launch_k(f_dw1_wide_kernel, grid, kThreads, 0, st, dh1, p, (long long)rows, N, part);
  ReduceArgs r;
  r.part = part; r.nblocks = grid; r.ld = 4 * N; r.nseg = 1;
  r.seg[0] = {dw1_kmajor, 4 * N, inv_gs};
  launch_reduce_partials(r, st);
  launch_k(f_dw1_transpose_wide_kernel, (N + 255) / 256, 256, 0, st, (const float*)dw1_kmajor, dw1, N);
}
void launch_f_dp_wide(const __half* dh1, const float* w1, float* dp, int64_t rows, int N, float scale,
                      cudaStream_t st) {
  launch_k(f_dp_wide_kernel, grid_for_rows(rows, 8 * 4, 148 * 4), kThreads, 0, st, dh1, w1, dp, (long long)rows, N / 256,
           scale);
}
void launch_f_unslab(const float* slab, int ngroups, const float* bias, float* out, int64_t rows, int OUT,
                     cudaStream_t st) {
  launch_k(f_unslab_kernel, grid_for_rows(rows, 4, 148 * 8), kThreads, 0, st, slab, ngroups, bias, out, (long long)rows,
           OUT);
}
void launch_f_out_loss_slab(const float* slab, int ngroups, const float* bias, const float* spectrum,
                            const float* metrics, __half* dout, int ld, int64_t rows, int S, int Mt, float* part,
                            float* db_out, float* loss_sums, float inv_gs, cudaStream_t st, float w_spec, float w_met) {
  const int OUT = S + Mt, ld_part = (OUT + 2 + 7) / 8 * 8;
  const int grid = grid_for_rows(rows, 16, kPartBlocks);
  launch_k(f_out_loss_slab_kernel, grid, kThreads, 0, st, slab, ngroups, bias, spectrum, metrics, dout, ld,
           (long long)rows, S, Mt, part, ld_part, w_spec, w_met);
  ReduceArgs r;
  r.part = part; r.nblocks = grid; r.ld = ld_part; r.nseg = 3;
  r.seg[0] = {db_out, OUT, inv_gs};
  r.seg[1] = {nullptr, ld_part - 2 - OUT, 0.f};
  r.seg[2] = {loss_sums, 2, 1.f};
  launch_reduce_partials(r, st);
}
void launch_f_input_grad_losses(const float* sums, double n_spec, double n_met, float* out, cudaStream_t st) {
  launch_k(f_input_grad_losses_kernel, 1, 32, 0, st, sums, n_spec, n_met, out);
}
void launch_f_train_losses(const float* sums, double n_spec, double n_met, float* out, cudaStream_t st) {
  launch_k(f_train_losses_kernel, 1, 32, 0, st, sums, n_spec, n_met, out);
}

}  // namespace pigan
