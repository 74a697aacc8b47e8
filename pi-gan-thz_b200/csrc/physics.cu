// K7 — batched physics metrics: resonance peak, FWHM-based Q, FoM and sensitivity S per spectrum.
// Follows calculate_peak_parameters (reference core/utils/data_loader.py:13-58) branch for branch and the
// S = (f/1.0)*(Q/100.0)*100 its callers add (data_loader.py:96,105).
//
// One warp per spectrum.  The row is read once with coalesced 128-byte warp loads into registers
// (element i lives in lane i%32, register i/32), the argmin (first occurrence, NumPy semantics) is a
// shuffle reduction, both half-depth crossings are found with one neighbour exchange + shuffle
// max/min, and the few scalar interpolation steps run in fp64 exactly as the NumPy reference does, so
// the only difference from the oracle on fp32 inputs is the final rounding of the outputs to fp32.
// HBM-bound: s*4 bytes in, 20 bytes out per spectrum.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "host_util.h"

namespace pigan {

__device__ __forceinline__ bool less_np(float a, float b) {
  // ordering np.argmin uses: NaN sorts before everything, otherwise plain <
  return (isnan(a) && !isnan(b)) || (a < b);
}

// Everything of one row that needs the whole warp: peak index, half-depth crossing pairs.  Element i of the row
// lives in lane i % 32, register i / 32.
//   fast path  (no NaN in the row, half-depth level exactly representable in fp32 — always true for the reference's
//              baseline 0): plain fp32 compares, the >= h / > h predicates of a lane's registers packed into bit
//              masks, neighbours fetched with two shuffles, first/last set bits with clz/ffs, warp REDUX min/max
//   exact path (anything else): the reference's comparisons one by one in fp64, NumPy's NaN-first argmin
template <int VPL>
__device__ __forceinline__ void row_scan(const float (&v)[VPL], int s, int lane, const int* peak_idx, long long row,
                                         float baseline, int& idx_out, int& lo_out, int& up_out, float& tmin_out) {
  const float kInf = __int_as_float(0x7f800000);
  bool has_nan = false;
#pragma unroll
  for (int j = 0; j < VPL; ++j) has_nan |= (v[j] != v[j]);
  has_nan = __any_sync(0xffffffffu, has_nan);

  // ---- peak index: given, or argmin with first-occurrence tie-break
  int idx;
  if (peak_idx != nullptr) {
    idx = peak_idx[row];
  } else if (!has_nan) {
    float bv = v[0];
    int bj = 0;
#pragma unroll
    for (int j = 1; j < VPL; ++j)
      if (v[j] < bv) { bv = v[j]; bj = j; }   // padding beyond s is +inf: never taken
    float gmin = bv;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, off));
    idx = __reduce_min_sync(0xffffffffu, bv == gmin ? bj * 32 + lane : 0x7fffffff);
  } else {
    float bv = v[0];
    int bi = lane;
#pragma unroll
    for (int j = 1; j < VPL; ++j) {
      const int i = j * 32 + lane;
      if (i < s && less_np(v[j], bv)) { bv = v[j]; bi = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      const bool take = less_np(ov, bv) || (!less_np(bv, ov) && oi < bi);
      if (take) { bv = ov; bi = oi; }
    }
    idx = bi;
  }
  const bool idx_ok = (idx >= 0 && idx < s);
  const int idx_c = idx_ok ? idx : 0;

  // t_min = t[idx] (broadcast from the owning lane/register without dynamic register indexing)
  float tmin_local = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
    if (j == (idx_c >> 5)) tmin_local = v[j];
  const float t_min = __shfl_sync(0xffffffffu, tmin_local, idx_c & 31);
  const double h = (double)t_min + ((double)baseline - (double)t_min) / 2.0;
  const float hf = (float)h;

  int lo = -1, up = 0x7fffffff;
  bool fast = false;
  if constexpr (VPL <= 16) fast = !has_nan && (double)hf == h;   // wider rows (s > 512) take the exact path
  if (fast) {
    // bit j of ge / gt: v[j] >= h / v[j] > h  (exact: h is an fp32 number)
    unsigned ge = 0, gt = 0;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      ge |= (v[j] >= hf ? 1u : 0u) << (j & 15);
      gt |= (v[j] > hf ? 1u : 0u) << (j & 15);
    }
    // predicates of element i+1: the next lane's same register, or lane 0's next register for lane 31
    const unsigned both = ge | (gt << 16);
    const unsigned nxt = __shfl_down_sync(0xffffffffu, both, 1);
    const unsigned first = __shfl_sync(0xffffffffu, both, 0);
    const unsigned nb = lane == 31 ? ((first >> 1) & 0x7fff7fffu) : nxt;
    // pairs (i, i+1) with i + 1 < s
    const int last_i = s - 2;                                   // largest valid pair start
    const int jmax_valid = last_i >= lane ? (last_i - lane) >> 5 : -1;
    const unsigned valid = jmax_valid >= 0 ? ((2u << jmax_valid) - 1u) : 0u;
    const unsigned cl = (ge ^ (nb & 0xffffu)) & valid;          // (a >= h) != (b >= h)
    const unsigned cu = (gt ^ (nb >> 16)) & valid;              // (a > h) != (b > h)
    // lower: largest i <= idx - 1 ; upper: smallest i >= idx + 1
    const int jl = (idx_c - 1 >= lane) ? (idx_c - 1 - lane) >> 5 : -1;
    const unsigned ml = jl >= 0 ? cl & ((2u << jl) - 1u) : 0u;
    if (ml) lo = (31 - __clz(ml)) * 32 + lane;
    const int ju = (idx_c + 1 - lane + 31) >> 5;                // smallest j with j*32+lane >= idx+1 (>= 0)
    const unsigned mu = ju < 32 ? cu & ~((1u << (ju < 0 ? 0 : ju)) - 1u) : 0u;
    if (mu) up = (__ffs(mu) - 1) * 32 + lane;
    lo = __reduce_max_sync(0xffffffffu, lo);
    up = __reduce_min_sync(0xffffffffu, up);
  } else {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * 32 + lane;
      float nxt = __shfl_down_sync(0xffffffffu, v[j], 1);
      float wrap = (j + 1 < VPL) ? v[j + 1] : kInf;
      wrap = __shfl_sync(0xffffffffu, wrap, 0);
      if (lane == 31) nxt = wrap;
      if (i + 1 < s) {
        const double a = (double)v[j], b = (double)nxt;
        const bool cl = (a >= h && b < h) || (a < h && b >= h);
        const bool cu = (a <= h && b > h) || (a > h && b <= h);
        if (cl && i <= idx_c - 1 && i > lo) lo = i;
        if (cu && i >= idx_c + 1 && i < up) up = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo = max(lo, __shfl_xor_sync(0xffffffffu, lo, off));
      up = min(up, __shfl_xor_sync(0xffffffffu, up, off));
    }
  }
  idx_out = idx;
  lo_out = lo;
  up_out = up;
  tmin_out = t_min;
}

// The scalar part of calculate_peak_parameters for one row (data_loader.py:27-58), run by the lane that holds the
// row's scan results: float64 interpolation of the two crossings, Q, FoM, S (and, BWD, the five gradient entries).
template <bool BWD>
__device__ __forceinline__ void row_finish(const float* __restrict__ t, long long row, int s,
                                           const double* __restrict__ freq, float baseline, int idx, int lo, int up,
                                           float t_min, int* __restrict__ out_idx, float* __restrict__ out_metrics,
                                           const float* __restrict__ grad_metrics, float* __restrict__ grad_spectra) {
  // t: the row's samples (global or shared memory)
      const double h = (double)t_min + ((double)baseline - (double)t_min) / 2.0;
      const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
      double f_res = kNaN, Q = kNaN, FoM = kNaN, S = kNaN;
      if (idx >= 0 && idx < s) {
        f_res = freq[idx];
        double f_lower = kNaN, f_upper = kNaN;
        if (lo >= 0) {
          const double ti = (double)t[lo], tj = (double)t[lo + 1];
          const double fi = freq[lo], fj = freq[lo + 1];
          f_lower = ((tj - ti) != 0.0) ? fi + (h - ti) * (fj - fi) / (tj - ti) : fi;
        }
        if (up != 0x7fffffff) {
          const double ti = (double)t[up], tj = (double)t[up + 1];
          const double fi = freq[up], fj = freq[up + 1];
          f_upper = ((tj - ti) != 0.0) ? fi + (h - ti) * (fj - fi) / (tj - ti) : fi;
        }
        if (!isnan(f_lower) && !isnan(f_upper) && f_upper > f_lower) {
          const double delta_f = f_upper - f_lower;
          if (delta_f > 1e-9) Q = f_res / delta_f;
          const double tm = (double)t_min;
          if (!isnan(tm) && fabs(tm) > 1e-6) FoM = isnan(Q) ? kNaN : Q / fabs(tm);
        }
        if (!isnan(Q)) S = (f_res / 1.0) * (Q / 100.0) * 100.0;
        if constexpr (BWD) {
          if (!isnan(Q)) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(grad_metrics) + row);   // d/d(f_res, Q, FoM, S)
            const double tm = (double)t_min, delta_f = f_upper - f_lower;
            const bool fom_ok = !isnan(FoM);
            // Q = f_res / (f_upper - f_lower), FoM = Q / |t_min|, S = f_res * Q; f_res = frequency[idx] is constant
            const double gq = (double)gm.y + (fom_ok ? (double)gm.z / fabs(tm) : 0.0) + (double)gm.w * f_res;
            const double g_fup = -gq * f_res / (delta_f * delta_f), g_flo = -g_fup;
            double g_h = 0.0;
            float* gr = grad_spectra + row * (long long)s;
            {
              const double ti = (double)t[lo], tj = (double)t[lo + 1], dt = tj - ti;
              if (dt != 0.0) {   // f_lower = f_i + (h - t_i) w / (t_j - t_i)
                const double w = freq[lo + 1] - freq[lo];
                g_h += g_flo * w / dt;
                gr[lo] += (float)(g_flo * w * (h - tj) / (dt * dt));
                gr[lo + 1] += (float)(-g_flo * w * (h - ti) / (dt * dt));
              }
            }
            {
              const double ti = (double)t[up], tj = (double)t[up + 1], dt = tj - ti;
              if (dt != 0.0) {
                const double w = freq[up + 1] - freq[up];
                g_h += g_fup * w / dt;
                gr[up] += (float)(g_fup * w * (h - tj) / (dt * dt));
                gr[up + 1] += (float)(-g_fup * w * (h - ti) / (dt * dt));
              }
            }
            // h = t_min + (baseline - t_min) / 2; FoM's own dependence on |t_min|
            double g_tmin = 0.5 * g_h;
            if (fom_ok) g_tmin += (double)gm.z * (-Q * (tm > 0.0 ? 1.0 : -1.0) / (tm * tm));
            gr[idx] += (float)g_tmin;
          }
        }
      }
      if (out_idx != nullptr) out_idx[row] = idx;   // 32 consecutive rows per warp: coalesced
      if (!BWD || out_metrics != nullptr) {
        float4 o = make_float4((float)f_res, (float)Q, (float)FoM, (float)S);
        *reinterpret_cast<float4*>(out_metrics + row * 4) = o;
      }
}

// A warp takes 32 consecutive rows: the warp-wide scan row by row (results parked in lane r for row r), then the
// scalar fp64 interpolation of all 32 rows in parallel, one lane per row — the serial tail costs one pass per 32 rows.
//
// BWD: also the vector-Jacobian product of (f_res, Q, FoM, S) with respect to the spectrum samples:
// with the branch decisions (peak index, the two crossing pairs) held fixed, the outputs depend on at most five
// samples of a row — t[idx] through the half-depth level and |t_min|, and the two samples of each crossing pair
// through the linear interpolation.  The warp clears the row of grad_spectra while it scans it, the lane that owns the
// row adds the five entries after the interpolation.  Rows whose Q is undefined get a zero gradient.
template <int VPL, bool BWD>
__global__ void __launch_bounds__(256) physics_metrics_kernel(
    const float* __restrict__ spectra, long long n, int s, const double* __restrict__ freq,
    const int* __restrict__ peak_idx, float baseline, int* __restrict__ out_idx,
    float* __restrict__ out_metrics, const float* __restrict__ grad_metrics, float* __restrict__ grad_spectra) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float kInf = __int_as_float(0x7f800000);

  for (long long base = warp0 * 32; base < n; base += warps_total * 32) {
    int my_idx = 0, my_lo = -1, my_up = 0x7fffffff;
    float my_tmin = 0.f;
    const int rows_here = (int)(n - base < 32 ? n - base : 32);
    float vn[VPL];   // next row, in flight while the current one is scanned
    {
      const float* t = spectra + base * (long long)s;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = j * 32 + lane;
        vn[j] = (j * 32 + 32 <= s || i < s) ? __ldg(t + i) : kInf;
      }
    }
    for (int r = 0; r < rows_here; ++r) {
      float v[VPL];
#pragma unroll
      for (int j = 0; j < VPL; ++j) v[j] = vn[j];
      if (r + 1 < rows_here) {
        const float* t = spectra + (base + r + 1) * (long long)s;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int i = j * 32 + lane;
          vn[j] = (j * 32 + 32 <= s || i < s) ? __ldg(t + i) : kInf;
        }
      }
      if constexpr (BWD) {
        float* gr = grad_spectra + (base + r) * (long long)s;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int i = j * 32 + lane;
          if (i < s) gr[i] = 0.f;
        }
      }
      int idx, lo, up;
      float tmin;
      row_scan<VPL>(v, s, lane, peak_idx, base + r, baseline, idx, lo, up, tmin);
      if (lane == r) { my_idx = idx; my_lo = lo; my_up = up; my_tmin = tmin; }
    }
    if constexpr (BWD) __syncwarp();   // the rows are cleared before their owners add to them
    if (lane < rows_here)
      row_finish<BWD>(spectra + (base + lane) * (long long)s, base + lane, s, freq, baseline, my_idx, my_lo, my_up,
                      my_tmin, out_idx, out_metrics, grad_metrics, grad_spectra);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Lane-per-row variant for rows of up to 256 samples (the dataset's 250).  The 32 rows of a warp's chunk are contiguous
// in memory (32 s floats), so lane 0 fetches the whole chunk with ONE bulk asynchronous copy (cp.async.bulk, the 1-D
// TMA path) into the warp's private two-stage ring in shared memory and signals an mbarrier; then every LANE scans its
// own row out of shared memory: a running (min, first index) over the row, and from the peak outwards the first
// half-depth crossing on each side.  No shuffles, no reductions, no bit masks: ~55 warp instructions per row instead of
// the ~350 of the warp-per-row kernel above, which profiles at 47 % of the HBM peak because it is ISSUE-bound, not
// latency-bound (staging its rows with bulk copies, 192 KB in flight per SM, made it 10 % slower: tools/phys_bench.py).
// Rows with a NaN, or a half-depth level that is not an fp32 number, take the reference's comparisons one by one in
// float64 (exact path, same lane).
constexpr int kRowsWarps = 6;     // 6 warps x 1 stage x 32 rows x 1000 B = 192 KB of shared memory per SM: a warp loads,
constexpr int kRowsStages = 1;    // then scans; the other warps' loads and scans overlap it

__device__ __forceinline__ uint32_t phys_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void phys_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void phys_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}

// one row, one thread: peak index (given, or NumPy argmin) and the two crossing pairs (data_loader.py:27-47)
__device__ __forceinline__ void row_scan_serial(const float* __restrict__ t, int s, const int* peak_idx, long long row,
                                                float baseline, int& idx_out, int& lo_out, int& up_out, float& tmin_out) {
  const float kInf = __int_as_float(0x7f800000);
  // running (min, first index) in four independent chains (element p -> chain p & 3; strict < keeps the first
  // occurrence inside a chain), 16 elements per trip with all loads issued up front: one lane walks its row alone,
  // so the shared-memory latency has to be covered by instruction-level parallelism
  float bv[4] = {kInf, kInf, kInf, kInf};
  int bi[4] = {0, 1, 2, 3};
  bool has_nan = false;
  int i = 0;
  if ((s & 1) == 0) {                   // rows start 8-byte aligned for even s: 64-bit loads
    const float2* t2 = reinterpret_cast<const float2*>(t);
    for (; i + 16 <= s; i += 16) {
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = t2[(i >> 1) + k];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        has_nan |= (v[k].x != v[k].x) | (v[k].y != v[k].y);
        const int c = (2 * k) & 3;
        if (v[k].x < bv[c]) { bv[c] = v[k].x; bi[c] = i + 2 * k; }
        if (v[k].y < bv[c + 1]) { bv[c + 1] = v[k].y; bi[c + 1] = i + 2 * k + 1; }
      }
    }
  }
  for (; i < s; ++i) {
    const float v0 = t[i];
    has_nan |= (v0 != v0);
    const int c = i & 3;
    // (dynamic chain index only in this short tail)
    if (c == 0) { if (v0 < bv[0]) { bv[0] = v0; bi[0] = i; } }
    else if (c == 1) { if (v0 < bv[1]) { bv[1] = v0; bi[1] = i; } }
    else if (c == 2) { if (v0 < bv[2]) { bv[2] = v0; bi[2] = i; } }
    else { if (v0 < bv[3]) { bv[3] = v0; bi[3] = i; } }
  }
  float b0 = bv[0];
  int idx = bi[0];
#pragma unroll
  for (int c = 1; c < 4; ++c)
    if (bv[c] < b0 || (bv[c] == b0 && bi[c] < idx)) { b0 = bv[c]; idx = bi[c]; }
  if (has_nan) {                        // np.argmin: the first NaN
    for (int k = 0; k < s; ++k)
      if (t[k] != t[k]) { idx = k; break; }
  }
  if (peak_idx != nullptr) idx = peak_idx[row];
  const bool idx_ok = idx >= 0 && idx < s;
  const int ic = idx_ok ? idx : 0;
  const float t_min = t[ic];
  const double h = (double)t_min + ((double)baseline - (double)t_min) / 2.0;
  const float hf = (float)h;
  int lo = -1, up = 0x7fffffff;
  if (!has_nan && (double)hf == h) {
    for (int k = ic - 1; k >= 0; --k)
      if ((t[k] >= hf) != (t[k + 1] >= hf)) { lo = k; break; }
    for (int k = ic + 1; k + 1 < s; ++k)
      if ((t[k] > hf) != (t[k + 1] > hf)) { up = k; break; }
  } else {
    for (int k = ic - 1; k >= 0; --k) {
      const double a = (double)t[k], b = (double)t[k + 1];
      if ((a >= h && b < h) || (a < h && b >= h)) { lo = k; break; }
    }
    for (int k = ic + 1; k + 1 < s; ++k) {
      const double a = (double)t[k], b = (double)t[k + 1];
      if ((a <= h && b > h) || (a > h && b <= h)) { up = k; break; }
    }
  }
  idx_out = idx;
  lo_out = lo;
  up_out = up;
  tmin_out = t_min;
}

template <bool BWD>
__global__ void __launch_bounds__(kRowsWarps * 32, 1) physics_metrics_rows_kernel(
    const float* __restrict__ spectra, long long n, int s, const double* __restrict__ freq,
    const int* __restrict__ peak_idx, float baseline, int* __restrict__ out_idx,
    float* __restrict__ out_metrics, const float* __restrict__ grad_metrics, float* __restrict__ grad_spectra) {
  extern __shared__ __align__(128) unsigned char phys_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t stage_pitch = ((uint32_t)(32 * s) * 4u + 127u) & ~127u;
  unsigned char* my = phys_smem + (size_t)warp * kRowsStages * stage_pitch;
  unsigned long long* bars =
      reinterpret_cast<unsigned long long*>(phys_smem + (size_t)kRowsWarps * kRowsStages * stage_pitch) + warp * kRowsStages;
  if (lane == 0) {
    for (int i = 0; i < kRowsStages; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(phys_smem_u32(bars + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long warps_total = (long long)gridDim.x * kRowsWarps;
  const long long warp0 = (long long)blockIdx.x * kRowsWarps + warp;
  const long long groups = (n + 31) / 32;      // n is a multiple of 4 here: every chunk is a multiple of 16 bytes
  auto issue = [&](long long k) {              // this warp's k-th chunk
    const long long g = warp0 + k * warps_total;
    if (g >= groups) return;
    const long long r0 = g * 32;
    const int rows = (int)(n - r0 < 32 ? n - r0 : 32);
    const int slot = (int)(k % kRowsStages);
    phys_bulk_load(phys_smem_u32(my + (size_t)slot * stage_pitch), spectra + r0 * (long long)s,
                   (uint32_t)(rows * s) * 4u, phys_smem_u32(bars + slot));
  };
  if (lane == 0)
    for (int k = 0; k < kRowsStages; ++k) issue(k);
  for (long long k = 0;; ++k) {
    const long long g = warp0 + k * warps_total;
    if (g >= groups) break;
    const long long r0 = g * 32;
    const int rows = (int)(n - r0 < 32 ? n - r0 : 32);
    const int slot = (int)(k % kRowsStages);
    if constexpr (BWD) {
      // clear the chunk's gradient rows (contiguous, 16-byte aligned: the host checks the base)
      float4* gz = reinterpret_cast<float4*>(grad_spectra + r0 * (long long)s);
      const int n4 = rows * s / 4;
      for (int i = lane; i < n4; i += 32) gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    phys_bar_wait(phys_smem_u32(bars + slot), (uint32_t)((k / kRowsStages) & 1));
    if constexpr (BWD) __syncwarp();     // the rows are cleared before their owners add to them
    if (lane < rows) {
      const float* t = reinterpret_cast<const float*>(my + (size_t)slot * stage_pitch) + lane * s;
      int idx, lo, up;
      float tmin;
      row_scan_serial(t, s, peak_idx, r0 + lane, baseline, idx, lo, up, tmin);
      row_finish<BWD>(t, r0 + lane, s, freq, baseline, idx, lo, up, tmin, out_idx, out_metrics, grad_metrics,
                      grad_spectra);
    }
    __syncwarp();                        // every lane has read the stage: it may be refilled
    if (lane == 0) issue(k + kRowsStages);
  }
}

}  // namespace pigan

using namespace pigan;

namespace {
bool phys_bulk_ok(const float* spectra, int s) {
  static const bool off = [] { const char* v = getenv("PIGAN_PHYS_BULK"); return v && v[0] == '0'; }();
  return !off && s >= 2 && s <= 256 && (reinterpret_cast<uintptr_t>(spectra) & 15u) == 0;
}
template <bool BWD>
int launch_phys_bulk(const float* spectra, int64_t n, int s, const double* freq, const int* peak_idx, float baseline,
                     int* out_idx, float* out_metrics, const float* grad_metrics, float* grad_spectra, cudaStream_t st) {
  const uint32_t stage_pitch = ((uint32_t)(32 * s) * 4u + 127u) & ~127u;
  const size_t smem = (size_t)kRowsWarps * kRowsStages * stage_pitch + (size_t)kRowsWarps * kRowsStages * 8;
  auto kern = physics_metrics_rows_kernel<BWD>;
  static bool attr_done = false;
  if (!attr_done) {
    PIGAN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const int64_t groups = ceil_div64(n, 32);
  const int64_t blocks_needed = ceil_div64(groups, kRowsWarps);
  const int grid = (int)(blocks_needed < sm_count() ? blocks_needed : sm_count());
  note_launch();
  kern<<<grid, kRowsWarps * 32, smem, st>>>(spectra, (long long)n, s, freq, peak_idx, baseline, out_idx, out_metrics,
                                            grad_metrics, grad_spectra);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
}  // namespace

extern "C" int pigan_physics_metrics(const float* spectra, int64_t n, int32_t s, const double* frequency,
                                     const int32_t* peak_idx, float baseline_transmission,
                                     int32_t* out_idx, float* out_metrics, void* stream) {
  PIGAN_CHECK_ARG(n >= 0);
  PIGAN_CHECK_ARG(s >= 1 && s <= 2048);
  if (n == 0) return PIGAN_OK;
  PIGAN_CHECK_ARG(spectra != nullptr && frequency != nullptr && out_metrics != nullptr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps_per_block = 8;
  const int64_t blocks_needed = ceil_div64(ceil_div64(n, 32), warps_per_block);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device");
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  const int vpl = (s + 31) / 32;
  if (phys_bulk_ok(spectra, s)) {
    // rows of up to 256 samples: the lane-per-row kernel on the leading multiple of 4 rows (its bulk copies move
    // multiples of 16 bytes), the warp-per-row kernel on the last 1-3
    const int64_t nb = n - n % 4;
    if (nb > 0) PIGAN_TRY((launch_phys_bulk<false>(spectra, nb, s, frequency, peak_idx, baseline_transmission, out_idx,
                                                   out_metrics, nullptr, nullptr, st)));
    if (nb == n) return PIGAN_OK;
    spectra += nb * s;
    if (peak_idx) peak_idx += nb;
    if (out_idx) out_idx += nb;
    out_metrics += nb * 4;
    n -= nb;
  }
#define PIGAN_LAUNCH_PHYS(V)                                                                        \
  note_launch(), physics_metrics_kernel<V, false><<<grid, 256, 0, st>>>(spectra, (long long)n, s, frequency, peak_idx, \
                                                  baseline_transmission, out_idx, out_metrics, nullptr, nullptr)
  if (vpl <= 8) PIGAN_LAUNCH_PHYS(8);
  else if (vpl <= 16) PIGAN_LAUNCH_PHYS(16);
  else if (vpl <= 32) PIGAN_LAUNCH_PHYS(32);
  else PIGAN_LAUNCH_PHYS(64);
#undef PIGAN_LAUNCH_PHYS
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_physics_metrics_backward(const float* spectra, int64_t n, int32_t s, const double* frequency,
                                              const int32_t* peak_idx, float baseline_transmission,
                                              const float* grad_metrics, float* grad_spectra, int32_t* out_idx,
                                              float* out_metrics, void* stream) {
  PIGAN_CHECK_ARG(n >= 0);
  PIGAN_CHECK_ARG(s >= 2 && s <= 2048);
  if (n == 0) return PIGAN_OK;
  PIGAN_CHECK_ARG(spectra != nullptr && frequency != nullptr && grad_metrics != nullptr && grad_spectra != nullptr);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(grad_metrics) & 15u) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device");
  const int64_t blocks_needed = ceil_div64(ceil_div64(n, 32), 8);
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  const int vpl = (s + 31) / 32;
  if (phys_bulk_ok(spectra, s) && (reinterpret_cast<uintptr_t>(grad_spectra) & 15u) == 0) {
    const int64_t nb = n - n % 4;
    if (nb > 0) PIGAN_TRY((launch_phys_bulk<true>(spectra, nb, s, frequency, peak_idx, baseline_transmission, out_idx,
                                                  out_metrics, grad_metrics, grad_spectra, st)));
    if (nb == n) return PIGAN_OK;
    spectra += nb * s;
    if (peak_idx) peak_idx += nb;
    if (out_idx) out_idx += nb;
    if (out_metrics) out_metrics += nb * 4;
    grad_metrics += nb * 4;
    grad_spectra += nb * s;
    n -= nb;
  }
#define PIGAN_LAUNCH_PHYS_BWD(V)                                                                                       \
  note_launch(), physics_metrics_kernel<V, true><<<grid, 256, 0, st>>>(spectra, (long long)n, s, frequency, peak_idx,  \
                                                  baseline_transmission, out_idx, out_metrics, grad_metrics,         \
                                                  grad_spectra)
  if (vpl <= 8) PIGAN_LAUNCH_PHYS_BWD(8);
  else if (vpl <= 16) PIGAN_LAUNCH_PHYS_BWD(16);
  else if (vpl <= 32) PIGAN_LAUNCH_PHYS_BWD(32);
  else PIGAN_LAUNCH_PHYS_BWD(64);
#undef PIGAN_LAUNCH_PHYS_BWD
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
