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
    if (lane < rows_here) {
      const long long row = base + lane;
      const float* t = spectra + row * (long long)s;
      const int idx = my_idx, lo = my_lo, up = my_up;
      const float t_min = my_tmin;
      const double h = (double)t_min + ((double)baseline - (double)t_min) / 2.0;
      const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
      double f_res = kNaN, Q = kNaN, FoM = kNaN, S = kNaN;
      if (idx >= 0 && idx < s) {
        f_res = freq[idx];
        double f_lower = kNaN, f_upper = kNaN;
        if (lo >= 0) {
          const double ti = (double)__ldg(t + lo), tj = (double)__ldg(t + lo + 1);
          const double fi = freq[lo], fj = freq[lo + 1];
          f_lower = ((tj - ti) != 0.0) ? fi + (h - ti) * (fj - fi) / (tj - ti) : fi;
        }
        if (up != 0x7fffffff) {
          const double ti = (double)__ldg(t + up), tj = (double)__ldg(t + up + 1);
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
              const double ti = (double)__ldg(t + lo), tj = (double)__ldg(t + lo + 1), dt = tj - ti;
              if (dt != 0.0) {   // f_lower = f_i + (h - t_i) w / (t_j - t_i)
                const double w = freq[lo + 1] - freq[lo];
                g_h += g_flo * w / dt;
                gr[lo] += (float)(g_flo * w * (h - tj) / (dt * dt));
                gr[lo + 1] += (float)(-g_flo * w * (h - ti) / (dt * dt));
              }
            }
            {
              const double ti = (double)__ldg(t + up), tj = (double)__ldg(t + up + 1), dt = tj - ti;
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
  }
}

}  // namespace pigan

using namespace pigan;

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
