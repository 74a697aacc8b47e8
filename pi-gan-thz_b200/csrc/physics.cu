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

template <int VPL>
__global__ void __launch_bounds__(256) physics_metrics_kernel(
    const float* __restrict__ spectra, long long n, int s, const double* __restrict__ freq,
    const int* __restrict__ peak_idx, float baseline, int* __restrict__ out_idx,
    float* __restrict__ out_metrics) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float kInf = __int_as_float(0x7f800000);

  for (; row < n; row += warps_total) {
    const float* t = spectra + row * (long long)s;
    float v[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * 32 + lane;
      v[j] = (i < s) ? __ldcs(t + i) : kInf;
    }

    // ---- peak index: given, or argmin with first-occurrence tie-break
    int idx;
    if (peak_idx != nullptr) {
      idx = peak_idx[row];
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

    // ---- crossings: flag pair (i, i+1); lower = largest i <= idx-1, upper = smallest i in [idx+1, s-2]
    int lo = -1, up = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = j * 32 + lane;
      // neighbour t[i+1]: next lane, or lane 0 of the next register row
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

    if (lane == 0) {
      const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
      double f_res = kNaN, Q = kNaN, FoM = kNaN, S = kNaN;
      if (idx_ok) {
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
      }
      if (out_idx != nullptr) out_idx[row] = idx;
      float4 o = make_float4((float)f_res, (float)Q, (float)FoM, (float)S);
      *reinterpret_cast<float4*>(out_metrics + row * 4) = o;
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
  const int64_t blocks_needed = ceil_div64(n, warps_per_block);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device");
  const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
  const int vpl = (s + 31) / 32;
#define PIGAN_LAUNCH_PHYS(V)                                                                        \
  note_launch(), physics_metrics_kernel<V><<<grid, 256, 0, st>>>(spectra, (long long)n, s, frequency, peak_idx,    \
                                                  baseline_transmission, out_idx, out_metrics)
  if (vpl <= 8) PIGAN_LAUNCH_PHYS(8);
  else if (vpl <= 16) PIGAN_LAUNCH_PHYS(16);
  else if (vpl <= 32) PIGAN_LAUNCH_PHYS(32);
  else PIGAN_LAUNCH_PHYS(64);
#undef PIGAN_LAUNCH_PHYS
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
