// Evaluator reductions on the device: the regression metrics of
// UnifiedEvaluator.calculate_metrics (core/evaluate/unified_evaluator.py:138-184) and the summary statistics of the
// structural-prediction loop (:393-405), computed from per-column / per-array sums so that (a) nothing is copied
// to the host per batch and (b) data-parallel ranks can add their sums before the final formulas.
// HBM-bound streaming kernels: every input element is read once; sums are fp64, two-stage and atomic-free
// (block partials -> fixed-order reduction), hence deterministic.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "../../include/pigan_b200.h"
#include "host_util.h"

namespace pigan {
namespace {

constexpr int kQ = 8;            // sums per column: y, p, y^2, p^2, y p, |y-p|, (y-p)^2, |(y-p)/(y+1e-8)|
constexpr int kMaxBlocks = 148 * 4;
constexpr int kSumm = 6;         // score summary sums: [viol>0], viol, err, err^2, cons, cons^2

// grid (row blocks, column tiles of 32); block 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) regression_partials_kernel(const float* __restrict__ y,
                                                                  const float* __restrict__ p, long long n, int C,
                                                                  double* __restrict__ part) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ double sm[8][32][kQ + 1];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  double s[kQ];
#pragma unroll
  for (int q = 0; q < kQ; ++q) s[q] = 0.0;
  if (c < C) {
    // 8 rows in flight per thread: the kernel is bound by memory-level parallelism, not by the fp64 adds
    constexpr int U = 8;
    const long long stride = (long long)gridDim.x * 8;
    for (long long r0 = (long long)blockIdx.x * 8 + ry; r0 < n; r0 += U * stride) {
      float yv[U], pv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * stride;
        yv[u] = r < n ? __ldg(y + r * C + c) : 0.f;
        pv[u] = r < n ? __ldg(p + r * C + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r0 + u * stride >= n) break;
        const float yf = yv[u], pf = pv[u];
        const float df = yf - pf;
        const float ape = fabsf(df / (yf + 1e-8f));   // float32 arithmetic as in numpy (unified_evaluator.py:182)
        const double yd = yf, pd = pf, d = (double)yf - (double)pf;
        s[0] += yd; s[1] += pd;
        s[2] = fma(yd, yd, s[2]); s[3] = fma(pd, pd, s[3]); s[4] = fma(yd, pd, s[4]);
        s[5] += fabs(d); s[6] = fma(d, d, s[6]); s[7] += (double)ape;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kQ; ++q) sm[ry][cx][q] = s[q];
  __syncthreads();
  if (ry == 0 && c < C) {
    double* o = part + ((size_t)blockIdx.x * C + c) * kQ;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][cx][q];
      o[q] = t;
    }
  }
}

// sums[i] (+)= sum_b part[b * width + i]
__global__ void reduce_f64_kernel(const double* __restrict__ part, int nblocks, int width, double* __restrict__ sums,
                                  int accumulate) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += part[(size_t)b * width + i];
  sums[i] = accumulate ? sums[i] + t : t;
}

// one block; out[6] = mse, mae, rmse, r2, pearson_r, mape (unified_evaluator.py:150-183)
__global__ void regression_finalize_kernel(const double* __restrict__ sums, double n, int C, double* __restrict__ out) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ double sm[5][256];
  double a_abs = 0.0, a_sq = 0.0, a_ape = 0.0, a_r2 = 0.0, a_pr = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double* s = sums + (size_t)c * kQ;
    a_abs += s[5]; a_sq += s[6]; a_ape += s[7];
    const double ss_tot = s[2] - s[0] * s[0] / n;     // sum (y - mean y)^2
    const double ss_p = s[3] - s[1] * s[1] / n;
    const double cov = s[4] - s[0] * s[1] / n;
    // sklearn r2_score (force_finite): a constant column scores 1 when it is reproduced exactly, else 0
    const double tiny = 1e-12 * fmax(s[2], 1e-300);
    a_r2 += ss_tot > tiny ? 1.0 - s[6] / ss_tot : (s[6] <= tiny ? 1.0 : 0.0);
    // scipy.stats.pearsonr: NaN for a constant input
    a_pr += (ss_tot > tiny && ss_p > 1e-12 * fmax(s[3], 1e-300)) ? cov / sqrt(ss_tot * ss_p) : nan("");
  }
  sm[0][threadIdx.x] = a_abs; sm[1][threadIdx.x] = a_sq; sm[2][threadIdx.x] = a_ape;
  sm[3][threadIdx.x] = a_r2; sm[4][threadIdx.x] = a_pr;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < 5; ++k)
      for (int i = 0; i < (int)blockDim.x; ++i) t[k] += sm[k][i];
    const double cnt = n * C;
    out[0] = t[1] / cnt;
    out[1] = t[0] / cnt;
    out[2] = sqrt(out[0]);
    out[3] = t[3] / C;
    out[4] = t[4] / C;
    out[5] = t[2] / cnt * 100.0;
  }
}

__global__ void __launch_bounds__(256) score_partials_kernel(const int32_t* __restrict__ viol,
                                                             const float* __restrict__ err,
                                                             const float* __restrict__ cons, long long n,
                                                             double* __restrict__ part) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ double sm[kSumm][8];
  double s[kSumm] = {0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = viol ? __ldg(viol + i) : 0;
    const double e = err ? (double)__ldg(err + i) : 0.0, c = cons ? (double)__ldg(cons + i) : 0.0;
    s[0] += v > 0 ? 1.0 : 0.0; s[1] += (double)v;
    s[2] += e; s[3] = fma(e, e, s[3]); s[4] += c; s[5] = fma(c, c, s[5]);
  }
#pragma unroll
  for (int q = 0; q < kSumm; ++q) {
    double t = s[q];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) sm[q][threadIdx.x >> 5] = t;
  }
  __syncthreads();
  if (threadIdx.x < kSumm) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sm[threadIdx.x][k];
    part[(size_t)blockIdx.x * kSumm + threadIdx.x] = t;
  }
}

// out[6] = violation rate, mean violations, mean / std (population, as np.std) of err and cons
__global__ void score_finalize_kernel(const double* __restrict__ s, double n, double* __restrict__ out) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x != 0) return;
  out[0] = s[0] / n;
  out[1] = s[1] / n;
  const double me = s[2] / n, mc = s[4] / n;
  out[2] = me;
  out[3] = sqrt(fmax(s[3] / n - me * me, 0.0));
  out[4] = mc;
  out[5] = sqrt(fmax(s[5] / n - mc * mc, 0.0));
}

int row_blocks(int64_t n, int rows_per_block) {
  int64_t b = (n + rows_per_block - 1) / rows_per_block;
  if (b < 1) b = 1;
  return (int)(b < kMaxBlocks ? b : kMaxBlocks);
}

}  // namespace
}  // namespace pigan

using namespace pigan;

extern "C" size_t pigan_eval_workspace_bytes(int32_t cols) {
  if (cols < 1) return 0;
  const size_t a = (size_t)kMaxBlocks * (size_t)cols * kQ * sizeof(double);
  const size_t b = (size_t)kMaxBlocks * kSumm * sizeof(double);
  return ((a > b ? a : b) + 255) & ~size_t(255);
}

extern "C" int pigan_regression_sums(const float* y_true, const float* y_pred, int64_t n, int32_t cols, double* sums,
                                     int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  PIGAN_CHECK_ARG(y_true && y_pred && sums && workspace && n >= 1 && cols >= 1);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0);
  if (workspace_bytes < pigan_eval_workspace_bytes(cols))
    return fail(PIGAN_ERR_WORKSPACE, "evaluator workspace too small: %zu < %zu", workspace_bytes,
                pigan_eval_workspace_bytes(cols));
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  const int gx = row_blocks(n, 8 * 8 * 4);
  launch_k(regression_partials_kernel, dim3(gx, (cols + 31) / 32), 256, 0, st, y_true, y_pred, (long long)n, (int)cols,
           part);
  const int width = cols * kQ;
  launch_k(reduce_f64_kernel, (width + 127) / 128, 128, 0, st, (const double*)part, gx, width, sums, (int)accumulate);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_regression_finalize(const double* sums, int64_t n_total, int32_t cols, double* out6, void* stream) {
  PIGAN_CHECK_ARG(sums && out6 && n_total >= 1 && cols >= 1);
  launch_k(regression_finalize_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), sums, (double)n_total, (int)cols,
           out6);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_score_summary_sums(const int32_t* violations, const float* recon_error, const float* consistency,
                                        int64_t n, double* sums, int32_t accumulate, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  PIGAN_CHECK_ARG(sums && workspace && n >= 1);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0);
  if (workspace_bytes < pigan_eval_workspace_bytes(1))
    return fail(PIGAN_ERR_WORKSPACE, "evaluator workspace too small");
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  const int gx = row_blocks(n, 256 * 8);
  launch_k(score_partials_kernel, gx, 256, 0, st, violations, recon_error, consistency, (long long)n, part);
  launch_k(reduce_f64_kernel, 1, 128, 0, st, (const double*)part, gx, kSumm, sums, (int)accumulate);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_score_summary_finalize(const double* sums, int64_t n_total, double* out6, void* stream) {
  PIGAN_CHECK_ARG(sums && out6 && n_total >= 1);
  launch_k(score_finalize_kernel, 1, 32, 0, static_cast<cudaStream_t>(stream), sums, (double)n_total, out6);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
