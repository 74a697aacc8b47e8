// Streaming (HBM-bound) kernels between the tensor-core GEMMs: casts, normalisation finalise/apply,
// the K=4 / N=4 layers that are too thin for UMMA, reductions, clip + Adam, weight packing.
// All of them read/write each element once with 16-byte vector accesses where the layout allows.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pigan {

struct BnFinalizeArgs {
  const float* colsum;    // sum over the (global) batch of the bias-free pre-activation
  const float* colsumsq;
  const float* bias_eff;  // constant added to every row (bias + centering correction)
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  float* mean;   // out: batch mean of the full pre-activation
  float* rstd;   // out
  float* scale;  // out: gamma * rstd
  float* shift;  // out: beta - mean * scale
  int C;
  double n;          // global batch size
  int num_updates;   // running-stat updates to apply (the reference runs G.forward twice per step, F8)
};

// spectrum centering / cast
void launch_center_vec(const float* x, int64_t rows, int S, int rows_used, float* cvec, int Kp, cudaStream_t st);
void launch_cast_center(const float* x, const float* cvec, __half* xc, int64_t rows, int S, int Kp, cudaStream_t st);
void launch_cast_center_noise(const float* target, const float* noise, float sigma, const float* cvec, __half* xc,
                              int64_t rows, int S, int Kp, cudaStream_t st);
// effective first-layer bias: out[i] = b[i] + sum_j c[j] * W[i*ld + j], j < S
void launch_bias_eff(const float* w, int ld, int S, const float* b, const float* cvec, float* out, int rows,
                     cudaStream_t st);
// weight packing
void launch_cast_pad(const float* src, int ld_src, int ncols, __half* dst, int ld_dst, int rows, cudaStream_t st);
void launch_transpose_cast(const float* src, int rows, int cols, int ld_src, __half* dst, int ld_dst, cudaStream_t st);
void launch_extract_cols(const float* src, int ld_src, int col0, int ncols, float* dst, int rows, cudaStream_t st);

// BatchNorm (generator)
void launch_bn_finalize(const BnFinalizeArgs& a, cudaStream_t st);
void launch_bn_eval_affine(const float* rm, const float* rv, const float* gamma, const float* beta, float* scale,
                           float* shift, int C, cudaStream_t st);
void launch_bn_relu_apply(const __half* h, const float* scale, const float* shift, __half* a, int64_t rows, int C,
                          cudaStream_t st);
// generator head: a2 = relu(bn2(h2)); p = tanh(a2 W3^T + b3); pden = denormalize(p); side operand rows for D
void launch_g_head_fwd(const __half* h2, const float* scale, const float* shift, const float* w3, const float* b3,
                       float* p_out, float* pden_out, const float* p_real, __half* paug, int64_t rows, int C,
                       cudaStream_t st);
struct GHeadBwdArgs {
  const float* p;         // [B,4] tanh output
  const float* dpden;     // [B,4] dL/dPden * GS (may be null)
  const float* dp_lc;     // [B,4] lambda_lc * dLC/dp * GS (may be null)
  float range_mult;       // lambda_range * GS / (4 * global batch)
  const __half* h2;       // [B,C]
  const float* scale;     // BN2 affine
  const float* shift;
  const float* mean;
  const float* rstd;
  const float* w3;        // [4,C]
  __half* dy2;            // out [B,C] (scaled)
  float* dw3;             // [4,C] += (unscaled)
  float* db3;             // [4]   +=
  float* sum_dy;          // [C] += (scaled)
  float* sum_dyx;         // [C] +=
  double* range_sum;      // += sum of clamp terms
  float inv_gs;
  int64_t rows;
  int C;
};
void launch_g_head_bwd(const GHeadBwdArgs& a, cudaStream_t st);
// dX = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)); bias / gamma / beta gradients
struct BnBwdArgs {
  const __half* dy;   // [B,C] scaled
  const __half* h;    // [B,C]
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* sum_dy;
  const float* sum_dyx;
  __half* dh;         // out [B,C] scaled (may alias dy)
  float* dbias;       // [C] += sum dh / GS
  float* dgamma;      // [C] = sum_dyx / GS
  float* dbeta;       // [C] = sum_dy / GS
  double inv_n;       // 1 / global batch
  float inv_gs;
  int64_t rows;
  int C;
};
void launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t st);

// forward model
void launch_f_l1(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb, __half* out,
                 int64_t rows, int C, cudaStream_t st);
void launch_ln_lrelu_apply(__half* h, const float* rowstats, int n_tiles, int tile_cols, const float* gamma,
                           const float* beta, int64_t rows, int N, cudaStream_t st);

// optimiser
void launch_sumsq(const float* g, int64_t n, double* out, cudaStream_t st);
struct AdamArgs {
  float* p;
  float* g;
  float* m;
  float* v;
  int64_t n;
  float lr, beta1, beta2, eps;
  double bias_c1, bias_c2;  // 1 - beta^t
  const double* total_sq;   // squared global grad norm
  float max_norm;
};
void launch_clip_adam(const AdamArgs& a, cudaStream_t st);
// rank-1 fix-ups of first-layer weight gradients: dw[i*ld + j] += db[i] * cvec[j] (j < S);
// dw[i*ld + S + e] += pcenter * db[i] (e < P)
void launch_dw_fixup(float* dw, int ld, int S, const float* db, const float* cvec, int P, float pcenter, int rows,
                     cudaStream_t st);

struct LossFinalizeArgs {
  const double* sums;  // see engine.cu kSum* indices
  float* out9;         // loss_history order
  double batch;        // global batch
  int S, Mt;
  float lam_recon, lam_phys_spec, lam_phys_metrics, lam_maxwell, lam_lc, lam_range, lam_kl;
};
void launch_loss_finalize(const LossFinalizeArgs& a, cudaStream_t st);

// candidate scoring helpers
void launch_count_violations(const float* p, int64_t rows, int P, int32_t* out, cudaStream_t st);

}  // namespace pigan
