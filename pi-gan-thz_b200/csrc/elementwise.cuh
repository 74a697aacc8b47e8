// Streaming (HBM-bound) kernels between the tensor-core GEMMs: casts, normalisation statistics / apply,
// the K=4 / N=4 layers that are too thin for UMMA, reductions over the batch (a thread owns columns here,
// so per-column sums are register accumulations), clip + Adam, weight packing.
// Every kernel reads/writes each element once with 16-byte vector accesses where the layout allows.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pigan {

constexpr float kParamCenter = 2.5f;

// Batch reductions (column sums over all rows) are two-stage and atomic-free: each of at most kPartBlocks blocks
// stores its partial column sums to one row of a scratch matrix, reduce_partials_kernel adds the rows in a fixed
// order.  The engine provides kPartBlocks * kPartCols floats of scratch.
constexpr int kPartBlocks = 148 * 4;
constexpr int kPartCols = 8 * 512;
struct ReduceSeg {
  float* dst;   // dst[c] += mult * column sum (nullptr: skip)
  int ncols;
  float mult;
};
struct ReduceArgs {
  const float* part;
  int nblocks;
  int ld;       // columns in use (sum of seg[].ncols)
  int nseg;
  ReduceSeg seg[8];
};

// ------------------------------------------------------------------------------------------ spectrum prep
// cvec[j] = mean of x[0:rows_used, j] (j < S), 0 for S <= j < Kp.  Any constant row works (the shift is undone
// in fp32 through the effective bias); the batch mean keeps the fp16 operand small.
void launch_center_vec(const float* x, int64_t rows, int S, int rows_used, float* cvec, int Kp, cudaStream_t st);
// xc[r, 0:S] = x[r] - cvec; xc[r, S:S+P] = params[r] - 2.5 (0 when params == null); xc[r, S+P : S+P+2] = 1
void launch_cast_center(const float* x, const float* cvec, const float* params, __half* xc, int64_t rows, int S,
                        int P, int Kp, cudaStream_t st);
// candidate spectra x = target + sigma * noise (unified_evaluator.py:453-455), centred on cvec
// target_stride = S: per-row base spectra x[r] + sigma * noise[r] (the evaluator's stability test, :453-455)
void launch_cast_center_noise(const float* target, const float* noise, float sigma, const float* cvec, __half* xc,
                              float* x_out, int64_t rows, int S, int P, int Kp, cudaStream_t st,
                              int64_t target_stride = 0);

// candidates with in-kernel Philox4x32-10 noise: z(first + r, j) depends only on (seed, global candidate index, j)
void launch_cast_center_philox(const float* target, float sigma, uint64_t seed, int64_t first, const float* cvec,
                               __half* xc, float* noise_out, int64_t rows, int S, int Kp, cudaStream_t st);
// running top-k of the inverse-design search
void launch_search_init(float* scores, int64_t* iota, int64_t* best_idx, float* params, int64_t total, int k,
                        cudaStream_t st);
void launch_fill_inf(float* s, int64_t n, cudaStream_t st);
void launch_search_gather(const int64_t* pos, const float* params, const int64_t* best_idx, int64_t group_base, int k,
                          int64_t* tmp_idx, float* tmp_params, cudaStream_t st);
void launch_search_commit(const float* sel_scores, const int64_t* tmp_idx, const float* tmp_params, int k,
                          float* scores, int64_t* best_idx, float* params, cudaStream_t st);

// ------------------------------------------------------------------------------------------ weight packing
// First layers (G: main.0, D: main.0): out[i, 0:S] = W[i, 0:S]; out[i, S:S+P] = W[i, S:S+P] (wp_cols = P) or 0;
// out[i, S+P] = hi(b_eff), out[i, S+P+1] = lo(b_eff) with
//   b_eff[i] = b[i] + sum_j cvec[j] W[i,j] + 2.5 * sum_e W[i, S+e]   (fp32; also written to b_eff_out)
// bias_cols = 0 leaves the two bias columns zero (generator: the stored pre-BatchNorm value stays bias-free so
// its fp16 rounding is relative to the batch spread, not to the constant offset; b_eff is applied in fp32).
void launch_pack_first_layer(const float* w, int ld_src, int S, int P, int wp_cols, int bias_cols, const float* b,
                             const float* cvec, __half* out, int Kp, float* b_eff_out, int rows, cudaStream_t st);
// all operand copies of one network's weights in a single launch (H1, H2 multiples of 32; w2 is [H2, H1] dense)
void launch_pack_net(const float* w1, int ld1, int S, int P, int wp_cols, int bias_cols, const float* b1,
                     const float* cvec, __half* w1h, int Kp, float* b_eff, int H1, const float* w2, int H2,
                     __half* w2h, __half* w2th, float* wp, cudaStream_t st);
// constant image of EpiHeadF1 (epilogues.cuh: kHeadImgFloats = 3584 floats); widths fixed at 256 / 4
void launch_head_consts(const float* scale2, const float* bias2, const float* w3, const float* b3, const float* fw1,
                        const float* fb1, const float* flnw, const float* flnb, float* out, cudaStream_t st);
void launch_cast_pad(const float* src, int ld_src, int ncols, __half* dst, int ld_dst, int rows, cudaStream_t st);
void launch_transpose_cast(const float* src, int rows, int cols, int ld_src, __half* dst, int ld_dst,
                           cudaStream_t st);
// wp[i][e] = W[i, S+e] (fp32, [rows_pad][4], zero beyond rows)
void launch_extract_wp(const float* w, int ld_src, int S, int P, float* wp, int rows, int rows_pad, cudaStream_t st);
void launch_copy_pad_f32(const float* src, int n, float* dst, int n_pad, cudaStream_t st);

// ------------------------------------------------------------------------------------------ BatchNorm (G)
// sum[c] += sum_r h[r,c], sumsq[c] += sum_r h[r,c]^2
void launch_colstats(const __half* h, int64_t rows, int C, float* sum, float* sumsq, float* part, cudaStream_t st);
void launch_sigmoid_bwd(float* p_inout, const float* grad_out, float scale, int64_t n, cudaStream_t st);
void launch_scale_copy(const float* src, float* dst, float mult, int64_t n, cudaStream_t st);
void launch_reduce_columns(const ReduceArgs& a, cudaStream_t st);   // dst[c] += mult * sum over the partial rows
struct BnFinalizeArgs {
  const float* sum;       // sums over the (global) batch of the stored (bias-free) pre-activation
  const float* sumsq;
  const float* offset;    // per-column constant the stored value omits (effective bias): enters running_mean only
  const float* gamma;
  const float* beta;
  float* running_mean;    // may be null
  float* running_var;
  long long* num_batches_tracked;
  float* mean;            // out: batch mean of the stored value
  float* rstd;            // out
  float* scale;           // out: gamma * rstd
  float* bias;            // out: beta - mean * scale   (affine applied to the stored value)
  int C;
  double n;               // global batch size
  int num_updates;        // running-stat updates to apply (the reference runs G.forward twice per step, F8)
};
void launch_bn_finalize(const BnFinalizeArgs& a, cudaStream_t st);
// reduction of [nblocks][2C] partial rows (sums | sums of squares) + finalize in one launch
void launch_bn_reduce_finalize(const BnFinalizeArgs& a, const float* part, int nblocks, cudaStream_t st);
// eval mode: y = gamma * (h + offset - running_mean) / sqrt(running_var + eps) + beta = scale * h + bias
void launch_bn_eval_affine(const float* rm, const float* rv, const float* gamma, const float* beta,
                           const float* offset, float* scale, float* bias, int C, cudaStream_t st);
void launch_bn_relu_apply(const __half* h, const float* scale, const float* bias, __half* a, int64_t rows, int C,
                          cudaStream_t st);
// generator head: a2 = relu(bn2(h2)); p = tanh(a2 W3^T + b3); pden = denormalize(p); also builds the fake-row
// tail of the spectrum operand (copy of xc[:, Kp-64:Kp] with the parameter columns replaced by pden - 2.5)
void launch_g_head_fwd(const __half* h2, const float* scale, const float* bias, const float* w3, const float* b3,
                       float* p_out, float* pden_out, const __half* xc, __half* tail_fake, int64_t rows, int C,
                       int Kp, int S, cudaStream_t st);
struct GHeadBwdArgs {
  const float* p;         // [B,4] tanh output
  const float* dpden;     // [B,4] dL/dPden * GS (may be null)
  const float* dp_lc;     // [B,4] lambda_lc * dLC/dp * GS (may be null)
  const float* dp_extra;  // [B,4] caller's extra d(loss)/dp, unscaled (may be null)
  float gs;               // gradient scale GS (= global batch)
  float range_mult;       // lambda_range * GS / (4 * global batch)
  const __half* h2;       // [B,C]
  const float* scale;     // BN2 affine
  const float* bias;
  const float* mean;
  const float* rstd;
  const float* w3;        // [4,C]
  __half* dy2;            // out [B,C] (apply pass): BatchNorm-2 input gradient dh2, scaled by GS
  float* dw3;             // [4,C] += (unscaled)
  float* db3;             // [4]   +=
  float* sum_dy;          // [C] += (scaled)
  float* sum_dyx;         // [C] +=
  double* range_sum;      // += sum of clamp terms
  float* dpre_part;       // scratch [blocks][8] of g_head_dpre_kernel (set by launch_g_head_bwd)
  float inv_gs;
  int64_t rows;
  int C;
  int ld;                 // row pitch of h2 / dy2 and of w3 / dw3 (set by launch_g_head_bwd: = C, or the full width when
                          // the launcher splits a layer wider than 1024 columns into segments)
  // apply pass only
  const float* gamma;     // BN2 weight
  float* dbias;           // [C] += sum dh / GS   (main.3.bias)
  float* dgamma;          // [C] += sum_dyx / GS
  float* dbeta;           // [C] += sum_dy / GS
  double inv_n;           // 1 / global batch
  float* part;            // partial-sum scratch (kPartBlocks x 8C floats)
  float* dpre;            // [B,4] scratch: dL/d(pre-tanh), written by the statistics pass
};
// apply = false: statistics pass (sum_dy, sum_dyx, dw3, db3, range_sum); apply = true: writes dy2 (see kernel)
void launch_g_head_bwd(const GHeadBwdArgs& a, bool apply, cudaStream_t st);
// relu-masked column sums for the BatchNorm backward: dy = da * [scale*h+bias > 0]
void launch_bn_bwd_stats(const __half* da, const __half* h, const float* scale, const float* bias,
                         const float* mean, const float* rstd, float* sum_dy, float* sum_dyx, int64_t rows, int C,
                         float* part, cudaStream_t st);
struct BnBwdArgs {
  const __half* dy;   // [B,C] scaled
  const __half* h;    // [B,C]
  int relu_mask;      // 1: dy still needs the ReLU mask (recomputed from h)
  const float* scale;
  const float* bias;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* sum_dy;
  const float* sum_dyx;
  __half* dh;         // out [B,C] scaled (may alias dy)
  float* dbias;       // [C] += sum dh / GS
  float* dgamma;      // [C] += sum_dyx / GS   (written by block 0)
  float* dbeta;       // [C] += sum_dy / GS
  double inv_n;       // 1 / global batch
  float inv_gs;
  int64_t rows;
  int C;
  int ld;             // row pitch of dy / h / dh (set by launch_bn_bwd_apply, see GHeadBwdArgs::ld)
  float* part;        // partial-sum scratch (kPartBlocks x C floats)
};
void launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t st);

// ------------------------------------------------------------------------------------------ discriminator
// dh2[r,c] = dlogit[r] * w3[c] * LeakyReLU'(z2[r,c]); dw3[c] += sum_r dlogit[r] z2[r,c] / GS;
// db2[c] += sum_r dh2[r,c] / GS; db3 += sum_r dlogit[r] / GS   (discriminator.py:24-26 backward)
void launch_d_l2_bwd(const __half* z2, const float* dlogit, const float* w3, __half* dh2, float* dw3, float* db2,
                     float* db3, int64_t rows, int C, float inv_gs, float* part, cudaStream_t st);

// ------------------------------------------------------------------------------------------ forward model
void launch_f_l1(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb, __half* out,
                 int64_t rows, int C, cudaStream_t st);
// in place: h = LeakyReLU(LayerNorm(h)) with per-row (sum, sumsq) partials over n_tiles column tiles
void launch_ln_lrelu_apply(__half* h, const float* rowstats, int n_tiles, const float* gamma, const float* beta,
                           int64_t rows, int N, cudaStream_t st);

// ------------------------------------------------------------------------------------------ optimiser
struct ZeroArgs {
  float* ptr[4];
  long long nfloat[4];
};
// zero up to 4 fp32 buffers (16-byte aligned) in one kernel
void launch_zero_buffers(float* const* ptrs, const int64_t* nfloat, int count, cudaStream_t st);
// stage 1 of the squared gradient norm: returns the number of fp64 block partials written to parts (<= 296)
int launch_sumsq(const float* g, int64_t n, double* parts, cudaStream_t st);
struct AdamArgs {
  float* p;
  float* g;
  float* m;
  float* v;
  int64_t n;
  float lr, beta1, beta2, eps;
  double bias_c1, bias_c2;  // 1 - beta^t
  const double* sq_parts;   // block partials of the squared global grad norm (launch_sumsq)
  int n_parts;
  float max_norm;
};
void launch_clip_adam(const AdamArgs& a, cudaStream_t st);
// second stage of the split-K weight-gradient GEMM (EpiWeightGradPartial): fixed-order sum of the k-split slabs
// fix_cvec != nullptr: also dw[i*ld + j] += db[i] * (j < fix_S ? fix_cvec[j] : 2.5) for j < fix_S + fix_P (the rank-1
// fix-up of a first layer whose operand is centred), db zero on entry
void launch_dw_reduce(const float* part, int tiles_m, int tiles_n, int splits, float* dw, int ld, int m_valid,
                      int n_valid, float scale, int bias_col, float* db, cudaStream_t st,
                      const float* fix_cvec = nullptr, int fix_S = 0, int fix_P = 0);
// rank-1 fix-ups of first-layer weight gradients: dw[i*ld + j] += db[i] * cvec[j] (j < S);
// dw[i*ld + S + e] += 2.5 * db[i] (e < P)
void launch_dw_fixup(float* dw, int ld, int S, int P, const float* db, const float* cvec, int rows,
                     cudaStream_t st);

struct LossFinalizeArgs {
  const double* sums;  // engine.cu kSum* indices
  float* out9;         // loss_history order
  double batch;        // global batch
  int S, Mt, P;
  float lam_recon, lam_phys_spec, lam_phys_metrics, lam_maxwell, lam_lc, lam_range, lam_kl;
};
void launch_loss_finalize(const LossFinalizeArgs& a, cudaStream_t st);

// ------------------------------------------------------------------------------------------ scoring
// violations[r] = #{j : p[r,j] < 0 or p[r,j] > 1}; consistency[r] = 1 / (1 + err[r])  (unified_evaluator.py:380,391)
void launch_score_finish(const float* p, const float* err, int64_t rows, int P, int32_t* violations,
                         float* consistency, cudaStream_t st);
// dout [rows, ld] fp16 = scale * g [rows, cols] (zero padded)
void launch_f_upstream_cast(const float* g, int cols, __half* dout, int ld, int64_t rows, float scale, cudaStream_t st);
// stability[r] = mean_j (p - p_noisy)^2, plausibility[r] = mean_j sigmoid(10 p - 5)   (unified_evaluator.py:458-468)
void launch_validation_scores(const float* p, const float* p_noisy, int64_t rows, int P, float* stability,
                              float* plausibility, cudaStream_t st);


// ------------------------------------------------------------------------------------------ surrogate training
// Counter-based Dropout of the training-mode forward model (see drop_keep8 in elementwise.cu)
struct DropoutArgs {
  unsigned long long seed;
  long long first_row;      // global index of local row 0 (data-parallel shards)
  unsigned int step;        // optimiser step: a fresh mask every step
  unsigned int thresh16;    // drop when 16 random bits < thresh16 (= round(p * 65536))
  float keep_scale;         // 1 / (1 - p)
};
void launch_f_l1_train(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb,
                       __half* xhat, __half* act, float* rstd, unsigned char* mask, unsigned char* keepbits,
                       int64_t rows, const DropoutArgs& dr, cudaStream_t st);
void launch_ln_train(__half* xhat, const float* rowstats, const float* gamma, const float* beta, __half* act,
                     float* rstd, unsigned char* mask, unsigned char* keepbits, int64_t rows, int N, int layer,
                     const DropoutArgs& dr, cudaStream_t st);
// loss + output-layer gradient; db_out[S+Mt] += inv_gs * column sums of dout, loss_sums[2] += sums of squares
void launch_f_out_loss(const float* out, const float* spectrum, const float* metrics, __half* dout, int ld,
                       int64_t rows, int S, int Mt, float* part, float* db_out, float* loss_sums, float inv_gs,
                       cudaStream_t st, float w_spec = 1.f, float w_met = 1.f);
// keepbits: [rows, N / 8] bytes written by the forward kernels (bit i = column 8 k + i kept).
// da -> dh in place (not for the first layer, p_in != null, whose dW1 comes out k-major [4][N] instead)
void launch_ln_bwd(__half* da, const __half* xhat, const float* rstd, const float* gamma, const float* beta,
                   const float* p_in, const unsigned char* keepbits, int64_t rows, int N, float keep_scale,
                   float* part, float* dgamma, float* dbeta, float* dbias, float* dw1_kmajor, float inv_gs,
                   cudaStream_t st, int store_dh = 0);
// dp[r, 0:4] = scale * dh1[r, :] . W1   (input gradient of the first layer; dh1 kept by launch_ln_bwd(store_dh = 1))
void launch_f_dp(const __half* dh1, const float* w1, float* dp, int64_t rows, float scale, cudaStream_t st);
void launch_f_dw1_transpose(const float* src, float* dw1, cudaStream_t st);
// ---- widened surrogate (BASELINE config 5): any hidden width N in {256, 512, 1024, 2048}, any even S / Mt with
// round_up(S + Mt, 64) <= 2560.  dr == nullptr: eval mode (act only).
// consts: 20 floats of scratch (closed-form LayerNorm statistics of the K = 4 layer, written by this call)
void launch_f_l1_wide(const float* p, const float* w1, const float* b1, const float* lnw, const float* lnb,
                      float* consts, __half* xhat, __half* act, float* rstd, unsigned char* mask,
                      unsigned char* keepbits, int64_t rows, int N, const DropoutArgs* dr, cudaStream_t st);
// dw1[N][4] += inv_gs * dh1^T p (dw1_kmajor: [4][N] scratch, zero on entry)
void launch_f_dw1_wide(const __half* dh1, const float* p, int64_t rows, int N, float* part, float* dw1_kmajor,
                       float* dw1, float inv_gs, cudaStream_t st);
void launch_f_dp_wide(const __half* dh1, const float* w1, float* dp, int64_t rows, int N, float scale,
                      cudaStream_t st);
// fp32 output-layer accumulators stored as [128 x 256] slabs (see slab_index in elementwise.cu)
void launch_f_unslab(const float* slab, int ngroups, const float* bias, float* out, int64_t rows, int OUT,
                     cudaStream_t st);
void launch_f_out_loss_slab(const float* slab, int ngroups, const float* bias, const float* spectrum,
                            const float* metrics, __half* dout, int ld, int64_t rows, int S, int Mt, float* part,
                            float* db_out, float* loss_sums, float inv_gs, cudaStream_t st, float w_spec = 1.f,
                            float w_met = 1.f);
// ---- widened PI-GAN step: generic-width versions of what the reference-width step fuses into GEMM epilogues
// layer 3 + Sigmoid + BCE of the discriminator on the stored activation z2 [rows, C] (EpiDiscL2's arithmetic)
void launch_d_logit_bce(const __half* z2, const float* w3, const float* b3, int64_t rows, int C, int64_t rows_a,
                        float label_a, float label_b, int64_t gap_begin, int64_t gap_end, double global_batch,
                        double* loss_sum, float* dlogit, float* prob_out, cudaStream_t st);
// the G-step's surrogate losses from the output layer's fp32 slabs (EpiFwdLoss' sums and dp_lc)
void launch_f_pigan_loss_slab(const float* slab, int ngroups, const float* bias, const float* spectrum,
                              const float* metrics, const float* p_norm, int64_t rows, int S, int Mt, int f1_idx,
                              int f2_idx, float lc_grad_mult, double* sums, float* dp_lc, cudaStream_t st);
void launch_f_input_grad_losses(const float* sums, double n_spec, double n_met, float* out, cudaStream_t st);
void launch_f_train_losses(const float* sums, double n_spec, double n_met, float* out, cudaStream_t st);

}  // namespace pigan
