/*
 * pigan_b200 — C ABI of the B200-native PI-GAN-THz hot path.
 *
 * The reference (jianghu105/PI-GAN-THz) has no FFI of its own: its hot path is Python calling ATen.
 * This header is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); every entry
 * point cites the reference code it replaces.  Conventions:
 *   - plain C types only; all tensor arguments are DEVICE pointers, row-major, contiguous;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it, nothing synchronises;
 *   - the caller owns every buffer, including the workspace handed to pigan_engine_create();
 *   - return value: PIGAN_OK (0) or a negative PIGAN_ERR_* code; pigan_last_error() has the text;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with PIGAN_ERR_CUDA.
 */
#ifndef PIGAN_B200_H
#define PIGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIGAN_ABI_VERSION 9

#define PIGAN_OK 0
#define PIGAN_ERR_INVALID (-1)     /* bad argument (null pointer, size, unsupported dimension) */
#define PIGAN_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define PIGAN_ERR_UNSUPPORTED (-3) /* shape outside what this build implements */
#define PIGAN_ERR_WORKSPACE (-4)   /* workspace too small */

int pigan_abi_version(void);
/* Thread-local, valid until the next failing call on the same thread. */
const char* pigan_last_error(void);

/* Number of CUDA kernels this library has launched in the calling process (all engines, all streams):
 * bench.py reports the difference across its timed region as gpu_launches. */
int64_t pigan_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Network dimensions.  Defaults are the reference's hard-coded widths
 * (core/models/generator.py:17-26, discriminator.py:21-28, forward_model.py:28-60, config/config.py:38-55).
 * ---------------------------------------------------------------------------------------------- */
typedef struct PiganDims {
  int32_t spectrum_dim; /* 250  cfg.SPECTRUM_DIM */
  int32_t param_dim;    /* 4    cfg.GENERATOR_OUTPUT_PARAM_DIM */
  int32_t metrics_dim;  /* 8    cfg.FORWARD_MODEL_OUTPUT_METRICS_DIM */
  int32_t g_hidden[2];  /* 512, 256 */
  int32_t d_hidden[2];  /* 512, 256 */
  int32_t f_hidden[5];  /* 256, 512, 1024, 512, 256 */
} PiganDims;

void pigan_default_dims(PiganDims* dims);

/* Flat fp32 parameter buffers.  Each network's trainable tensors live in ONE contiguous device buffer in
 * nn.Module.state_dict() order (Linear weight [out,in] then bias, norm weight then bias), which is what the
 * host shim exposes as nn.Parameter views, what Adam/clip run over and what NCCL all-reduces.
 *   generator      : main.0.{weight,bias} main.1.{weight,bias} main.3.{weight,bias} main.4.{weight,bias}
 *                    main.6.{weight,bias}                                   (262 404 floats at defaults)
 *   discriminator  : main.0.{weight,bias} main.2.{weight,bias} main.4.{weight,bias}          (262 145)
 *   forward model  : model.{0,1,4,5,8,9,12,13,16,17,20}.{weight,bias}                      (1 385 730)
 * BatchNorm buffers (generator): running_mean1[h1] running_var1[h1] running_mean2[h2] running_var2[h2]
 * contiguous fp32, plus num_batches_tracked[2] int64.                                               */
int64_t pigan_generator_param_count(const PiganDims* dims);
int64_t pigan_discriminator_param_count(const PiganDims* dims);
int64_t pigan_forward_model_param_count(const PiganDims* dims);
int64_t pigan_generator_bn_buffer_count(const PiganDims* dims);

/* ------------------------------------------------------------------------------------------------
 * Physics metrics — replaces calculate_peak_parameters (core/utils/data_loader.py:13-58) plus the
 * sensitivity S its callers derive (data_loader.py:96,105), batched: one warp per spectrum.
 *   spectra    [n, s] fp32 transmission (dB)
 *   frequency  [s] fp64 (the reference's np.linspace grid, data_loader.py:124)
 *   peak_idx   [n] int32 resonance index per spectrum, or NULL: argmin of the row (first occurrence)
 *   out_idx    [n] int32 peak index used (may be NULL)
 *   out_metrics[n, 4] fp32: f_res, Q, FoM, S  (NaN where the reference returns NaN)
 * ---------------------------------------------------------------------------------------------- */
int pigan_physics_metrics(const float* spectra, int64_t n, int32_t s, const double* frequency,
                          const int32_t* peak_idx, float baseline_transmission, int32_t* out_idx,
                          float* out_metrics, void* stream);

/* Differentiable physics metrics (the reference has no such term, F3): the vector-Jacobian product of
 * the four metrics above with respect to the spectrum, for loss terms on Q / FoM / S of a predicted spectrum.  The
 * branch decisions of calculate_peak_parameters (peak index, the two half-depth crossing pairs) are held fixed, as
 * autograd does on a differentiable restatement: Q depends on t[idx] (through the half-depth level), and on the two
 * samples of each crossing pair (through the linear interpolation, data_loader.py:24,36); FoM = Q / |t_min| adds its
 * own t[idx] term; f_res = frequency[idx] has zero gradient.
 *   grad_metrics [n, 4] fp32  dL/d(f_res, Q, FoM, S) (16-byte aligned)
 *   grad_spectra [n, s] fp32  out: dL/d(spectra) — at most five non-zeros per row, all zero where Q is NaN
 *   out_idx, out_metrics      optional forward outputs as in pigan_physics_metrics                               */
int pigan_physics_metrics_backward(const float* spectra, int64_t n, int32_t s, const double* frequency,
                                   const int32_t* peak_idx, float baseline_transmission, const float* grad_metrics,
                                   float* grad_spectra, int32_t* out_idx, float* out_metrics, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Engine: owns nothing but a view of the caller's workspace (activations, fp16 operand copies of the
 * weights, reduction scratch).  One engine per process / GPU / stream; calls are not re-entrant on the
 * same engine.  max_batch bounds the rows any later call may pass.
 *
 * dims: NULL or pigan_default_dims() = the reference widths: every entry point below.  Other dims (BASELINE config 5:
 * hidden 2048, 2048-point spectra; the reference's Generator / Discriminator / ForwardModel stacks at other widths)
 * select generic-width code paths:
 *   - param_dim 4, even spectrum_dim / metrics_dim with their sum padded to 64 <= 2560, every f_hidden width one of
 *     256 / 512 / 1024 / 2048: the surrogate's entry points - pigan_engine_load_forward_model,
 *     pigan_forward_model_forward / _vjp / _input_grad, pigan_fwd_train_step[_phase];
 *   - in addition spectrum_dim a multiple of 64 (<= 2048) and g_hidden / d_hidden widths of 256 / 512 / 1024 / 2048: the
 *     PI-GAN step, pigan_train_step[_phase], with fp32 spectrum / params_denorm inputs (no prepared operand).
 * The stand-alone generator / discriminator forward and backward entry points, scoring, model validation and the
 * inverse-design search run at the reference widths only and return PIGAN_ERR_UNSUPPORTED otherwise.
 * Anything else: pigan_engine_workspace_bytes returns 0 and pigan_engine_create PIGAN_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------------------------------- */
typedef struct PiganEngine PiganEngine;

size_t pigan_engine_workspace_bytes(const PiganDims* dims, int64_t max_batch);
int pigan_engine_create(PiganEngine** out, const PiganDims* dims, int64_t max_batch, void* workspace,
                        size_t workspace_bytes, void* stream);
int pigan_engine_destroy(PiganEngine* engine);
/* Frozen forward surrogate: packs fp16 operand copies of model.{4,8,12,16,20}.weight once
 * (forward_model_pretrained.pth layout, pretrain_fwd_model.py:148-150; train_pigan.py:374-377). */
int pigan_engine_load_forward_model(PiganEngine* engine, const float* f_params, void* stream);

/* Spectra are centred on a constant row before the fp16 cast of the first-layer operand (the shift is undone
 * exactly, in fp32, through the layer's effective bias).  By default each call centres on the mean of its first
 * rows.  Data-parallel training must use ONE row on all ranks — the BatchNorm sums the ranks exchange are sums of
 * the centred pre-activations — so the host sets it here: center [S] fp32 device memory that stays valid while
 * set; NULL restores the default. */
int pigan_engine_set_spectrum_center(PiganEngine* engine, const float* center);

/* Builds the fp16 operand of the first layers from fp32 data (device pointers): out[r, 0:S] = spectrum[r] - center,
 * out[r, S:S+P] = params_denorm[r] - 2.5 (0 when params_denorm is NULL), out[r, S+P : S+P+2] = 1, rest 0; 256
 * columns per row.  Rows are independent: call it once for a whole dataset and slice batches out of the result. */
int pigan_prepare_spectrum_operand(const float* spectrum, const float* params_denorm, const float* center, int64_t n,
                                   int32_t spectrum_dim, int32_t param_dim, void* out_operand, void* stream);

/* Generator.forward (core/models/generator.py:28-33).  training != 0: BatchNorm uses batch statistics
 * and updates bn_buffers / num_batches_tracked (one update); training == 0: running statistics.
 *   spectrum [n,S] fp32  ->  out_params_norm [n,P] fp32 in (-1,1) */
int pigan_generator_forward(PiganEngine* engine, const float* g_params, float* g_bn_buffers,
                            int64_t* g_num_batches_tracked, const float* spectrum, int64_t n, int32_t training,
                            float* out_params_norm, void* stream);
/* Discriminator.forward (core/models/discriminator.py:30-39): spectrum [n,S], params [n,P] (raw 2.2..2.8)
 * -> out_prob [n] (= [n,1]) */
int pigan_discriminator_forward(PiganEngine* engine, const float* d_params, const float* spectrum,
                                const float* params, int64_t n, float* out_prob, void* stream);
/* Backward passes of the two trainable modules for callers that drive them through autograd (the drop-in modules'
 * torch.autograd.Function wrappers) instead of the fused train step; what the reference gets from loss.backward()
 * on Generator / Discriminator (train_pigan.py:141,185 and every variant trainer).  Both recompute the forward pass
 * from the inputs (train-mode BatchNorm batch statistics for the generator; running buffers untouched), then run the
 * backward kernels of pigan_train_step.  grad_scale s > 0: the fp16 gradient tensors inside hold s x the gradient -
 * choose s so that s * max|upstream| is about 1.  Outputs are unscaled and OVERWRITTEN.
 *   generator:      g_grads[G params] = d sum(grad_params_norm * G(spectrum)) / d(g_params);  n >= 2
 *   discriminator:  d_grads[D params] = d sum(grad_prob * D(spectrum, params)) / d(d_params);
 *                   grad_params [n,P] (may be NULL) = the same derivative with respect to `params`           */
int pigan_generator_backward(PiganEngine* engine, const float* g_params, const float* spectrum, int64_t n,
                             const float* grad_params_norm, float grad_scale, float* g_grads, void* stream);
int pigan_discriminator_backward(PiganEngine* engine, const float* d_params, const float* spectrum,
                                 const float* params, int64_t n, const float* grad_prob, float grad_scale,
                                 float* d_grads, float* grad_params, void* stream);
/* ForwardModel.forward in eval mode (core/models/forward_model.py:62-76): params_norm [n,P] ->
 * out [n, S+Mt] fp32 (columns [0,S) spectrum, [S,S+Mt) metrics, one buffer as in the reference) */
int pigan_forward_model_forward(PiganEngine* engine, const float* params_norm, int64_t n, float* out,
                                void* stream);

/* One iteration of train_pigan's inner loop (core/train/train_pigan.py:114-187): D-step + G-step with the
 * frozen surrogate, the 7 losses of :174-181 (weights = config/config.py:79-88), clip_grad_norm_(1.0) and
 * Adam(betas 0.5/0.999) for both networks.  Phases let the host interleave the data-parallel all-reduces:
 *   phase 0  G forward up to BatchNorm-1 statistics     -> reduce bn_sums[0 .. 2*h1)
 *   phase 1  ... up to BatchNorm-2 statistics           -> reduce bn_sums[2*h1 .. 2*h1+2*h2)
 *   phase 2  G head, D-step forward/backward            -> reduce d_grads
 *   phase 3  D clip+Adam; G-step D/F forward, losses, head backward -> reduce bn_bwd_sums[0 .. 2*h2)
 *   phase 4  BatchNorm-2 backward, dW2, dX              -> reduce bn_bwd_sums[2*h2 .. 2*h2+2*h1) and loss_sums[0:8]
 *                                                          (all loss numerators are complete after phase 3; any
 *                                                          point before phase 6 will do)
 *   phase 5  BatchNorm-1 backward, dW1                  -> reduce g_grads
 *   phase 6  G clip+Adam, loss finalisation
 * With one process, pigan_train_step runs phases 0..6 back to back. */
typedef struct PiganTrainArgs {
  /* batch (device, fp32): the 5-tuple of MetamaterialDataset.__getitem__ minus the unused members */
  const float* spectrum;      /* [B,S]  real_spectrum */
  const float* params_denorm; /* [B,P]  real_params_denorm (raw 2.2..2.8) */
  const float* metrics_norm;  /* [B,Mt] real_metrics_norm */
  int64_t batch;              /* rows on this rank */
  int64_t global_batch;       /* rows over all ranks (== batch without data parallelism) */
  /* generator state */
  float* g_params;
  float* g_grads;
  float* g_exp_avg;
  float* g_exp_avg_sq;
  float* g_bn_buffers;
  int64_t* g_num_batches_tracked; /* [2] */
  /* discriminator state */
  float* d_params;
  float* d_grads;
  float* d_exp_avg;
  float* d_exp_avg_sq;
  /* optimiser (per call: schedulers run on the host per epoch, train_pigan.py:61-62,252-253) */
  float lr_g, lr_d;
  int64_t step; /* Adam step count t of this iteration, 1-based */
  /* loss weights, cfg.LAMBDA_* */
  float lambda_recon, lambda_physics_spectrum, lambda_physics_metrics, lambda_maxwell, lambda_lc,
      lambda_param_range, lambda_bnn_kl;
  int32_t f1_idx, f2_idx; /* dataset.metric_name_to_idx['f1'/'f2'] */
  /* outputs */
  float* losses; /* [9] device: d, g, adv, recon_spec, recon_metrics, maxwell, lc, param_range, bnn_kl */
  /* optional, instead of spectrum + params_denorm: the fp16 first-layer operand [B,256] built once per dataset by
   * pigan_prepare_spectrum_operand (the analogue of MetamaterialDataset's one-time normalisation,
   * data_loader.py:185-219) and the row [S] it was centred on.  Halves the bytes a step needs from the host.
   * The batch must then be a multiple of 128 rows (PIGAN_ERR_INVALID otherwise). */
  const void* spectrum_operand;
  const float* spectrum_center;
  /* optional extra gradient into the generator output (the physics-metric loss): dp_extra [B,P] =
   * d(extra loss)/d(params_norm) of a loss that is a MEAN over the global batch, already multiplied by its weight;
   * read by phase 3 (the generator-head backward), so it may be computed between phases 2 and 3 from
   * pigan_engine_generator_output.  NULL (the default, = the reference's loss) adds nothing. */
  const float* dp_extra;
  /* bit 0: keep every kernel of the step on the caller's stream (no second stream for the surrogate chain), so the
   * caller may use the engine (pigan_forward_model_forward, pigan_forward_model_vjp) between the phases */
  int32_t flags;
} PiganTrainArgs;

int pigan_train_step(PiganEngine* engine, const PiganTrainArgs* args, void* stream);
int pigan_train_step_phase(PiganEngine* engine, const PiganTrainArgs* args, int32_t phase, void* stream);
/* [B,P] generator output params_norm of the current step (valid after phase 2) */
float* pigan_engine_generator_output(PiganEngine* engine);
/* Device buffers the host all-reduces between phases (fp32 unless noted); valid for the engine's lifetime. */
float* pigan_engine_bn_sums(PiganEngine* engine);      /* [2*h1 + 2*h2]  sum, sumsq per BatchNorm */
float* pigan_engine_bn_bwd_sums(PiganEngine* engine);  /* [2*h2 + 2*h1]  sum dy, sum dy*xhat */
double* pigan_engine_loss_sums(PiganEngine* engine);   /* [16] fp64 */

/* ------------------------------------------------------------------------------------------------
 * Data-parallel exchange over NVLink peer memory (csrc/dp.cu).  The reference is single-process; under data
 * parallelism the step exchanges BatchNorm sums, loss numerators and gradients between the phases of
 * pigan_train_step_phase.  Every rank allocates one exchange region (pigan_dp_alloc: a cudaMalloc of its own so
 * that it can be exported), publishes its cudaIpc handle to the others (any host channel), opens theirs and
 * creates a context from the table of mapped regions (own region at index `rank`).
 *   pigan_dp_allreduce_small  in-place sum over ranks of a buffer of <= 8 KB (fp32 or fp64), one CTA, one launch
 *   pigan_dp_allreduce_grads  dst = sum over ranks of gradient slot (net, epoch & 1); the gradients of a step are
 *                             written straight into that slot (PiganTrainArgs.g_grads / d_grads point into it)
 * `channel` identifies the exchange point inside a step (0..15, the same on every rank), `epoch` >= 1 grows by one
 * per step.  Ranks sum in rank order, so all replicas get bit-identical results.  Waits are bounded (5 s -> trap).
 * ---------------------------------------------------------------------------------------------- */
typedef struct PiganDp PiganDp;
size_t pigan_dp_region_bytes(int64_t max_grad_floats);
size_t pigan_dp_grad_slot_offset(int64_t max_grad_floats, int32_t net /*0 G, 1 D*/, int32_t parity);
int pigan_dp_alloc(size_t bytes, void** out);
int pigan_dp_free(void* region);
int pigan_dp_ipc_export(void* region, void* handle64);
int pigan_dp_ipc_open(const void* handle64, void** out_mapped);
int pigan_dp_ipc_close(void* mapped);
int pigan_dp_create(PiganDp** out, int32_t world, int32_t rank, void* const* region_of_rank, int64_t max_grad_floats);
int pigan_dp_destroy(PiganDp* dp);
int pigan_dp_allreduce_small(PiganDp* dp, void* buf, int32_t n, int32_t is_double, int32_t channel, uint32_t epoch,
                             void* stream);
/* fp32 buffer a[0:na] and fp64 buffer b[0:nb] summed over ranks in ONE exchange (na + 2 nb <= 2048 words) */
int pigan_dp_allreduce_small2(PiganDp* d, float* a, int32_t na, double* b, int32_t nb, int32_t channel, uint32_t epoch,
                              void* stream);
int pigan_dp_allreduce_grads(PiganDp* dp, int32_t net, float* dst, int64_t n, int32_t channel, uint32_t epoch,
                             double* sumsq, void* stream);

/* Optional instrumentation (replaces the reference's time.time() ETA bookkeeping, train_pigan.py:113,218, as
 * the only timing hook): between _begin and _end every engine call records CUDA events on the caller's stream
 * at the boundaries of its kernel sections.  sections_csv = NULL or "" records all sections, else only the
 * comma-separated names.  _end synchronises on the recorded events and writes one line per section,
 * "<name> <count> <total_ms>\n", NUL-terminated, into `report`.  Off by default. */
int pigan_engine_profile_begin(PiganEngine* engine, const char* sections_csv);
int pigan_engine_profile_end(PiganEngine* engine, char* report, size_t report_bytes);

/* Inverse-design scoring — loop body of UnifiedEvaluator.evaluate_structural_prediction
 * (core/evaluate/unified_evaluator.py:376-392): G(eval) -> range violations -> F(eval) -> per-candidate
 * mean((x - recon)^2) -> 1/(1+err).  Candidates are either given (spectra [n,S]) or generated as
 * target + sigma * noise (unified_evaluator.py:453-455) with noise [n,S] passed explicitly; in the second
 * form the error is measured against `target` (the design goal).  Any output pointer may be NULL. */
int pigan_score_candidates(PiganEngine* engine, const float* g_params, const float* g_bn_buffers,
                           const float* spectra, const float* target, const float* noise, float sigma, int64_t n,
                           float* out_params_norm, int32_t* out_violations, float* out_recon_error,
                           float* out_consistency, void* stream);
/* Model-validation loop body — UnifiedEvaluator.evaluate_model_validation (core/evaluate/unified_evaluator.py:439-468)
 * for n rows in one call: out_cycle_error[r] = mean((x_r - F(G(x_r)).spectrum)^2), out_stability[r] =
 * mean((G(x_r) - G(x_r + sigma * noise_r))^2), out_plausibility[r] = mean(sigmoid(10 G(x_r) - 5)); noise [n,S] is
 * passed explicitly (the reference draws torch.randn_like).  out_params_norm [n,4] may be NULL. */
int pigan_validate_model(PiganEngine* engine, const float* g_params, const float* g_bn_buffers, const float* spectra,
                         const float* noise, float sigma, int64_t n, float* out_params_norm, float* out_cycle_error,
                         float* out_stability, float* out_plausibility, void* stream);
/* Inverse-design search (BASELINE config 4): scores `count` candidates cand_i = target + sigma * z_i, i = first_candidate
 * ... first_candidate + count - 1, through G(eval) -> F(eval) -> mean((target - recon)^2) and keeps the k best.
 * z_i is drawn in-kernel (Philox4x32-10 keyed by `seed`, counter = (i, column block)), so candidate i gets the same
 * noise whatever the chunking, sharding or number of ranks — a rank passes its own [first_candidate, +count) range
 * and the k-sized results of all ranks are merged with pigan_topk_smallest after one all-gather.
 *   out_scores [k] ascending (+inf where fewer than k candidates exist), out_indices [k] global candidate ids
 *   (-1 for empty slots), out_params_norm [k,4] generator outputs of the winners.
 *   noise_dump: NULL, or [count, S] fp32 receiving z (tests replay it through pigan_score_candidates).
 * workspace: pigan_search_workspace_bytes(engine, k) bytes of device memory, 256-byte aligned; k <= 4096. */
size_t pigan_search_workspace_bytes(const PiganEngine* engine, int32_t k);
int pigan_inverse_design_search(PiganEngine* engine, const float* g_params, const float* g_bn_buffers,
                                const float* target, float sigma, uint64_t seed, int64_t first_candidate,
                                int64_t count, int32_t k, float* out_scores, int64_t* out_indices,
                                float* out_params_norm, float* noise_dump, void* workspace, size_t workspace_bytes,
                                void* stream);

/* ------------------------------------------------------------------------------------------------
 * Forward-surrogate training step — replaces the loop body of pretrain_forward_model,
 * /root/reference/core/train/pretrain_fwd_model.py:68-92: F in train mode (forward_model.py:28-57, Dropout 0.2),
 * loss = MSE(spectrum) + MSE(metrics) (:80-84), backward, clip_grad_norm_(1.0) (:90), Adam(lr) (:43, :91).
 * Dropout keep-masks are counter based (Philox4x32-10 keyed by dropout_seed, counter = (first_row + row,
 * layer * 4096 + column / 8, step); a column is dropped when its 16 random bits < round(p * 65536)), so the
 * same (seed, step, global row) gives the same mask for any sharding; torch's own RNG stream is not reproduced.
 * Parameters / Adam state: flat fp32 buffers in state_dict order (model.0.weight, model.0.bias, model.1.weight,
 * ... model.20.bias), 1 385 730 floats for the reference widths.
 *   losses [3] (device): total, spectrum, metrics — means over global_batch rows.
 *   loss_sums: NULL, or [2] device floats that receive this rank's sums of squares in phase 0 (all-reduce them,
 *              and f_grads, between the phases when training data-parallel).
 *   mask_dump: NULL, or [sum_i H_i][...] bytes laid out layer after layer, layer i as [batch, H_i] keep flags (tests).
 * The step reuses the engine's surrogate activations and packed weights: afterwards the frozen surrogate is
 * unloaded, call pigan_engine_load_forward_model again before pigan_train_step / scoring.
 * workspace: pigan_fwd_train_workspace_bytes(engine) bytes of device memory, 256-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct PiganFwdTrainArgs {
  const float* params_norm;   /* [batch, param_dim]    inputs of F (data_loader.py:185-196 normalisation) */
  const float* spectrum;      /* [batch, spectrum_dim] regression target 1 */
  const float* metrics_norm;  /* [batch, metrics_dim]  regression target 2 */
  int64_t batch;              /* local rows */
  int64_t global_batch;       /* rows over all ranks (loss means and gradient scale) */
  int64_t first_row;          /* global index of local row 0 (Dropout counter) */
  float* f_params;
  float* f_grads;             /* out: unclipped after phase 0, clipped after phase 1 (as torch leaves .grad) */
  float* f_exp_avg;
  float* f_exp_avg_sq;
  float lr;
  int64_t step;               /* Adam t, 1-based; also the Dropout counter word */
  float beta1, beta2, eps;    /* torch defaults 0.9, 0.999, 1e-8 (pretrain_fwd_model.py:43) */
  float max_norm;             /* 1.0 (pretrain_fwd_model.py:90) */
  float dropout_p;            /* 0.2 (forward_model.py:33); 0 disables Dropout */
  uint64_t dropout_seed;
  float* losses;
  float* loss_sums;
  uint8_t* mask_dump;
} PiganFwdTrainArgs;
size_t pigan_fwd_train_workspace_bytes(const PiganEngine* engine);
/* Vector-Jacobian product of the surrogate with frozen weights (eval mode): out_dp [n,P] = J^T grad_out for
 * grad_out [n, S+Mt] = dL/d(F(params_norm)) (spectrum columns first, then metrics), the general form of
 * pigan_forward_model_input_grad (UnifiedTrainer semantics, core/train/unified_trainer.py:240-256,325: a loss on
 * F(G(x)) that reaches G through F).  Same workspace and engine-state caveats as pigan_fwd_train_step. */
int pigan_forward_model_vjp(PiganEngine* engine, const float* f_params, const float* params_norm,
                            const float* grad_out, int64_t n, float* out_dp, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Gradient of the surrogate's regression loss with respect to its INPUT, weights frozen (the
 * physics-loss gradient of UnifiedTrainer.train_pigan_step, /root/reference/core/train/unified_trainer.py:240-256,
 * 325).  F runs in eval mode (Dropout = identity) on params_norm [n, param_dim];
 *   L = w_spectrum * MSE(F(p).spectrum, spectrum) + w_metrics * MSE(F(p).metrics, metrics_norm)   (means over n rows)
 *   out_dp [n, param_dim] = dL/d(params_norm) (16-byte aligned); out_losses [2] (device) = the two unweighted MSEs.
 * f_params: flat parameters as for pigan_engine_load_forward_model (passing the loaded buffer keeps it loaded).
 * workspace: pigan_fwd_train_workspace_bytes(engine). */
int pigan_forward_model_input_grad(PiganEngine* engine, const float* f_params, const float* params_norm,
                                   const float* spectrum, const float* metrics_norm, int64_t n, float w_spectrum,
                                   float w_metrics, float* out_dp, float* out_losses, void* workspace,
                                   size_t workspace_bytes, void* stream);
int pigan_fwd_train_step(PiganEngine* engine, const PiganFwdTrainArgs* args, void* workspace, size_t workspace_bytes,
                         void* stream);
/* phase 0: forward + loss + backward (local gradients, already divided by global_batch); phase 1: clip + Adam */
int pigan_fwd_train_step_phase(PiganEngine* engine, const PiganFwdTrainArgs* args, int32_t phase, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluator reductions on the device.
 * Regression metrics of UnifiedEvaluator.calculate_metrics, /root/reference/core/evaluate/unified_evaluator.py:
 * 138-184 (sklearn mean_squared_error / mean_absolute_error / r2_score with uniform column average, mean of the
 * per-column scipy pearsonr, MAPE with the reference's +1e-8 in float32):
 *   pigan_regression_sums      y_true, y_pred [n, cols] fp32 row-major -> sums[cols][8] fp64 (sum y, p, y^2, p^2,
 *                              y p, |y-p|, (y-p)^2, |(y-p)/(y+1e-8)|); accumulate != 0 adds to what sums holds
 *                              (batches / all-reduce over ranks happen on the sums)
 *   pigan_regression_finalize  sums + total row count -> out[6] fp64 = mse, mae, rmse, r2, pearson_r, mape
 * Summary of the structural-prediction loop (:393-405):
 *   pigan_score_summary_sums      violations int32[n], recon_error, consistency fp32[n] (any may be NULL) ->
 *                                 sums[6] fp64 (count(v>0), sum v, sum e, sum e^2, sum c, sum c^2)
 *   pigan_score_summary_finalize  -> out[6] = param_range_violation_rate, avg_param_violations,
 *                                 reconstruction_error_mean, _std (population, np.std), consistency_score_mean, _std
 * workspace: pigan_eval_workspace_bytes(cols) bytes of device memory, 16-byte aligned (cols = 1 for the summary).
 * ---------------------------------------------------------------------------------------------- */
size_t pigan_eval_workspace_bytes(int32_t cols);
int pigan_regression_sums(const float* y_true, const float* y_pred, int64_t n, int32_t cols, double* sums,
                          int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);
int pigan_regression_finalize(const double* sums, int64_t n_total, int32_t cols, double* out6, void* stream);
int pigan_score_summary_sums(const int32_t* violations, const float* recon_error, const float* consistency, int64_t n,
                             double* sums, int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);
int pigan_score_summary_finalize(const double* sums, int64_t n_total, double* out6, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-device data pipeline: the dataset lives in HBM; batches are row gathers, synthetic
 * spectra are generated where they are consumed.  Replaces DataLoader(shuffle=True, num_workers=4) + per-batch
 * H2D of /root/reference/core/train/train_pigan.py:114-121, 351-357.
 *   pigan_gather_rows       dst[i, :] = src[index[i], :] for i < count; rows are row_bytes wide (any array:
 *                           fp16 operand rows of 512 B, spectra of 1000 B, metrics of 32 B ...), index int64 on the
 *                           device.  Indices outside [0, n_rows) write zeros and set *out_of_range (device int32,
 *                           optional) instead of reading out of bounds.
 *   pigan_generate_spectra  the spectrum part of generate_single_terahertz_spectrum_and_params,
 *                           /root/reference/core/utils/data_loader.py:62-80, for n rows: two Gaussian dips whose
 *                           centre / depth / width depend on (r1, r2, w, g), tanh baseline, optional linear offset,
 *                           noise_level * z, min(., 0).  params_denorm [n,4] given, or NULL: drawn U(2.2, 2.8)
 *                           here and written to params_out [n,4].  z is counter based (Philox4x32-10 keyed by seed;
 *                           counter = (first_index + row, (col % 32) * 64 + (col / 32) / 4, 1), Box-Muller word
 *                           (col / 32) % 4), so row i of the dataset is the same for any batching or sharding;
 *                           numpy's RNG stream is not reproduced.  noise_dump: NULL or [n, S] receiving z (tests).
 *                           The peak search of :83-110 (scipy.find_peaks) is not part of this call; per-spectrum
 *                           metrics come from pigan_physics_metrics.
 * ---------------------------------------------------------------------------------------------- */
int pigan_gather_rows(const void* src, int64_t n_rows, int32_t row_bytes, const int64_t* index, int64_t count,
                      void* dst, int32_t* out_of_range, void* stream);
int pigan_generate_spectra(const float* params_denorm, float* params_out, const float* frequency, int64_t n,
                           int32_t spectrum_dim, float noise_level, uint64_t seed, int64_t first_index,
                           int32_t apply_offset, float* out_spectrum, float* noise_dump, void* stream);

/* k smallest of scores[n] (k <= 4096, n < 2^32), ascending, ties by position; NaN sorts last.
 * out_indices[i] = in_indices[pos] when in_indices is given (merging gathered shard results), else
 * index_base + pos.  workspace: pigan_topk_workspace_bytes(n, k) bytes of device memory, 16-byte aligned. */
size_t pigan_topk_workspace_bytes(int64_t n, int32_t k);
int pigan_topk_smallest(const float* scores, const int64_t* in_indices, int64_t n, int32_t k, int64_t index_base,
                        float* out_scores, int64_t* out_indices, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Profiling aid of the Linear+LayerNorm GEMM epilogue: a device buffer [4][64][5] of int64 that receives clock64
 * stamps of CTA 0 (unit start, after pass 1, after the partial exchange, after pass 2) for the four hidden layers of the
 * surrogate; NULL switches the trace off (tools/ln_trace.py).  Test hooks of the GEMM core live in
 * include/pigan_b200_debug.h and in a separate library (lib/libpigan_b200_test.so). */
int pigan_engine_trace_layernorm(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* PIGAN_B200_H */
