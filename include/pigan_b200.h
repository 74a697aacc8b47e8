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

#define PIGAN_ABI_VERSION 1

#define PIGAN_OK 0
#define PIGAN_ERR_INVALID (-1)     /* bad argument (null pointer, size, unsupported dimension) */
#define PIGAN_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define PIGAN_ERR_UNSUPPORTED (-3) /* shape outside what this build implements */
#define PIGAN_ERR_WORKSPACE (-4)   /* workspace too small */

int pigan_abi_version(void);
/* Thread-local, valid until the next failing call on the same thread. */
const char* pigan_last_error(void);

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

/* ------------------------------------------------------------------------------------------------
 * Test hooks for the tcgen05 GEMM core (not part of the reference surface; used by tests/ only).
 *   gemm_tn: C[M,N] fp32 = A[M,K] * B[N,K]^T       A,B fp16 row-major, K multiple of 8; `variant` picks
 *            the tile configuration (0: 256x1, 1: 256x2, 2: 144x2, 3: 128x1 accumulators; 10-12: probes
 *            that skip the stores, for pipeline timing)
 *   linear : out[M,N] fp16 = act(A * B^T + bias) through the production store epilogue (TMA-store
 *            staging); a_tail [M,64] optionally replaces the last 64 columns of A; bias (zero-padded to a
 *            multiple of 256) and rowstats [M][ceil(N/256)][2] = per-row (sum, sum of squares) per tile
 *            are optional
 *   gemm_nt: C[M,n_valid] fp32 += A[Kd,M]^T * B[Kd,N], split over k_splits CTAs; rows of B wrap modulo
 *            b_wrap_rows; b_tail [Kd - tail_from_row, 64] optionally replaces B's last 64 columns for
 *            reduction rows >= tail_from_row; column bias_col of the product accumulates into db[M]
 * ---------------------------------------------------------------------------------------------- */
int pigan_debug_gemm_tn(const void* a, const void* b, float* c, int32_t m, int32_t n, int32_t k,
                        int32_t variant, void* stream);
int pigan_debug_linear(const void* a, const void* a_tail, const void* b, const float* bias, void* out_f16,
                       float* rowstats, int32_t m, int32_t n, int32_t k, int32_t leaky, void* stream);
int pigan_debug_gemm_nt(const void* a, const void* b, const void* b_tail, float* c, int32_t kd, int32_t m,
                        int32_t n, int32_t k_splits, int32_t b_wrap_rows, int32_t tail_from_row,
                        int32_t n_valid, int32_t bias_col, float* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIGAN_B200_H */
