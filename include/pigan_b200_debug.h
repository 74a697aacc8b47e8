/* Test hooks of the tcgen05 GEMM core of libpigan_b200.  NOT part of the product: these entry points are built into a
 * separate library, pi-gan-thz_b200/lib/libpigan_b200_test.so (csrc/debug_gemm.cu + the host utilities), which only
 * tests/ and tools/ load (pi-gan-thz_b200/native_test.py). */
#ifndef PIGAN_B200_DEBUG_H
#define PIGAN_B200_DEBUG_H
#include "pigan_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

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
/* on != 0: pigan_debug_linear uses the streamed-operand kernel even where production keeps the weights resident */
int pigan_debug_force_streamed(int32_t on);
int pigan_debug_linear(const void* a, const void* a_tail, const void* b, const float* bias, void* out_f16,
                       float* rowstats, int32_t m, int32_t n, int32_t k, int32_t leaky, void* stream);
/* same as pigan_debug_linear through the two-CTA (cta_group::2) kernel */
int pigan_debug_linear2(const void* a, const void* a_tail, const void* b, const float* bias, void* out_f16,
                        float* rowstats, int32_t m, int32_t n, int32_t k, int32_t leaky, void* stream);
int pigan_debug_gemm_nt(const void* a, const void* b, const void* b_tail, float* c, int32_t kd, int32_t m,
                        int32_t n, int32_t k_splits, int32_t b_wrap_rows, int32_t tail_from_row,
                        int32_t n_valid, int32_t bias_col, float* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIGAN_B200_DEBUG_H */
