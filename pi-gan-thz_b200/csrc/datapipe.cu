// On-device data pipeline: the pieces that replace DataLoader(num_workers=4) + per-batch H2D
// (core/train/train_pigan.py:114-121, 351-357) once the dataset lives in HBM —
//   pigan_gather_rows       a shuffled batch = rows of the resident arrays picked by a device index vector
//   pigan_generate_spectra  the synthetic spectrum formula of generate_single_terahertz_spectrum_and_params
//                           (core/utils/data_loader.py:62-80) batched, with counter-based noise
// Both are HBM-bound streaming kernels (bytes written once, read once).
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pigan_b200.h"
#include "host_util.h"

namespace pigan {
namespace {

template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ src, const long long* __restrict__ index,
                                                          long long count, int vec_per_row, long long n_rows,
                                                          V* __restrict__ dst, int* __restrict__ bad) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long total = count * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int v = (int)(i - r * vec_per_row);
    const long long s = __ldg(index + r);
    if (s < 0 || s >= n_rows) {   // never read out of bounds: flag it, leave zeros
      if (bad) *bad = 1;
      dst[i] = V{};
      continue;
    }
    dst[i] = src[s * vec_per_row + v];
  }
}

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
__device__ __forceinline__ void normal4(uint4 u, float* z) {   // Box-Muller, as the candidate-noise generator
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)u.x + 0.5f) * k, u1 = (float)u.y * k;
  const float u2 = ((float)u.z + 0.5f) * k, u3 = (float)u.w * k;
  const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

// A warp per spectrum; lane l owns columns l, l + 32, l + 64, ...  (coalesced 128-byte stores).
// Noise z(row, col): Philox4x32-10 keyed by seed, counter = (global row, (col % 32) * 64 + (col / 32) / 4, 1),
// word (col / 32) % 4 of the Box-Muller quadruple.  Structure parameters, when drawn here: counter
// (global row, 0, 2), word e -> 2.2 + 0.6 * (u + 0.5) / 2^32.
template <int NQ>   // NQ = ceil(S / 128): a lane owns up to 4 * NQ columns
__global__ void __launch_bounds__(256) generate_spectra_kernel(const float* __restrict__ params_in,
                                                               float* __restrict__ params_out,
                                                               const float* __restrict__ freq, long long n, int S,
                                                               float noise_level, unsigned long long seed,
                                                               long long first, int apply_offset,
                                                               float* __restrict__ out, float* __restrict__ noise_out) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  // the baseline terms depend on the column only: once per lane for its (up to 16) columns
  constexpr int kMaxCols = 4 * NQ;
  float fq[kMaxCols], base[kMaxCols];
#pragma unroll
  for (int k = 0; k < kMaxCols; ++k) {
    const int col = k * 32 + lane;
    fq[k] = col < S ? __ldg(freq + col) : 0.f;
    base[k] = -0.5f * (tanhf((fq[k] - 1.5f) * 2.0f) + 1.0f);                // data_loader.py:74
    if (apply_offset) base[k] += -0.5f + 0.5f * (fq[k] / 3.0f);             // :75-77
  }
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += wstride) {
    const long long gi = first + row;
    float4 pr;
    if (params_in) {
      pr = __ldg(reinterpret_cast<const float4*>(params_in) + row);
    } else {
      const uint4 u = philox4x32_10(make_uint4((unsigned int)gi, (unsigned int)(gi >> 32), 0u, 2u), key);
      const float k = 2.3283064365386963e-10f;
      pr = make_float4(2.2f + 0.6f * (((float)u.x + 0.5f) * k), 2.2f + 0.6f * (((float)u.y + 0.5f) * k),
                       2.2f + 0.6f * (((float)u.z + 0.5f) * k), 2.2f + 0.6f * (((float)u.w + 0.5f) * k));
    }
    if (params_out && lane == 0) *reinterpret_cast<float4*>(params_out + row * 4) = pr;
    const float r1 = pr.x - 2.5f, r2 = pr.y - 2.5f, w = pr.z - 2.5f, g = pr.w - 2.5f;
    const float c1 = 0.870f + r1 * 0.05f + w * 0.03f;          // data_loader.py:64
    const float m1 = -12.657f + r2 * 1.5f - g * 1.0f;           // :65
    const float w1 = 0.08f + fabsf(r1 * 0.02f);                 // :66
    const float c2 = 2.115f + r2 * 0.07f + g * 0.04f;           // :69
    const float m2 = -11.763f + r1 * 1.0f - w * 0.8f;           // :70
    const float w2 = 0.15f + fabsf(r2 * 0.03f);                 // :71
    const float i1 = 1.0f / (2.0f * w1 * w1), i2 = 1.0f / (2.0f * w2 * w2);
#pragma unroll
    for (int q = 0; q < kMaxCols / 4; ++q) {
      if (q * 128 >= S) break;
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (noise_level != 0.f || noise_out)
        normal4(philox4x32_10(make_uint4((unsigned int)gi, (unsigned int)(gi >> 32), (unsigned int)(lane * 64 + q), 1u),
                              key), z);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = (q * 4 + i) * 32 + lane;
        if (col >= S) break;
        const float f = fq[q * 4 + i];
        float t = m1 * expf(-((f - c1) * (f - c1)) * i1);                   // :67
        t += m2 * expf(-((f - c2) * (f - c2)) * i2);                        // :72-73
        t += base[q * 4 + i];                                               // :74-77
        t += noise_level * z[i];                                            // :78-79
        out[row * S + col] = fminf(t, 0.f);                                 // :80
        if (noise_out) noise_out[row * S + col] = z[i];
      }
    }
  }
}

template <typename V>
void launch_gather(const void* src, const int64_t* index, int64_t count, int row_bytes, int64_t n_rows, void* dst,
                   int* bad, cudaStream_t st) {
  const int vpr = row_bytes / (int)sizeof(V);
  long long total = (long long)count * vpr;
  long long blocks = (total + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  launch_k(gather_rows_kernel<V>, (int)blocks, 256, 0, st, static_cast<const V*>(src),
           reinterpret_cast<const long long*>(index), (long long)count, vpr, (long long)n_rows, static_cast<V*>(dst), bad);
}

}  // namespace
}  // namespace pigan

using namespace pigan;

extern "C" int pigan_gather_rows(const void* src, int64_t n_rows, int32_t row_bytes, const int64_t* index,
                                 int64_t count, void* dst, int32_t* out_of_range, void* stream) {
  PIGAN_CHECK_ARG(src && index && dst && n_rows >= 1 && count >= 0 && row_bytes >= 1);
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  if (count == 0) return PIGAN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)row_bytes;
  if ((al & 15u) == 0) launch_gather<uint4>(src, index, count, row_bytes, n_rows, dst, out_of_range, st);
  else if ((al & 7u) == 0) launch_gather<uint2>(src, index, count, row_bytes, n_rows, dst, out_of_range, st);
  else if ((al & 3u) == 0) launch_gather<uint32_t>(src, index, count, row_bytes, n_rows, dst, out_of_range, st);
  else if ((al & 1u) == 0) launch_gather<uint16_t>(src, index, count, row_bytes, n_rows, dst, out_of_range, st);
  else launch_gather<uint8_t>(src, index, count, row_bytes, n_rows, dst, out_of_range, st);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

extern "C" int pigan_generate_spectra(const float* params_denorm, float* params_out, const float* frequency, int64_t n,
                                      int32_t spectrum_dim, float noise_level, uint64_t seed, int64_t first_index,
                                      int32_t apply_offset, float* out_spectrum, float* noise_dump, void* stream) {
  PIGAN_CHECK_ARG(frequency && out_spectrum && n >= 1 && spectrum_dim >= 1 && spectrum_dim <= 512 && first_index >= 0);
  PIGAN_CHECK_ARG(params_denorm == nullptr || (reinterpret_cast<uintptr_t>(params_denorm) & 15u) == 0);
  PIGAN_CHECK_ARG(params_out == nullptr || (reinterpret_cast<uintptr_t>(params_out) & 15u) == 0);
  if (sm_count() <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  long long blocks = (n + 8 * 4 - 1) / (8 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  auto go = [&](auto kern) {
    launch_k(kern, (int)blocks, 256, 0, static_cast<cudaStream_t>(stream), params_denorm, params_out, frequency,
             (long long)n, (int)spectrum_dim, noise_level, (unsigned long long)seed, (long long)first_index,
             (int)apply_offset, out_spectrum, noise_dump);
  };
  if (spectrum_dim <= 128) go(generate_spectra_kernel<1>);
  else if (spectrum_dim <= 256) go(generate_spectra_kernel<2>);
  else go(generate_spectra_kernel<4>);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
