// Microbenchmark: TMEM -> register drain rate of tcgen05.ld.32x32b as a function of
//   * the load shape (.x16 / .x32 / .x64 / .x128 = registers per thread per instruction),
//   * the number of loads issued back to back before ONE tcgen05.wait::ld (BATCH = loads in flight per warp),
//   * the number of warps per SM sub-partition (4 / 8 / 16 warps per CTA, one CTA per SM).
// Round 1's ldtm_bench.cu only had one load in flight per warp (load, wait, use), which measures
// latency x loads-in-flight, not bandwidth; this one separates the two.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench2 ldtm_bench2.cu && ./ldtm_bench2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pi-gan-thz_b200/csrc/ptx.cuh"
#include "ldtm_shapes.cuh"
using namespace pigan;

template <int W>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t* r) {
  if constexpr (W == 16) ldtm_x16(taddr, r);
  else if constexpr (W == 32) ldtm_x32(taddr, r);
  else if constexpr (W == 64) ldtm_x64(taddr, r);
  else ldtm_x128(taddr, r);
}

// USE: 0 = fold every register with a 3-input add (16 IADD3 per 32 registers), 1 = epilogue-like math per element
// (bias add, LeakyReLU, fp16 pack, 16-byte shared-memory stores into a swizzled staging tile)
template <int W, int BATCH, int USE>
__global__ void __launch_bounds__((W * BATCH >= 128 ? 8 : 16) * 32, 1) k(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  extern __shared__ __align__(16) uint4 stage[];  // 64 KB (dynamic)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  const int q = warp & 3;
  const uint32_t tacc = base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * 256);
  uint32_t acc = 0;
  float facc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int c = 0; c < 256; c += W * BATCH) {
      uint32_t r[BATCH][W];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) ldtm<W>(tacc + c + b * W, r[b]);
      tmem_ld_wait();
      if constexpr (USE == 0) {
#pragma unroll
        for (int b = 0; b < BATCH; ++b)
#pragma unroll
          for (int i = 0; i < W; i += 2) acc += r[b][i] + r[b][i + 1];
      } else {
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
#pragma unroll
          for (int i = 0; i < W; i += 8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = __uint_as_float(r[b][i + j]) + 0.25f;
              y[j] = fmaxf(x, 0.2f * x);
            }
            uint4 u;
            __half2 h;
            h = __floats2half2_rn(y[0], y[1]); u.x = *(uint32_t*)&h;
            h = __floats2half2_rn(y[2], y[3]); u.y = *(uint32_t*)&h;
            h = __floats2half2_rn(y[4], y[5]); u.z = *(uint32_t*)&h;
            h = __floats2half2_rn(y[6], y[7]); u.w = *(uint32_t*)&h;
            const int chunk = ((c + b * W + i) >> 3) & 7;
            stage[(threadIdx.x & 511) * 8 + (chunk ^ (threadIdx.x & 7))] = u;
          }
        }
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + stage[threadIdx.x].x + (uint32_t)facc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

template <int W, int BATCH, int USE>
void run() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
  for (int warps = 4; warps <= (W * BATCH >= 128 ? 8 : 16); warps *= 2) {
    const int iters = 200;
    cudaFuncSetAttribute(k<W, BATCH, USE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<W, BATCH, USE><<<148, warps * 32, 65536>>>(out, cyc, 5);  // warm-up
    k<W, BATCH, USE><<<148, warps * 32, 65536>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    // every warp reads 32 lanes x 256 columns x 4 B = 32 KB per iteration
    const double bytes = (double)warps * 32768.0 * iters;
    printf("x%-3d batch %d use %d warps %2d: %7.0f cycles per warp pass over 32x256 fp32; %6.1f B/clk per SM; "
           "%6.0f cycles per 128x256 tile\n", W, BATCH, USE, warps, (double)mx / iters, bytes / (double)mx,
           131072.0 / (bytes / (double)mx));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  printf("# tcgen05.ld.32x32b drain rate, one CTA per SM on 148 SMs (CTA 0 reported)\n");
  run<16, 1, 0>(); run<16, 2, 0>(); run<16, 4, 0>(); run<16, 8, 0>();
  run<32, 1, 0>(); run<32, 2, 0>(); run<32, 4, 0>();
  run<64, 1, 0>(); run<64, 2, 0>();
  run<128, 1, 0>();
  printf("# with epilogue-like math (bias, LeakyReLU, fp16 pack, STS.128)\n");
  run<16, 1, 1>(); run<32, 1, 1>(); run<32, 2, 1>(); run<64, 1, 1>(); run<64, 2, 1>(); run<128, 1, 1>();
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
