// Microbenchmark: how fast can epilogue warps drain TMEM?  (tools/micro, not part of the library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench ldtm_bench.cu && ./ldtm_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pi-gan-thz_b200/csrc/ptx.cuh"
using namespace pigan;

// mode 0: LDTM.x32 + wait only; 1: + 3 flops/elem; 2: + 6 flops/elem; 3: 6 flops, next load issued before math (2 reg sets)
// 4: x16 loads, 6 flops; 5: 6 flops + cvt/pack + STS
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cyc, int groups, int iters) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) uint4 stage[2048];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < 4 * groups) {
    const int q = warp & 3;
    const uint32_t tacc = base + ((uint32_t)(q * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 3) {
        float a[32], b[32];
        tmem_ld32(tacc, a);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 256; c += 64) {
          tmem_ld32(tacc + c + 32, b);
#pragma unroll
          for (int i = 0; i < 32; ++i) { float x = a[i] + 1.f; x = (x - acc) * 1.5f; x = fmaf(x, 2.f, 3.f); acc += fmaxf(x, 0.2f * x); }
          tmem_ld_wait();
          if (c + 64 < 256) tmem_ld32(tacc + c + 64, a);
#pragma unroll
          for (int i = 0; i < 32; ++i) { float x = b[i] + 1.f; x = (x - acc) * 1.5f; x = fmaf(x, 2.f, 3.f); acc += fmaxf(x, 0.2f * x); }
          tmem_ld_wait();
        }
      } else if (MODE == 4) {
#pragma unroll 1
        for (int c = 0; c < 256; c += 16) {
          float v[16];
          tmem_ld16(tacc + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) { float x = v[i] + 1.f; x = (x - 0.5f) * 1.5f; x = fmaf(x, 2.f, 3.f); acc += fmaxf(x, 0.2f * x); }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
          float v[32];
          tmem_ld32(tacc + c, v);
          tmem_ld_wait();
          if (MODE == 0) { acc += v[lane]; }
          if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { float x = v[i] + 1.f; acc += x; acc = fmaf(x, x, acc); }
          }
          if (MODE == 2 || MODE == 5) {
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { float x = v[i] + 1.f; x = (x - 0.5f) * 1.5f; x = fmaf(x, 2.f, 3.f); y[i] = fmaxf(x, 0.2f * x); }
            if (MODE == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) acc += y[i];
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u;
                __half2 h;
                h = __floats2half2_rn(y[8*i], y[8*i+1]); u.x = *(uint32_t*)&h;
                h = __floats2half2_rn(y[8*i+2], y[8*i+3]); u.y = *(uint32_t*)&h;
                h = __floats2half2_rn(y[8*i+4], y[8*i+5]); u.z = *(uint32_t*)&h;
                h = __floats2half2_rn(y[8*i+6], y[8*i+7]); u.w = *(uint32_t*)&h;
                stage[(threadIdx.x & 127) * 8 + ((i + c / 32 * 4) & 7 ^ (threadIdx.x & 7))] = u;
              }
            }
          }
        }
      }
    }
    t1 = clock64();
  }
  if (lane == 0 && warp < 4 * groups) cyc[blockIdx.x * 8 + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)stage[threadIdx.x].x;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
  for (int groups = 1; groups <= 2; ++groups) {
    const int iters = 64;
    k<MODE><<<148, 256>>>(out, cyc, groups, iters);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s groups %d: %6.0f cycles per 128x256 tile pass (%.1f per column; %.1f B/clk per SM)\n", name, groups,
           (double)h[0] / iters, (double)h[0] / iters / 256, groups * 131072.0 * iters / (double)h[0]);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("0: LDTM.x32 + wait");
  run<1>("1: + 3 flops/elem (serial acc)");
  run<2>("2: + 6 flops/elem");
  run<3>("3: 6 flops, loads overlapped with math");
  run<4>("4: LDTM.x16, 6 flops");
  run<5>("5: 6 flops + pack + STS.128");
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
