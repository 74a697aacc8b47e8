// Host-side helpers shared by the C-ABI translation units: error reporting without exceptions
// across the ABI, cached device properties, TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pigan_b200.h"

namespace pigan {

void set_last_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define PIGAN_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::pigan::fail(PIGAN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define PIGAN_CHECK_ARG(cond)                                                            \
  do {                                                                                   \
    if (!(cond))                                                                         \
      return ::pigan::fail(PIGAN_ERR_INVALID, "invalid argument: %s (%s:%d)", #cond,     \
                           __FILE__, __LINE__);                                          \
  } while (0)

#define PIGAN_TRY(expr)                                                                  \
  do {                                                                                   \
    int _s = (expr);                                                                     \
    if (_s != PIGAN_OK) return _s;                                                       \
  } while (0)

void note_launch();          // every kernel launch of the library passes through here (pigan_launch_count)
bool debug_sync_enabled();  // PIGAN_DEBUG_SYNC=1: synchronise after every GEMM launch and name the kernel that failed
bool pdl_enabled();         // programmatic dependent launch between consecutive kernels (PIGAN_PDL=0 disables)
int sm_count();  // multiprocessors of the current device (cached per device)

// Row-major 2-D fp16 tensor [outer, inner] with row pitch ld_elems; box = [box_outer, box_inner],
// 128B swizzle (box_inner * 2 bytes must be <= 128), zero fill out of bounds.
int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer);

// Row-major fp16 [rows, cols] as the TMA-store target of an epilogue warp: box = 32 rows x 32 columns, 64B swizzle.
int make_tmap_f16_store(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems);

// Row-major 2-D fp32 tensor, box = [box_outer, 32 floats] (128 B rows), 128B swizzle: TMA-store target of fp32 tiles.
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                     uint32_t box_outer);

// Launches `kern` as a programmatic dependent of the previous kernel on the stream (when enabled): its blocks may be
// scheduled while that kernel drains.  Every kernel launched this way starts with griddepcontrol.wait (pdl_wait()).
template <class... Exp, class... Act>
inline void launch_k(void (*kern)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  note_launch();
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<Act&&>(args)...);   // errors surface at the caller's cudaGetLastError
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pigan
