#include "host_util.h"
#include <cstdlib>

#include <cudaTypedefs.h>
#include <stdarg.h>

#include <atomic>

namespace pigan {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("PIGAN_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}
bool debug_sync_enabled() {
  static const bool on = [] {
    const char* v = getenv("PIGAN_DEBUG_SYNC");
    return v && v[0] == '1';
  }();
  return on;
}
int sm_count() {
  static int cached_dev = -1;
  static int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  auto fn = encode_fn();
  if (fn == nullptr) return fail(PIGAN_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld_elems * 2) % 16 != 0)
    return fail(PIGAN_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple pitch");
  if (box_inner * 2 > 128 || box_outer > 256)
    return fail(PIGAN_ERR_INVALID, "TMA box out of range");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PIGAN_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return PIGAN_OK;
}

// TMA-store target of the epilogues' per-warp staging: fp16 [rows, cols], box = 32 rows x 32 columns (64-byte rows),
// 64B swizzle (epilogues.cuh: WarpStager)
int make_tmap_f16_store(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems) {
  auto fn = encode_fn();
  if (fn == nullptr) return fail(PIGAN_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld_elems * 2) % 16 != 0)
    return fail(PIGAN_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PIGAN_ERR_CUDA, "cuTensorMapEncodeTiled(store) failed (CUresult %d)", (int)r);
  return PIGAN_OK;
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                     uint32_t box_outer) {
  auto fn = encode_fn();
  if (fn == nullptr) return fail(PIGAN_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld_elems * 4) % 16 != 0 || box_outer > 256)
    return fail(PIGAN_ERR_INVALID, "fp32 TMA operand must be 16-byte aligned with a 16-byte multiple pitch");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 4};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PIGAN_ERR_CUDA, "cuTensorMapEncodeTiled(fp32) failed (CUresult %d)", (int)r);
  return PIGAN_OK;
}

}  // namespace pigan

extern "C" int pigan_abi_version(void) { return PIGAN_ABI_VERSION; }
extern "C" const char* pigan_last_error(void) { return pigan::g_last_error.c_str(); }
namespace pigan { long long launch_count(); }
extern "C" int64_t pigan_launch_count(void) { return pigan::launch_count(); }

#include "layout.h"

extern "C" void pigan_default_dims(PiganDims* d) {
  if (d == nullptr) return;
  d->spectrum_dim = 250;
  d->param_dim = 4;
  d->metrics_dim = 8;
  d->g_hidden[0] = 512; d->g_hidden[1] = 256;
  d->d_hidden[0] = 512; d->d_hidden[1] = 256;
  d->f_hidden[0] = 256; d->f_hidden[1] = 512; d->f_hidden[2] = 1024; d->f_hidden[3] = 512; d->f_hidden[4] = 256;
}
static PiganDims dims_or_default(const PiganDims* d) {
  PiganDims x;
  if (d) x = *d; else pigan_default_dims(&x);
  return x;
}
extern "C" int64_t pigan_generator_param_count(const PiganDims* d) { return pigan::GenLayout(dims_or_default(d)).total; }
extern "C" int64_t pigan_discriminator_param_count(const PiganDims* d) { return pigan::DiscLayout(dims_or_default(d)).total; }
extern "C" int64_t pigan_forward_model_param_count(const PiganDims* d) { return pigan::FwdLayout(dims_or_default(d)).total; }
extern "C" int64_t pigan_generator_bn_buffer_count(const PiganDims* d) { return pigan::GenLayout(dims_or_default(d)).bn_total; }
