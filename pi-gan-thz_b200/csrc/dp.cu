// Data-parallel exchange over NVLink peer memory: one-shot all-reduce kernels for the batch-coupled buffers of the
// PI-GAN step (BatchNorm forward/backward sums, loss numerators, gradients).  The reference has no distributed code
// (SURVEY 2.2); under data parallelism the step needs 8 small all-reduces per iteration, all of them
// latency-bound: 3-4 KB for the BatchNorm sums, 1 MB for each network's gradients.  Instead of a library call per
// reduction every rank maps every other rank's exchange region (cudaIpc, NVLink P2P through NVSwitch) and
//   small buffers: PUSHES its values into a slot of its own in every peer's region as 8-byte (word, epoch) pairs - the
//                  pair is one store, so the epoch doubles as the "data valid" flag and no fence or separate flag is
//                  needed (the LL idea of NCCL) - then polls its local slots until every peer's pairs carry this
//                  epoch and sums them in rank order (one CTA, one launch, result in place).  One NVLink one-way
//                  latency per exchange instead of fence + flag + remote read round trip (PIGAN_DP_PULL=1 keeps the
//                  pull variant: copy into the own slot, flag, read the peers' slots);
//   gradients    : live in the exchange region already (the weight-gradient kernels write there); each CTA waits for
//                  the peers' flags, then sums its chunk straight out of the peers' memory into the local buffer
//                  Adam reads, and accumulates the squared norm clip_grad_norm_ needs on the way.
// Sums run in rank order on every rank, so the replicas stay bit-identical.  Spins are bounded: a protocol bug traps.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "host_util.h"

namespace pigan {
namespace {

constexpr int kMaxWorld = 16;
constexpr int kChannels = 16;            // exchange points per step
constexpr int kSmallWords = 2048;                           // 32-bit words per small exchange (8 KB of payload)
constexpr int kSmallSrcBytes = kSmallWords * 8;             // one source rank's (word, epoch) pairs
constexpr int kSmallSlotBytes = kSmallSrcBytes * kMaxWorld; // per channel and parity: [source rank][word]
constexpr size_t kFlagBytes = 4096;      // [kChannels][kMaxWorld] uint32
constexpr size_t kSmallBytes = (size_t)kChannels * 2 * kSmallSlotBytes;

struct PeerTable {
  uint8_t* base[kMaxWorld];
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {   // never from a stale L1 line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// flags[channel][src_rank] in the region of the rank that waits
__device__ __forceinline__ uint32_t* flag_ptr(uint8_t* base, int channel, int src) {
  return reinterpret_cast<uint32_t*>(base) + channel * kMaxWorld + src;
}
__device__ __forceinline__ void publish(const PeerTable& t, int channel, uint32_t epoch) {
  __threadfence_system();
  for (int p = 0; p < t.world; ++p) st_release_sys(flag_ptr(t.base[p], channel, t.rank), epoch);
}
// epochs only grow; a peer may already be one exchange ahead on this channel's next use, hence >=
__device__ __forceinline__ void wait_peer(const PeerTable& t, int channel, int src, uint32_t epoch) {
  const uint32_t* f = flag_ptr(t.base[t.rank], channel, src);
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
    if ((++spins & 0xFFu) == 0) {
      const uint64_t now = timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 5000000000ull) {  // 5 s: a rank died or the schedule diverged
        printf("pigan dp: rank %d waited 5 s for rank %d on channel %d epoch %u\n", t.rank, src, channel, epoch);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void st_pair_sys(uint2* p, uint32_t w, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(w), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 ld_pair_sys(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
// Push variant.  Slot of (channel, parity) in every region: [source rank][word] pairs.  A slot is rewritten two
// epochs later, which a peer can only reach after it has consumed this epoch's exchange of the other parity, i.e.
// after this rank finished reading.  exchange_one: element i of a T buffer occupies words [w0, w0 + sizeof(T)/4).
template <typename T>
__device__ __forceinline__ void exchange_one(const PeerTable& t, T* __restrict__ buf, int i, int w0, int channel,
                                             uint32_t epoch, size_t slot_off) {
  constexpr int WORDS = sizeof(T) / 4;
  const T mine = buf[i];
  uint32_t w[WORDS];
  memcpy(w, &mine, sizeof(T));
  for (int p = 0; p < t.world; ++p) {
    if (p == t.rank) continue;
    uint2* dst = reinterpret_cast<uint2*>(t.base[p] + slot_off + (size_t)t.rank * kSmallSrcBytes) + w0;
#pragma unroll
    for (int k = 0; k < WORDS; ++k) st_pair_sys(dst + k, w[k], epoch);
  }
  T s = 0;
  for (int r = 0; r < t.world; ++r) {
    T v = mine;
    if (r != t.rank) {
      const uint2* src = reinterpret_cast<const uint2*>(t.base[t.rank] + slot_off + (size_t)r * kSmallSrcBytes) + w0;
      uint32_t g[WORDS];
#pragma unroll
      for (int k = 0; k < WORDS; ++k) {
        uint2 pr = ld_pair_sys(src + k);
        uint64_t t0 = 0;
        uint32_t spins = 0;
        while (pr.y != epoch) {
          if ((++spins & 0xFFFu) == 0) {
            const uint64_t now = timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 5000000000ull) {  // 5 s: a rank died or the schedule diverged
              printf("pigan dp: rank %d waited 5 s for rank %d on channel %d epoch %u (word %d)\n", t.rank, r, channel,
                     epoch, w0 + k);
              __trap();
            }
          }
          pr = ld_pair_sys(src + k);
        }
        g[k] = pr.x;
      }
      memcpy(&v, g, sizeof(T));
    }
    s += v;
  }
  buf[i] = s;
}
// a[0:na] (fp32) and b[0:nb] (fp64) in ONE exchange (either may be empty): words [0, na) | [na, na + 2 nb)
__global__ void __launch_bounds__(1024) allreduce_small_push_kernel(PeerTable t, float* __restrict__ a, int na,
                                                                    double* __restrict__ b, int nb, int channel,
                                                                    uint32_t epoch, size_t slot_off) {
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent (launch_k)
  for (int i = threadIdx.x; i < na + nb; i += blockDim.x) {
    if (i < na) exchange_one<float>(t, a, i, i, channel, epoch, slot_off);
    else exchange_one<double>(t, b, i - na, na + 2 * (i - na), channel, epoch, slot_off);
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) allreduce_small_kernel(PeerTable t, T* __restrict__ buf, int n, int channel,
                                                               uint32_t epoch, size_t slot_off) {
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent (launch_k)
  T* mine = reinterpret_cast<T*>(t.base[t.rank] + slot_off);
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = buf[i];
  __syncthreads();
  if (threadIdx.x == 0) publish(t, channel, epoch);
  if (threadIdx.x < t.world) wait_peer(t, channel, threadIdx.x, epoch);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    T s = 0;
    for (int r = 0; r < t.world; ++r) s += reinterpret_cast<const volatile T*>(t.base[r] + slot_off)[i];
    buf[i] = s;
  }
}

// src_off: offset of this exchange's gradient slot inside every rank's region; dst: local reduced gradients
__global__ void __launch_bounds__(256) allreduce_grads_kernel(PeerTable t, size_t src_off, float* __restrict__ dst,
                                                              long long n, int channel, uint32_t epoch,
                                                              double* __restrict__ sumsq) {
  __shared__ float red[8];
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent (launch_k)
  if (blockIdx.x == 0 && threadIdx.x == 0) publish(t, channel, epoch);   // earlier kernels of the stream are done
  if (threadIdx.x < t.world) wait_peer(t, channel, threadIdx.x, epoch);
  __syncthreads();
  float acc = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < t.world; ++r) {
      const float4 v = ld_peer_f4(reinterpret_cast<const float4*>(t.base[r] + src_off) + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(dst)[i] = s;
    acc = fmaf(s.x, s.x, fmaf(s.y, s.y, fmaf(s.z, s.z, fmaf(s.w, s.w, acc))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float s = 0.f;
    for (int r = 0; r < t.world; ++r) s += reinterpret_cast<const volatile float*>(t.base[r] + src_off)[i];
    dst[i] = s;
    acc = fmaf(s, s, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && sumsq != nullptr) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    atomicAdd(sumsq, (double)s);
  }
}

}  // namespace
}  // namespace pigan

using namespace pigan;

struct PiganDp {
  PeerTable t;
  size_t region_bytes;
  size_t grad_floats;   // capacity of one gradient slot
};

// Layout of a rank's exchange region: [flags 4 KB][small slots: channel x parity x 16 KB][gradient slots: 4 x cap]
extern "C" size_t pigan_dp_region_bytes(int64_t max_grad_floats) {
  const size_t cap = ((size_t)max_grad_floats * sizeof(float) + 255) & ~size_t(255);
  return kFlagBytes + kSmallBytes + 4 * cap;
}
extern "C" size_t pigan_dp_grad_slot_offset(int64_t max_grad_floats, int32_t net, int32_t parity) {
  const size_t cap = ((size_t)max_grad_floats * sizeof(float) + 255) & ~size_t(255);
  return kFlagBytes + kSmallBytes + (size_t)(net * 2 + parity) * cap;
}

extern "C" int pigan_dp_alloc(size_t bytes, void** out) {
  PIGAN_CHECK_ARG(out != nullptr && bytes > 0);
  PIGAN_CUDA_OK(cudaMalloc(out, bytes));   // a whole allocation of its own: cudaIpc exports allocations, not views
  PIGAN_CUDA_OK(cudaMemset(*out, 0, bytes));
  PIGAN_CUDA_OK(cudaDeviceSynchronize());
  return PIGAN_OK;
}
extern "C" int pigan_dp_free(void* p) {
  PIGAN_CUDA_OK(cudaFree(p));
  return PIGAN_OK;
}
extern "C" int pigan_dp_ipc_export(void* p, void* handle64) {
  PIGAN_CHECK_ARG(p && handle64);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  PIGAN_CUDA_OK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), p));
  return PIGAN_OK;
}
extern "C" int pigan_dp_ipc_open(const void* handle64, void** out) {
  PIGAN_CHECK_ARG(handle64 && out);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  PIGAN_CUDA_OK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return PIGAN_OK;
}
extern "C" int pigan_dp_ipc_close(void* p) {
  PIGAN_CUDA_OK(cudaIpcCloseMemHandle(p));
  return PIGAN_OK;
}

extern "C" int pigan_dp_create(PiganDp** out, int32_t world, int32_t rank, void* const* region_of_rank,
                               int64_t max_grad_floats) {
  PIGAN_CHECK_ARG(out && region_of_rank && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world);
  PiganDp* d = new (std::nothrow) PiganDp();
  if (!d) return fail(PIGAN_ERR_INVALID, "out of host memory");
  d->t.world = world;
  d->t.rank = rank;
  for (int r = 0; r < kMaxWorld; ++r) d->t.base[r] = r < world ? static_cast<uint8_t*>(region_of_rank[r]) : nullptr;
  d->grad_floats = (size_t)max_grad_floats;
  d->region_bytes = pigan_dp_region_bytes(max_grad_floats);
  *out = d;
  return PIGAN_OK;
}
extern "C" int pigan_dp_destroy(PiganDp* d) {
  delete d;
  return PIGAN_OK;
}

// In-place sum over ranks of buf[0:n] (fp32: is_double = 0, fp64: 1); n * elem <= 16 KB.  `channel` names the
// exchange point inside a step (all ranks use the same one), `epoch` the step (strictly increasing per channel).
extern "C" int pigan_dp_allreduce_small(PiganDp* d, void* buf, int32_t n, int32_t is_double, int32_t channel,
                                        uint32_t epoch, void* stream) {
  PIGAN_CHECK_ARG(d && buf && n >= 1 && channel >= 0 && channel < kChannels && epoch >= 1);
  PIGAN_CHECK_ARG((size_t)n * (is_double ? 2 : 1) <= (size_t)kSmallWords);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t slot = kFlagBytes + ((size_t)channel * 2 + (epoch & 1u)) * kSmallSlotBytes;
  static const bool pull = [] { const char* v = getenv("PIGAN_DP_PULL"); return v && v[0] == '1'; }();
  if (!pull) {
    if (is_double)
      launch_k(allreduce_small_push_kernel, 1, 1024, 0, st, d->t, static_cast<float*>(nullptr), 0,
               static_cast<double*>(buf), n, channel, epoch, slot);
    else
      launch_k(allreduce_small_push_kernel, 1, 1024, 0, st, d->t, static_cast<float*>(buf), n,
               static_cast<double*>(nullptr), 0, channel, epoch, slot);
    PIGAN_CUDA_OK(cudaGetLastError());
    return PIGAN_OK;
  }
  if (is_double)
    launch_k(allreduce_small_kernel<double>, 1, 1024, 0, st, d->t, static_cast<double*>(buf), n, channel, epoch, slot);
  else
    launch_k(allreduce_small_kernel<float>, 1, 1024, 0, st, d->t, static_cast<float*>(buf), n, channel, epoch, slot);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

// One exchange for an fp32 buffer and an fp64 buffer together (BatchNorm backward sums + loss numerators): in-place
// sums over ranks of a[0:na] and b[0:nb]; na + 2 nb <= 2048 words.  Push variant only.
extern "C" int pigan_dp_allreduce_small2(PiganDp* d, float* a, int32_t na, double* b, int32_t nb, int32_t channel,
                                         uint32_t epoch, void* stream) {
  PIGAN_CHECK_ARG(d && na >= 0 && nb >= 0 && na + nb >= 1 && (na == 0 || a) && (nb == 0 || b));
  PIGAN_CHECK_ARG(channel >= 0 && channel < kChannels && epoch >= 1 && (size_t)na + 2 * (size_t)nb <= (size_t)kSmallWords);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t slot = kFlagBytes + ((size_t)channel * 2 + (epoch & 1u)) * kSmallSlotBytes;
  launch_k(allreduce_small_push_kernel, 1, 1024, 0, st, d->t, a, (int)na, b, (int)nb, channel, epoch, slot);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}

// dst[0:n] = sum over ranks of the gradient slot (net, epoch & 1) of every rank's region; *sumsq += |dst|^2.
extern "C" int pigan_dp_allreduce_grads(PiganDp* d, int32_t net, float* dst, int64_t n, int32_t channel,
                                        uint32_t epoch, double* sumsq, void* stream) {
  PIGAN_CHECK_ARG(d && dst && n >= 1 && (size_t)n <= d->grad_floats && (net == 0 || net == 1));
  PIGAN_CHECK_ARG(channel >= 0 && channel < kChannels && epoch >= 1);
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t off = pigan_dp_grad_slot_offset((int64_t)d->grad_floats, net, (int32_t)(epoch & 1u));
  int grid = (int)((n / 4 + 255) / 256);
  if (grid > sm_count()) grid = sm_count();
  if (grid < 1) grid = 1;
  launch_k(allreduce_grads_kernel, grid, 256, 0, st, d->t, off, dst, (long long)n, channel, epoch, sumsq);
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
