// K8 — k smallest scores with their candidate indices, entirely on the device (no host round trip, so the
// scoring loop stays graph-capturable).  The reference has no ranking step: this is the
// "final top-k gather" of the inverse-design path; each rank / chunk runs it on its shard, the k-sized
// results are concatenated (all_gather) and the same kernel merges them.
//
// Radix select over 64-bit keys (order-preserving score bits << 32 | position) — unique keys, so ties
// break by position and the result is deterministic.  8 byte-wise histogram passes over data that is
// L2-resident for the chunk sizes used (<= a few M scores), a compaction of the k winners and one
// single-block bitonic sort.
#include "host_util.h"

namespace pigan {
namespace {

constexpr int kMaxK = 4096;

struct TopkState {
  unsigned long long prefix;  // key bits decided so far (high bytes)
  unsigned int k_rem;         // rank still to locate inside the current prefix bucket
  unsigned int count;         // compaction cursor
};

__device__ __forceinline__ unsigned int score_key(float s) {
  unsigned int b = __float_as_uint(s);
  if (s != s) return 0xFFFFFFFFu;                    // NaN sorts last
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending float order as unsigned order
}
__device__ __forceinline__ unsigned long long full_key(float s, long long i) {
  return ((unsigned long long)score_key(s) << 32) | (unsigned long long)(unsigned int)i;
}

__global__ void topk_init_kernel(TopkState* st, unsigned int* hist, int k) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) {
    st->prefix = 0ull;
    st->k_rem = (unsigned int)k;
    st->count = 0u;
  }
}

// histogram of byte `pass` (0 = most significant) over keys matching the decided prefix
__global__ void topk_hist_kernel(const float* __restrict__ scores, long long n, int pass, const TopkState* st,
                                 unsigned int* hist) {
  __shared__ unsigned int sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned long long prefix = st->prefix;
  const int shift = 56 - 8 * pass;
  const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = full_key(scores[i], i);
    if ((key & mask) == prefix) atomicAdd(&sh[(unsigned int)(key >> shift) & 0xFFu], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// pick the bucket holding the k_rem-th key, extend the prefix, clear the histogram for the next pass
__global__ void topk_pick_kernel(TopkState* st, unsigned int* hist, int pass) {
  if (threadIdx.x == 0) {
    unsigned int k = st->k_rem, acc = 0;
    int b = 0;
    for (; b < 256; ++b) {
      if (acc + hist[b] >= k) break;
      acc += hist[b];
    }
    if (b > 255) b = 255;
    st->k_rem = k - acc;
    st->prefix |= (unsigned long long)b << (56 - 8 * pass);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
}

// after 8 passes prefix == the k-th smallest key: gather every key <= it
__global__ void topk_collect_kernel(const float* __restrict__ scores, long long n, TopkState* st,
                                    unsigned long long* keys, int k) {
  const unsigned long long kth = st->prefix;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = full_key(scores[i], i);
    if (key <= kth) {
      const unsigned int slot = atomicAdd(&st->count, 1u);
      if (slot < (unsigned int)k) keys[slot] = key;
    }
  }
}

__global__ void topk_sort_kernel(const unsigned long long* keys, int k, const float* __restrict__ scores,
                                 const long long* __restrict__ in_indices, long long index_base, float* out_scores,
                                 long long* out_indices) {
  extern __shared__ unsigned long long sk[];
  int n2 = 1;
  while (n2 < k) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) sk[i] = i < k ? keys[i] : ~0ull;
  __syncthreads();
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const unsigned long long a = sk[i], b = sk[j];
          if ((a > b) == up) {
            sk[i] = b;
            sk[j] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const long long pos = (long long)(sk[i] & 0xFFFFFFFFull);
    out_scores[i] = scores[pos];
    out_indices[i] = in_indices ? in_indices[pos] : index_base + pos;
  }
}

}  // namespace
}  // namespace pigan

using namespace pigan;

extern "C" size_t pigan_topk_workspace_bytes(int64_t n, int32_t k) {
  (void)n;
  if (k < 1 || k > kMaxK) return 0;
  return 256 + 256 * sizeof(unsigned int) + (size_t)k * sizeof(unsigned long long) + 256;
}

extern "C" int pigan_topk_smallest(const float* scores, const int64_t* in_indices, int64_t n, int32_t k,
                                   int64_t index_base, float* out_scores, int64_t* out_indices, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  PIGAN_CHECK_ARG(scores && out_scores && out_indices && workspace);
  PIGAN_CHECK_ARG(k >= 1 && k <= kMaxK && n >= k && n < (1ll << 32));
  PIGAN_CHECK_ARG(workspace_bytes >= pigan_topk_workspace_bytes(n, k));
  PIGAN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w = static_cast<uint8_t*>(workspace);
  TopkState* state = reinterpret_cast<TopkState*>(w);
  unsigned int* hist = reinterpret_cast<unsigned int*>(w + 256);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(w + 256 + 256 * sizeof(unsigned int));
  int grid = (int)((n + 1023) / 1024);
  const int cap = sm_count() * 8;
  if (cap <= 0) return fail(PIGAN_ERR_CUDA, "no CUDA device");
  if (grid > cap) grid = cap;
  note_launch(), topk_init_kernel<<<1, 256, 0, st>>>(state, hist, k);
  for (int pass = 0; pass < 8; ++pass) {
    note_launch(), topk_hist_kernel<<<grid, 256, 0, st>>>(scores, (long long)n, pass, state, hist);
    note_launch(), topk_pick_kernel<<<1, 256, 0, st>>>(state, hist, pass);
  }
  note_launch(), topk_collect_kernel<<<grid, 256, 0, st>>>(scores, (long long)n, state, keys, k);
  int n2 = 1;
  while (n2 < k) n2 <<= 1;
  note_launch(), topk_sort_kernel<<<1, 1024, n2 * sizeof(unsigned long long), st>>>(
      keys, k, scores, reinterpret_cast<const long long*>(in_indices), (long long)index_base, out_scores,
      reinterpret_cast<long long*>(out_indices));
  PIGAN_CUDA_OK(cudaGetLastError());
  return PIGAN_OK;
}
