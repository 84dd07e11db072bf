// Full descending order of score vectors — the device form of np.argsort(-scores) and of the
// 1-based rank lookup built from it (Tool/rank_chunks_optimized.py:225-235).  Scores are packed
// with their index into order-preserving 64-bit keys (ties -> lower index first, deterministic,
// where the reference's non-stable sort leaves them unspecified) and sorted with a bitonic
// network: shared-memory stages for strides < 2048, one global pass per larger stride.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

constexpr int kSortTile = 2048;  // keys per CTA in the shared-memory stages (1024 threads x 2)

__global__ void rank_keys_init_kernel(const float* __restrict__ scores, long long n, long long n2, uint64_t* __restrict__ keys) {
  const int q = blockIdx.y;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += static_cast<long long>(gridDim.x) * blockDim.x)
    keys[q * n2 + i] = i < n ? make_key(scores[q * n + i], static_cast<uint32_t>(i)) : 0ull;  // padding sorts last
}

// descending order overall: in block (i & k) == 0 the larger key goes to the lower address
__device__ __forceinline__ void cmp_swap(uint64_t& a, uint64_t& b, bool desc) {
  if ((a < b) == desc) {
    const uint64_t t = a;
    a = b;
    b = t;
  }
}

// all stages k <= kSortTile (start), or the tail j < kSortTile of a larger stage k
__global__ void __launch_bounds__(kSortTile / 2) rank_sort_smem_kernel(uint64_t* __restrict__ keys, long long n2, long long k_from,
                                                                       long long k_to) {
  __shared__ uint64_t sk[kSortTile];
  const int q = blockIdx.y;
  uint64_t* base = keys + q * n2 + static_cast<long long>(blockIdx.x) * kSortTile;
  const long long g0 = static_cast<long long>(blockIdx.x) * kSortTile;
  const int span = static_cast<int>(std::min<long long>(kSortTile, n2));
  for (int i = threadIdx.x; i < span; i += blockDim.x) sk[i] = base[i];
  __syncthreads();
  for (long long k = k_from; k <= k_to; k <<= 1) {
    for (int j = static_cast<int>(std::min<long long>(k >> 1, kSortTile >> 1)); j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < span / 2; t += blockDim.x) {
        const int i = 2 * t - (t & (j - 1));  // index with bit j clear
        const bool desc = ((g0 + i) & k) == 0;
        cmp_swap(sk[i], sk[i + j], desc);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < span; i += blockDim.x) base[i] = sk[i];
}

__global__ void rank_sort_global_kernel(uint64_t* __restrict__ keys, long long n2, long long k, long long j) {
  const int q = blockIdx.y;
  uint64_t* base = keys + q * n2;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n2 / 2; t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = 2 * t - (t & (j - 1));
    uint64_t a = base[i], b = base[i + j];
    const bool desc = (i & k) == 0;
    if ((a < b) == desc) {
      base[i] = b;
      base[i + j] = a;
    }
  }
}

__global__ void rank_extract_kernel(const uint64_t* __restrict__ keys, long long n, long long n2, int* __restrict__ order,
                                    int* __restrict__ rank1) {
  const int q = blockIdx.y;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int idx = static_cast<int>(key_index(keys[q * n2 + r]));
    if (order) order[q * n + r] = idx;
    if (rank1) rank1[q * n + idx] = static_cast<int>(r) + 1;
  }
}

static long long next_pow2_ll(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace ss

using namespace ss;

extern "C" size_t ss_rank_order_workspace_bytes(int n_queries, int64_t n) {
  if (n_queries <= 0 || n <= 0) return 0;
  return static_cast<size_t>(n_queries) * static_cast<size_t>(next_pow2_ll(n)) * 8 + 256;
}

extern "C" int ss_rank_order(const float* scores, int n_queries, int64_t n, void* workspace, size_t workspace_bytes,
                             int32_t* out_order, int32_t* out_rank1, void* stream) {
  if (!scores || !workspace || (!out_order && !out_rank1)) return fail(SS_ERR_INVALID_ARG, "ss_rank_order: null pointer");
  if (n_queries <= 0 || n <= 0 || n > 0x7FFFFFFFll || n_queries > 65535) return fail(SS_ERR_INVALID_ARG, "ss_rank_order: bad sizes");
  if (workspace_bytes < ss_rank_order_workspace_bytes(n_queries, n)) return fail(SS_ERR_WORKSPACE, "ss_rank_order: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint64_t* keys = reinterpret_cast<uint64_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  const long long n2 = next_pow2_ll(n);
  const int gx = static_cast<int>(std::max<long long>(1, std::min<long long>((n2 + 255) / 256, static_cast<long long>(sm_count()) * 8)));
  rank_keys_init_kernel<<<dim3(gx, n_queries), 256, 0, st>>>(scores, n, n2, keys);
  SS_CUDA_CHECK(cudaGetLastError());
  const int tiles = static_cast<int>(std::max<long long>(1, n2 / kSortTile));
  rank_sort_smem_kernel<<<dim3(tiles, n_queries), kSortTile / 2, 0, st>>>(keys, n2, 2, std::min<long long>(n2, kSortTile));
  SS_CUDA_CHECK(cudaGetLastError());
  for (long long k = static_cast<long long>(kSortTile) * 2; k <= n2; k <<= 1) {
    for (long long j = k >> 1; j >= kSortTile; j >>= 1) {
      rank_sort_global_kernel<<<dim3(gx, n_queries), 256, 0, st>>>(keys, n2, k, j);
      SS_CUDA_CHECK(cudaGetLastError());
    }
    rank_sort_smem_kernel<<<dim3(tiles, n_queries), kSortTile / 2, 0, st>>>(keys, n2, k, k);
    SS_CUDA_CHECK(cudaGetLastError());
  }
  rank_extract_kernel<<<dim3(gx, n_queries), 256, 0, st>>>(keys, n, n2, out_order, out_rank1);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
