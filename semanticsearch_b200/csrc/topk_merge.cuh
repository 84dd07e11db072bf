// Block-level k-way merge of best-first key lists, shared by the standalone merge kernel (K6) and
// the last-CTA epilogue fused into the cosine kernels.
#pragma once

#include "ss_common.cuh"

namespace ss {

struct MergeOut {
  uint64_t* keys;      // [k_out] or nullptr
  float* scores;       // [k_out] or nullptr
  long long* indices;  // [k_out] or nullptr
};

// All threads of the block call this.  `lists` may point to shared or global memory; list p
// starts at lists + p * list_stride and holds k_in keys sorted best-first (0 = empty).
// `scratch` is a shared uint64_t[2] owned by the caller.  Outputs are fully written (empty slots
// get key 0 / -inf / -1).
__device__ __forceinline__ void block_merge_lists(const uint64_t* lists, int n_lists, int k_in, long long list_stride,
                                                  int k_out, const MergeOut& out, uint64_t* scratch) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int j = tid; j < k_out; j += nthr) {
    if (out.keys) out.keys[j] = 0ull;
    if (out.scores) out.scores[j] = -INFINITY;
    if (out.indices) out.indices[j] = -1;
  }
  if (tid == 0) scratch[0] = 0ull;
  __syncthreads();
  // Phase 1: prune.  The k_out-th best list head is a lower bound of the global k_out-th best key.
  if (n_lists >= k_out) {
    for (int p = tid; p < n_lists; p += nthr) {
      const uint64_t h = lists[static_cast<size_t>(p) * list_stride];
      int better = 0;
      for (int p2 = 0; p2 < n_lists; ++p2) better += (lists[static_cast<size_t>(p2) * list_stride] > h) ? 1 : 0;
      if (better == k_out - 1) scratch[0] = h;  // keys are unique, so exactly one head has this rank (or all are empty)
    }
  }
  __syncthreads();
  const uint64_t floor_key = scratch[0];
  // Phase 2: exact global rank of every surviving candidate by binary search in the other lists.
  const int lim = min(k_in, k_out);  // element i of a sorted list already has i better keys
  for (int c = tid; c < n_lists * lim; c += nthr) {
    const int pl = c / lim, i = c - pl * lim;
    const uint64_t key = lists[static_cast<size_t>(pl) * list_stride + i];
    if (key == 0ull || key < floor_key) continue;
    int rank = i;
    for (int p2 = 0; p2 < n_lists && rank < k_out; ++p2) {
      if (p2 == pl) continue;
      const uint64_t* l2 = lists + static_cast<size_t>(p2) * list_stride;
      if (!(l2[0] > key)) continue;  // common case: nothing in that list beats this key
      int lo = 1, hi = k_in;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (l2[mid] > key) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k_out) {
      if (out.keys) out.keys[rank] = key;
      if (out.scores) out.scores[rank] = key_score(key);
      if (out.indices) out.indices[rank] = key_index(key);
    }
  }
}

}  // namespace ss
