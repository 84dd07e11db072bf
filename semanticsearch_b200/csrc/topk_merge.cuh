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

constexpr int kMergeSurvivorCap = 2048;  // shared-memory scratch (keys) the fast path needs

__device__ __forceinline__ void merge_emit(const MergeOut& out, int rank, uint64_t key) {
  if (out.keys) out.keys[rank] = key;
  if (out.scores) out.scores[rank] = key_score(key);
  if (out.indices) out.indices[rank] = key_index(key);
}

// All threads of the block sort a[0..n) (n a power of two, shared memory) in descending order.
__device__ __forceinline__ void block_bitonic_desc(uint64_t* a, int n) {
  for (int k2 = 2; k2 <= n; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const uint64_t x = a[i], y = a[p];
        const bool desc = (i & k2) == 0;
        if (desc ? (x < y) : (x > y)) {
          a[i] = y;
          a[p] = x;
        }
      }
      __syncthreads();
    }
  }
}

// All threads of the block call this.  `lists` may point to shared or global memory; list p
// starts at lists + p * list_stride and holds k_in keys sorted best-first (0 = empty).
// `scratch` is a shared uint64_t[2] owned by the caller; `surv` is an optional shared scratch of
// kMergeSurvivorCap keys.  Outputs are fully written (empty slots get key 0 / -inf / -1).
//
//   1. prune: with at least k_out lists, the k_out-th best list head is a lower bound ("floor")
//      of the global k_out-th best key, so only each list's prefix >= floor can matter;
//   2. fast path: those prefixes (found by binary search, typically a few keys per list) are
//      gathered into `surv` and bitonic-sorted by the whole block;
//   3. fallback (no scratch, or more survivors than it holds): exact global rank of every
//      surviving candidate by binary search in the other lists.
__device__ __forceinline__ void block_merge_lists(const uint64_t* lists, int n_lists, int k_in, long long list_stride,
                                                  int k_out, const MergeOut& out, uint64_t* scratch, uint64_t* surv = nullptr) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int j = tid; j < k_out; j += nthr) {
    if (out.keys) out.keys[j] = 0ull;
    if (out.scores) out.scores[j] = -INFINITY;
    if (out.indices) out.indices[j] = -1;
  }
  if (tid == 0) {
    scratch[0] = 0ull;
    scratch[1] = 0ull;  // survivor count (low 32 bits), overflow flag (bit 32)
  }
  __syncthreads();
  // Phase 1: prune.
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
  const int lim = min(k_in, k_out);  // element i of a sorted list already has i better keys
  if (surv != nullptr) {
    // Phase 2 (fast path): gather every list's prefix >= max(floor, 1) into surv, any order.
    const uint64_t f = floor_key > 0ull ? floor_key : 1ull;
    unsigned int* s_count = reinterpret_cast<unsigned int*>(&scratch[1]);
    for (int p = tid; p < n_lists; p += nthr) {
      const uint64_t* l = lists + static_cast<size_t>(p) * list_stride;
      int lo = 0, hi = lim;  // first index whose key < f
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (l[mid] >= f) lo = mid + 1; else hi = mid;
      }
      if (lo > 0) {
        const unsigned int pos = atomicAdd(s_count, static_cast<unsigned int>(lo));
        if (pos + lo <= static_cast<unsigned int>(kMergeSurvivorCap)) {
          for (int i = 0; i < lo; ++i) surv[pos + i] = l[i];
        } else {
          s_count[1] = 1u;
        }
      }
    }
    __syncthreads();
    const unsigned int total = s_count[0];
    const bool overflow = s_count[1] != 0u;
    if (!overflow) {
      int n = 32;
      while (n < static_cast<int>(total)) n <<= 1;
      for (int i = static_cast<int>(total) + tid; i < n; i += nthr) surv[i] = 0ull;
      __syncthreads();
      block_bitonic_desc(surv, n);
      for (int j = tid; j < k_out && j < static_cast<int>(total); j += nthr) merge_emit(out, j, surv[j]);
      return;
    }
  }
  // Phase 3 (fallback): exact global rank of every surviving candidate.
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
    if (rank < k_out) merge_emit(out, rank, key);
  }
}

}  // namespace ss
