// K6 — k-way merge of best-first key lists (per-CTA partials, or per-GPU lists after all-gather)
// and row inverse norms.  The distributed form of np.argsort(-scores)[:k]
// (Tool/rank_chunks_optimized.py:225): every candidate's global rank is the sum, over the other
// lists, of the number of better keys — found by binary search because each list is sorted.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

__global__ void __launch_bounds__(256) topk_merge_kernel(const uint64_t* __restrict__ keys_in, int n_lists, int k_in,
                                                         long long query_stride, long long list_stride, int k_out,
                                                         uint64_t* __restrict__ out_keys, float* __restrict__ out_scores,
                                                         long long* __restrict__ out_indices) {
  const int q = blockIdx.x;
  const uint64_t* base = keys_in + static_cast<size_t>(q) * query_stride;
  for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
    if (out_keys) out_keys[static_cast<size_t>(q) * k_out + j] = 0ull;
    if (out_scores) out_scores[static_cast<size_t>(q) * k_out + j] = -INFINITY;
    if (out_indices) out_indices[static_cast<size_t>(q) * k_out + j] = -1;
  }
  __syncthreads();
  const int lim = min(k_in, k_out);  // element i of a sorted list has at least i better keys
  for (int c = threadIdx.x; c < n_lists * lim; c += blockDim.x) {
    const int pl = c / lim, i = c - pl * lim;
    const uint64_t key = base[static_cast<size_t>(pl) * list_stride + i];
    if (key == 0ull) continue;
    int rank = i;
    for (int p2 = 0; p2 < n_lists && rank < k_out; ++p2) {
      if (p2 == pl) continue;
      const uint64_t* l2 = base + static_cast<size_t>(p2) * list_stride;
      int lo = 0, hi = k_in;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (l2[mid] > key) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k_out) {
      const size_t o = static_cast<size_t>(q) * k_out + rank;
      if (out_keys) out_keys[o] = key;
      if (out_scores) out_scores[o] = key_score(key);
      if (out_indices) out_indices[o] = key_index(key);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) row_inv_norms_kernel(const T* __restrict__ rows, long long n_rows, int dim,
                                                            float zero_value, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp; r < n_rows; r += nwarps) {
    const T* row = rows + static_cast<size_t>(r) * dim;
    float ssq = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float x = to_f32<T>(row[c]);
      ssq = fmaf(x, x, ssq);
    }
    ssq = warp_sum(ssq);
    if (lane == 0) out[r] = ssq > 0.f ? 1.0f / sqrtf(ssq) : zero_value;
  }
}

}  // namespace ss

using namespace ss;

extern "C" int ss_topk_merge(const uint64_t* keys_in, int n_lists, int n_queries, int k_in, int64_t query_stride,
                             int64_t list_stride, int k_out, uint64_t* out_keys, float* out_scores, int64_t* out_indices,
                             void* stream) {
  if (!keys_in) return fail(SS_ERR_INVALID_ARG, "ss_topk_merge: null input");
  if (n_lists <= 0 || n_queries <= 0 || k_in <= 0 || k_out <= 0)
    return fail(SS_ERR_INVALID_ARG, "ss_topk_merge: sizes must be positive");
  topk_merge_kernel<<<n_queries, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      keys_in, n_lists, k_in, query_stride, list_stride, k_out, out_keys, out_scores,
      reinterpret_cast<long long*>(out_indices));
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}

extern "C" int ss_row_inv_norms(const void* rows, int64_t n_rows, int dim, int dtype, float zero_value, float* out_inv_norms,
                                void* stream) {
  if (!rows || !out_inv_norms) return fail(SS_ERR_INVALID_ARG, "ss_row_inv_norms: null pointer");
  if (n_rows <= 0 || dim <= 0 || !dtype_ok(dtype)) return fail(SS_ERR_INVALID_ARG, "ss_row_inv_norms: bad shape or dtype");
  const int blocks = static_cast<int>(std::min<long long>((n_rows + 7) / 8, static_cast<long long>(sm_count()) * 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case SS_F32: row_inv_norms_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(rows), n_rows, dim, zero_value, out_inv_norms); break;
    case SS_BF16: row_inv_norms_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(rows), n_rows, dim, zero_value, out_inv_norms); break;
    default: row_inv_norms_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(rows), n_rows, dim, zero_value, out_inv_norms); break;
  }
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
