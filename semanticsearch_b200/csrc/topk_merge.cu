// K6 — k-way merge of best-first key lists (per-CTA partials, or per-GPU lists after all-gather)
// and row inverse norms.  The distributed form of np.argsort(-scores)[:k]
// (Tool/rank_chunks_optimized.py:225): every candidate's global rank is the sum, over the other
// lists, of the number of better keys — found by binary search because each list is sorted.
#include <algorithm>

#include "ss_common.cuh"
#include "topk_merge.cuh"

namespace ss {

// One CTA per query.  Lists are staged in shared memory when they fit (they always do for the
// shapes of BASELINE.json: 148 x 100 keys = 118 KB at most), followed by the survivor scratch of
// the fast path.
constexpr int kMergeThreads = 512;

__global__ void __launch_bounds__(kMergeThreads) topk_merge_kernel(const uint64_t* __restrict__ keys_in, int n_lists, int k_in,
                                                                   long long query_stride, long long list_stride, int k_out,
                                                                   uint64_t* __restrict__ out_keys, float* __restrict__ out_scores,
                                                                   long long* __restrict__ out_indices, int stage_in_smem) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  __shared__ uint64_t scratch[2];
  const int q = blockIdx.x;
  const uint64_t* base = keys_in + static_cast<size_t>(q) * query_stride;
  MergeOut out;
  out.keys = out_keys ? out_keys + static_cast<size_t>(q) * k_out : nullptr;
  out.scores = out_scores ? out_scores + static_cast<size_t>(q) * k_out : nullptr;
  out.indices = out_indices ? out_indices + static_cast<size_t>(q) * k_out : nullptr;
  uint64_t* surv = reinterpret_cast<uint64_t*>(merge_smem);
  if (stage_in_smem) {
    uint64_t* sl = surv + kMergeSurvivorCap;
    const int total = n_lists * k_in;
    if (list_stride == k_in && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
      // contiguous [list][k] block: flat 16-byte copy, four loads in flight per thread
      const int nvec = total >> 1;
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(base);
      ulonglong2* dst = reinterpret_cast<ulonglong2*>(sl);
      int c = threadIdx.x;
      for (; c + 3 * kMergeThreads < nvec; c += 4 * kMergeThreads) {
        const ulonglong2 v0 = src[c], v1 = src[c + kMergeThreads], v2 = src[c + 2 * kMergeThreads], v3 = src[c + 3 * kMergeThreads];
        dst[c] = v0;
        dst[c + kMergeThreads] = v1;
        dst[c + 2 * kMergeThreads] = v2;
        dst[c + 3 * kMergeThreads] = v3;
      }
      for (; c < nvec; c += kMergeThreads) dst[c] = src[c];
      if ((total & 1) && threadIdx.x == 0) sl[total - 1] = base[total - 1];
    } else {
      for (int c = threadIdx.x; c < total; c += blockDim.x) {
        const int pl = c / k_in, i = c - pl * k_in;
        sl[c] = base[static_cast<size_t>(pl) * list_stride + i];
      }
    }
    __syncthreads();
    block_merge_lists(sl, n_lists, k_in, k_in, k_out, out, scratch, surv);
  } else {
    block_merge_lists(base, n_lists, k_in, list_stride, k_out, out, scratch, surv);
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
  const size_t bytes = static_cast<size_t>(n_lists) * k_in * 8;
  const size_t surv_bytes = static_cast<size_t>(kMergeSurvivorCap) * 8;
  const int stage = bytes + surv_bytes + 1024 <= smem_optin() ? 1 : 0;
  const size_t dyn = surv_bytes + (stage ? bytes : 0);
  if (dyn > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
  topk_merge_kernel<<<n_queries, kMergeThreads, dyn, static_cast<cudaStream_t>(stream)>>>(
      keys_in, n_lists, k_in, query_stride, list_stride, k_out, out_keys, out_scores,
      reinterpret_cast<long long*>(out_indices), stage);
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
