// C99 rank transform of per-document similarity matrices (Method/Semantic_Splitter_Optimized.py:171-192).
//
//   global: R[i,j] = #{k : S[i,k] < S[i,j]} + #{k : S[k,j] < S[i,j]}            (:189-192)
//   local : R[i,j] = #{(a,b) in the clipped m x m window around (i,j) : S[a,b] < S[i,j]} / window size  (:171-186)
//
// Pure compare-and-count (integer work): one CTA per (document, row); the row is staged in shared
// memory, the column walk is coalesced across threads.  The reference does this with an n^3
// boolean broadcast (134 MB at n = 512) or 262 144 Python iterations per document.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

__global__ void __launch_bounds__(256) c99_rank_kernel(const float* __restrict__ S_all, const int* __restrict__ offsets,
                                                       const long long* __restrict__ s_offsets, const int* __restrict__ row_doc,
                                                       int local_mode, int half, float* __restrict__ R_all) {
  extern __shared__ float c99_row[];  // the whole row i (global mode) or the window rows' needed span is read from global
  const int grow = blockIdx.x;         // global row index in the concatenated batch
  const int doc = row_doc[grow];
  const int row_base = offsets[doc];
  const int n = offsets[doc + 1] - row_base;
  const int i = grow - row_base;
  const float* S = S_all + s_offsets[doc];
  float* R = R_all + s_offsets[doc];
  if (!local_mode) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) c99_row[k] = S[static_cast<size_t>(i) * n + k];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const float v = c99_row[j];
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        cnt += (c99_row[k] < v) ? 1 : 0;                              // row i: shared-memory broadcast
        cnt += (S[static_cast<size_t>(k) * n + j] < v) ? 1 : 0;       // column j: coalesced across threads
      }
      R[static_cast<size_t>(i) * n + j] = static_cast<float>(cnt);
    }
  } else {
    const int i0 = max(0, i - half), i1 = min(n, i + half + 1);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int j0 = max(0, j - half), j1 = min(n, j + half + 1);
      const float v = S[static_cast<size_t>(i) * n + j];
      int cnt = 0;
      for (int a = i0; a < i1; ++a)
        for (int b = j0; b < j1; ++b) cnt += (S[static_cast<size_t>(a) * n + b] < v) ? 1 : 0;
      const int denom = (i1 - i0) * (j1 - j0);
      R[static_cast<size_t>(i) * n + j] = static_cast<float>(static_cast<double>(cnt) / static_cast<double>(denom > 0 ? denom : 1));
    }
  }
}

__global__ void c99_row_doc_kernel(const int* __restrict__ offsets, int n_docs, int* __restrict__ row_doc) {
  const int doc = blockIdx.x;
  if (doc >= n_docs) return;
  for (int r = offsets[doc] + threadIdx.x; r < offsets[doc + 1]; r += blockDim.x) row_doc[r] = doc;
}

}  // namespace ss

using namespace ss;

extern "C" int ss_c99_rank_matrix(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, int total_rows,
                                  int max_doc_rows, int use_local_rank, int mask_size, int32_t* workspace_rows, float* out_R,
                                  void* stream) {
  if (!S || !offsets || !s_offsets || !workspace_rows || !out_R) return fail(SS_ERR_INVALID_ARG, "ss_c99_rank_matrix: null pointer");
  if (n_docs <= 0 || total_rows <= 0 || max_doc_rows <= 0) return fail(SS_ERR_INVALID_ARG, "ss_c99_rank_matrix: sizes must be positive");
  const size_t smem = static_cast<size_t>(max_doc_rows) * sizeof(float);
  if (smem + 1024 > smem_optin()) return fail(SS_ERR_UNSUPPORTED, "ss_c99_rank_matrix: document too long for the shared-memory row");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c99_row_doc_kernel<<<n_docs, 128, 0, st>>>(offsets, n_docs, workspace_rows);
  SS_CUDA_CHECK(cudaGetLastError());
  if (smem > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(c99_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int m = std::max(3, mask_size | 1);
  c99_rank_kernel<<<total_rows, 256, smem, st>>>(S, offsets, reinterpret_cast<const long long*>(s_offsets), workspace_rows,
                                                use_local_rank ? 1 : 0, m / 2, out_R);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
