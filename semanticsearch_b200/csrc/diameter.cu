// a12 — diameter-bounded splitting of a chunk's sentences, the controller's optional "enforce_diameter" stage
// (data_process/simple_chunk_controller.py:571-594, called from :596-625 on sim = emb @ emb.T, :614).
//
//   diameter([a, b)) = 1 - min_{i != j in [a, b)} S[i][j]                         (:576-586)
//   a span whose diameter exceeds the threshold is cut after the position of its lowest adjacent similarity
//   S[i][i + 1] (first minimum, np.argmin) and both halves are treated the same way, left half first (:588-594).
//
// One CTA per document walks the recursion with an explicit stack in shared memory (left half on top, so spans come
// out in ascending order).  The minimum of a span is a block reduction over its off-diagonal entries; the matrices
// are small (a chunk holds a few dozen sentences), so the n^2 re-reads of nested spans stay in L1 / L2.
#include "ss_common.cuh"

namespace ss {

constexpr int kDiamThreads = 256;
constexpr int kDiamMaxRows = 4096;  // stack of (start, end) pairs in shared memory

__global__ void __launch_bounds__(kDiamThreads) diameter_split_kernel(const float* __restrict__ S_all, const int* __restrict__ offsets,
                                                                      const long long* __restrict__ s_offsets, double threshold,
                                                                      int* __restrict__ out_ends, int* __restrict__ out_n_spans,
                                                                      double* __restrict__ out_diameter) {
  extern __shared__ int diam_stack[];  // [2 * n]
  __shared__ unsigned int s_min;
  __shared__ unsigned long long s_adj;
  __shared__ int s_top, s_emitted;
  const int doc = blockIdx.x;
  const int row_base = offsets[doc];
  const int n = offsets[doc + 1] - row_base;
  const float* S = S_all + s_offsets[doc];
  int* ends = out_ends + row_base;
  if (n <= 0) {
    if (threadIdx.x == 0) {
      out_n_spans[doc] = 0;
      out_diameter[doc] = 0.0;
    }
    return;
  }
  if (threadIdx.x == 0) {
    diam_stack[0] = 0;
    diam_stack[1] = n;
    s_top = 1;
    s_emitted = 0;
  }
  __syncthreads();
  bool first = true;
  while (true) {
    __syncthreads();
    const int top = s_top;
    if (top == 0) break;
    const int a = diam_stack[2 * (top - 1)], b = diam_stack[2 * (top - 1) + 1];
    const int len = b - a;
    if (threadIdx.x == 0) {
      s_min = 0xFFFFFFFFu;
      s_adj = ~0ull;
    }
    __syncthreads();
    bool split = false;
    if (len >= 2) {
      unsigned int mn = 0xFFFFFFFFu;
      for (long long t = threadIdx.x; t < static_cast<long long>(len) * len; t += kDiamThreads) {
        const int i = static_cast<int>(t / len), j = static_cast<int>(t - static_cast<long long>(i) * len);
        if (i != j) mn = min(mn, float_to_ordered(S[static_cast<size_t>(a + i) * n + (a + j)]));
      }
      mn = __reduce_min_sync(0xffffffffu, mn);
      if ((threadIdx.x & 31) == 0) atomicMin(&s_min, mn);
      unsigned long long adj = ~0ull;  // (ordered similarity, position): the first minimum wins, like np.argmin
      for (int i = a + threadIdx.x; i < b - 1; i += kDiamThreads)
        adj = min(adj, (static_cast<unsigned long long>(float_to_ordered(S[static_cast<size_t>(i) * n + i + 1])) << 32) |
                           static_cast<unsigned int>(i));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) adj = min(adj, __shfl_xor_sync(0xffffffffu, adj, o));
      if ((threadIdx.x & 31) == 0) atomicMin(&s_adj, adj);
      __syncthreads();
      const double diam = 1.0 - static_cast<double>(ordered_to_float(s_min));  // 1.0 - float(min_sim), in float64 (:583-586)
      if (first && threadIdx.x == 0) out_diameter[doc] = diam;
      split = diam > threshold;
    } else if (first && threadIdx.x == 0) {
      out_diameter[doc] = 0.0;
    }
    first = false;
    __syncthreads();
    if (threadIdx.x == 0) {
      if (!split) {
        ends[s_emitted++] = b;  // the span [previous end, b) is final
        s_top = top - 1;
      } else {
        const int cut = static_cast<int>(s_adj & 0xFFFFFFFFull) + 1;
        diam_stack[2 * (top - 1)] = cut;  // right half below ...
        diam_stack[2 * (top - 1) + 1] = b;
        diam_stack[2 * top] = a;          // ... left half on top: it is split first
        diam_stack[2 * top + 1] = cut;
        s_top = top + 1;
      }
    }
  }
  if (threadIdx.x == 0) out_n_spans[doc] = s_emitted;
}

}  // namespace ss

using namespace ss;

extern "C" int ss_diameter_split(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, int max_doc_rows,
                                 double threshold, int32_t* out_span_ends, int32_t* out_n_spans, double* out_diameter, void* stream) {
  if (!S || !offsets || !s_offsets || !out_span_ends || !out_n_spans || !out_diameter)
    return fail(SS_ERR_INVALID_ARG, "ss_diameter_split: null pointer");
  if (n_docs <= 0 || max_doc_rows <= 0) return fail(SS_ERR_INVALID_ARG, "ss_diameter_split: sizes must be positive");
  if (max_doc_rows > kDiamMaxRows) return fail(SS_ERR_UNSUPPORTED, "ss_diameter_split: documents of more than 4096 rows are not supported");
  const size_t smem = static_cast<size_t>(2 * max_doc_rows + 2) * sizeof(int);
  diameter_split_kernel<<<n_docs, kDiamThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      S, offsets, reinterpret_cast<const long long*>(s_offsets), threshold, out_span_ends, out_n_spans, out_diameter);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
