// K8 — segmented ranking of query groups: cosine scores, cosine / BM25 ranks, reciprocal-rank fusion,
// the fused order and the percentile thresholds of every query group of a block in ONE launch.
//
// Replaces, for all groups at once, the per-query body of OptimizedRanker.rank_single_query_optimized and
// the labelling that follows (Tool/rank_chunks_optimized.py):
//   cosine_similarity(q, chunks)[0]                                   (:215-216)
//   np.argsort(-cosine) / np.argsort(-bm25) and the 1-based rank lookups   (:225-235)
//   rrf = 1/(k + rank_cos) + 1/(k + rank_bm25), k = 60                (:238-239)
//   sort_values(by="rrf_score", ascending=False)                      (:250)
//   np.percentile(rrf, upper) / np.percentile(rrf, lower)             (:518-519)
// The reference runs one Python task per query over a few hundred chunks; here one CTA owns one group
// (rows offsets[g]..offsets[g+1] of the chunk-embedding matrix, query g), everything stays in shared
// memory between the steps, and ties are resolved lower-row-first where the reference's non-stable sorts
// leave them unspecified.  BM25 itself is lexical and stays on the host; its scores come in as an input.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

constexpr int kRrfThreads = 512;
constexpr int kRrfMaxRows = 8192;

struct RrfParams {
  const float* chunks;   // [total_rows][dim]
  const int* offsets;    // [n_groups + 1]
  const float* queries;  // [n_groups][dim]
  const float* bm25;     // [total_rows] or nullptr
  int dim;
  double k_rrf, q_hi, q_lo;  // quantiles in [0, 1]
  float* out_cos;        // [total_rows]
  int* out_rank_cos;     // [total_rows] 1-based, inside the group
  int* out_rank_bm25;    // [total_rows] or nullptr
  double* out_rrf;       // [total_rows]
  int* out_order;        // [total_rows] local row indices of the group, best fused score first
  double* out_thr;       // [n_groups][2] = np.percentile(rrf, upper), np.percentile(rrf, lower)
};

__device__ __forceinline__ uint64_t rrf_f64_to_ordered(double d) {
  const uint64_t b = static_cast<uint64_t>(__double_as_longlong(d));
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double rrf_ordered_to_f64(uint64_t o) {
  const uint64_t b = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
  return __longlong_as_double(static_cast<long long>(b));
}

// Block-wide bitonic sort of n (power of two) packed keys, descending.
__device__ __forceinline__ void rrf_sort_keys_desc(uint64_t* a, int n) {
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

// Same network on (64-bit value, row index) pairs: value descending, equal values -> lower row first.
__device__ __forceinline__ void rrf_sort_pairs_desc(uint64_t* v, int* ix, int n) {
  for (int k2 = 2; k2 <= n; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const uint64_t x = v[i], y = v[p];
        const int xi = ix[i], yi = ix[p];
        const bool x_first = x > y || (x == y && static_cast<unsigned int>(xi) < static_cast<unsigned int>(yi));
        const bool desc = (i & k2) == 0;
        if (desc != x_first) {
          v[i] = y;
          v[p] = x;
          ix[i] = yi;
          ix[p] = xi;
        }
      }
      __syncthreads();
    }
  }
}

// np.percentile(values, q * 100) (linear) read from a DESCENDING sorted key array of m values.
__device__ __forceinline__ double rrf_quantile_desc(const uint64_t* desc_sorted, int m, double q) {
  const double vi = __dmul_rn(static_cast<double>(m - 1), q);
  const double fl = floor(vi);
  int lo = static_cast<int>(fl);
  lo = max(0, min(lo, m - 1));
  const int hi = min(lo + 1, m - 1);
  const double g = __dsub_rn(vi, fl);
  const double a = rrf_ordered_to_f64(desc_sorted[m - 1 - lo]), b = rrf_ordered_to_f64(desc_sorted[m - 1 - hi]);
  const double diff = __dsub_rn(b, a);  // numpy's _lerp, without FMA contraction
  double r = __dadd_rn(a, __dmul_rn(diff, g));
  if (g >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, g)));
  return r;
}

__global__ void __launch_bounds__(kRrfThreads) segmented_rank_rrf_kernel(const RrfParams p) {
  extern __shared__ __align__(16) unsigned char rrf_smem[];
  __shared__ float s_inv_q;
  const int g = blockIdx.x;
  const int row0 = p.offsets[g];
  const int n = p.offsets[g + 1] - row0;
  if (n <= 0) {
    if (threadIdx.x < 2) p.out_thr[2 * g + threadIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  int n2 = 2;
  while (n2 < n) n2 <<= 1;
  uint64_t* keys = reinterpret_cast<uint64_t*>(rrf_smem);        // [n2]
  int* rank_c = reinterpret_cast<int*>(keys + n2);                // [n2]
  int* rank_b = rank_c + n2;                                      // [n2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* q = p.queries + static_cast<size_t>(g) * p.dim;

  // ---- query norm (sklearn: zero norm -> 1) ----
  if (warp == 0) {
    float ss = 0.f;
    for (int c = lane; c < p.dim; c += 32) ss = fmaf(q[c], q[c], ss);
    ss = warp_sum(ss);
    if (lane == 0) s_inv_q = ss > 0.f ? 1.0f / sqrtf(ss) : 1.0f;
  }
  __syncthreads();
  const float inv_q = s_inv_q;

  // ---- cosine of every chunk row of the group: one warp per row ----
  const bool vec = (p.dim & 3) == 0 && (reinterpret_cast<uintptr_t>(p.chunks) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.queries) & 15) == 0;
  for (int r = warp; r < n; r += nwarps) {
    const float* c = p.chunks + static_cast<size_t>(row0 + r) * p.dim;
    float dot = 0.f, ssq = 0.f;
    if (vec) {
      for (int i = lane * 4; i < p.dim; i += 128) {
        const float4 cv = *reinterpret_cast<const float4*>(c + i);
        const float4 qv = *reinterpret_cast<const float4*>(q + i);
        dot = fmaf(cv.x, qv.x, fmaf(cv.y, qv.y, fmaf(cv.z, qv.z, fmaf(cv.w, qv.w, dot))));
        ssq = fmaf(cv.x, cv.x, fmaf(cv.y, cv.y, fmaf(cv.z, cv.z, fmaf(cv.w, cv.w, ssq))));
      }
    } else {
      for (int i = lane; i < p.dim; i += 32) {
        dot = fmaf(c[i], q[i], dot);
        ssq = fmaf(c[i], c[i], ssq);
      }
    }
    dot = warp_sum(dot);
    ssq = warp_sum(ssq);
    if (lane == 0) {
      const float cosv = (dot * inv_q) * (ssq > 0.f ? 1.0f / sqrtf(ssq) : 1.0f);
      p.out_cos[row0 + r] = cosv;
      keys[r] = make_key(cosv, static_cast<uint32_t>(r));
    }
  }
  for (int i = n + threadIdx.x; i < n2; i += blockDim.x) keys[i] = 0ull;
  __syncthreads();
  rrf_sort_keys_desc(keys, n2);
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const int idx = static_cast<int>(key_index(keys[r]));
    rank_c[idx] = r + 1;
    p.out_rank_cos[row0 + idx] = r + 1;
  }
  __syncthreads();

  // ---- BM25 ranks (scores computed on the host) ----
  const bool has_bm = p.bm25 != nullptr;
  if (has_bm) {
    for (int r = threadIdx.x; r < n2; r += blockDim.x) keys[r] = r < n ? make_key(p.bm25[row0 + r], static_cast<uint32_t>(r)) : 0ull;
    __syncthreads();
    rrf_sort_keys_desc(keys, n2);
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const int idx = static_cast<int>(key_index(keys[r]));
      rank_b[idx] = r + 1;
      if (p.out_rank_bm25) p.out_rank_bm25[row0 + idx] = r + 1;
    }
    __syncthreads();
  }

  // ---- reciprocal-rank fusion in fp64 (the rank arrays are float64 in the reference), fused order ----
  for (int r = threadIdx.x; r < n2; r += blockDim.x) {
    if (r < n) {
      double v = __ddiv_rn(1.0, __dadd_rn(p.k_rrf, static_cast<double>(rank_c[r])));
      if (has_bm) v = __dadd_rn(v, __ddiv_rn(1.0, __dadd_rn(p.k_rrf, static_cast<double>(rank_b[r]))));
      p.out_rrf[row0 + r] = v;
      keys[r] = rrf_f64_to_ordered(v);
      rank_c[r] = r;  // rank_c now carries the row index next to its fused score
    } else {
      keys[r] = 0ull;
      rank_c[r] = 0x7fffffff;
    }
  }
  __syncthreads();
  rrf_sort_pairs_desc(keys, rank_c, n2);
  for (int r = threadIdx.x; r < n; r += blockDim.x) p.out_order[row0 + r] = rank_c[r];
  if (threadIdx.x == 0) p.out_thr[2 * g] = rrf_quantile_desc(keys, n, p.q_hi);
  if (threadIdx.x == 32) p.out_thr[2 * g + 1] = rrf_quantile_desc(keys, n, p.q_lo);
}

}  // namespace ss

using namespace ss;

extern "C" int ss_segmented_rank_rrf(const float* chunks, int dim, const int32_t* offsets, int n_groups, int max_group_rows,
                                     const float* queries, const float* bm25_scores, double k_rrf, double upper_percentile,
                                     double lower_percentile, float* out_cos, int32_t* out_rank_cos, int32_t* out_rank_bm25,
                                     double* out_rrf, int32_t* out_order, double* out_thr, void* stream) {
  if (!chunks || !offsets || !queries || !out_cos || !out_rank_cos || !out_rrf || !out_order || !out_thr)
    return fail(SS_ERR_INVALID_ARG, "ss_segmented_rank_rrf: null pointer");
  if (dim <= 0 || n_groups <= 0 || max_group_rows <= 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_rank_rrf: sizes must be positive");
  if (max_group_rows > kRrfMaxRows) return fail(SS_ERR_UNSUPPORTED, "ss_segmented_rank_rrf: groups larger than 8192 chunks are not supported");
  if (!(upper_percentile >= 0.0 && upper_percentile <= 100.0 && lower_percentile >= 0.0 && lower_percentile <= 100.0))
    return fail(SS_ERR_INVALID_ARG, "ss_segmented_rank_rrf: percentiles must be in [0, 100]");
  int n2 = 2;
  while (n2 < max_group_rows) n2 <<= 1;
  const size_t smem = static_cast<size_t>(n2) * (8 + 4 + 4);
  if (smem > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(segmented_rank_rrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  RrfParams p;
  p.chunks = chunks;
  p.offsets = offsets;
  p.queries = queries;
  p.bm25 = bm25_scores;
  p.dim = dim;
  p.k_rrf = k_rrf;
  p.q_hi = upper_percentile / 100.0;
  p.q_lo = lower_percentile / 100.0;
  p.out_cos = out_cos;
  p.out_rank_cos = out_rank_cos;
  p.out_rank_bm25 = out_rank_bm25;
  p.out_rrf = out_rrf;
  p.out_order = out_order;
  p.out_thr = out_thr;
  segmented_rank_rrf_kernel<<<n_groups, kRrfThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
