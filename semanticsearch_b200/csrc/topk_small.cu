// K9 — cosine top-k for many queries over a SMALL corpus (the reference's own scale: 100 queries x
// 10 000 chunks x 384 fp32, BASELINE config 1).
//
// Replaces the per-query loop of cosine_similarity(q, chunks)[0] + np.argsort(-s)[:k]
// (Tool/rank_chunks_optimized.py:215-216,225) when the corpus is L2-resident.  The streaming kernel
// (K1) is built for corpora that take milliseconds to read: on 15 MB its per-CTA prologue, per-warp
// candidate lists and list merges dominate (0.43 ms for config 1).  Here the work is two flat steps:
//
//   1. score tile kernel : S[q][r] = (q . c) / (|q| |c|) for a 64 x 64 (query, row) tile per CTA — fp32
//      FFMA with 4 x 4 register micro-tiles, K chunks of 16 through double-buffered shared memory,
//      both norms accumulated while streaming (sklearn zero rule), any input dtype;
//   2. select kernel     : one CTA per query finds the k-th largest score exactly with a 3-pass
//      (11/11/10-bit) radix select over its N scores (warp-aggregated histogram increments), gathers
//      everything above it plus the lowest-index ties in row order, and bitonic-sorts the k keys.
//
// The B x N score matrix lives in the caller's workspace (4 MB for config 1) and never leaves L2.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

constexpr int kSmTile = 64;
constexpr int kSmBK = 16;
constexpr int kSmPad = 4;
constexpr int kSelThreads = 512;
constexpr int kSelBins = 2048;

struct SmallParams {
  const void* corpus;
  const void* queries;
  int n_rows, n_queries, dim;
  float* scores;  // [n_queries][n_rows]
};

template <typename T>
__device__ __forceinline__ void load4(const T* base, int dim, bool row_ok, int k, float (&v)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = (row_ok && k + e < dim) ? to_f32<T>(base[k + e]) : 0.f;
}

template <typename TC, typename TQ>
__global__ void __launch_bounds__(256) small_scores_kernel(const SmallParams p) {
  __shared__ __align__(16) float As[2][kSmBK][kSmTile + kSmPad];  // queries
  __shared__ __align__(16) float Bs[2][kSmBK][kSmTile + kSmPad];  // corpus rows
  __shared__ float inv_a[kSmTile], inv_b[kSmTile];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int q0 = blockIdx.y * kSmTile, r0 = blockIdx.x * kSmTile;
  const bool qa_ok = q0 + lrow < p.n_queries, rb_ok = r0 + lrow < p.n_rows;
  const TQ* pa = static_cast<const TQ*>(p.queries) + static_cast<size_t>(qa_ok ? q0 + lrow : 0) * p.dim;
  const TC* pb = static_cast<const TC*>(p.corpus) + static_cast<size_t>(rb_ok ? r0 + lrow : 0) * p.dim;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ssq_a = 0.f, ssq_b = 0.f;
  const int nk = (p.dim + kSmBK - 1) / kSmBK;
  float va[4], vb[4];
  load4<TQ>(pa, p.dim, qa_ok, lk, va);
  load4<TC>(pb, p.dim, rb_ok, lk, vb);
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[buf][lk + e][lrow] = va[e];
      Bs[buf][lk + e][lrow] = vb[e];
      ssq_a = fmaf(va[e], va[e], ssq_a);
      ssq_b = fmaf(vb[e], vb[e], ssq_b);
    }
    __syncthreads();
    if (kc + 1 < nk) {
      load4<TQ>(pa, p.dim, qa_ok, (kc + 1) * kSmBK + lk, va);
      load4<TC>(pb, p.dim, rb_ok, (kc + 1) * kSmBK + lk, vb);
    }
#pragma unroll
    for (int kk = 0; kk < kSmBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  ssq_a += __shfl_xor_sync(0xffffffffu, ssq_a, 1);
  ssq_a += __shfl_xor_sync(0xffffffffu, ssq_a, 2);
  ssq_b += __shfl_xor_sync(0xffffffffu, ssq_b, 1);
  ssq_b += __shfl_xor_sync(0xffffffffu, ssq_b, 2);
  if ((tid & 3) == 0) {
    inv_a[lrow] = ssq_a > 0.f ? 1.0f / sqrtf(ssq_a) : 1.0f;  // sklearn: zero norm -> divide by 1
    inv_b[lrow] = ssq_b > 0.f ? 1.0f / sqrtf(ssq_b) : 1.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= p.n_queries) continue;
    float* out = p.scores + static_cast<size_t>(q) * p.n_rows;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + tx * 4 + j;
      if (r < p.n_rows) out[r] = (acc[i][j] * inv_a[ty * 4 + i]) * inv_b[tx * 4 + j];
    }
  }
}

// ---- per-query exact selection ------------------------------------------------------------------
__device__ __forceinline__ void sel_hist_add(unsigned int* hist, unsigned int bin, bool valid, int lane) {
  const unsigned int key = valid ? bin : 0xFFFFFFFFu;
  const unsigned int peers = __match_any_sync(0xffffffffu, key);
  if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], static_cast<unsigned int>(__popc(peers)));
}

__global__ void __launch_bounds__(kSelThreads) small_select_kernel(const float* __restrict__ scores, int n_rows, int k,
                                                                   uint32_t index_base, uint64_t* __restrict__ out_keys,
                                                                   float* __restrict__ out_scores, long long* __restrict__ out_indices) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  __shared__ unsigned int hist[kSelBins];
  __shared__ unsigned int s_prefix, s_want;
  __shared__ unsigned int warp_gt[kSelThreads / 32], warp_eq[kSelThreads / 32];
  uint64_t* sel = reinterpret_cast<uint64_t*>(sel_smem);  // [kpad2] selected keys
  const int q = blockIdx.x;
  const float* s = scores + static_cast<size_t>(q) * n_rows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kk = min(k, n_rows);  // rows that exist
  int kpad = 2;
  while (kpad < k) kpad <<= 1;

  // ---- k-th largest ordered score value by radix select (descending rank kk - 1) ----
  if (tid == 0) {
    s_prefix = 0u;
    s_want = static_cast<unsigned int>(kk - 1);  // 0-based rank from the top
  }
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  unsigned int known = 0u;
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = tid; i < kSelBins; i += kSelThreads) hist[i] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const int sh = shifts[pass];
    const unsigned int dmask = (1u << widths[pass]) - 1u;
    for (int i0 = 0; i0 < n_rows; i0 += kSelThreads) {  // warp-uniform trip count: the aggregation needs all lanes
      const int i = i0 + tid;
      float v = i < n_rows ? s[i] : 0.f;
      if (!(v == v)) v = -INFINITY;  // NaN ranks last, like make_key
      const unsigned int o = float_to_ordered(v);
      sel_hist_add(hist, (o >> sh) & dmask, i < n_rows && (o & known) == prefix, lane);
    }
    __syncthreads();
    if (warp == 0) {  // walk the bins from the top to the one holding the wanted rank
      const unsigned int want = s_want;
      unsigned int run = 0u;
      const int nb = 1 << widths[pass];
      for (int b0 = nb - 32; b0 >= 0; b0 -= 32) {
        const unsigned int c = hist[b0 + (31 - lane)];  // lane 0 = highest bin of the group
        unsigned int incl = c;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
          const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o2);
          if (lane >= o2) incl += up;
        }
        const unsigned int excl = run + incl - c;
        const bool here = (want >= excl) && (want < excl + c);
        const unsigned int bal = __ballot_sync(0xffffffffu, here);
        if (bal) {
          const int src = __ffs(bal) - 1;
          const unsigned int e = __shfl_sync(0xffffffffu, excl, src);
          if (lane == 0) {
            s_prefix = prefix | (static_cast<unsigned int>(b0 + (31 - src)) << sh);
            s_want = want - e;
          }
          break;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    known |= dmask << sh;
    __syncthreads();
  }
  const unsigned int T = s_prefix;  // ordered bits of the kk-th largest score

  // ---- ordered gather.  Every warp owns a contiguous range of rows: it counts its scores above / equal
  // to T, the block prefix-sums the counts once, and the warp then writes its survivors in row order —
  // everything above T, and the lowest-index rows equal to T for the remaining slots. ----
  const int per_warp = (n_rows + kSelThreads / 32 - 1) / (kSelThreads / 32);
  const int w_lo = min(n_rows, warp * per_warp), w_hi = min(n_rows, w_lo + per_warp);
  unsigned int c_gt = 0, c_eq = 0;
  for (int i = w_lo + lane; i < w_hi; i += 32) {
    float v = s[i];
    if (!(v == v)) v = -INFINITY;
    const unsigned int o = float_to_ordered(v);
    c_gt += o > T ? 1u : 0u;
    c_eq += o == T ? 1u : 0u;
  }
  c_gt = __reduce_add_sync(0xffffffffu, c_gt);
  c_eq = __reduce_add_sync(0xffffffffu, c_eq);
  if (lane == 0) {
    warp_gt[warp] = c_gt;
    warp_eq[warp] = c_eq;
  }
  for (int i = tid; i < kpad; i += kSelThreads) sel[i] = 0ull;
  __syncthreads();
  unsigned int n_gt = 0, off_gt = 0, off_eq = 0;
  for (int w = 0; w < kSelThreads / 32; ++w) {
    n_gt += warp_gt[w];
    if (w < warp) {
      off_gt += warp_gt[w];
      off_eq += warp_eq[w];
    }
  }
  const unsigned int eq_take = static_cast<unsigned int>(kk) - n_gt;
  const unsigned int lt = (1u << lane) - 1u;
  for (int i0 = w_lo; i0 < w_hi; i0 += 32) {
    const int i = i0 + lane;
    float v = i < w_hi ? s[i] : 0.f;
    if (!(v == v)) v = -INFINITY;
    const unsigned int o = float_to_ordered(v);
    const bool gt = i < w_hi && o > T, eq = i < w_hi && o == T;
    const unsigned int m_gt = __ballot_sync(0xffffffffu, gt), m_eq = __ballot_sync(0xffffffffu, eq);
    if (gt) sel[off_gt + __popc(m_gt & lt)] = make_key(v, index_base + static_cast<uint32_t>(i));
    if (eq) {
      const unsigned int pos = off_eq + __popc(m_eq & lt);
      if (pos < eq_take) sel[n_gt + pos] = make_key(v, index_base + static_cast<uint32_t>(i));
    }
    off_gt += __popc(m_gt);
    off_eq += __popc(m_eq);
  }
  __syncthreads();

  // ---- sort the kk keys (descending) and emit; slots past the corpus size stay empty ----
  for (int k2 = 2; k2 <= kpad; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (kpad >> 1); t += kSelThreads) {
        const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int b = a | j;
        const uint64_t x = sel[a], y = sel[b];
        const bool desc = (a & k2) == 0;
        if (desc ? (x < y) : (x > y)) {
          sel[a] = y;
          sel[b] = x;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += kSelThreads) {
    const uint64_t key = j < kk ? sel[j] : 0ull;
    const size_t o = static_cast<size_t>(q) * k + j;
    if (out_keys) out_keys[o] = key;
    if (out_scores) out_scores[o] = key ? key_score(key) : -INFINITY;
    if (out_indices) out_indices[o] = key ? key_index(key) : -1;
  }
}

template <typename TC>
static cudaError_t launch_small_scores(const SmallParams& p, int query_dtype, cudaStream_t st) {
  const dim3 grid((p.n_rows + kSmTile - 1) / kSmTile, (p.n_queries + kSmTile - 1) / kSmTile);
  switch (query_dtype) {
    case SS_F32: small_scores_kernel<TC, float><<<grid, 256, 0, st>>>(p); break;
    case SS_BF16: small_scores_kernel<TC, __nv_bfloat16><<<grid, 256, 0, st>>>(p); break;
    default: small_scores_kernel<TC, __half><<<grid, 256, 0, st>>>(p); break;
  }
  return cudaGetLastError();
}

}  // namespace ss

using namespace ss;

extern "C" size_t ss_cosine_topk_small_workspace_bytes(int64_t n_rows, int n_queries) {
  if (n_rows <= 0 || n_queries <= 0) return 0;
  return align_up(static_cast<size_t>(n_rows) * n_queries * 4, 256) + 256;
}

extern "C" int ss_cosine_topk_small(const void* corpus, int64_t n_rows, int dim, int corpus_dtype, const void* queries, int n_queries,
                                    int query_dtype, int k, uint32_t index_base, void* workspace, size_t workspace_bytes,
                                    uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream) {
  if (!corpus || !queries || !workspace) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_small: null pointer");
  if (n_rows <= 0 || dim <= 0 || n_queries <= 0 || k <= 0) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_small: sizes must be positive");
  if (!dtype_ok(corpus_dtype) || !dtype_ok(query_dtype)) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_small: bad dtype");
  if (k > 4096) return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_small: k > 4096 is not supported");
  if (n_rows > (1ll << 24) || static_cast<uint64_t>(index_base) + static_cast<uint64_t>(n_rows) > 0xFFFFFFFFull)
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_small: corpus too large for the small-corpus path");
  if (workspace_bytes < ss_cosine_topk_small_workspace_bytes(n_rows, n_queries))
    return fail(SS_ERR_WORKSPACE, "ss_cosine_topk_small: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SmallParams p;
  p.corpus = corpus;
  p.queries = queries;
  p.n_rows = static_cast<int>(n_rows);
  p.n_queries = n_queries;
  p.dim = dim;
  p.scores = reinterpret_cast<float*>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  cudaError_t e;
  {
    ProfileScope prof(st);
    switch (corpus_dtype) {
      case SS_F32: e = launch_small_scores<float>(p, query_dtype, st); break;
      case SS_BF16: e = launch_small_scores<__nv_bfloat16>(p, query_dtype, st); break;
      default: e = launch_small_scores<__half>(p, query_dtype, st); break;
    }
  }
  if (e != cudaSuccess) return cuda_fail(e, "small_scores launch");
  int kpad = 2;
  while (kpad < k) kpad <<= 1;
  small_select_kernel<<<n_queries, kSelThreads, static_cast<size_t>(kpad) * 8, st>>>(p.scores, p.n_rows, k, index_base, out_keys, out_scores,
                                                                               reinterpret_cast<long long*>(out_indices));
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
