// K10 — C99 divisive cut search on per-document rank matrices
// (Method/Semantic_Splitter_Optimized.py:194-238; the profile knee :239-264 stays with the caller).
//
// The reference evaluates, for every segment [a, b) of the current segmentation and every admissible cut
// a+m <= c <= b-m, gain = 0.5 * (mean R[a:c,a:c] + mean R[c:b,c:b]) - mean R[a:b,a:b] with one ndarray.mean()
// per block (O(n^2) work per candidate, ~2 s per 512-sentence document), takes the first best cut in
// (segment-list order, ascending c) order, splits, and repeats until the gain falls under
// max(min_gain, 0.1 * |mean of the segment|) or no candidate is left.
//
// One CTA per document:
//   1. a float64 summed-area table P of R is built in global memory (L2-resident, 2 MB at n = 512) with the
//      summation order of np.cumsum(np.cumsum(R, axis=0), axis=1): sequential down the columns (coalesced),
//      then sequential along the rows through 32 x 32 shared-memory tiles (coalesced loads/stores, one row per
//      lane).  For the default global rank matrix every entry is an integer < 2^11, so every block sum is exact.
//   2. each thread owns candidate positions c; it keeps its segment [a, b), the segment's position in the
//      reference's segment list and the gain of cutting at c in registers.  A split only invalidates the gains
//      of the positions inside the split segment; everything else is reused, so a round is one block arg-max
//      (gain desc, list position asc, c asc = the reference's first-best rule) plus O(segment) recomputation.
//   3. optionally the inside-density profile D(#segments) (:194-204,234) is produced per round: per-segment
//      totals in parallel, summed by one thread in ascending segment order like the reference's loop.
// All float64 arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction), so the cuts, their
// order and the profile are bit-identical to the float64 host statement of the same search.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

constexpr int kCutThreads = 256;
constexpr int kCutWarps = kCutThreads / 32;
constexpr int kCutTileLd = 33;
constexpr int kCutTile = 32 * kCutTileLd;                                   // doubles per warp
constexpr size_t kCutSmemBytes = static_cast<size_t>(kCutWarps) * kCutTile * 8;  // 67 584 B, reused by the search
constexpr int kCutMaxRows = 2048;

struct CutParams {
  const float* R;
  const int* offsets;
  const long long* s_offsets;
  const long long* sat_offsets;
  double* sat;
  const int* min_chunk;  // per document, or null -> min_chunk_all
  const int* max_cuts;   // per document (-1 = unlimited), or null -> max_cuts_all
  int min_chunk_all, max_cuts_all;
  double min_gain;
  int by_gain;
  int* out_cuts;         // [total_rows]: document d's cuts in pick order at offsets[d]
  int* out_n_cuts;       // [n_docs]; -1 = document not processed (too long, or min_chunk < 1)
  double* out_profile;   // [total_rows] or null: D_series of document d at offsets[d] (n_cuts + 1 values)
};

__device__ __forceinline__ double sat_total(const double* P, int ld, int a, int b) {
  const double bb = P[static_cast<size_t>(b) * ld + b], ab = P[static_cast<size_t>(a) * ld + b];
  const double ba = P[static_cast<size_t>(b) * ld + a], aa = P[static_cast<size_t>(a) * ld + a];
  return __dadd_rn(__dsub_rn(__dsub_rn(bb, ab), ba), aa);
}
__device__ __forceinline__ double sat_mean(const double* P, int ld, int a, int b) {
  const int len = b - a;
  return __ddiv_rn(sat_total(P, ld, a, b), static_cast<double>(len * len));
}

// (gain desc, list position asc, cut asc): true when x is better than y.
struct CutBest {
  double gain;
  int idx, c, a, b;
};
__device__ __forceinline__ bool cut_better(const CutBest& x, const CutBest& y) {
  if (x.gain != y.gain) return x.gain > y.gain;
  if (x.idx != y.idx) return x.idx < y.idx;
  return x.c < y.c;
}
__device__ __forceinline__ CutBest cut_shfl_down(const CutBest& v, int d) {
  CutBest o;
  o.gain = __shfl_down_sync(0xffffffffu, v.gain, d);
  o.idx = __shfl_down_sync(0xffffffffu, v.idx, d);
  o.c = __shfl_down_sync(0xffffffffu, v.c, d);
  o.a = __shfl_down_sync(0xffffffffu, v.a, d);
  o.b = __shfl_down_sync(0xffffffffu, v.b, d);
  return o;
}

template <int CPT>
__global__ void __launch_bounds__(kCutThreads) c99_divisive_kernel(const CutParams p) {
  extern __shared__ __align__(16) unsigned char cut_smem[];
  __shared__ CutBest w_best[kCutWarps];
  __shared__ int dec[5];  // stop, list position, cut, a, b
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int doc = blockIdx.x;
  const int row_base = p.offsets[doc];
  const int n = p.offsets[doc + 1] - row_base;
  const int m = p.min_chunk ? p.min_chunk[doc] : p.min_chunk_all;
  const int max_cuts = p.max_cuts ? p.max_cuts[doc] : p.max_cuts_all;
  if (m < 1 || n > CPT * kCutThreads) {
    if (tid == 0) p.out_n_cuts[doc] = -1;
    return;
  }
  if (n < 2 * m) {  // :165-166
    if (tid == 0) p.out_n_cuts[doc] = 0;
    return;
  }
  const float* __restrict__ R = p.R + p.s_offsets[doc];
  double* P = p.sat + p.sat_offsets[doc];  // written and re-read by this CTA: no __restrict__ / read-only path
  const int ld = n + 1;

  // ---- summed-area table: P[i+1][j+1] = sum R[0..i][0..j] ------------------------------------------------
  for (int t = tid; t <= n; t += kCutThreads) {
    P[t] = 0.0;
    P[static_cast<size_t>(t) * ld] = 0.0;
  }
  for (int j = tid; j < n; j += kCutThreads) {  // np.cumsum(axis=0): sequential down column j
    double acc = 0.0;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = R[static_cast<size_t>(i + u) * n + j];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc = __dadd_rn(acc, static_cast<double>(v[u]));
        P[static_cast<size_t>(i + u + 1) * ld + j + 1] = acc;
      }
    }
    for (; i < n; ++i) {
      acc = __dadd_rn(acc, static_cast<double>(R[static_cast<size_t>(i) * n + j]));
      P[static_cast<size_t>(i + 1) * ld + j + 1] = acc;
    }
  }
  __syncthreads();
  {
    double* tile = reinterpret_cast<double*>(cut_smem) + warp * kCutTile;
    for (int g = warp; g * 32 < n; g += kCutWarps) {  // np.cumsum(axis=1): sequential along each row, 32 rows per warp
      const int row0 = g * 32;
      const int rows = min(32, n - row0);
      double carry = 0.0;
      for (int jc = 0; jc < n; jc += 32) {
        const int cols = min(32, n - jc);
        if (lane < cols)
          for (int r = 0; r < rows; ++r) tile[r * kCutTileLd + lane] = P[static_cast<size_t>(row0 + r + 1) * ld + jc + lane + 1];
        __syncwarp();
        if (lane < rows)
          for (int c = 0; c < cols; ++c) {
            carry = __dadd_rn(carry, tile[lane * kCutTileLd + c]);
            tile[lane * kCutTileLd + c] = carry;
          }
        __syncwarp();
        if (lane < cols)
          for (int r = 0; r < rows; ++r) P[static_cast<size_t>(row0 + r + 1) * ld + jc + lane + 1] = tile[r * kCutTileLd + lane];
        __syncwarp();
      }
    }
  }
  __syncthreads();

  // ---- divisive search -------------------------------------------------------------------------------------
  int* bnd[2] = {reinterpret_cast<int*>(cut_smem), reinterpret_cast<int*>(cut_smem) + (kCutMaxRows + 8)};  // sorted boundaries
  double* seg_tot = reinterpret_cast<double*>(cut_smem + 2 * (kCutMaxRows + 8) * sizeof(int));
  int cur = 0, n_bnd = 2;
  const bool want_profile = p.out_profile != nullptr;
  if (tid == 0) {
    bnd[0][0] = 0;
    bnd[0][1] = n;
    if (want_profile) p.out_profile[row_base] = __ddiv_rn(__dadd_rn(0.0, sat_total(P, ld, 0, n)), static_cast<double>(static_cast<long long>(n) * n));
  }
  int seg_a[CPT], seg_b[CPT], seg_idx[CPT];
  double gain[CPT];
  unsigned alive = 0u, dirty = 0u;
#pragma unroll
  for (int r = 0; r < CPT; ++r) {
    const int c = tid + r * kCutThreads;
    seg_a[r] = 0;
    seg_b[r] = n;
    seg_idx[r] = 0;
    gain[r] = -INFINITY;
    if (c > 0 && c < n) {
      alive |= 1u << r;
      dirty |= 1u << r;
    }
  }
  int n_segs = 1, n_cuts = 0;
  for (;;) {
    CutBest best;
    best.gain = -INFINITY;
    best.idx = 0x7fffffff;
    best.c = 0x7fffffff;
    best.a = best.b = 0;
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
      if (!((alive >> r) & 1u)) continue;
      const int c = tid + r * kCutThreads;
      const int a = seg_a[r], b = seg_b[r];
      if ((dirty >> r) & 1u) {
        double g = -INFINITY;
        if (b - a >= 2 * m && c >= a + m && c <= b - m) {  // :210-214
          const double whole = sat_mean(P, ld, a, b);
          g = __dsub_rn(__dmul_rn(0.5, __dadd_rn(sat_mean(P, ld, a, c), sat_mean(P, ld, c, b))), whole);  // :219
        }
        gain[r] = g;
      }
      CutBest mine;
      mine.gain = gain[r];
      mine.idx = seg_idx[r];
      mine.c = c;
      mine.a = a;
      mine.b = b;
      if (mine.gain > -INFINITY && cut_better(mine, best)) best = mine;
    }
    dirty = 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const CutBest o = cut_shfl_down(best, d);
      if (cut_better(o, best)) best = o;
    }
    if (lane == 0) w_best[warp] = best;
    __syncthreads();
    if (warp == 0) {
      if (lane < kCutWarps) best = w_best[lane];
      else {
        best.gain = -INFINITY;
        best.idx = best.c = 0x7fffffff;
      }
#pragma unroll
      for (int d = kCutWarps / 2; d > 0; d >>= 1) {
        const CutBest o = cut_shfl_down(best, d);
        if (cut_better(o, best)) best = o;
      }
      if (lane == 0) {
        bool stop = !(best.gain > -INFINITY) || (max_cuts >= 0 && n_cuts >= max_cuts);  // :225
        if (!stop && p.by_gain) {
          const double floor_gain = fmax(p.min_gain, __dmul_rn(0.1, fabs(sat_mean(P, ld, best.a, best.b))));  // :224
          stop = best.gain < floor_gain;                                                                      // :228
        }
        dec[0] = stop ? 1 : 0;
        dec[1] = best.idx;
        dec[2] = best.c;
        dec[3] = best.a;
        dec[4] = best.b;
        if (!stop) p.out_cuts[row_base + n_cuts] = best.c;
      }
    }
    __syncthreads();
    if (dec[0]) break;
    const int s_idx = dec[1], pos = dec[2];
    // the split segment leaves the list, its halves are appended: [a, pos) at n_segs - 1, [pos, b) at n_segs (:230-231)
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
      if (!((alive >> r) & 1u)) continue;
      const int c = tid + r * kCutThreads;
      if (seg_idx[r] == s_idx) {
        if (c < pos) {
          seg_b[r] = pos;
          seg_idx[r] = n_segs - 1;
          dirty |= 1u << r;
        } else if (c > pos) {
          seg_a[r] = pos;
          seg_idx[r] = n_segs;
          dirty |= 1u << r;
        } else {
          alive &= ~(1u << r);  // a boundary is never a candidate again
        }
      } else if (seg_idx[r] > s_idx) {
        seg_idx[r] -= 1;
      }
    }
    ++n_segs;
    ++n_cuts;
    if (want_profile) {  // D_series.append(_inside_density(R, sorted(segs)))  (:234)
      const int* src = bnd[cur];
      int* dst = bnd[cur ^ 1];
      for (int t = tid; t < n_bnd; t += kCutThreads) {
        const int v = src[t];
        dst[t + (v > pos ? 1 : 0)] = v;
        if (v < pos && src[t + 1] > pos) dst[t + 1] = pos;  // src[n_bnd - 1] = n > pos, so t + 1 is in range here
      }
      cur ^= 1;
      ++n_bnd;
      __syncthreads();
      const int* bs = bnd[cur];
      for (int t = tid; t < n_bnd - 1; t += kCutThreads) seg_tot[t] = sat_total(P, ld, bs[t], bs[t + 1]);
      __syncthreads();
      if (tid == 0) {
        double tot = 0.0;
        long long area = 0;
        for (int t = 0; t < n_bnd - 1; ++t) {
          const long long len = bs[t + 1] - bs[t];
          tot = __dadd_rn(tot, seg_tot[t]);
          area += len * len;
        }
        p.out_profile[row_base + n_cuts] = __ddiv_rn(tot, static_cast<double>(area));
      }
    }
    // dec[] is rewritten only after the next round's first barrier, which every thread reaches after reading it
  }
  if (tid == 0) p.out_n_cuts[doc] = n_cuts;
}

}  // namespace ss

using namespace ss;

extern "C" int ss_c99_divisive_cuts(const float* R, const int32_t* offsets, const int64_t* s_offsets, const int64_t* sat_offsets,
                                    int n_docs, int max_doc_rows, const int32_t* min_chunk, int min_chunk_all,
                                    const int32_t* max_cuts, int max_cuts_all, double min_gain, int stop_by_gain,
                                    double* sat_workspace, int32_t* out_cuts, int32_t* out_n_cuts, double* out_profile, void* stream) {
  if (!R || !offsets || !s_offsets || !sat_offsets || !sat_workspace || !out_cuts || !out_n_cuts)
    return fail(SS_ERR_INVALID_ARG, "ss_c99_divisive_cuts: null pointer");
  if (n_docs <= 0 || max_doc_rows <= 0) return fail(SS_ERR_INVALID_ARG, "ss_c99_divisive_cuts: sizes must be positive");
  if (!min_chunk && min_chunk_all < 1) return fail(SS_ERR_INVALID_ARG, "ss_c99_divisive_cuts: min_chunk must be >= 1");
  if (max_doc_rows > kCutMaxRows) return fail(SS_ERR_UNSUPPORTED, "ss_c99_divisive_cuts: documents longer than 2048 sentences are not supported");
  CutParams p;
  p.R = R;
  p.offsets = offsets;
  p.s_offsets = reinterpret_cast<const long long*>(s_offsets);
  p.sat_offsets = reinterpret_cast<const long long*>(sat_offsets);
  p.sat = sat_workspace;
  p.min_chunk = min_chunk;
  p.max_cuts = max_cuts;
  p.min_chunk_all = min_chunk_all;
  p.max_cuts_all = max_cuts_all;
  p.min_gain = min_gain;
  p.by_gain = stop_by_gain ? 1 : 0;
  p.out_cuts = out_cuts;
  p.out_n_cuts = out_n_cuts;
  p.out_profile = out_profile;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto launch = [&](auto kernel) -> int {
    SS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCutSmemBytes)));
    kernel<<<n_docs, kCutThreads, kCutSmemBytes, st>>>(p);
    SS_CUDA_CHECK(cudaGetLastError());
    return SS_OK;
  };
  if (max_doc_rows <= 2 * kCutThreads) return launch(c99_divisive_kernel<2>);
  if (max_doc_rows <= 4 * kCutThreads) return launch(c99_divisive_kernel<4>);
  return launch(c99_divisive_kernel<8>);
}
