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
//   1. a float64 summed-area table P of R is built in global memory (2 MB at n = 512) in ONE sweep over 32 x 32
//      tiles along anti-diagonals, with the summation order of np.cumsum(np.cumsum(R, axis=0), axis=1): inside a
//      tile one lane per column runs down the rows, then one lane per row runs along the columns; column and row
//      carries pass between tiles through shared memory, one barrier per diagonal.  R is read once, P written
//      once.  For the default global rank matrix every entry is an integer < 2^11, so every block sum is exact.
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
constexpr int kCutTile = 32 * 32;  // doubles per warp
// tiles (64 KB, reused by the search) + column and row carries of the table sweep
constexpr size_t cut_smem_bytes(int cpt) { return static_cast<size_t>(kCutWarps) * kCutTile * 8 + 2 * static_cast<size_t>(cpt) * kCutThreads * 8; }
constexpr int kCutMaxRows = 4096;  // 16 candidate cuts per thread; the reference corpus's longest document has 3939 sentences

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
  double gain, whole;  // whole = mean of the candidate's segment (the stop rule needs the winner's)
  int idx, c;
};
__device__ __forceinline__ bool cut_better(const CutBest& x, const CutBest& y) {
  if (x.gain != y.gain) return x.gain > y.gain;
  if (x.idx != y.idx) return x.idx < y.idx;
  return x.c < y.c;
}
__device__ __forceinline__ CutBest cut_shfl_down(const CutBest& v, int d) {
  CutBest o;
  o.gain = __shfl_down_sync(0xffffffffu, v.gain, d);
  o.whole = __shfl_down_sync(0xffffffffu, v.whole, d);
  o.idx = __shfl_down_sync(0xffffffffu, v.idx, d);
  o.c = __shfl_down_sync(0xffffffffu, v.c, d);
  return o;
}

template <int CPT>
__global__ void __launch_bounds__(kCutThreads, CPT == 2 ? 3 : 1) c99_divisive_kernel(const CutParams p) {
  extern __shared__ __align__(16) unsigned char cut_smem[];
  __shared__ CutBest w_best[kCutWarps];
  __shared__ int dec[3];  // stop, list position, cut
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
  // One sweep over 32 x 32 tiles along anti-diagonals: tile (b, jc) continues column sums from tile (b-1, jc) and row
  // sums from tile (b, jc-1), both finished one diagonal earlier, so R is read once and P written once while every
  // addition happens in np.cumsum(np.cumsum(R, 0), 1) order.  The warp's next tile of R is prefetched into registers.
  for (int t = tid; t <= n; t += kCutThreads) {
    P[t] = 0.0;
    P[static_cast<size_t>(t) * ld] = 0.0;
  }
  {
    const int T = (n + 31) >> 5;
    double* tile = reinterpret_cast<double*>(cut_smem) + warp * kCutTile;  // [32][32], column index XOR row (conflict-free both ways)
    double* colcarry = reinterpret_cast<double*>(cut_smem) + kCutWarps * kCutTile;
    double* rowcarry = colcarry + CPT * kCutThreads;
    float v[32], vn[32];
    auto load_tile = [&](float (&dst)[32], int tb, int tjc) {
      const int row0 = tb * 32, col = tjc * 32 + lane;
      const int rows = min(32, n - row0);
#pragma unroll
      for (int r = 0; r < 32; ++r) dst[r] = (r < rows && col < n) ? R[static_cast<size_t>(row0 + r) * n + col] : 0.f;
    };
    bool have = false;
    for (int d = 0; d < 2 * T - 1; ++d) {
      const int b_lo = max(0, d - T + 1);
      const int cnt = min(d, T - 1) - b_lo + 1;
      for (int t = warp; t < cnt; t += kCutWarps) {
        const int tb = b_lo + t, tjc = d - tb;
        if (!have) load_tile(v, tb, tjc);
        int nb, njc;
        bool nvalid;
        if (t + kCutWarps < cnt) {
          nb = tb + kCutWarps;
          njc = d - nb;
          nvalid = true;
        } else {
          const int d2 = d + 1;
          const int b_lo2 = max(0, d2 - T + 1);
          nvalid = d2 < 2 * T - 1 && warp < min(d2, T - 1) - b_lo2 + 1;
          nb = b_lo2 + warp;
          njc = d2 - nb;
        }
        if (nvalid) load_tile(vn, nb, njc);
        const int row0 = tb * 32, col0 = tjc * 32;
        const int rows = min(32, n - row0), cols = min(32, n - col0);
        if (lane < cols) {  // np.cumsum(axis=0): lane = column, sequential down the rows
          double cc = tb ? colcarry[col0 + lane] : 0.0;
#pragma unroll
          for (int r = 0; r < 32; ++r)
            if (r < rows) {
              cc = __dadd_rn(cc, static_cast<double>(v[r]));
              tile[r * 32 + (lane ^ r)] = cc;
            }
          colcarry[col0 + lane] = cc;
        }
        __syncwarp();
        if (lane < rows) {  // np.cumsum(axis=1): lane = row, sequential along the columns
          double rc = tjc ? rowcarry[row0 + lane] : 0.0;
          for (int c = 0; c < cols; ++c) {
            rc = __dadd_rn(rc, tile[lane * 32 + (c ^ lane)]);
            tile[lane * 32 + (c ^ lane)] = rc;
          }
          rowcarry[row0 + lane] = rc;
        }
        __syncwarp();
        if (lane < cols)
          for (int r = 0; r < rows; ++r) P[static_cast<size_t>(row0 + r + 1) * ld + col0 + lane + 1] = tile[r * 32 + (lane ^ r)];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = vn[r];
        have = nvalid;
      }
      __syncthreads();  // the carries of this diagonal are visible to the next one
    }
  }

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
  double gain[CPT], whole_of[CPT];
  unsigned alive = 0u, dirty = 0u;
#pragma unroll
  for (int r = 0; r < CPT; ++r) {
    const int c = tid + r * kCutThreads;
    seg_a[r] = 0;
    seg_b[r] = n;
    seg_idx[r] = 0;
    gain[r] = -INFINITY;
    whole_of[r] = 0.0;
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
    best.whole = 0.0;
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
          whole_of[r] = whole;
        }
        gain[r] = g;
      }
      CutBest mine;
      mine.gain = gain[r];
      mine.whole = whole_of[r];
      mine.idx = seg_idx[r];
      mine.c = c;
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
          const double floor_gain = fmax(p.min_gain, __dmul_rn(0.1, fabs(best.whole)));  // :224
          stop = best.gain < floor_gain;                                                                      // :228
        }
        dec[0] = stop ? 1 : 0;
        dec[1] = best.idx;
        dec[2] = best.c;
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
  if (max_doc_rows > kCutMaxRows) return fail(SS_ERR_UNSUPPORTED, "ss_c99_divisive_cuts: documents longer than 4096 sentences are not supported");
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
  auto launch = [&](auto kernel, size_t smem) -> int {
    SS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kernel<<<n_docs, kCutThreads, smem, st>>>(p);
    SS_CUDA_CHECK(cudaGetLastError());
    return SS_OK;
  };
  if (max_doc_rows <= 2 * kCutThreads) return launch(c99_divisive_kernel<2>, cut_smem_bytes(2));
  if (max_doc_rows <= 4 * kCutThreads) return launch(c99_divisive_kernel<4>, cut_smem_bytes(4));
  if (max_doc_rows <= 8 * kCutThreads) return launch(c99_divisive_kernel<8>, cut_smem_bytes(8));
  return launch(c99_divisive_kernel<16>, cut_smem_bytes(16));
}
