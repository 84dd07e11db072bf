// K5 — semantic-splitter similarity passes over ragged documents.
//
//  (a) ss_segmented_adjacent_cosine: adj[r] = cos(E[r], E[r+1]) for every consecutive row pair of
//      the concatenated sentence-embedding matrix, one streaming pass (rows are L2-normalised on
//      the fly; zero rows give 0).  Replaces `_embed` + the per-sentence dot loop at
//      Method/Semantic_Splitter_Optimized.py:140-152,412.
//  (b) ss_segmented_percentile: per document, distance d = 1 - adj, the P-th percentile threshold
//      (numpy "linear" interpolation) and breakpoint flags d > thr (BASELINE.json config 3), plus
//      the robust statistics of Splitter:417-437 (median-of-3 smoothing :340-356, median, MAD,
//      P25/P75) computed from exact order statistics of an in-shared-memory bitonic sort.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

// ------------------------------------------------------------------------------------------
// (a) adjacent cosine — same TMA bulk-copy ring as K1, tiles overlap by one row
// ------------------------------------------------------------------------------------------
constexpr int kAdjConsumerWarps = 16;
constexpr int kAdjThreads = (kAdjConsumerWarps + 1) * 32;
constexpr int kAdjMaxStages = 8;
constexpr int kAdjR = 3;  // adjacent pairs per warp iteration: 3 dots + 4 sums of squares = 7 values

struct AdjParams {
  const void* rows;
  long long n_rows;
  int dim;
  int tile_rows;  // pairs per tile; the tile loads tile_rows + 1 rows
  long long n_tiles;
  int stages;
  uint32_t tile_bytes;
  float* out;  // [n_rows]
};

template <typename T>
__global__ void __launch_bounds__(kAdjThreads, 1) adjacent_cosine_kernel(const AdjParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NP = Pairs<T>::NP;
  constexpr int R = kAdjR;
  constexpr int NV = 2 * R + 1;
  constexpr int NVP = 8;
  constexpr int SH = 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = static_cast<uint32_t>(p.dim) * sizeof(T);
  const int chunks = row_bytes / 16;
  unsigned char* tiles = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(p.stages) * p.tile_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kAdjMaxStages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kAdjConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kAdjConsumerWarps) {
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        mbar_wait(&empty_bar[s], phase ^ 1u);
        const long long row0 = t * p.tile_rows;
        const long long rows = min(static_cast<long long>(p.tile_rows) + 1, p.n_rows - row0);
        const uint32_t bytes = static_cast<uint32_t>(rows) * row_bytes;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_copy_g2s(tiles + static_cast<size_t>(s) * p.tile_bytes,
                      static_cast<const unsigned char*>(p.rows) + static_cast<size_t>(row0) * row_bytes, bytes, &full_bar[s]);
        if (++s == p.stages) {
          s = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    const int cw = warp;
    const int vidx = lane >> SH;  // 0..R-1: dot(row i, row i+1); R..2R: ssq(row i)
    const int groups_per_tile = (p.tile_rows + R - 1) / R;
    int s = 0;
    uint32_t phase = 0;
    long long it = 0;
    for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      mbar_wait(&full_bar[s], phase);
      const long long row0 = t * p.tile_rows;
      const int rows = static_cast<int>(min(static_cast<long long>(p.tile_rows) + 1, p.n_rows - row0));  // loaded rows
      const int pairs = min(p.tile_rows, rows - 1);  // pairs this tile owns (the last row of the matrix owns none)
      const unsigned char* tile = tiles + static_cast<size_t>(s) * p.tile_bytes;
      const int ngroups = (pairs + R - 1) / R;
      int g = static_cast<int>((cw + kAdjConsumerWarps - (it * groups_per_tile) % kAdjConsumerWarps) % kAdjConsumerWarps);
      for (; g < ngroups; g += kAdjConsumerWarps) {
        const int r0 = g * R;
        const uint4* rowp[R + 1];
#pragma unroll
        for (int rr = 0; rr <= R; ++rr)
          rowp[rr] = reinterpret_cast<const uint4*>(tile + static_cast<size_t>(min(r0 + rr, rows - 1)) * row_bytes);
        unsigned long long acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0ull;
        for (int c = lane; c < chunks; c += 32) {
          unsigned long long x[R + 1][NP];
#pragma unroll
          for (int rr = 0; rr <= R; ++rr) {
            const uint4 raw = rowp[rr][c];
            Pairs<T>::unpack(raw, x[rr]);
          }
#pragma unroll
          for (int e = 0; e < NP; ++e) {
#pragma unroll
            for (int rr = 0; rr <= R; ++rr) acc[R + rr] = ffma2(x[rr][e], x[rr][e], acc[R + rr]);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) acc[rr] = ffma2(x[rr][e], x[rr + 1][e], acc[rr]);
          }
        }
        float vals[NVP];
#pragma unroll
        for (int i = 0; i < NVP; ++i) vals[i] = (i < NV) ? sum2(acc[i]) : 0.f;
        const float mine = transpose_reduce<NVP>(vals, lane);
        const int i = vidx < R ? vidx : 0;
        const float ssq_a = __shfl_sync(0xffffffffu, mine, (R + i) << SH);
        const float ssq_b = __shfl_sync(0xffffffffu, mine, (R + i + 1) << SH);
        if (vidx < R && (lane & ((1 << SH) - 1)) == 0 && r0 + vidx < pairs) {
          const float inv_a = ssq_a > 0.f ? 1.0f / sqrtf(ssq_a) : 0.f;
          const float inv_b = ssq_b > 0.f ? 1.0f / sqrtf(ssq_b) : 0.f;
          p.out[row0 + r0 + vidx] = (mine * inv_a) * inv_b;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (++s == p.stages) {
        s = 0;
        phase ^= 1u;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) p.out[p.n_rows - 1] = 0.f;
}

// Plain-load fallback for rows that are not 16-byte multiples.
template <typename T>
__global__ void __launch_bounds__(256) adjacent_cosine_generic_kernel(const T* __restrict__ rows, long long n_rows, int dim,
                                                                      float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp; r < n_rows; r += nwarps) {
    if (r == n_rows - 1) {
      if (lane == 0) out[r] = 0.f;
      continue;
    }
    const T* a = rows + static_cast<size_t>(r) * dim;
    const T* b = a + dim;
    float dot = 0.f, sa = 0.f, sb = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float x = to_f32<T>(a[c]), y = to_f32<T>(b[c]);
      dot = fmaf(x, y, dot);
      sa = fmaf(x, x, sa);
      sb = fmaf(y, y, sb);
    }
    dot = warp_sum(dot);
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    if (lane == 0) {
      const float ia = sa > 0.f ? 1.0f / sqrtf(sa) : 0.f;
      const float ib = sb > 0.f ? 1.0f / sqrtf(sb) : 0.f;
      out[r] = (dot * ia) * ib;
    }
  }
}

// ------------------------------------------------------------------------------------------
// (b) per-document order statistics
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t f64_to_ordered(double d) {
  const uint64_t b = static_cast<uint64_t>(__double_as_longlong(d));
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_to_f64(uint64_t o) {
  const uint64_t b = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
  return __longlong_as_double(static_cast<long long>(b));
}

// numpy's linear-interpolation quantile on a sorted array (numpy/lib/_function_base_impl.py
// `_quantile` + `_lerp`): virtual index (m-1)*q, gamma = frac, a + (b-a)*g, and for g >= 0.5
// b - (b-a)*(1-g).
__device__ __forceinline__ double np_quantile_sorted(const uint64_t* sorted, int m, double q) {
  const double vi = __dmul_rn(static_cast<double>(m - 1), q);
  double fl = floor(vi);
  int lo = static_cast<int>(fl);
  lo = max(0, min(lo, m - 1));
  const int hi = min(lo + 1, m - 1);
  const double g = __dsub_rn(vi, fl);
  const double a = ordered_to_f64(sorted[lo]), b = ordered_to_f64(sorted[hi]);
  // numpy rounds the product and the sum separately: keep the compiler from fusing them into an FMA
  const double diff = __dsub_rn(b, a);
  double r = __dadd_rn(a, __dmul_rn(diff, g));
  if (g >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, g)));
  return r;
}
__device__ __forceinline__ double np_median_sorted(const uint64_t* sorted, int m) {
  if (m & 1) return ordered_to_f64(sorted[m / 2]);
  return __dadd_rn(ordered_to_f64(sorted[m / 2 - 1]), ordered_to_f64(sorted[m / 2])) / 2.0;  // np.mean of the two middle values
}

// In-place ascending bitonic sort of `n2` (power of two) keys in shared memory.
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* a, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Short documents (at most 33 sentences = 32 adjacent pairs — most of the reference corpus): one WARP per
// document, one pair per lane, order statistics from a shuffle bitonic sort.  Same arithmetic and outputs
// as the CTA kernel below.
// ------------------------------------------------------------------------------------------------
constexpr int kPctSmallWarps = 8;

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// ascending sort of one key per lane
__device__ __forceinline__ uint64_t warp_sort_asc_u64(uint64_t v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = shfl_u64(v, lane ^ j);
      const bool up = (lane & k) == 0;           // this block sorts ascending
      const bool lower = (lane & j) == 0;        // this lane keeps the smaller key of an ascending pair
      const bool take_min = lower == up;
      v = take_min ? (v < o ? v : o) : (v > o ? v : o);
    }
  }
  return v;
}
// np.quantile (linear) / np.median on the warp-sorted keys (lane i holds order statistic i); all lanes get the result
__device__ __forceinline__ double warp_quantile(uint64_t sorted, int m, double q) {
  const double vi = __dmul_rn(static_cast<double>(m - 1), q);
  const double fl = floor(vi);
  int lo = static_cast<int>(fl);
  lo = max(0, min(lo, m - 1));
  const int hi = min(lo + 1, m - 1);
  const double g = __dsub_rn(vi, fl);
  const double a = ordered_to_f64(shfl_u64(sorted, lo)), b = ordered_to_f64(shfl_u64(sorted, hi));
  const double diff = __dsub_rn(b, a);
  double r = __dadd_rn(a, __dmul_rn(diff, g));
  if (g >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, g)));
  return r;
}
__device__ __forceinline__ double warp_median(uint64_t sorted, int m) {
  if (m & 1) return ordered_to_f64(shfl_u64(sorted, m / 2));
  return __dadd_rn(ordered_to_f64(shfl_u64(sorted, m / 2 - 1)), ordered_to_f64(shfl_u64(sorted, m / 2))) / 2.0;
}

__global__ void __launch_bounds__(kPctSmallWarps * 32) segmented_percentile_small_kernel(
    float* __restrict__ adj, const int* __restrict__ offsets, int n_docs, double q, double* __restrict__ out_thr,
    unsigned char* __restrict__ out_flags, double* __restrict__ out_stats, float* __restrict__ out_smooth) {
  const int lane = threadIdx.x & 31;
  const int doc = blockIdx.x * kPctSmallWarps + (threadIdx.x >> 5);
  if (doc >= n_docs) return;
  const int a0 = offsets[doc], a1 = offsets[doc + 1];
  const int m = a1 - a0 - 1;  // adjacent pairs
  if (m < 1 || m > 32) return;  // 0/1-sentence and longer documents belong to the CTA kernel
  const bool live = lane < m;
  const float x = live ? adj[a0 + lane] : 0.f;
  const double d = 1.0 - static_cast<double>(x);
  const uint64_t sorted = warp_sort_asc_u64(live ? f64_to_ordered(d) : ~0ull, lane);
  const double thr = warp_quantile(sorted, m, q);
  if (lane == 0) {
    out_thr[doc] = thr;
    adj[a1 - 1] = 0.f;  // the slot of a document's last sentence held a cross-document cosine
    if (out_flags) out_flags[a1 - 1] = 0;
    if (out_smooth) out_smooth[a1 - 1] = 0.f;
  }
  if (out_flags && live) out_flags[a0 + lane] = d > thr ? 1 : 0;
  if (!out_stats && !out_smooth) return;
  // median-of-3 smoothing with edge replication (Splitter:340-356); unchanged if m < 3
  double v = static_cast<double>(x);
  {
    const float xl = __shfl_sync(0xffffffffu, x, max(lane - 1, 0));
    const float xr = __shfl_sync(0xffffffffu, x, min(lane + 1, m - 1));
    if (m >= 3) {
      const double l = static_cast<double>(xl), r = static_cast<double>(xr);
      v = fmax(fmin(l, v), fmin(fmax(l, v), r));
    }
  }
  if (out_smooth && live) out_smooth[a0 + lane] = static_cast<float>(v);
  if (!out_stats) return;
  const uint64_t s2 = warp_sort_asc_u64(live ? f64_to_ordered(v) : ~0ull, lane);
  const double med = warp_median(s2, m);
  const double p25 = warp_quantile(s2, m, 0.25), p75 = warp_quantile(s2, m, 0.75);
  const uint64_t s3 = warp_sort_asc_u64(live ? f64_to_ordered(fabs(v - med)) : ~0ull, lane);
  const double mad = warp_median(s3, m) + 1e-9;
  if (lane == 0) {
    double* st = out_stats + static_cast<size_t>(doc) * 4;
    st[0] = med;
    st[1] = mad;
    st[2] = p25;
    st[3] = p75;
  }
}

// One CTA per document.  Shared: keys[n2] (u64) + vals[n2] (double).
__device__ __forceinline__ void percentile_doc(float* __restrict__ adj, const int* __restrict__ offsets, int doc, double q,
                                               uint64_t* keys, double* base, double* s_med_p, double* __restrict__ out_thr,
                                               unsigned char* __restrict__ out_flags, double* __restrict__ out_stats,
                                               float* __restrict__ out_smooth) {
  double& s_med = *s_med_p;
  const int a0 = offsets[doc], a1 = offsets[doc + 1];
  const int n = a1 - a0;
  const int m = n - 1;  // adjacent pairs
  if (m >= 1 && m <= 32) return;  // handled by segmented_percentile_small_kernel (one warp per document)
  if (out_flags)
    for (int i = threadIdx.x; i < n; i += blockDim.x) out_flags[a0 + i] = 0;
  if (out_smooth)
    for (int i = threadIdx.x; i < n; i += blockDim.x) out_smooth[a0 + i] = (i < m) ? adj[a0 + i] : 0.f;
  if (m < 1) {
    if (threadIdx.x == 0) {
      if (n == 1) adj[a0] = 0.f;
      out_thr[doc] = __longlong_as_double(0x7FF8000000000000ll);
      if (out_stats)
        for (int j = 0; j < 4; ++j) out_stats[static_cast<size_t>(doc) * 4 + j] = 0.0;
    }
    return;
  }
  int n2 = 1;
  while (n2 < m) n2 <<= 1;

  // ---- P-th percentile of d = 1 - adj (fp64, exact) and breakpoint flags -------------------
  for (int i = threadIdx.x; i < n2; i += blockDim.x)
    keys[i] = (i < m) ? f64_to_ordered(1.0 - static_cast<double>(adj[a0 + i])) : ~0ull;
  __syncthreads();
  bitonic_sort_smem(keys, n2);
  const double thr = np_quantile_sorted(keys, m, q);
  __syncthreads();
  if (threadIdx.x == 0) {
    out_thr[doc] = thr;
    adj[a1 - 1] = 0.f;  // the slot of a document's last sentence held a cross-document cosine
  }
  if (out_flags)
    for (int i = threadIdx.x; i < m; i += blockDim.x)
      out_flags[a0 + i] = (1.0 - static_cast<double>(adj[a0 + i])) > thr ? 1 : 0;
  if (!out_stats && !out_smooth) return;

  // ---- median-of-3 smoothing with edge replication (Splitter:340-356); unchanged if m < 3 ----
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    double v = static_cast<double>(adj[a0 + i]);
    if (m >= 3) {
      const double l = static_cast<double>(adj[a0 + max(i - 1, 0)]);
      const double r = static_cast<double>(adj[a0 + min(i + 1, m - 1)]);
      v = fmax(fmin(l, v), fmin(fmax(l, v), r));  // median of three
    }
    base[i] = v;
    if (out_smooth) out_smooth[a0 + i] = static_cast<float>(v);
  }
  __syncthreads();
  if (!out_stats) return;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) keys[i] = (i < m) ? f64_to_ordered(base[i]) : ~0ull;
  __syncthreads();
  bitonic_sort_smem(keys, n2);
  if (threadIdx.x == 0) {
    const double med = np_median_sorted(keys, m);
    s_med = med;
    double* st = out_stats + static_cast<size_t>(doc) * 4;
    st[0] = med;
    st[2] = np_quantile_sorted(keys, m, 0.25);
    st[3] = np_quantile_sorted(keys, m, 0.75);
  }
  __syncthreads();
  const double med = s_med;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) keys[i] = (i < m) ? f64_to_ordered(fabs(base[i] - med)) : ~0ull;
  __syncthreads();
  bitonic_sort_smem(keys, n2);
  if (threadIdx.x == 0) out_stats[static_cast<size_t>(doc) * 4 + 1] = np_median_sorted(keys, m) + 1e-9;
}

// One CTA per document (a persistent, document-striding variant measured slower: long documents unbalance it).
__global__ void __launch_bounds__(128) segmented_percentile_kernel(float* __restrict__ adj, const int* __restrict__ offsets,
                                                                   int n_docs, double q, int n2_max,
                                                                   double* __restrict__ out_thr, unsigned char* __restrict__ out_flags,
                                                                   double* __restrict__ out_stats, float* __restrict__ out_smooth) {
  extern __shared__ __align__(16) unsigned char pct_smem[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(pct_smem);
  double* base = reinterpret_cast<double*>(keys + n2_max);
  __shared__ double s_med;
  if (blockIdx.x < n_docs) percentile_doc(adj, offsets, blockIdx.x, q, keys, base, &s_med, out_thr, out_flags, out_stats, out_smooth);
}

}  // namespace ss

using namespace ss;

extern "C" int ss_segmented_adjacent_cosine(const void* rows, int64_t n_rows, int dim, int dtype, float* out_adj, void* stream) {
  if (!rows || !out_adj) return fail(SS_ERR_INVALID_ARG, "ss_segmented_adjacent_cosine: null pointer");
  if (n_rows <= 0 || dim <= 0 || !dtype_ok(dtype)) return fail(SS_ERR_INVALID_ARG, "ss_segmented_adjacent_cosine: bad shape or dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t row_bytes = static_cast<size_t>(dim) * dtype_size(dtype);
  const size_t smem_cap = smem_optin();
  const bool bulk = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(rows) % 16 == 0) && row_bytes * (kAdjR + 1) <= 96 * 1024;
  if (bulk) {
    AdjParams p;
    p.rows = rows;
    p.n_rows = n_rows;
    p.dim = dim;
    p.tile_rows = static_cast<int>(std::max<size_t>(kAdjR, (48 * 1024) / row_bytes / kAdjR * kAdjR));
    p.tile_bytes = static_cast<uint32_t>(align_up((p.tile_rows + 1) * row_bytes, 128));
    const size_t fixed = 2 * kAdjMaxStages * 8 + 128;
    p.stages = static_cast<int>(std::min<size_t>(4, (smem_cap - fixed) / p.tile_bytes));
    if (p.stages >= 2) {
      const long long pairs = std::max<long long>(1, n_rows - 1);
      p.n_tiles = (pairs + p.tile_rows - 1) / p.tile_rows;
      p.out = out_adj;
      const size_t smem = fixed + static_cast<size_t>(p.stages) * p.tile_bytes;
      const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(sm_count(), p.n_tiles)));
      cudaError_t e = cudaSuccess;
      ProfileScope prof(st);
      switch (dtype) {
        case SS_F32:
          e = cudaFuncSetAttribute(adjacent_cosine_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
          if (e == cudaSuccess) adjacent_cosine_kernel<float><<<grid, kAdjThreads, smem, st>>>(p);
          break;
        case SS_BF16:
          e = cudaFuncSetAttribute(adjacent_cosine_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
          if (e == cudaSuccess) adjacent_cosine_kernel<__nv_bfloat16><<<grid, kAdjThreads, smem, st>>>(p);
          break;
        default:
          e = cudaFuncSetAttribute(adjacent_cosine_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
          if (e == cudaSuccess) adjacent_cosine_kernel<__half><<<grid, kAdjThreads, smem, st>>>(p);
          break;
      }
      if (e != cudaSuccess) return cuda_fail(e, "adjacent_cosine_kernel attribute");
      SS_CUDA_CHECK(cudaGetLastError());
      return SS_OK;
    }
  }
  const int blocks = static_cast<int>(std::min<long long>((n_rows + 7) / 8, static_cast<long long>(sm_count()) * 8));
  ProfileScope prof(st);
  switch (dtype) {
    case SS_F32: adjacent_cosine_generic_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(rows), n_rows, dim, out_adj); break;
    case SS_BF16: adjacent_cosine_generic_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(rows), n_rows, dim, out_adj); break;
    default: adjacent_cosine_generic_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(rows), n_rows, dim, out_adj); break;
  }
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}

extern "C" int ss_segmented_percentile(float* adj, const int32_t* offsets, int n_docs, int max_doc_rows, double pct,
                                       double* out_thr, uint8_t* out_flags, double* out_stats, float* out_smooth, void* stream) {
  if (!adj || !offsets || !out_thr) return fail(SS_ERR_INVALID_ARG, "ss_segmented_percentile: null pointer");
  if (n_docs <= 0 || max_doc_rows <= 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_percentile: sizes must be positive");
  if (!(pct >= 0.0 && pct <= 100.0)) return fail(SS_ERR_INVALID_ARG, "ss_segmented_percentile: pct must be in [0, 100]");
  int n2 = 1;
  while (n2 < std::max(1, max_doc_rows - 1)) n2 <<= 1;
  const size_t smem = static_cast<size_t>(n2) * 16;
  if (smem + 1024 > smem_optin())
    return fail(SS_ERR_UNSUPPORTED, "ss_segmented_percentile: documents longer than 8193 sentences are not supported");
  if (smem > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(segmented_percentile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  segmented_percentile_small_kernel<<<(n_docs + kPctSmallWarps - 1) / kPctSmallWarps, kPctSmallWarps * 32, 0,
                                      static_cast<cudaStream_t>(stream)>>>(adj, offsets, n_docs, pct / 100.0, out_thr, out_flags, out_stats,
                                                                           out_smooth);
  segmented_percentile_kernel<<<n_docs, 128, smem, static_cast<cudaStream_t>(stream)>>>(adj, offsets, n_docs, pct / 100.0, n2, out_thr,
                                                                                       out_flags, out_stats, out_smooth);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
