// K4 — semantic-grouping "threshold pass" over a ragged batch of similarity matrices.
//
// One CTA per document performs, on device, the data-parallel part of semantic_grouping_main
// (Method/Semantic_Grouping_Optimized.py):
//   1. mu / sigma over all n*n entries of S                                      (:102-103)
//   2. sim_sharp = 1/(1+exp(-((S-mu)/sigma)/tau)) in fp32, diagonal := 0         (:104-111)
//      centrality = rowsum / max(n-1, 1)                                         (:115)
//   3. exact order statistics of the positive entries -> np.quantile(vals, q) for q = 0.80,
//      0.65, 0.60 (linear interpolation, fp64) and 0.1 * std(vals)               (:351-355,458-462,534-537,559-561)
//      via a 3-pass (11/11/10-bit) radix select on the fp32 bit patterns
//   4. per-row top-(k_eff+1) neighbours, value-descending / index-ascending       (:270-283, k_eff :347)
// Documents of up to 512 sentences take a register-resident row pass: one warp loads a row of S once,
// sharpens it, writes sim_sharp, feeds the first radix histogram (warp-aggregated increments) and
// selects the row's neighbours by bisection on the ordered bits + ballot compaction — steps 2, the
// first pass of 3 and 4 cost ONE read of S instead of ~35 passes over sim_sharp.
// The floor filter + symmetrisation of the kNN graph and the sequential clustering that
// follows stay on the host, which receives sim_sharp, centrality, thresholds and neighbour lists.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

constexpr int kGrpThreads = 512;
constexpr int kGrpWarps = kGrpThreads / 32;
constexpr int kRadixBins = 2048;
constexpr int kNumQ = 3;
constexpr int kKnnWidth = 33;  // k_eff <= 32 -> at most 33 neighbours per row

struct GroupParams {
  const float* S;              // packed per-document matrices
  const int* offsets;          // [n_docs + 1] row offsets
  const long long* s_offsets;  // [n_docs + 1] element offsets
  int n_docs;
  float tau;
  int knn_mode;  // 0: auto k = clamp(round(0.06 n), 5, 32); >0: explicit k; -1: max(5, min(20, n-1))
  float* sharp;          // packed like S
  double* centrality;    // [total_rows]
  double* doc_stats;     // [n_docs][8]: mu, sigma, q80, q65, q60, 0.1*std(pos), count(pos), k_eff
  int* knn_idx;          // [total_rows][33], -1 padded
  float* knn_val;        // [total_rows][33]
};

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kGrpWarps; ++w) t += red[w];
  return t;
}

// numpy `_lerp` (numpy/lib/_function_base_impl.py) for the "linear" quantile method.
__device__ __forceinline__ double np_lerp(double a, double b, double g) {
  // numpy rounds the product and the sum separately: keep the compiler from fusing them into an FMA
  const double diff = __dsub_rn(b, a);
  double r = __dadd_rn(a, __dmul_rn(diff, g));
  if (g >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, g)));
  return r;
}

constexpr int kRowRegs = 16;               // fast path: a whole row (n <= 512) lives in 16 registers per lane
constexpr int kFastMaxN = 32 * kRowRegs;

// One warp sorts a[0..64) (shared memory) in descending order.
__device__ __forceinline__ void warp_bitonic64_desc(uint64_t* a, int lane) {
#pragma unroll 1
  for (int k2 = 2; k2 <= 64; k2 <<= 1) {
#pragma unroll 1
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const int i = ((lane & ~(j - 1)) << 1) | (lane & (j - 1));
      const int q = i | j;
      const uint64_t x = a[i], y = a[q];
      const bool desc = (i & k2) == 0;
      if (desc ? (x < y) : (x > y)) {
        a[i] = y;
        a[q] = x;
      }
      __syncwarp();
    }
  }
}

// Warp-aggregated histogram increment: lanes that hit the same bin elect one lane to add their count,
// so the saturated ends of the sigmoid (thousands of values in one bin) do not serialise on one address.
__device__ __forceinline__ void hist_add_aggregated(unsigned int* hist, unsigned int bin, bool valid, int lane) {
  const unsigned int key = valid ? bin : 0xFFFFFFFFu;
  const unsigned int peers = __match_any_sync(0xffffffffu, key);
  if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], static_cast<unsigned int>(__popc(peers)));
}

__global__ void __launch_bounds__(kGrpThreads) group_threshold_kernel(const GroupParams p) {
  __shared__ double red[kGrpWarps];
  __shared__ unsigned int hist[kNumQ][kRadixBins];
  __shared__ uint64_t knn_scr[kGrpWarps][64];
  __shared__ unsigned int t_prefix[kNumQ], t_rank[kNumQ];  // per target: determined high bits, remaining rank
  __shared__ unsigned int t_cnt_le[kNumQ], t_min_above[kNumQ];
  __shared__ float s_mu, s_sigma;

  const int doc = blockIdx.x;
  const int row_base = p.offsets[doc];
  const int n = p.offsets[doc + 1] - row_base;
  double* st = p.doc_stats + static_cast<size_t>(doc) * 8;
  if (n < 1) {
    if (threadIdx.x < 8) st[threadIdx.x] = 0.0;
    return;
  }
  const float* S = p.S + p.s_offsets[doc];
  float* sharp = p.sharp + p.s_offsets[doc];
  const long long nn = static_cast<long long>(n) * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- k for the kNN graph (Grouping:347,349 and :273) -----------------------------------------
  int kk;
  if (p.knn_mode == 0) kk = max(5, min(32, static_cast<int>(rint(static_cast<double>(n) * 0.06))));
  else if (p.knn_mode > 0) kk = p.knn_mode;
  else kk = max(5, min(20, n - 1));
  const int k_eff = max(1, min(kk, n - 1));
  const int width = min(min(k_eff + 1, n), kKnnWidth);

  // ---- 1. mean / std over all n^2 entries (fp64 accumulation, rounded to fp32 like numpy's result) ----
  double s1 = 0.0, s2 = 0.0;
  for (long long i = tid; i < nn; i += kGrpThreads) {
    const double v = static_cast<double>(S[i]);
    s1 += v;
    s2 += v * v;
  }
  for (int i = tid; i < kRadixBins; i += kGrpThreads) hist[0][i] = 0u;  // pass-0 histogram (shared by the three targets)
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (tid == 0) {
    const double mean = s1 / static_cast<double>(nn);
    const double var = fmax(s2 / static_cast<double>(nn) - mean * mean, 0.0);
    s_mu = static_cast<float>(mean);
    s_sigma = static_cast<float>(sqrt(var)) + 1e-9f;  // np.float32(std) + 1e-9 stays float32
  }
  __syncthreads();
  const float mu = s_mu, sigma = s_sigma, tau = p.tau;

  // ---- 2. sharpen, zero the diagonal, row sums, first radix histogram, kNN lists -----------------
  double pos_s1 = 0.0, pos_s2 = 0.0;
  unsigned int pos_cnt = 0;
  if (n <= kFastMaxN) {
    // Fast path: one warp per row, the row stays in registers for all four products of the pass.
    uint64_t* scr = knn_scr[warp];
    for (int r = warp; r < n; r += kGrpWarps) {
      const float* srow = S + static_cast<size_t>(r) * n;
      float* orow = sharp + static_cast<size_t>(r) * n;
      uint32_t o[kRowRegs];  // order-preserving bits of the sharpened value; 0 = column past the end
      double rs = 0.0;
#pragma unroll
      for (int j = 0; j < kRowRegs; ++j) {
        o[j] = 0u;
        if (32 * j < n) {
          const int c = lane + 32 * j;
          bool pos = false;
          unsigned int bits = 0u;
          if (c < n) {
            const float z = (srow[c] - mu) / sigma;
            float v = 1.0f / (1.0f + expf(-(z / tau)));
            if (c == r) v = 0.f;
            orow[c] = v;
            rs += static_cast<double>(v);
            o[j] = float_to_ordered(v);
            if (v > 0.f) {
              pos = true;
              bits = __float_as_uint(v);
              pos_s1 += static_cast<double>(v);
              pos_s2 += static_cast<double>(v) * static_cast<double>(v);
              ++pos_cnt;
            }
          }
          hist_add_aggregated(hist[0], bits >> 21, pos, lane);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      if (lane == 0) {
        const float rsum = static_cast<float>(rs);
        p.centrality[row_base + r] = static_cast<double>(rsum / static_cast<float>(max(n - 1, 1)));
      }
      // top-`width` of the row (value desc, index asc): the width-th largest value T by bisection on the
      // ordered bits (every real column has bit 31 set), then everything above T plus the lowest-index ties
      uint32_t T = 0x80000000u;
#pragma unroll 1
      for (int bit = 30; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int j = 0; j < kRowRegs; ++j) c += (o[j] >= cand) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= width) T = cand;
      }
      int c_gt = 0;
#pragma unroll
      for (int j = 0; j < kRowRegs; ++j) c_gt += (o[j] > T) ? 1 : 0;
      c_gt = __reduce_add_sync(0xffffffffu, c_gt);
      scr[lane] = 0ull;
      scr[lane + 32] = 0ull;
      __syncwarp();
      int base_gt = 0, taken_eq = 0;
      const unsigned int lt_mask = (1u << lane) - 1u;
#pragma unroll
      for (int j = 0; j < kRowRegs; ++j) {
        if (32 * j < n) {
          const uint64_t key = (static_cast<uint64_t>(o[j]) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(lane + 32 * j));
          const unsigned int m_gt = __ballot_sync(0xffffffffu, o[j] > T);
          const unsigned int m_eq = __ballot_sync(0xffffffffu, o[j] == T);
          if (o[j] > T) scr[base_gt + __popc(m_gt & lt_mask)] = key;
          base_gt += __popc(m_gt);
          const int room = width - c_gt - taken_eq;  // ties are admitted in ascending column order
          const int my_rank = __popc(m_eq & lt_mask);
          if (o[j] == T && my_rank < room) scr[c_gt + taken_eq + my_rank] = key;
          taken_eq += min(__popc(m_eq), max(room, 0));
        }
      }
      __syncwarp();
      warp_bitonic64_desc(scr, lane);
      int* oi = p.knn_idx + static_cast<size_t>(row_base + r) * kKnnWidth;
      float* ov = p.knn_val + static_cast<size_t>(row_base + r) * kKnnWidth;
      for (int s = lane; s < kKnnWidth; s += 32) {
        const uint64_t key = scr[s];
        const bool ok = s < width;
        oi[s] = ok ? static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull)) : -1;
        ov[s] = ok ? ordered_to_float(static_cast<uint32_t>(key >> 32)) : 0.f;
      }
      __syncwarp();
    }
  } else {
    // Generic path for very long documents: rows are re-read from memory.
    for (int r = warp; r < n; r += kGrpWarps) {
      const float* srow = S + static_cast<size_t>(r) * n;
      float* orow = sharp + static_cast<size_t>(r) * n;
      double rs = 0.0;
      for (int c0 = 0; c0 < n; c0 += 32) {
        const int c = c0 + lane;
        bool pos = false;
        unsigned int bits = 0u;
        if (c < n) {
          const float z = (srow[c] - mu) / sigma;
          float v = 1.0f / (1.0f + expf(-(z / tau)));
          if (c == r) v = 0.f;
          orow[c] = v;
          rs += static_cast<double>(v);
          if (v > 0.f) {
            pos = true;
            bits = __float_as_uint(v);
            pos_s1 += static_cast<double>(v);
            pos_s2 += static_cast<double>(v) * static_cast<double>(v);
            ++pos_cnt;
          }
        }
        hist_add_aggregated(hist[0], bits >> 21, pos, lane);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      if (lane == 0) {
        const float rsum = static_cast<float>(rs);
        p.centrality[row_base + r] = static_cast<double>(rsum / static_cast<float>(max(n - 1, 1)));
      }
      __syncwarp();  // this warp's writes to orow are visible to its own re-reads below
      int* oi = p.knn_idx + static_cast<size_t>(row_base + r) * kKnnWidth;
      float* ov = p.knn_val + static_cast<size_t>(row_base + r) * kKnnWidth;
      uint64_t prev = ~0ull;
      for (int s = 0; s < width; ++s) {  // successive maxima
        uint64_t best = 0ull;
        for (int c = lane; c < n; c += 32) {
          const uint64_t key = (static_cast<uint64_t>(float_to_ordered(orow[c])) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(c));
          if (key < prev && key > best) best = key;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const uint64_t other = __shfl_xor_sync(0xffffffffu, best, off);
          best = other > best ? other : best;
        }
        if (lane == 0) {
          oi[s] = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(best & 0xFFFFFFFFull));
          ov[s] = ordered_to_float(static_cast<uint32_t>(best >> 32));
        }
        prev = best;
      }
      for (int s = width + lane; s < kKnnWidth; s += 32) {
        oi[s] = -1;
        ov[s] = 0.f;
      }
    }
  }
  pos_s1 = block_sum(pos_s1, red);
  pos_s2 = block_sum(pos_s2, red);
  const double m_d = block_sum(static_cast<double>(pos_cnt), red);
  const unsigned int m = static_cast<unsigned int>(m_d + 0.5);
  __syncthreads();  // sharp[] and the pass-0 histogram written by this CTA are visible to the whole CTA

  // ---- 3. radix select of the lower order statistic of each quantile ---------------------------
  const double qs[kNumQ] = {0.80, 0.65, 0.60};
  unsigned int lo_rank[kNumQ];
  double gamma[kNumQ];
#pragma unroll
  for (int t = 0; t < kNumQ; ++t) {
    const double vi = __dmul_rn(m > 0 ? static_cast<double>(m - 1) : 0.0, qs[t]);
    const double fl = floor(vi);
    lo_rank[t] = static_cast<unsigned int>(fl);
    gamma[t] = __dsub_rn(vi, fl);
  }
  double q_out[kNumQ] = {0.0, 0.0, 0.0};
  if (m > 0) {
    if (tid < kNumQ) {
      t_prefix[tid] = 0u;
      t_rank[tid] = lo_rank[tid];
    }
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    unsigned int known_mask = 0u;
    for (int pass = 0; pass < 3; ++pass) {
      const int sh = shifts[pass];
      const unsigned int dmask = (1u << widths[pass]) - 1u;
      if (pass > 0) {
        for (int i = tid; i < kNumQ * kRadixBins; i += kGrpThreads) (&hist[0][0])[i] = 0u;
        __syncthreads();
        const unsigned int pf0 = t_prefix[0], pf1 = t_prefix[1], pf2 = t_prefix[2];
        for (long long i = tid; i < nn; i += kGrpThreads) {
          const unsigned int b = __float_as_uint(sharp[i]);
          if (b == 0u) continue;  // diagonal / underflowed zeros are not "positive values"
          const unsigned int hi = b & known_mask, dg = (b >> sh) & dmask;
          if (hi == pf0) atomicAdd(&hist[0][dg], 1u);
          if (hi == pf1) atomicAdd(&hist[1][dg], 1u);
          if (hi == pf2) atomicAdd(&hist[2][dg], 1u);
        }
      }
      __syncthreads();
      // one warp per target walks its histogram (pass 0: the shared one) to the bin that holds the wanted rank
      if (warp < kNumQ) {
        const unsigned int* h = pass == 0 ? hist[0] : hist[warp];
        const unsigned int want = t_rank[warp];
        unsigned int run = 0u;
        const int nb = 1 << widths[pass];
        for (int b0 = 0; b0 < nb; b0 += 32) {
          const unsigned int c = h[b0 + lane];
          unsigned int incl = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
          }
          const unsigned int excl = run + incl - c;
          const bool here = (want >= excl) && (want < excl + c);
          const unsigned int bal = __ballot_sync(0xffffffffu, here);
          if (bal) {
            const int src = __ffs(bal) - 1;
            const unsigned int e = __shfl_sync(0xffffffffu, excl, src);
            if (lane == 0) {
              t_prefix[warp] |= static_cast<unsigned int>(b0 + src) << sh;
              t_rank[warp] = want - e;
            }
            break;
          }
          run += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      known_mask |= dmask << sh;
      __syncthreads();
    }
    // ---- upper order statistic: next distinct value unless duplicates already cover rank lo+1 ----
    if (tid < kNumQ) {
      t_cnt_le[tid] = 0u;
      t_min_above[tid] = 0xFFFFFFFFu;
    }
    __syncthreads();
    const unsigned int v0 = t_prefix[0], v1 = t_prefix[1], v2 = t_prefix[2];
    unsigned int c0 = 0, c1 = 0, c2 = 0, a0 = 0xFFFFFFFFu, a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
    for (long long i = tid; i < nn; i += kGrpThreads) {
      const unsigned int b = __float_as_uint(sharp[i]);
      if (b == 0u) continue;
      if (b <= v0) ++c0; else a0 = min(a0, b);
      if (b <= v1) ++c1; else a1 = min(a1, b);
      if (b <= v2) ++c2; else a2 = min(a2, b);
    }
    c0 = __reduce_add_sync(0xffffffffu, c0);
    c1 = __reduce_add_sync(0xffffffffu, c1);
    c2 = __reduce_add_sync(0xffffffffu, c2);
    a0 = __reduce_min_sync(0xffffffffu, a0);
    a1 = __reduce_min_sync(0xffffffffu, a1);
    a2 = __reduce_min_sync(0xffffffffu, a2);
    if (lane == 0) {
      atomicAdd(&t_cnt_le[0], c0);
      atomicAdd(&t_cnt_le[1], c1);
      atomicAdd(&t_cnt_le[2], c2);
      atomicMin(&t_min_above[0], a0);
      atomicMin(&t_min_above[1], a1);
      atomicMin(&t_min_above[2], a2);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kNumQ; ++t) {
      const unsigned int vlo = t_prefix[t];
      unsigned int vhi = vlo;
      if (lo_rank[t] + 1 < m && t_cnt_le[t] <= lo_rank[t] + 1) vhi = t_min_above[t];
      q_out[t] = np_lerp(static_cast<double>(__uint_as_float(vlo)), static_cast<double>(__uint_as_float(vhi)), gamma[t]);
    }
  }

  if (tid == 0) {
    st[0] = static_cast<double>(mu);
    st[1] = static_cast<double>(sigma);
    st[2] = q_out[0];
    st[3] = q_out[1];
    st[4] = q_out[2];
    double sd = 0.0;
    if (m > 0) {
      const double mean = pos_s1 / m_d;
      sd = sqrt(fmax(pos_s2 / m_d - mean * mean, 0.0));
    }
    st[5] = 0.1 * sd;
    st[6] = m_d;
    st[7] = static_cast<double>(kk);
  }
}

}  // namespace ss

using namespace ss;

extern "C" int ss_group_threshold_pass(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, float tau,
                                       int knn_mode, float* out_sharp, double* out_centrality, double* out_doc_stats,
                                       int32_t* out_knn_idx, float* out_knn_val, void* stream) {
  if (!S || !offsets || !s_offsets || !out_sharp || !out_centrality || !out_doc_stats || !out_knn_idx || !out_knn_val)
    return fail(SS_ERR_INVALID_ARG, "ss_group_threshold_pass: null pointer");
  if (n_docs <= 0 || !(tau > 0.f)) return fail(SS_ERR_INVALID_ARG, "ss_group_threshold_pass: n_docs and tau must be positive");
  if (knn_mode > 32) return fail(SS_ERR_UNSUPPORTED, "ss_group_threshold_pass: knn_k > 32 is not supported");
  GroupParams p;
  p.S = S;
  p.offsets = offsets;
  p.s_offsets = reinterpret_cast<const long long*>(s_offsets);
  p.n_docs = n_docs;
  p.tau = tau;
  p.knn_mode = knn_mode;
  p.sharp = out_sharp;
  p.centrality = out_centrality;
  p.doc_stats = out_doc_stats;
  p.knn_idx = out_knn_idx;
  p.knn_val = out_knn_val;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfileScope prof(st);
    group_threshold_kernel<<<n_docs, kGrpThreads, 0, st>>>(p);
  }
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
