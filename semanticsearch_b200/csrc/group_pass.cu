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
  int symmetric;  // caller promises S == S^T bit for bit (K3's output): order statistics and moments of the positive values
                  // are taken over the strict upper triangle only (the full multiset is that one, twice)
  float* sharp;          // packed like S
  double* centrality;    // [total_rows]
  double* doc_stats;     // [n_docs][8]: mu, sigma, q80, q65, q60, 0.1*std(pos), count(pos), k_eff
  int* knn_idx;          // [total_rows][33], -1 padded
  float* knn_val;        // [total_rows][33]
};

template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
  constexpr int kGrpWarps = NT / 32;
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

constexpr int kFastMaxN = 512;  // fast path: a whole row lives in 4..16 registers per lane
constexpr int kListCap = 2048;  // candidates per quantile the shared-memory selection can hold

// Descending bitonic sort of 64 keys held two per lane (sort index g = 2 * lane + r): the distance-1 stages are a
// register swap, the others one shuffle per key.  After the call key g of the sorted order is k[g & 1] of lane g >> 1.
template <typename K>
__device__ __forceinline__ void warp_sort64_desc(K& k0, K& k1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
    const bool desc_blk = ((lane * 2) & k) == 0;  // k = 64: always (the whole sequence ends descending)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 2) {
        const int lane_mask = j >> 1;
        const bool keep_max = desc_blk == ((lane & lane_mask) == 0);
        const K o0 = __shfl_xor_sync(0xffffffffu, k0, lane_mask), o1 = __shfl_xor_sync(0xffffffffu, k1, lane_mask);
        k0 = keep_max ? max(k0, o0) : min(k0, o0);
        k1 = keep_max ? max(k1, o1) : min(k1, o1);
      } else {
        const K hi = max(k0, k1), lo = min(k0, k1);
        k0 = desc_blk ? hi : lo;
        k1 = desc_blk ? lo : hi;
      }
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

// One warp walks a histogram to the bin that holds 0-based rank `want`: returns the bin, the rank
// inside it and its population (all lanes get the result).
__device__ __forceinline__ void walk_hist(const unsigned int* h, int nb, unsigned int want, int lane, unsigned int& bin,
                                          unsigned int& rank_in, unsigned int& count) {
  unsigned int run = 0u;
  bin = 0u;
  rank_in = 0u;
  count = 0u;
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
      bin = static_cast<unsigned int>(b0 + src);
      rank_in = want - __shfl_sync(0xffffffffu, excl, src);
      count = __shfl_sync(0xffffffffu, c, src);
      return;
    }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// All threads of the block sort a[0..n) (n a power of two, shared memory) in ascending order.
__device__ __forceinline__ void block_bitonic_asc_u32(unsigned int* a, int n) {
  for (int k2 = 2; k2 <= n; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int q = i | j;
        const unsigned int x = a[i], y = a[q];
        const bool asc = (i & k2) == 0;
        if (asc ? (x > y) : (x < y)) {
          a[i] = y;
          a[q] = x;
        }
      }
      __syncthreads();
    }
  }
}

struct RowCtx {
  const float* S;
  float* sharp;
  int n, row_base, width;
  float mu, sigma, tau;
  unsigned int* hist;  // linear-bin histogram of the positive sharpened values
  uint64_t* scr;       // this warp's 64-key scratch
  double* centrality;
  int* knn_idx;
  float* knn_val;
};

// Monotone 2048-way binning of a sharpened value in [0, 1] (a multiply and a truncating conversion: 1.0f lands in bin 2047;
// the selection only needs the map to be non-decreasing and identical in the histogram and the gather pass).
__device__ __forceinline__ unsigned int lin_bin(float v) { return __float2uint_rz(v * 2047.9998f); }

// One warp per row with the row held in NREG registers per lane (n <= 32 * NREG).
template <int NREG, int kGrpWarps, bool SYM>
__device__ __forceinline__ void row_pass_regs(const RowCtx& cx, int warp, int lane, double& pos_s1, double& pos_s2,
                                              unsigned int& pos_cnt) {
  const float* S = cx.S;
  float* sharp = cx.sharp;
  const int n = cx.n, row_base = cx.row_base, width = cx.width;
  const float mu = cx.mu;
  const float zscale = 1.4426950408889634f / (cx.sigma * cx.tau);  // log2(e) / (sigma * tau)
  uint64_t* scr = cx.scr;
  unsigned int c_lo = 0u, c_hi = 0u;  // this LANE's counts of bins 0 and 2047 (the saturated ends of the sigmoid), flushed once
    for (int r = warp; r < n; r += kGrpWarps) {
      const float* srow = S + static_cast<size_t>(r) * n;
      float* orow = sharp + static_cast<size_t>(r) * n;
      uint32_t o[NREG];  // order-preserving bits of the sharpened value; 0 = column past the end
      double rs = 0.0, ps1 = 0.0;
#pragma unroll
      for (int j = 0; j < NREG; ++j) {  // the whole row in flight before any of it is consumed
        const int c = lane + 32 * j;
        o[j] = (32 * j < n && c < n) ? __float_as_uint(srow[c]) : 0u;
      }
#pragma unroll
      for (int j = 0; j < NREG; ++j) {
        const float s_raw = __uint_as_float(o[j]);
        o[j] = 0u;
        if (32 * j < n) {
          const int c = lane + 32 * j;
          float v = 0.f;
          if (c < n) {
            // sigmoid(((S - mu) / sigma) / tau) with one FMA, ex2.approx and rcp.approx: a few ulp from
            // the reference's fp32 expression (Grouping:105-106), far inside the 1e-5 bound; saturates
            // to exactly 0 / 1 like numpy's exp overflow does
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-((s_raw - mu) * zscale)));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(1.0f + e));
            if (c == r) v = 0.f;
            orow[c] = v;
            rs += static_cast<double>(v);
            o[j] = __float_as_uint(v) | 0x80000000u;    // float_to_ordered of a non-negative value
          }
          // moments and histogram of the "positive values": with a symmetric S only the strict upper triangle is
          // counted (register blocks left of the diagonal skip the work altogether, a warp-uniform test)
          if (!SYM || 32 * j + 31 > r) {
            const bool take = v > 0.f && (!SYM || c > r);   // v == 0 for columns past the end
            if (take) {
              const double vd = static_cast<double>(v);  // float64 sums: 0.1 * std(vals) stays within 1e-9 of numpy's
              ps1 += vd;
              pos_s2 = fma(vd, vd, pos_s2);
              ++pos_cnt;
            }
            const unsigned int bin = lin_bin(v);
            c_lo += (take && bin == 0u) ? 1u : 0u;
            c_hi += (take && bin == 2047u) ? 1u : 0u;
            if (take && bin - 1u < 2046u) atomicAdd(&cx.hist[bin], 1u);
          }
        }
      }
      pos_s1 += ps1;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      if (lane == 0) {
        const float rsum = static_cast<float>(rs);
        cx.centrality[row_base + r] = static_cast<double>(rsum / static_cast<float>(max(n - 1, 1)));
      }
      // top-`width` of the row (value desc, index asc).  Cheap bound first: each lane's two largest values
      // form a 64-value sample; its width-th largest L is a lower bound of the row's width-th largest, and
      // usually only a few more than `width` columns reach L — gather those and sort them.
      const unsigned int lt_mask = (1u << lane) - 1u;
      uint32_t m1 = 0u, m2 = 0u;
#pragma unroll
      for (int j = 0; j < NREG; ++j) {
        m2 = max(m2, min(m1, o[j]));
        m1 = max(m1, o[j]);
      }
      warp_sort64_desc<uint32_t>(m1, m2, lane);  // width <= n real columns are in the sample, so L is a real value
      const uint32_t L = __shfl_sync(0xffffffffu, ((width - 1) & 1) ? m2 : m1, (width - 1) >> 1);
      const uint32_t vmax = __shfl_sync(0xffffffffu, m1, 0);  // the row's largest value (sorted slot 0)
      int c_ge = 0;
#pragma unroll
      for (int j = 0; j < NREG; ++j) c_ge += (o[j] >= L) ? 1 : 0;
      c_ge = __reduce_add_sync(0xffffffffu, c_ge);
      int* oi = cx.knn_idx + static_cast<size_t>(row_base + r) * kKnnWidth;
      float* ov = cx.knn_val + static_cast<size_t>(row_base + r) * kKnnWidth;
      if (c_ge <= 64 && vmax - L < (1u << 26) - 1u) {
        // Usual case: at most 64 candidates whose ordered bits span less than 2^26 above L.  The compaction below lists them in
        // ascending column order, so (value - L) << 6 | (63 - position) is a 32-bit key with the order "value descending,
        // column ascending": the final sort runs on 32-bit keys (half the shuffles and compares of the 64-bit network).
        uint32_t* keys32 = reinterpret_cast<uint32_t*>(scr);
        uint16_t* cols16 = reinterpret_cast<uint16_t*>(keys32 + 64);
        scr[lane] = 0ull;  // keys32[2 * lane], keys32[2 * lane + 1]: 0 = empty slot, sorts last
        __syncwarp();
        int base = 0;
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
          if (32 * j < n) {
            const unsigned int m_ge = __ballot_sync(0xffffffffu, o[j] >= L);
            if (o[j] >= L) {
              const int pos = base + __popc(m_ge & lt_mask);
              keys32[pos] = (((o[j] - L) << 6) | static_cast<uint32_t>(63 - pos)) + 1u;
              cols16[pos] = static_cast<uint16_t>(lane + 32 * j);
            }
            base += __popc(m_ge);
          }
        }
        __syncwarp();
        uint32_t k0 = keys32[2 * lane], k1 = keys32[2 * lane + 1];
        warp_sort64_desc<uint32_t>(k0, k1, lane);
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // lane holds sorted slots 2 * lane and 2 * lane + 1
          const int sl = 2 * lane + h;
          const uint32_t key = (h ? k1 : k0) - 1u;
          if (sl < kKnnWidth) {
            const bool ok = sl < width;  // c_ge >= width: the first `width` sorted slots are real candidates
            oi[sl] = ok ? static_cast<int>(cols16[63 - (key & 63u)]) : -1;
            ov[sl] = ok ? ordered_to_float(L + (key >> 6)) : 0.f;
          }
        }
        __syncwarp();
        continue;
      }
      scr[lane] = 0ull;
      scr[lane + 32] = 0ull;
      __syncwarp();
      if (c_ge <= 64) {
        int base = 0;
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
          if (32 * j < n) {
            const unsigned int m_ge = __ballot_sync(0xffffffffu, o[j] >= L);
            if (o[j] >= L)
              scr[base + __popc(m_ge & lt_mask)] = (static_cast<uint64_t>(o[j]) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(lane + 32 * j));
            base += __popc(m_ge);
          }
        }
      } else {
        // crowded (many ties at the bound): the width-th largest value T by bisection over the whole row,
        // then everything above T plus the lowest-index ties
        uint32_t T = 0x80000000u;
#pragma unroll 1
        for (int bit = 30; bit >= 0; --bit) {
          const uint32_t cand = T | (1u << bit);
          int c = 0;
#pragma unroll
          for (int j = 0; j < NREG; ++j) c += (o[j] >= cand) ? 1 : 0;
          c = __reduce_add_sync(0xffffffffu, c);
          if (c >= width) T = cand;
        }
        int c_gt = 0;
#pragma unroll
        for (int j = 0; j < NREG; ++j) c_gt += (o[j] > T) ? 1 : 0;
        c_gt = __reduce_add_sync(0xffffffffu, c_gt);
        int base_gt = 0, taken_eq = 0;
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
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
      }
      __syncwarp();
      unsigned long long k0 = scr[2 * lane], k1 = scr[2 * lane + 1];
      warp_sort64_desc<unsigned long long>(k0, k1, lane);
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // lane holds sorted slots 2 * lane and 2 * lane + 1
        const int sl = 2 * lane + h;
        const unsigned long long key = h ? k1 : k0;
        if (sl < kKnnWidth) {
          const bool ok = sl < width;
          oi[sl] = ok ? static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull)) : -1;
          ov[sl] = ok ? ordered_to_float(static_cast<uint32_t>(key >> 32)) : 0.f;
        }
      }
      __syncwarp();
    }
  c_lo = __reduce_add_sync(0xffffffffu, c_lo);
  c_hi = __reduce_add_sync(0xffffffffu, c_hi);
  if (lane == 0) {
    if (c_lo) atomicAdd(&cx.hist[0], c_lo);
    if (c_hi) atomicAdd(&cx.hist[2047], c_hi);
  }
}


// ------------------------------------------------------------------------------------------------
// Tiny documents (n <= 32 sentences — three quarters of the reference corpus, median 10): one WARP per
// document, lane r owns row r of S in registers.  Same arithmetic and outputs as the CTA kernel below,
// without its fixed costs (2048-bin histograms, block barriers, shared-memory sorts).
// ------------------------------------------------------------------------------------------------
constexpr int kSmallMaxN = 32;
constexpr int kSmallWarps = 8;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// value (fp32 bit pattern of a positive float) with at least `want` positive entries >= it, i.e. the want-th largest
template <int NMAX>
__device__ __forceinline__ uint32_t small_select_desc(const uint32_t (&o)[NMAX], int n, int want) {
  uint32_t t = 0u;
#pragma unroll 1
  for (int bit = 29; bit >= 0; --bit) {  // sharpened values are <= 1.0f = 0x3F800000: bits 31 and 30 are never set
    const uint32_t cand = t | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) c += (j < n && o[j] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= want) t = cand;
  }
  return t;
}

template <int NMAX>
__device__ __forceinline__ void group_small_doc(const GroupParams& p, int doc, int row_base, int n, int lane) {
  const float* S = p.S + p.s_offsets[doc];
  float* sharp = p.sharp + p.s_offsets[doc];
  const bool live = lane < n;
  const double nn = static_cast<double>(n) * n;

  int kk;
  if (p.knn_mode == 0) kk = max(5, min(32, static_cast<int>(rint(static_cast<double>(n) * 0.06))));
  else if (p.knn_mode > 0) kk = p.knn_mode;
  else kk = max(5, min(20, n - 1));
  const int k_eff = max(1, min(kk, n - 1));
  const int width = min(min(k_eff + 1, n), kKnnWidth);

  float x[NMAX];
  double s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    x[j] = (live && j < n) ? S[lane * n + j] : 0.f;
    const double d = static_cast<double>(x[j]);
    s1 += d;
    s2 += d * d;
  }
  s1 = warp_sum_f64(s1);
  s2 = warp_sum_f64(s2);
  const double mean = s1 / nn;
  const float mu = static_cast<float>(mean);
  const float sigma = static_cast<float>(sqrt(fmax(s2 / nn - mean * mean, 0.0))) + 1e-9f;
  const float zscale = 1.4426950408889634f / (sigma * p.tau);

  uint32_t o[NMAX];  // bit pattern of the positive sharpened values, 0 = not a "positive value"
  double rs = 0.0, pos_s1 = 0.0, pos_s2 = 0.0;
  int pos_cnt = 0;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    float v = 0.f;
    if (live && j < n) {
      v = __frcp_rn(1.0f + exp2f(-((x[j] - mu) * zscale)));  // same expression as the CTA kernel's row pass
      if (j == lane) v = 0.f;
      sharp[lane * n + j] = v;
      rs += static_cast<double>(v);
    }
    x[j] = v;
    o[j] = v > 0.f ? __float_as_uint(v) : 0u;
    if (v > 0.f) {
      pos_s1 += static_cast<double>(v);
      pos_s2 += static_cast<double>(v) * static_cast<double>(v);
      ++pos_cnt;
    }
  }
  if (live) p.centrality[row_base + lane] = static_cast<double>(static_cast<float>(rs) / static_cast<float>(max(n - 1, 1)));
  pos_s1 = warp_sum_f64(pos_s1);
  pos_s2 = warp_sum_f64(pos_s2);
  const int m = __reduce_add_sync(0xffffffffu, pos_cnt);
  const double m_d = static_cast<double>(m);

  // quantiles of the positive values: numpy's linear interpolation between order statistics lo and lo + 1
  const double qs[kNumQ] = {0.80, 0.65, 0.60};
  double q_out[kNumQ] = {0.0, 0.0, 0.0};
  if (m > 0) {
#pragma unroll 1
    for (int t = 0; t < kNumQ; ++t) {
      const double vi = __dmul_rn(static_cast<double>(m - 1), qs[t]);
      const double fl = floor(vi);
      const int lo = static_cast<int>(fl);
      const double gamma = __dsub_rn(vi, fl);
      const uint32_t vlo = small_select_desc<NMAX>(o, n, m - lo);      // ascending rank lo == (m - lo)-th largest
      uint32_t vhi = vlo;
      if (lo + 1 < m) {  // order statistic lo + 1: vlo again if its duplicates reach that rank, else the next larger value
        int c_gt = 0;
        uint32_t above = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
          if (j < n && o[j] > vlo) {
            ++c_gt;
            above = min(above, o[j]);
          }
        }
        c_gt = __reduce_add_sync(0xffffffffu, c_gt);
        above = __reduce_min_sync(0xffffffffu, above);
        if (c_gt > m - lo - 2) vhi = above;
      }
      q_out[t] = np_lerp(static_cast<double>(__uint_as_float(vlo)), static_cast<double>(__uint_as_float(vhi)), gamma);
    }
  }

  // row `lane`: its `width` best neighbours, value descending / index ascending, by successive maxima on registers
  if (live) {
    int* oi = p.knn_idx + static_cast<size_t>(row_base + lane) * kKnnWidth;
    float* ov = p.knn_val + static_cast<size_t>(row_base + lane) * kKnnWidth;
    uint64_t prev = ~0ull;
#pragma unroll 1
    for (int sidx = 0; sidx < width; ++sidx) {
      uint64_t best = 0ull;
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        const uint64_t key = (static_cast<uint64_t>(float_to_ordered(x[j])) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(j));
        if (j < n && key < prev && key > best) best = key;
      }
      oi[sidx] = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(best & 0xFFFFFFFFull));
      ov[sidx] = ordered_to_float(static_cast<uint32_t>(best >> 32));
      prev = best;
    }
    for (int sidx = width; sidx < kKnnWidth; ++sidx) {
      oi[sidx] = -1;
      ov[sidx] = 0.f;
    }
  }
  if (lane == 0) {
    double* st = p.doc_stats + static_cast<size_t>(doc) * 8;
    st[0] = static_cast<double>(mu);
    st[1] = static_cast<double>(sigma);
    st[2] = q_out[0];
    st[3] = q_out[1];
    st[4] = q_out[2];
    double sd = 0.0;
    if (m > 0) {
      const double pm = pos_s1 / m_d;
      sd = sqrt(fmax(pos_s2 / m_d - pm * pm, 0.0));
    }
    st[5] = 0.1 * sd;
    st[6] = m_d;
    st[7] = static_cast<double>(kk);
  }
}

__global__ void __launch_bounds__(kSmallWarps * 32) group_threshold_small_kernel(const GroupParams p) {
  const int lane = threadIdx.x & 31;
  const int doc = blockIdx.x * kSmallWarps + (threadIdx.x >> 5);
  if (doc >= p.n_docs) return;
  const int row_base = p.offsets[doc];
  const int n = p.offsets[doc + 1] - row_base;
  if (n < 1 || n > kSmallMaxN) return;  // empty and larger documents belong to the CTA kernels
  if (n <= 8) group_small_doc<8>(p, doc, row_base, n, lane);
  else if (n <= 16) group_small_doc<16>(p, doc, row_base, n, lane);
  else group_small_doc<32>(p, doc, row_base, n, lane);
}

// NT = 512 threads for documents of more than 128 sentences, 128 threads for 33..128: a mid-size document
// keeps only a few warps busy, and every block barrier costs the idle ones.
template <int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 5) group_threshold_kernel(const GroupParams p, int n_min, int n_max) {
  constexpr int kGrpThreads = NT;
  constexpr int kGrpWarps = NT / 32;
  __shared__ double red[kGrpWarps];
  __shared__ unsigned int hist[kNumQ][kRadixBins];
  __shared__ uint64_t knn_scr[kGrpWarps][64];
  __shared__ unsigned int t_prefix[kNumQ], t_rank[kNumQ];  // per target: determined high bits, remaining rank
  __shared__ unsigned int t_cnt_le[kNumQ], t_min_above[kNumQ], t_fill[kNumQ];
  __shared__ float s_mu, s_sigma;

  const int doc = blockIdx.x;
  const int row_base = p.offsets[doc];
  const int n = p.offsets[doc + 1] - row_base;
  double* st = p.doc_stats + static_cast<size_t>(doc) * 8;
  if (n < 1) {
    if (threadIdx.x < 8) st[threadIdx.x] = 0.0;
    return;
  }
  if (n <= kSmallMaxN) return;  // handled by group_threshold_small_kernel (one warp per document)
  if (n < n_min || n > n_max) return;  // the other block size's share
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
  {
    // eight independent loads per thread and iteration: the pass is a pure stream over S and would
    // otherwise be bound by load latency (one 4-byte load in flight per thread)
    double t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};
    long long i0 = 0;
    for (; i0 + 8 * kGrpThreads <= nn; i0 += 8 * kGrpThreads) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = S[i0 + q * kGrpThreads + tid];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double d = static_cast<double>(v[q]);
        t1[q & 1] += d;
        t2[q & 1] += d * d;
      }
    }
    for (long long i = i0 + tid; i < nn; i += kGrpThreads) {
      const double d = static_cast<double>(S[i]);
      t1[0] += d;
      t2[0] += d * d;
    }
    s1 = t1[0] + t1[1];
    s2 = t2[0] + t2[1];
  }
  for (int i = tid; i < kRadixBins; i += kGrpThreads) hist[0][i] = 0u;  // pass-0 histogram (shared by the three targets)
  s1 = block_sum<NT>(s1, red);
  s2 = block_sum<NT>(s2, red);
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
  // symmetric S on the register path: the histogram, the moments and the gather pass below see the strict upper triangle
  // only; every count, sum and rank of the full multiset is twice / half of what they see
  const bool sym = p.symmetric != 0 && n <= kFastMaxN;
  if (n <= kFastMaxN) {
    // Fast path: one warp per row, the row stays in registers for all four products of the pass.
    RowCtx cx;
    cx.S = S; cx.sharp = sharp; cx.n = n; cx.row_base = row_base; cx.width = width;
    cx.mu = mu; cx.sigma = sigma; cx.tau = tau; cx.hist = hist[0]; cx.scr = knn_scr[warp];
    cx.centrality = p.centrality; cx.knn_idx = p.knn_idx; cx.knn_val = p.knn_val;
    if (sym) {
      if (n <= 128) row_pass_regs<4, kGrpWarps, true>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
      else if (n <= 256) row_pass_regs<8, kGrpWarps, true>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
      else if (n <= 384) row_pass_regs<12, kGrpWarps, true>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
      else row_pass_regs<16, kGrpWarps, true>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
    } else {
      if (n <= 128) row_pass_regs<4, kGrpWarps, false>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
      else if (n <= 256) row_pass_regs<8, kGrpWarps, false>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
      else if (n <= 384) row_pass_regs<12, kGrpWarps, false>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
      else row_pass_regs<16, kGrpWarps, false>(cx, warp, lane, pos_s1, pos_s2, pos_cnt);
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
        hist_add_aggregated(hist[0], lin_bin(__uint_as_float(bits)), pos, lane);
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
  pos_s1 = block_sum<NT>(pos_s1, red) * (sym ? 2.0 : 1.0);
  pos_s2 = block_sum<NT>(pos_s2, red) * (sym ? 2.0 : 1.0);
  const double m_d = block_sum<NT>(static_cast<double>(pos_cnt), red) * (sym ? 2.0 : 1.0);
  const unsigned int m = static_cast<unsigned int>(m_d + 0.5);
  const int rsh = sym ? 1 : 0;  // rank r of the full multiset = rank r >> rsh of what the histogram counted
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
    // 3a. walk the linear-bin histogram: bin, rank inside the bin and population of the bin per target
    if (warp < kNumQ) {
      unsigned int bin, rin, cnt;
      walk_hist(hist[0], kRadixBins, lo_rank[warp] >> rsh, lane, bin, rin, cnt);
      if (lane == 0) {
        t_prefix[warp] = bin;
        t_rank[warp] = rin;
        t_cnt_le[warp] = cnt;
        t_min_above[warp] = 0xFFFFFFFFu;
      }
    }
    __syncthreads();
    const unsigned int b0 = t_prefix[0], b1 = t_prefix[1], b2 = t_prefix[2];
    const unsigned int n0 = t_cnt_le[0], n1 = t_cnt_le[1], n2 = t_cnt_le[2];
    if (n0 <= kListCap && n1 <= kListCap && n2 <= kListCap) {
      // 3b. fast path: ONE more pass over sim_sharp gathers each target bin's values into shared memory
      // (the histogram is dead now and becomes the three lists) and tracks the smallest value above each bin
      unsigned int* lists = &hist[0][0];
      __syncthreads();
      if (tid < kNumQ) t_fill[tid] = 0u;
      __syncthreads();
      unsigned int a0 = 0xFFFFFFFFu, a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
      auto visit = [&](unsigned int b) {
        if (b == 0u) return;  // diagonal / underflowed zeros are not "positive values"
        const unsigned int bin = lin_bin(__uint_as_float(b));
        if (bin == b0) lists[atomicAdd(&t_fill[0], 1u)] = b; else if (bin > b0) a0 = min(a0, b);
        if (bin == b1) lists[kListCap + atomicAdd(&t_fill[1], 1u)] = b; else if (bin > b1) a1 = min(a1, b);
        if (bin == b2) lists[2 * kListCap + atomicAdd(&t_fill[2], 1u)] = b; else if (bin > b2) a2 = min(a2, b);
      };
      if (sym) {
        // strict upper triangle, one warp per row, four independent loads per lane and step
        for (int r = warp; r < n - 1; r += kGrpWarps) {
          const float* row = sharp + static_cast<size_t>(r) * n;
          for (int c0 = r + 1; c0 < n; c0 += 128) {
            unsigned int bb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c = c0 + q * 32 + lane;
              bb[q] = c < n ? __float_as_uint(row[c]) : 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) visit(bb[q]);
          }
        }
      } else {
        for (long long i0 = 0; i0 < nn; i0 += 8 * kGrpThreads) {
          unsigned int bb[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const long long i = i0 + q * kGrpThreads + tid;
            bb[q] = i < nn ? __float_as_uint(sharp[i]) : 0u;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) visit(bb[q]);
        }
      }
      a0 = __reduce_min_sync(0xffffffffu, a0);
      a1 = __reduce_min_sync(0xffffffffu, a1);
      a2 = __reduce_min_sync(0xffffffffu, a2);
      if (lane == 0) {
        atomicMin(&t_min_above[0], a0);
        atomicMin(&t_min_above[1], a1);
        atomicMin(&t_min_above[2], a2);
      }
      __syncthreads();
      // 3c. sort each list (positive floats order like their bit patterns) and read the two order statistics
#pragma unroll 1
      for (int t = 0; t < kNumQ; ++t) {
        const unsigned int cnt = t_cnt_le[t];
        unsigned int* l = lists + t * kListCap;
        int np2 = 2;
        while (np2 < static_cast<int>(cnt)) np2 <<= 1;
        for (int i = static_cast<int>(cnt) + tid; i < np2; i += kGrpThreads) l[i] = 0xFFFFFFFFu;
        __syncthreads();
        block_bitonic_asc_u32(l, np2);
        const unsigned int rin = t_rank[t];
        const unsigned int vlo = l[rin];
        unsigned int vhi = vlo;
        // order statistic lo + 1 of the full multiset: the same counted value when both fall on one of its two copies
        if (lo_rank[t] + 1 < m && ((lo_rank[t] + 1) >> rsh) != (lo_rank[t] >> rsh)) vhi = (rin + 1 < cnt) ? l[rin + 1] : t_min_above[t];
        q_out[t] = np_lerp(static_cast<double>(__uint_as_float(vlo)), static_cast<double>(__uint_as_float(vhi)), gamma[t]);
      }
    } else {
      // 3d. fallback (a bin too crowded for shared memory, e.g. thousands of identical similarities):
      // 3-pass (11/11/10-bit) radix select on the fp32 bit patterns, every pass re-reading sim_sharp
      __syncthreads();
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
        for (int i = tid; i < kNumQ * kRadixBins; i += kGrpThreads) (&hist[0][0])[i] = 0u;
        __syncthreads();
        const unsigned int pf0 = t_prefix[0], pf1 = t_prefix[1], pf2 = t_prefix[2];
        for (long long i0 = 0; i0 < nn; i0 += kGrpThreads) {  // warp-uniform trip count: the aggregation needs all lanes
          const long long i = i0 + tid;
          const unsigned int b = i < nn ? __float_as_uint(sharp[i]) : 0u;
          const unsigned int hi = b & known_mask, dg = (b >> sh) & dmask;
          hist_add_aggregated(hist[0], dg, b != 0u && hi == pf0, lane);
          hist_add_aggregated(hist[1], dg, b != 0u && hi == pf1, lane);
          hist_add_aggregated(hist[2], dg, b != 0u && hi == pf2, lane);
        }
        __syncthreads();
        if (warp < kNumQ) {
          unsigned int bin, rin, cnt;
          walk_hist(hist[warp], 1 << widths[pass], t_rank[warp], lane, bin, rin, cnt);
          if (lane == 0) {
            t_prefix[warp] |= bin << sh;
            t_rank[warp] = rin;
          }
        }
        known_mask |= dmask << sh;
        __syncthreads();
      }
      // upper order statistic: next distinct value unless duplicates already cover rank lo+1
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
                                       int knn_mode, int s_is_symmetric, float* out_sharp, double* out_centrality,
                                       double* out_doc_stats, int32_t* out_knn_idx, float* out_knn_val, void* stream) {
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
  p.symmetric = s_is_symmetric != 0;
  p.sharp = out_sharp;
  p.centrality = out_centrality;
  p.doc_stats = out_doc_stats;
  p.knn_idx = out_knn_idx;
  p.knn_val = out_knn_val;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfileScope prof(st);
    group_threshold_kernel<512><<<n_docs, 512, 0, st>>>(p, 129, 0x7fffffff);
    group_threshold_kernel<128><<<n_docs, 128, 0, st>>>(p, kSmallMaxN + 1, 128);
    group_threshold_small_kernel<<<(n_docs + kSmallWarps - 1) / kSmallWarps, kSmallWarps * 32, 0, st>>>(p);
  }
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
