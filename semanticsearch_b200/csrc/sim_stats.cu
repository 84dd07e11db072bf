// Similarity-distribution statistics of the strict upper triangle of each document's S matrix.
//
// Device form of analyze_similarity_distribution (Method/semantic_common.py:250-270): values
// >= 1 - 1e-5 are dropped, then min / max / mean / std and np.percentile at
// [10, 25, 50, 75, 80, 85, 90, 95].  numpy's percentile on a float32 array works entirely in
// float32 (q = p / float32(100), virtual index (m-1)*q, gamma and the lerp), which is restated
// here; the order statistics come from a 3-pass radix select over order-preserving keys, eight
// targets at a time.  One CTA per document.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

constexpr int kStatThreads = 512;
constexpr int kStatWarps = kStatThreads / 32;
constexpr int kStatBins = 2048;
constexpr int kStatQ = 8;
constexpr int kStatOut = 13;  // count, min, max, mean, std, p10, p25, p50, p75, p80, p85, p90, p95

__device__ __forceinline__ double stat_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kStatWarps; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(kStatThreads) sim_distribution_kernel(const float* __restrict__ S_all, const int* __restrict__ offsets,
                                                                        const long long* __restrict__ s_offsets, int n_docs, float cut,
                                                                        double* __restrict__ out_all) {
  extern __shared__ unsigned int stat_hist[];  // [kStatQ][kStatBins]
  __shared__ double red[kStatWarps];
  __shared__ unsigned int t_prefix[kStatQ], t_rank[kStatQ], t_cnt_le[kStatQ], t_min_above[kStatQ];
  __shared__ unsigned int s_min, s_max, s_max_all;

  const int doc = blockIdx.x;
  const int n = offsets[doc + 1] - offsets[doc];
  double* out = out_all + static_cast<size_t>(doc) * kStatOut;
  if (n < 2) {
    if (threadIdx.x < kStatOut) out[threadIdx.x] = threadIdx.x == 0 ? -1.0 : 0.0;  // count -1: "None" in the reference
    return;
  }
  const float* S = S_all + s_offsets[doc];
  const long long ne = static_cast<long long>(n) * (n - 1) / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    s_min = 0xFFFFFFFFu;
    s_max = 0u;
    s_max_all = 0u;
  }
  __syncthreads();

  // ---- pass A: count / min / max / moments of the kept values ---------------------------------
  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  unsigned int mn = 0xFFFFFFFFu, mx = 0u, mx_all = 0u;
  for (int i = warp; i < n - 1; i += kStatWarps) {  // one warp per row keeps the loads coalesced
    const float* row = S + static_cast<size_t>(i) * n;
    for (int j = i + 1 + lane; j < n; j += 32) {
      const float v = row[j];
      const unsigned int key = float_to_ordered(v);
      mx_all = max(mx_all, key);
      if (v < cut) {
        s1 += static_cast<double>(v);
        s2 += static_cast<double>(v) * static_cast<double>(v);
        cnt += 1.0;
        mn = min(mn, key);
        mx = max(mx, key);
      }
    }
  }
  atomicMin(&s_min, mn);
  atomicMax(&s_max, mx);
  atomicMax(&s_max_all, mx_all);
  s1 = stat_block_sum(s1, red);
  s2 = stat_block_sum(s2, red);
  cnt = stat_block_sum(cnt, red);
  __syncthreads();
  const unsigned int m = static_cast<unsigned int>(cnt + 0.5);
  if (m == 0) {  // everything filtered: the reference reports max(sims) for every key (:257-260)
    if (tid < kStatOut) out[tid] = tid == 0 ? 0.0 : static_cast<double>(ordered_to_float(s_max_all));
    return;
  }

  // ---- float32 virtual indices, exactly like numpy on a float32 array --------------------------
  const float pcts[kStatQ] = {10.f, 25.f, 50.f, 75.f, 80.f, 85.f, 90.f, 95.f};
  unsigned int lo_rank[kStatQ];
  float gamma[kStatQ];
#pragma unroll
  for (int t = 0; t < kStatQ; ++t) {
    const float q = __fdiv_rn(pcts[t], 100.0f);
    const float vi = __fmul_rn(static_cast<float>(m - 1), q);
    const float fl = floorf(vi);
    lo_rank[t] = min(static_cast<unsigned int>(fl), m - 1);
    gamma[t] = __fsub_rn(vi, fl);
  }
  if (tid < kStatQ) {
    t_prefix[tid] = 0u;
    t_rank[tid] = lo_rank[tid];
  }
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  unsigned int known_mask = 0u;
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = tid; i < kStatQ * kStatBins; i += kStatThreads) stat_hist[i] = 0u;
    __syncthreads();
    unsigned int pf[kStatQ];
#pragma unroll
    for (int t = 0; t < kStatQ; ++t) pf[t] = t_prefix[t];
    const int sh = shifts[pass];
    const unsigned int dmask = (1u << widths[pass]) - 1u;
    for (int i = warp; i < n - 1; i += kStatWarps) {
      const float* row = S + static_cast<size_t>(i) * n;
      for (int j = i + 1 + lane; j < n; j += 32) {
        const float v = row[j];
        if (!(v < cut)) continue;
        const unsigned int b = float_to_ordered(v);
        const unsigned int hi = b & known_mask, dg = (b >> sh) & dmask;
        // targets that share a prefix share a histogram row (the first of them)
        unsigned int done = 0u;
#pragma unroll
        for (int t = 0; t < kStatQ; ++t) {
          if (hi == pf[t] && !((done >> t) & 1u)) {
            atomicAdd(&stat_hist[t * kStatBins + dg], 1u);
#pragma unroll
            for (int t2 = t + 1; t2 < kStatQ; ++t2)
              if (pf[t2] == pf[t]) done |= 1u << t2;
          }
        }
      }
    }
    __syncthreads();
    if (warp < kStatQ) {
      // use the histogram row of the first target with the same prefix
      int src_row = warp;
      for (int t = 0; t < warp; ++t)
        if (t_prefix[t] == t_prefix[warp]) {
          src_row = t;
          break;
        }
      const unsigned int want = t_rank[warp];
      unsigned int run = 0u, found_bin = 0u, found_excl = 0u;
      bool found = false;
      const int nb = 1 << widths[pass];
      for (int b0 = 0; b0 < nb && !found; b0 += 32) {
        const unsigned int c = stat_hist[src_row * kStatBins + b0 + lane];
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
          found_excl = __shfl_sync(0xffffffffu, excl, src);
          found_bin = static_cast<unsigned int>(b0 + src);
          found = true;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      __syncwarp();
      // all targets read their source rows before anyone updates a prefix
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (lane == 0 && found) {
        t_prefix[warp] |= found_bin << sh;
        t_rank[warp] = want - found_excl;
      }
    }
    known_mask |= dmask << sh;
    __syncthreads();
  }
  if (tid < kStatQ) {
    t_cnt_le[tid] = 0u;
    t_min_above[tid] = 0xFFFFFFFFu;
  }
  __syncthreads();
  unsigned int vt[kStatQ], cle[kStatQ], mab[kStatQ];
#pragma unroll
  for (int t = 0; t < kStatQ; ++t) {
    vt[t] = t_prefix[t];
    cle[t] = 0u;
    mab[t] = 0xFFFFFFFFu;
  }
  for (int i = warp; i < n - 1; i += kStatWarps) {
    const float* row = S + static_cast<size_t>(i) * n;
    for (int j = i + 1 + lane; j < n; j += 32) {
      const float v = row[j];
      if (!(v < cut)) continue;
      const unsigned int b = float_to_ordered(v);
#pragma unroll
      for (int t = 0; t < kStatQ; ++t) {
        if (b <= vt[t]) ++cle[t]; else mab[t] = min(mab[t], b);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kStatQ; ++t) {
    atomicAdd(&t_cnt_le[t], cle[t]);
    atomicMin(&t_min_above[t], mab[t]);
  }
  __syncthreads();
  if (tid == 0) {
    const double mean = s1 / cnt;
    const double var = fmax(s2 / cnt - mean * mean, 0.0);
    out[0] = cnt;
    out[1] = static_cast<double>(ordered_to_float(s_min));
    out[2] = static_cast<double>(ordered_to_float(s_max));
    out[3] = static_cast<double>(static_cast<float>(mean));       // np.mean of float32 -> float32
    out[4] = static_cast<double>(static_cast<float>(sqrt(var)));  // np.std of float32 -> float32
    for (int t = 0; t < kStatQ; ++t) {
      const unsigned int klo = t_prefix[t];
      unsigned int khi = klo;
      if (lo_rank[t] + 1 < m && t_cnt_le[t] <= lo_rank[t] + 1) khi = t_min_above[t];
      const float a = ordered_to_float(klo), b = ordered_to_float(khi), g = gamma[t];
      const float diff = __fsub_rn(b, a);
      float r = __fadd_rn(a, __fmul_rn(diff, g));
      if (g >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, g)));
      out[5 + t] = static_cast<double>(r);
    }
  }
}

}  // namespace ss

using namespace ss;

extern "C" int ss_similarity_distribution(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, float eps,
                                          double* out_stats, void* stream) {
  if (!S || !offsets || !s_offsets || !out_stats) return fail(SS_ERR_INVALID_ARG, "ss_similarity_distribution: null pointer");
  if (n_docs <= 0) return fail(SS_ERR_INVALID_ARG, "ss_similarity_distribution: n_docs must be positive");
  const size_t smem = static_cast<size_t>(kStatQ) * kStatBins * sizeof(unsigned int);
  SS_CUDA_CHECK(cudaFuncSetAttribute(sim_distribution_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const float cut = static_cast<float>(1.0 - static_cast<double>(eps));  // the Python float 1 - 1e-5 compared as float32
  sim_distribution_kernel<<<n_docs, kStatThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      S, offsets, reinterpret_cast<const long long*>(s_offsets), n_docs, cut, out_stats);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
