// §8f-2 — block sums of sim_sharp over cluster member lists, and the co-association matrix of a label sweep.
//
// The reference's clustering stage evaluates `_mean_between(A, B)` / `_mean_within(A)` with pure-Python double loops
// over sim_sharp (Method/Semantic_Grouping_Optimized.py:118-130), once per candidate pair inside the merge / refine
// loops (:444-553), and one mean per (sentence, cluster) in the reassignment pass (:566-588).  Every one of those
// numbers is a quotient of two sums this file produces for ALL clusters of a document at once:
//
//   rowsum[x][g] = sum_{y in members_g} sim_sharp[x][y]          (float64, members in list order)
//   block[a][b]  = sum_{x in members_a} rowsum[x][b]             (float64, members in list order)
//
//   _mean_between(A, B) = block[A][B] / (|A| |B|)
//   _mean_within(A)     = block[A][A] / (|A| (|A| - 1))          sim_sharp is bit-symmetric with a zero diagonal, so
//                                                                  the i < j pairs are half of the full block
//   reassignment mean of sentence x against cluster g = rowsum[x][g] / |g|
//
// Member lists are multisets (the reference's merge step can list a sentence twice, :478-487, and np.ix_ then counts
// it twice); sums over unions are additive, so the host derives every merged-candidate mean from one launch.
// Summation order is fixed (lane-strided partials + butterfly), so results are reproducible run to run.
//
// ss_group_coassociation is the consensus matrix of the Louvain resolution sweep (:231-241):
//   C[i][j] = #{l : labels[l][i] == labels[l][j]} / L  for i != j, 0 on the diagonal.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

struct BlockSumParams {
  const float* sharp;            // packed per-document matrices (layout of K3 / K4)
  const int* offsets;            // [n_docs + 1] row offsets
  const long long* s_offsets;    // [n_docs + 1] element offsets of each document's matrix
  const int* group_prefix;       // [n_docs + 1] first group of each document
  const int* member_prefix;      // [n_groups + 1] first member of each group
  const int* members;            // document-local row indices
  const long long* rs_offsets;   // [n_docs] start of the document's rowsum block (n x G doubles)
  const long long* blk_offsets;  // [n_docs] start of the document's block matrix (G x G doubles)
  int n_docs;
  double* rowsum;
  double* block;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kBlkThreads = 256;

// grid = total rows; one CTA per sentence row x, one warp per group in turn
__global__ void __launch_bounds__(kBlkThreads) group_rowsum_kernel(const BlockSumParams p, const int* __restrict__ row_doc) {
  const int grow = blockIdx.x;
  const int doc = row_doc[grow];
  const int row_base = p.offsets[doc];
  const int n = p.offsets[doc + 1] - row_base;
  const int x = grow - row_base;
  const int g0 = p.group_prefix[doc], G = p.group_prefix[doc + 1] - g0;
  if (G <= 0) return;
  const float* srow = p.sharp + p.s_offsets[doc] + static_cast<long long>(x) * n;
  double* out = p.rowsum + p.rs_offsets[doc] + static_cast<long long>(x) * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < G; g += kBlkThreads / 32) {
    const int m0 = p.member_prefix[g0 + g], m1 = p.member_prefix[g0 + g + 1];
    double acc = 0.0;
    for (int i = m0 + lane; i < m1; i += 32) {
      const int y = p.members[i];
      if (y >= 0 && y < n) acc += static_cast<double>(__ldg(srow + y));
    }
    acc = warp_sum(acc);
    if (lane == 0) out[g] = acc;
  }
}

// grid = documents; one warp per (a, b) pair in turn
__global__ void __launch_bounds__(kBlkThreads) group_block_kernel(const BlockSumParams p) {
  const int doc = blockIdx.x;
  const int n = p.offsets[doc + 1] - p.offsets[doc];
  const int g0 = p.group_prefix[doc], G = p.group_prefix[doc + 1] - g0;
  if (G <= 0) return;
  const double* rs = p.rowsum + p.rs_offsets[doc];
  double* out = p.block + p.blk_offsets[doc];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pair = warp; pair < G * G; pair += kBlkThreads / 32) {
    const int a = pair / G, b = pair - a * G;
    const int m0 = p.member_prefix[g0 + a], m1 = p.member_prefix[g0 + a + 1];
    double acc = 0.0;
    for (int i = m0 + lane; i < m1; i += 32) {
      const int x = p.members[i];
      if (x >= 0 && x < n) acc += rs[static_cast<long long>(x) * G + b];
    }
    acc = warp_sum(acc);
    if (lane == 0) out[pair] = acc;
  }
}

__global__ void __launch_bounds__(256) coassociation_kernel(const int* __restrict__ labels, int n_labelings, int n,
                                                            double* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n) * n) return;
  const int i = static_cast<int>(idx / n), j = static_cast<int>(idx - static_cast<long long>(i) * n);
  double c = 0.0;
  if (i != j) {
    int same = 0;
    for (int l = 0; l < n_labelings; ++l) same += (labels[static_cast<long long>(l) * n + i] == labels[static_cast<long long>(l) * n + j]) ? 1 : 0;
    c = static_cast<double>(same) / static_cast<double>(n_labelings);
  }
  out[idx] = c;
}

}  // namespace ss

using namespace ss;

extern "C" int ss_group_block_sums(const float* sharp, const int32_t* offsets, const int64_t* s_offsets, int n_docs, int total_rows,
                                   const int32_t* row_doc, const int32_t* group_prefix, const int32_t* member_prefix,
                                   const int32_t* members, const int64_t* rowsum_offsets, const int64_t* block_offsets,
                                   double* out_rowsum, double* out_block, void* stream) {
  if (!sharp || !offsets || !s_offsets || !row_doc || !group_prefix || !member_prefix || !members || !rowsum_offsets ||
      !block_offsets || !out_rowsum || !out_block)
    return fail(SS_ERR_INVALID_ARG, "ss_group_block_sums: null pointer");
  if (n_docs <= 0 || total_rows <= 0) return fail(SS_ERR_INVALID_ARG, "ss_group_block_sums: n_docs and total_rows must be positive");
  BlockSumParams p;
  p.sharp = sharp;
  p.offsets = offsets;
  p.s_offsets = reinterpret_cast<const long long*>(s_offsets);
  p.group_prefix = group_prefix;
  p.member_prefix = member_prefix;
  p.members = members;
  p.rs_offsets = reinterpret_cast<const long long*>(rowsum_offsets);
  p.blk_offsets = reinterpret_cast<const long long*>(block_offsets);
  p.n_docs = n_docs;
  p.rowsum = out_rowsum;
  p.block = out_block;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfileScope prof(st);
    group_rowsum_kernel<<<total_rows, kBlkThreads, 0, st>>>(p, row_doc);
    group_block_kernel<<<n_docs, kBlkThreads, 0, st>>>(p);
  }
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}

extern "C" int ss_group_coassociation(const int32_t* labels, int n_labelings, int n, double* out_C, void* stream) {
  if (!labels || !out_C) return fail(SS_ERR_INVALID_ARG, "ss_group_coassociation: null pointer");
  if (n_labelings <= 0 || n <= 0) return fail(SS_ERR_INVALID_ARG, "ss_group_coassociation: sizes must be positive");
  const long long total = static_cast<long long>(n) * n;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  coassociation_kernel<<<static_cast<unsigned int>((total + 255) / 256), 256, 0, st>>>(labels, n_labelings, n, out_C);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
