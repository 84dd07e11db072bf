// C99 rank transform of per-document similarity matrices (Method/Semantic_Splitter_Optimized.py:171-192).
//
//   global: R[i,j] = #{k : S[i,k] < S[i,j]} + #{k : S[k,j] < S[i,j]}            (:189-192)
//   local : R[i,j] = #{(a,b) in the clipped m x m window around (i,j) : S[a,b] < S[i,j]} / window size  (:171-186)
//
// Pure compare-and-count (integer work).  The reference does this with an n^3 boolean broadcast
// (134 MB at n = 512) or 262 144 Python iterations per document.
//
// Global mode, documents of up to 2048 sentences: #{k : x[k] < x[j]} along a line (row or column) is the
// lower bound of x[j] in the sorted line, so each warp sorts one line in registers (bitonic network,
// 32 * NPL values, element g = lane * NPL + r: the short-distance stages are register swaps, only the
// long-distance ones shuffle), parks the sorted line in shared memory and binary-searches it once per
// element: O(n^2 log^2 n) per document instead of O(n^3).  Columns go first, through a shared-memory panel of
// 32 (8 for n > 512) adjacent columns that the CTA fills with cp.async and stores back row-segment by
// row-segment (whole sectors); rows follow, each warp adding its ranks to R with coalesced accesses.
// Local mode and longer documents: one CTA per (document, row) counting kernel.
#include <algorithm>
#include <cstdlib>

#include "ss_common.cuh"

namespace ss {

__global__ void __launch_bounds__(256) c99_rank_kernel(const float* __restrict__ S_all, const int* __restrict__ offsets,
                                                       const long long* __restrict__ s_offsets, const int* __restrict__ row_doc,
                                                       int local_mode, int half, float* __restrict__ R_all) {
  extern __shared__ float c99_row[];  // the whole row i (global mode) or the window rows' needed span is read from global
  const int grow = blockIdx.x;         // global row index in the concatenated batch
  const int doc = row_doc[grow];
  const int row_base = offsets[doc];
  const int n = offsets[doc + 1] - row_base;
  const int i = grow - row_base;
  const float* S = S_all + s_offsets[doc];
  float* R = R_all + s_offsets[doc];
  if (!local_mode) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) c99_row[k] = S[static_cast<size_t>(i) * n + k];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const float v = c99_row[j];
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        cnt += (c99_row[k] < v) ? 1 : 0;                              // row i: shared-memory broadcast
        cnt += (S[static_cast<size_t>(k) * n + j] < v) ? 1 : 0;       // column j: coalesced across threads
      }
      R[static_cast<size_t>(i) * n + j] = static_cast<float>(cnt);
    }
  } else {
    const int i0 = max(0, i - half), i1 = min(n, i + half + 1);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int j0 = max(0, j - half), j1 = min(n, j + half + 1);
      const float v = S[static_cast<size_t>(i) * n + j];
      int cnt = 0;
      for (int a = i0; a < i1; ++a)
        for (int b = j0; b < j1; ++b) cnt += (S[static_cast<size_t>(a) * n + b] < v) ? 1 : 0;
      const int denom = (i1 - i0) * (j1 - j0);
      R[static_cast<size_t>(i) * n + j] = static_cast<float>(static_cast<double>(cnt) / static_cast<double>(denom > 0 ? denom : 1));
    }
  }
}

// ---- global mode by sorting -------------------------------------------------------------------------------
constexpr int kRankWarps = 8;
constexpr int kRankThreads = kRankWarps * 32;
constexpr int kRankSortMaxRows = 2048;

__device__ __forceinline__ int sorted_slot(int g) { return g + (g >> 5); }  // lane-major stores without bank conflicts

// Ascending bitonic sort of 32 * NPL values held as v[r] of every lane, sort index g = lane * NPL + r.
template <int NPL>
__device__ __forceinline__ void warp_sort_asc(float (&v)[NPL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * NPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= NPL) {
        const int lane_mask = j / NPL;
        // ascending block when bit k of g is clear; k > j >= NPL, so that bit is a lane bit (k = 32 * NPL: always clear)
        const bool up = ((lane * NPL) & k) == 0;
        const bool keep_min = up == ((lane & lane_mask) == 0);
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const float o = __shfl_xor_sync(0xffffffffu, v[r], lane_mask);
          v[r] = keep_min ? fminf(v[r], o) : fmaxf(v[r], o);
        }
      } else {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const int pr = r ^ j;
          if (pr > r) {
            const bool up = (((lane * NPL) + r) & k) == 0;
            const float lo = fminf(v[r], v[pr]), hi = fmaxf(v[r], v[pr]);
            v[r] = up ? lo : hi;
            v[pr] = up ? hi : lo;
          }
        }
      }
    }
  }
}

// Strict ranks of one line: src[e] for e < n (global or shared memory) -> dst[e] = #{k : src[k] < src[e]} (+ the previous
// dst[e] when ACC).  dst may alias src (the column panel is ranked in place).  `sorted`: 33 * NPL floats of warp-private
// shared memory.
template <int NPL, bool ACC>
__device__ __forceinline__ void rank_line(const float* src, float* dst, int n, float* sorted, int lane) {
  float v[NPL];
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int e = r * 32 + lane;
    v[r] = e < n ? src[e] : INFINITY;
  }
  float prev[ACC ? NPL : 1];
  if constexpr (ACC) {  // issued now, consumed after the sort
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
      const int e = r * 32 + lane;
      prev[r] = e < n ? dst[e] : 0.f;
    }
  }
  warp_sort_asc<NPL>(v, lane);
#pragma unroll
  for (int r = 0; r < NPL; ++r) sorted[sorted_slot(lane * NPL + r)] = v[r];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int e = r * 32 + lane;
    v[r] = e < n ? src[e] : INFINITY;  // re-read (cache / shared memory) instead of holding a second register copy
  }
  __syncwarp();                        // every lane has its values before an in-place dst is written
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const int e = r * 32 + lane;
    if (e < n) {
      int pos = 0;
#pragma unroll
      for (int step = 16 * NPL; step > 0; step >>= 1)
        if (sorted[sorted_slot(pos + step - 1)] < v[r]) pos += step;
      if constexpr (ACC) dst[e] = static_cast<float>(pos) + prev[r];
      else dst[e] = static_cast<float>(pos);
    }
  }
  __syncwarp();  // `sorted` is reused by the warp's next line
}

template <int MAXNPL, bool ACC>
__device__ __forceinline__ void rank_line_any(const float* src, float* dst, int n, float* sorted, int lane) {
  if (n <= 32) return rank_line<1, ACC>(src, dst, n, sorted, lane);
  if (n <= 64) return rank_line<2, ACC>(src, dst, n, sorted, lane);
  if (n <= 128) return rank_line<4, ACC>(src, dst, n, sorted, lane);
  if constexpr (MAXNPL >= 16) {
    if (n <= 256) return rank_line<8, ACC>(src, dst, n, sorted, lane);
    if (n <= 512) return rank_line<16, ACC>(src, dst, n, sorted, lane);
  }
  if constexpr (MAXNPL >= 64) {
    if (n <= 1024) return rank_line<32, ACC>(src, dst, n, sorted, lane);
    return rank_line<64, ACC>(src, dst, n, sorted, lane);
  }
}

// R[i][j] += #{k : S[i][k] < S[i][j]} (second pass: coalesced read-modify-write): one warp per row of the concatenated batch.
template <int MAXNPL, bool ACC>
__global__ void __launch_bounds__(kRankThreads) c99_rank_rows_kernel(const float* __restrict__ S_all, const int* __restrict__ offsets,
                                                                     const long long* __restrict__ s_offsets,
                                                                     const int* __restrict__ row_doc, int total_rows,
                                                                     float* __restrict__ R_all) {
  extern __shared__ float rank_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grow = blockIdx.x * kRankWarps + warp;
  if (grow >= total_rows) return;
  const int doc = row_doc[grow];
  const int base = offsets[doc];
  const int n = offsets[doc + 1] - base;
  const size_t line = static_cast<size_t>(s_offsets[doc]) + static_cast<size_t>(grow - base) * n;
  rank_line_any<MAXNPL, ACC>(S_all + line, R_all + line, n, rank_smem + warp * (33 * MAXNPL), lane);
}

// R[i][j] = #{k : S[k][j] < S[i][j]} (first pass: plain stores): one CTA per PW adjacent columns of the concatenated
// batch (split at document ends).  The panel is filled with 4-byte cp.async copies — every thread has its whole share
// in flight at once — and written back without reading R.
template <int MAXNPL, int PW>
__global__ void __launch_bounds__(kRankThreads) c99_rank_cols_kernel(const float* __restrict__ S_all, const int* __restrict__ offsets,
                                                                     const long long* __restrict__ s_offsets,
                                                                     const int* __restrict__ row_doc, int total_rows,
                                                                     float* __restrict__ R_all) {
  extern __shared__ float rank_smem[];
  float* panel = rank_smem;                                       // [PW][n | 1]
  float* sorted = rank_smem + PW * (32 * MAXNPL + 1) + (threadIdx.x >> 5) * (33 * MAXNPL);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int g = blockIdx.x * PW;
  const int end = min(total_rows, g + PW);
  while (g < end) {
    const int doc = row_doc[g];
    const int base = offsets[doc];
    const int n = offsets[doc + 1] - base;
    const int j0 = g - base;
    const int w = min(end, base + n) - g;
    const int ldp = n | 1;
    const float* S = S_all + s_offsets[doc];
    float* R = R_all + s_offsets[doc];
    if (w == PW) {
      for (int idx = threadIdx.x; idx < n * PW; idx += kRankThreads) {
        const int k = idx / PW, c = idx % PW;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(panel + c * ldp + k)), "l"(S + static_cast<size_t>(k) * n + j0 + c) : "memory");
      }
    } else {
      for (int idx = threadIdx.x; idx < n * w; idx += kRankThreads) {
        const int k = idx / w, c = idx - k * w;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(panel + c * ldp + k)), "l"(S + static_cast<size_t>(k) * n + j0 + c) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int c = warp; c < w; c += kRankWarps) rank_line_any<MAXNPL, false>(panel + c * ldp, panel + c * ldp, n, sorted, lane);
    __syncthreads();
    if (w == PW) {
#pragma unroll 4
      for (int idx = threadIdx.x; idx < n * PW; idx += kRankThreads) {
        const int k = idx / PW, c = idx % PW;
        R[static_cast<size_t>(k) * n + j0 + c] = panel[c * ldp + k];
      }
    } else {
      for (int idx = threadIdx.x; idx < n * w; idx += kRankThreads) {
        const int k = idx / w, c = idx - k * w;
        R[static_cast<size_t>(k) * n + j0 + c] = panel[c * ldp + k];
      }
    }
    __syncthreads();
    g += w;
  }
}

// Symmetric S: #{k : S[k][j] < S[i][j]} = #{k : S[j][k] < S[j][i]}, i.e. the column ranks are the transposed row ranks.
// R holds the row ranks; CTA (doc, I) adds the transposed tile to every tile pair (I, J >= I) in place through shared
// memory (both directions coalesced).
__global__ void __launch_bounds__(256) c99_rank_symmetrize_kernel(const int* __restrict__ offsets, const long long* __restrict__ s_offsets,
                                                                  float* __restrict__ R_all) {
  __shared__ float ta[32][33], tb[32][33];
  const int doc = blockIdx.x, I = blockIdx.y;
  const int n = offsets[doc + 1] - offsets[doc];
  const int T = (n + 31) >> 5;
  if (I >= T) return;
  float* R = R_all + s_offsets[doc];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 row groups of 4 rows
  for (int J = I; J < T; ++J) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty * 4 + q;
      const int ia = I * 32 + r, ja = J * 32 + tx;  // tile (I, J)
      const int ib = J * 32 + r, jb = I * 32 + tx;  // tile (J, I)
      ta[r][tx] = (ia < n && ja < n) ? R[static_cast<size_t>(ia) * n + ja] : 0.f;
      tb[r][tx] = (ib < n && jb < n) ? R[static_cast<size_t>(ib) * n + jb] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty * 4 + q;
      const int ia = I * 32 + r, ja = J * 32 + tx;
      const int ib = J * 32 + r, jb = I * 32 + tx;
      if (ia < n && ja < n) R[static_cast<size_t>(ia) * n + ja] = ta[r][tx] + tb[tx][r];
      if (J != I && ib < n && jb < n) R[static_cast<size_t>(ib) * n + jb] = tb[r][tx] + ta[tx][r];
    }
    __syncthreads();
  }
}

template <int MAXNPL, int PW>
static int launch_rank_sorted(const float* S, const int32_t* offsets, const long long* s_offsets, const int* row_doc, int n_docs,
                              int total_rows, int max_doc_rows, bool symmetric, float* R, cudaStream_t st) {
  const size_t smem_rows = static_cast<size_t>(kRankWarps) * 33 * MAXNPL * sizeof(float);
  const size_t smem_cols = smem_rows + static_cast<size_t>(PW) * (32 * MAXNPL + 1) * sizeof(float);
  const int row_blocks = (total_rows + kRankWarps - 1) / kRankWarps;
  if (symmetric) {
    if (smem_rows > 48 * 1024)
      SS_CUDA_CHECK(cudaFuncSetAttribute(c99_rank_rows_kernel<MAXNPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_rows)));
    c99_rank_rows_kernel<MAXNPL, false><<<row_blocks, kRankThreads, smem_rows, st>>>(S, offsets, s_offsets, row_doc, total_rows, R);
    SS_CUDA_CHECK(cudaGetLastError());
    const dim3 grid(static_cast<unsigned>(n_docs), static_cast<unsigned>((max_doc_rows + 31) / 32));
    c99_rank_symmetrize_kernel<<<grid, 256, 0, st>>>(offsets, s_offsets, R);
    SS_CUDA_CHECK(cudaGetLastError());
    return SS_OK;
  }
  if (smem_rows > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(c99_rank_rows_kernel<MAXNPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_rows)));
  if (smem_cols > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(c99_rank_cols_kernel<MAXNPL, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_cols)));
  c99_rank_cols_kernel<MAXNPL, PW><<<(total_rows + PW - 1) / PW, kRankThreads, smem_cols, st>>>(S, offsets, s_offsets, row_doc, total_rows, R);
  SS_CUDA_CHECK(cudaGetLastError());
  c99_rank_rows_kernel<MAXNPL, true><<<row_blocks, kRankThreads, smem_rows, st>>>(S, offsets, s_offsets, row_doc, total_rows, R);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}

// ---- local mode on shared-memory tiles --------------------------------------------------------------------
// R[i][j] = #{(a, b) in the clipped (2H+1)^2 window : S[a][b] < S[i][j]} / window size (Splitter:171-186; the controller's default
// preset, simple_chunk_controller.py:1451).  One CTA per 32 x 64 tile of one document: the tile and its halo are staged in
// shared memory once (out-of-document entries are +inf and never count), warp w owns 8 adjacent columns, lane l row l.
// For every window row a thread loads the 8 + 2H values its 8 outputs share and compares each against the centres it
// belongs to: 25 shared-memory loads per output instead of 121 global ones, conflict-free (odd row pitch).
constexpr int kLrRows = 32, kLrCols = 64, kLrCpt = 8, kLrThreads = 256;

__global__ void c99_tile_list_kernel(const int* __restrict__ offsets, int n_docs, int* __restrict__ count, int2* __restrict__ list) {
  const int doc = blockIdx.x * blockDim.x + threadIdx.x;
  if (doc >= n_docs) return;
  const int n = offsets[doc + 1] - offsets[doc];
  const int t = (n + kLrRows - 1) / kLrRows;
  if (t <= 0) return;
  const int base = atomicAdd(count, t);
  for (int i = 0; i < t; ++i) list[base + i] = make_int2(doc, i);
}

template <int H>
__global__ void __launch_bounds__(kLrThreads) c99_local_rank_kernel(const float* __restrict__ S_all, const int* __restrict__ offsets,
                                                                    const long long* __restrict__ s_offsets,
                                                                    const int* __restrict__ count, const int2* __restrict__ list,
                                                                    float* __restrict__ R_all) {
  constexpr int TR = kLrRows + 2 * H, LD = (kLrCols + 2 * H) | 1;
  __shared__ float tile[TR * LD];
  if (static_cast<int>(blockIdx.x) >= *count) return;
  const int2 e = list[blockIdx.x];
  const int doc = e.x;
  const int n = offsets[doc + 1] - offsets[doc];
  const int i_base = e.y * kLrRows, j_base = static_cast<int>(blockIdx.y) * kLrCols;
  if (j_base >= n) return;
  const float* S = S_all + s_offsets[doc];
  float* R = R_all + s_offsets[doc];
  for (int t = threadIdx.x; t < TR * (kLrCols + 2 * H); t += kLrThreads) {
    const int a = t / (kLrCols + 2 * H), b = t - a * (kLrCols + 2 * H);
    const int gi = i_base - H + a, gj = j_base - H + b;
    tile[a * LD + b] = (gi >= 0 && gi < n && gj >= 0 && gj < n) ? S[static_cast<size_t>(gi) * n + gj] : INFINITY;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = i_base + lane, c0 = j_base + w * kLrCpt;
  if (i >= n || c0 >= n) return;
  float ctr[kLrCpt];
  int cnt[kLrCpt];
#pragma unroll
  for (int m = 0; m < kLrCpt; ++m) {
    ctr[m] = tile[(lane + H) * LD + w * kLrCpt + m + H];
    cnt[m] = 0;
  }
#pragma unroll
  for (int a = 0; a <= 2 * H; ++a) {
    const float* row = tile + (lane + a) * LD + w * kLrCpt;
    float v[kLrCpt + 2 * H];
#pragma unroll
    for (int k = 0; k < kLrCpt + 2 * H; ++k) v[k] = row[k];
#pragma unroll
    for (int m = 0; m < kLrCpt; ++m)
#pragma unroll
      for (int b = 0; b <= 2 * H; ++b) cnt[m] += (v[m + b] < ctr[m]) ? 1 : 0;
  }
  const int rows_in = min(n, i + H + 1) - max(0, i - H);
#pragma unroll
  for (int m = 0; m < kLrCpt; ++m) {
    const int j = c0 + m;
    if (j < n) {
      const int denom = rows_in * (min(n, j + H + 1) - max(0, j - H));
      R[static_cast<size_t>(i) * n + j] = static_cast<float>(static_cast<double>(cnt[m]) / static_cast<double>(denom > 0 ? denom : 1));
    }
  }
}

template <int H>
static int launch_local_rank(const float* S, const int32_t* offsets, const long long* s_offsets, int n_docs, int total_rows,
                             int max_doc_rows, int32_t* workspace, float* out_R, cudaStream_t st) {
  int* count = workspace + total_rows;                       // after the row -> document map
  int2* list = reinterpret_cast<int2*>(workspace + ((static_cast<size_t>(total_rows) + 3) & ~static_cast<size_t>(1)));  // 8-byte aligned
  SS_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
  c99_tile_list_kernel<<<(n_docs + 255) / 256, 256, 0, st>>>(offsets, n_docs, count, list);
  SS_CUDA_CHECK(cudaGetLastError());
  const unsigned int max_tiles = static_cast<unsigned int>((total_rows + kLrRows - 1) / kLrRows + n_docs);
  dim3 grid(max_tiles, static_cast<unsigned int>((max_doc_rows + kLrCols - 1) / kLrCols));
  c99_local_rank_kernel<H><<<grid, kLrThreads, 0, st>>>(S, offsets, s_offsets, count, list, out_R);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}

__global__ void c99_row_doc_kernel(const int* __restrict__ offsets, int n_docs, int* __restrict__ row_doc) {
  const int doc = blockIdx.x;
  if (doc >= n_docs) return;
  for (int r = offsets[doc] + threadIdx.x; r < offsets[doc + 1]; r += blockDim.x) row_doc[r] = doc;
}

}  // namespace ss

using namespace ss;

extern "C" int ss_c99_rank_matrix(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, int total_rows,
                                  int max_doc_rows, int use_local_rank, int mask_size, int32_t* workspace_rows, float* out_R,
                                  void* stream) {
  if (!S || !offsets || !s_offsets || !workspace_rows || !out_R) return fail(SS_ERR_INVALID_ARG, "ss_c99_rank_matrix: null pointer");
  if (n_docs <= 0 || total_rows <= 0 || max_doc_rows <= 0) return fail(SS_ERR_INVALID_ARG, "ss_c99_rank_matrix: sizes must be positive");
  const size_t smem = static_cast<size_t>(max_doc_rows) * sizeof(float);
  if (smem + 1024 > smem_optin()) return fail(SS_ERR_UNSUPPORTED, "ss_c99_rank_matrix: document too long for the shared-memory row");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c99_row_doc_kernel<<<n_docs, 128, 0, st>>>(offsets, n_docs, workspace_rows);
  SS_CUDA_CHECK(cudaGetLastError());
  const bool symmetric = (use_local_rank & 2) != 0;  // bit 1: the caller guarantees S == S^T bit for bit (output of K3)
  use_local_rank &= 1;
  if (!use_local_rank && max_doc_rows <= kRankSortMaxRows && !getenv("SS_C99_RANK_COUNTING")) {
    const long long* so = reinterpret_cast<const long long*>(s_offsets);
    if (max_doc_rows <= 128) return launch_rank_sorted<4, 32>(S, offsets, so, workspace_rows, n_docs, total_rows, max_doc_rows, symmetric, out_R, st);
    if (max_doc_rows <= 512) return launch_rank_sorted<16, 32>(S, offsets, so, workspace_rows, n_docs, total_rows, max_doc_rows, symmetric, out_R, st);
    return launch_rank_sorted<64, 8>(S, offsets, so, workspace_rows, n_docs, total_rows, max_doc_rows, symmetric, out_R, st);
  }
  const int m = std::max(3, mask_size | 1);
  if (use_local_rank && !getenv("SS_C99_RANK_COUNTING")) {
    const long long* so = reinterpret_cast<const long long*>(s_offsets);
    switch (m / 2) {  // the tiled kernel is compiled for the common window sizes (11 = the reference's default)
      case 1: return launch_local_rank<1>(S, offsets, so, n_docs, total_rows, max_doc_rows, workspace_rows, out_R, st);
      case 2: return launch_local_rank<2>(S, offsets, so, n_docs, total_rows, max_doc_rows, workspace_rows, out_R, st);
      case 3: return launch_local_rank<3>(S, offsets, so, n_docs, total_rows, max_doc_rows, workspace_rows, out_R, st);
      case 4: return launch_local_rank<4>(S, offsets, so, n_docs, total_rows, max_doc_rows, workspace_rows, out_R, st);
      case 5: return launch_local_rank<5>(S, offsets, so, n_docs, total_rows, max_doc_rows, workspace_rows, out_R, st);
      case 6: return launch_local_rank<6>(S, offsets, so, n_docs, total_rows, max_doc_rows, workspace_rows, out_R, st);
      case 7: return launch_local_rank<7>(S, offsets, so, n_docs, total_rows, max_doc_rows, workspace_rows, out_R, st);
      default: break;  // larger windows: the counting kernel below
    }
  }
  if (smem > 48 * 1024)
    SS_CUDA_CHECK(cudaFuncSetAttribute(c99_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  c99_rank_kernel<<<total_rows, 256, smem, st>>>(S, offsets, reinterpret_cast<const long long*>(s_offsets), workspace_rows,
                                                use_local_rank ? 1 : 0, m / 2, out_R);
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
