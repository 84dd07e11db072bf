// K3 — segmented (ragged) sentence x sentence similarity matrices, S_d = En_d . En_d^T.
//
// Replaces create_similarity_matrix's arithmetic (Method/semantic_common.py:158-164,186-191), the
// C99 similarity at Method/Semantic_Splitter_Optimized.py:169 and the controller's diagnostic
// recomputations (data_process/simple_chunk_controller.py:614,682,743) for a whole batch of
// documents in one launch: no per-document H2D/D2H, no per-document cuBLAS launch.
//
// Work decomposition: every document contributes T(T+1)/2 upper-triangular 64x64 output tiles
// (T = ceil(n/64)); one CTA per tile, located by binary search in a prefix array.  Each CTA runs
// an fp32 FFMA micro-kernel (4x4 outputs per thread, K chunks of 16 through double-buffered
// shared memory), accumulates the row norms of the rows it streams, scales the tile by
// 1/(|e_i| |e_j|) and writes it — and its transpose for off-diagonal tiles — with coalesced rows.
// fp32 end to end: the 1e-5 parity tolerance rules out plain TF32/BF16 tensor-core products.
#include <algorithm>

#include "ss_common.cuh"

namespace ss {

constexpr int kSimTile = 64;
constexpr int kSimBK = 16;
constexpr int kSimPad = 4;

struct SimParams {
  const float* rows;
  int dim;
  const int* offsets;         // [n_docs + 1] row offsets
  const long long* s_offsets; // [n_docs + 1] element offsets into out (prefix sums of n^2)
  const int* tile_prefix;     // [n_docs + 1] prefix sums of T(T+1)/2
  int n_docs;
  float* out;
};

__device__ __forceinline__ float4 load_row4(const float* base, int dim, bool row_ok, int k, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok) return v;
  if (vec) {
    if (k < dim) v = *reinterpret_cast<const float4*>(base + k);  // dim % 4 == 0: whole vector in range
  } else {
    if (k + 0 < dim) v.x = base[k + 0];
    if (k + 1 < dim) v.y = base[k + 1];
    if (k + 2 < dim) v.z = base[k + 2];
    if (k + 3 < dim) v.w = base[k + 3];
  }
  return v;
}

__global__ void __launch_bounds__(256) segmented_simmatrix_kernel(const SimParams p) {
  __shared__ __align__(16) float As[2][kSimBK][kSimTile + kSimPad];
  __shared__ __align__(16) float Bs[2][kSimBK][kSimTile + kSimPad];
  __shared__ float Cs[kSimTile][kSimTile + 1];
  __shared__ float inv_a[kSimTile], inv_b[kSimTile];

  // ---- locate (document, tile row, tile column) ----------------------------------------------
  const int tile = blockIdx.x;
  int lo = 0, hi = p.n_docs;  // last doc with tile_prefix[doc] <= tile
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.tile_prefix[mid] <= tile) lo = mid; else hi = mid;
  }
  const int doc = lo;
  const int row_base = p.offsets[doc];
  const int n = p.offsets[doc + 1] - row_base;
  const int T = (n + kSimTile - 1) / kSimTile;
  int l = tile - p.tile_prefix[doc];
  int ti = 0;
  while (l >= T - ti) {
    l -= T - ti;
    ++ti;
  }
  const int tj = ti + l;
  if (n < 1) return;

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int lrow = tid >> 2;           // tile row this thread loads
  const int lk = (tid & 3) * 4;        // k offset inside the 16-wide chunk
  const bool vec = (p.dim & 3) == 0;
  const int ra = ti * kSimTile + lrow, rb = tj * kSimTile + lrow;
  const bool ra_ok = ra < n, rb_ok = rb < n;
  const float* pa = p.rows + static_cast<size_t>(row_base + (ra_ok ? ra : 0)) * p.dim;
  const float* pb = p.rows + static_cast<size_t>(row_base + (rb_ok ? rb : 0)) * p.dim;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ssq_a = 0.f, ssq_b = 0.f;

  const int nk = (p.dim + kSimBK - 1) / kSimBK;
  float4 va = load_row4(pa, p.dim, ra_ok, lk, vec);
  float4 vb = load_row4(pb, p.dim, rb_ok, lk, vec);
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    As[buf][lk + 0][lrow] = va.x; As[buf][lk + 1][lrow] = va.y; As[buf][lk + 2][lrow] = va.z; As[buf][lk + 3][lrow] = va.w;
    Bs[buf][lk + 0][lrow] = vb.x; Bs[buf][lk + 1][lrow] = vb.y; Bs[buf][lk + 2][lrow] = vb.z; Bs[buf][lk + 3][lrow] = vb.w;
    ssq_a = fmaf(va.x, va.x, fmaf(va.y, va.y, fmaf(va.z, va.z, fmaf(va.w, va.w, ssq_a))));
    ssq_b = fmaf(vb.x, vb.x, fmaf(vb.y, vb.y, fmaf(vb.z, vb.z, fmaf(vb.w, vb.w, ssq_b))));
    __syncthreads();  // buffer `buf` is full; the other buffer is free (its readers passed the previous barrier)
    if (kc + 1 < nk) {
      va = load_row4(pa, p.dim, ra_ok, (kc + 1) * kSimBK + lk, vec);
      vb = load_row4(pb, p.dim, rb_ok, (kc + 1) * kSimBK + lk, vec);
    }
#pragma unroll
    for (int kk = 0; kk < kSimBK; ++kk) {
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
  // ---- row norms: the 4 threads that loaded one row are adjacent lanes ---------------------------
  ssq_a += __shfl_xor_sync(0xffffffffu, ssq_a, 1);
  ssq_a += __shfl_xor_sync(0xffffffffu, ssq_a, 2);
  ssq_b += __shfl_xor_sync(0xffffffffu, ssq_b, 1);
  ssq_b += __shfl_xor_sync(0xffffffffu, ssq_b, 2);
  if ((tid & 3) == 0) {
    // zero rows stay zero (reference: norm 0 -> 1e-9, and 0 / 1e-9 == 0)
    inv_a[lrow] = ssq_a > 0.f ? 1.0f / sqrtf(ssq_a) : 0.f;
    inv_b[lrow] = ssq_b > 0.f ? 1.0f / sqrtf(ssq_b) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j] * (inv_a[ty * 4 + i] * inv_b[tx * 4 + j]);  // symmetric in (i, j)
  __syncthreads();

  float* S = p.out + p.s_offsets[doc];
  const int c_lane = tid & 63, r_grp = tid >> 6;
  // direct tile: rows ti*64.., columns tj*64..
  for (int r = r_grp; r < kSimTile; r += 4) {
    const int gr = ti * kSimTile + r, gc = tj * kSimTile + c_lane;
    if (gr < n && gc < n) S[static_cast<size_t>(gr) * n + gc] = Cs[r][c_lane];
  }
  if (ti != tj) {  // mirrored tile: S[c][r] = S[r][c]
    for (int c = r_grp; c < kSimTile; c += 4) {
      const int gr = tj * kSimTile + c, gc = ti * kSimTile + c_lane;
      if (gr < n && gc < n) S[static_cast<size_t>(gr) * n + gc] = Cs[c_lane][c];
    }
  }
}

}  // namespace ss

using namespace ss;

extern "C" int ss_segmented_plan_host(const int32_t* offsets_host, int n_docs, int64_t* s_offsets_host,
                                      int32_t* tile_prefix_host, int64_t* total_tiles, int32_t* max_doc_rows) {
  if (!offsets_host || n_docs < 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_plan_host: bad arguments");
  int64_t s = 0, t = 0;
  int32_t mx = 0;
  for (int d = 0; d < n_docs; ++d) {
    const int64_t n = static_cast<int64_t>(offsets_host[d + 1]) - offsets_host[d];
    if (n < 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_plan_host: offsets must be non-decreasing");
    if (s_offsets_host) s_offsets_host[d] = s;
    if (tile_prefix_host) tile_prefix_host[d] = static_cast<int32_t>(t);
    const int64_t T = (n + kSimTile - 1) / kSimTile;
    s += n * n;
    t += T * (T + 1) / 2;
    if (t > 0x7FFFFFFF) return fail(SS_ERR_UNSUPPORTED, "ss_segmented_plan_host: more than 2^31 tiles in one batch");
    mx = std::max<int32_t>(mx, static_cast<int32_t>(n));
  }
  if (s_offsets_host) s_offsets_host[n_docs] = s;
  if (tile_prefix_host) tile_prefix_host[n_docs] = static_cast<int32_t>(t);
  if (total_tiles) *total_tiles = t;
  if (max_doc_rows) *max_doc_rows = mx;
  return SS_OK;
}

extern "C" int ss_segmented_simmatrix(const float* rows, int dim, const int32_t* offsets, const int64_t* s_offsets,
                                      const int32_t* tile_prefix, int n_docs, int64_t total_tiles, float* out_S, void* stream) {
  if (!rows || !offsets || !s_offsets || !tile_prefix || !out_S) return fail(SS_ERR_INVALID_ARG, "ss_segmented_simmatrix: null pointer");
  if (dim <= 0 || n_docs <= 0 || total_tiles < 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_simmatrix: bad sizes");
  if (total_tiles == 0) return SS_OK;
  if (total_tiles > 0x7FFFFFFF) return fail(SS_ERR_UNSUPPORTED, "ss_segmented_simmatrix: too many tiles");
  if ((dim & 3) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) != 0)
    return fail(SS_ERR_INVALID_ARG, "ss_segmented_simmatrix: rows must be 16-byte aligned");
  SimParams p;
  p.rows = rows;
  p.dim = dim;
  p.offsets = offsets;
  p.s_offsets = reinterpret_cast<const long long*>(s_offsets);
  p.tile_prefix = tile_prefix;
  p.n_docs = n_docs;
  p.out = out_S;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfileScope prof(st);
    segmented_simmatrix_kernel<<<static_cast<unsigned int>(total_tiles), 256, 0, st>>>(p);
  }
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}
