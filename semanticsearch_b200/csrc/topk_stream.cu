// K1 — streaming cosine similarity with a fused per-warp top-k epilogue (memory-bound path).
//
// Replaces sklearn.cosine_similarity(q, C)[0] + np.argsort(-s) at
// Tool/rank_chunks_optimized.py:215-216,225-235 for small query batches.
//
// Data path (per CTA, one CTA per SM, persistent over corpus tiles):
//   HBM --cp.async.bulk (TMA engine, 1-D, contiguous rows)--> smem ring (S stages, mbarrier
//   full/empty) --LDS.128--> 15 consumer warps, one corpus row per warp at a time:
//   dot(q_b, c) for up to BT queries and sum(c^2) in one pass, butterfly reduce, score =
//   dot * rsqrt(sum c^2) (queries are pre-normalised in fp32), compare against the warp's
//   current k-th best, rare insert into a per-warp list in smem.  At the end the CTA merges its
//   warps' lists and writes k keys per query; ss_topk_merge() folds the per-CTA lists.
// The score matrix never exists in memory; the corpus is read exactly once per query group.
#include <algorithm>
#include <cstring>

#include "ss_common.cuh"

namespace ss {

constexpr int kStreamThreads = 512;                      // 1 producer warp + 15 consumer warps
constexpr int kConsumerWarps = kStreamThreads / 32 - 1;  // 15
constexpr int kMaxStages = 8;
constexpr int kMaxQReg = 6;  // B=1 fast path keeps the query in registers for <= 6 chunks/lane

struct StreamParams {
  const void* corpus;
  long long n_rows;
  int dim;
  const float* qn;  // [n_queries][dim] fp32, L2-normalised
  int n_queries;
  int k;
  int kpad;  // k rounded up to 32
  uint32_t index_base;
  int tile_rows;
  long long n_tiles;
  int stages;
  uint32_t tile_bytes;  // smem bytes reserved per stage (multiple of 128)
  uint64_t* partial;    // [n_queries][gridDim.x][k]
};

// ------------------------------------------------------------------------------------------
// Query preparation: Qn = q / max-norm rule (sklearn: zero norm -> divide by 1)
// ------------------------------------------------------------------------------------------
template <typename TQ>
__global__ void prep_queries_kernel(const TQ* __restrict__ q, int n_queries, int dim, float* __restrict__ qn) {
  const int b = blockIdx.x;
  if (b >= n_queries) return;
  const TQ* row = q + static_cast<size_t>(b) * dim;
  float ss = 0.f;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    const float v = to_f32<TQ>(row[c]);
    ss = fmaf(v, v, ss);
  }
  __shared__ float red[32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  float norm = sqrtf(red[0]);
  if (norm == 0.f) norm = 1.f;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) qn[static_cast<size_t>(b) * dim + c] = to_f32<TQ>(row[c]) / norm;
}

// ------------------------------------------------------------------------------------------
// Per-warp top-k list in shared memory (unsorted, threshold = current minimum)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t list_min(const uint64_t* lst, int kpad, int lane) {
  uint64_t m = ~0ull;
  for (int j = lane; j < kpad; j += 32) {
    const uint64_t v = lst[j];
    m = v < m ? v : m;
  }
  return warp_min_u64(m);
}

// Warp-uniform call: replace the list minimum with `key`, return the new minimum.
__device__ __noinline__ uint64_t list_insert(uint64_t* lst, int kpad, uint64_t key, int lane) {
  uint64_t m = ~0ull;
  int pos = 0x7fffffff;
  for (int j = lane; j < kpad; j += 32) {
    const uint64_t v = lst[j];
    if (v < m) {
      m = v;
      pos = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t om = __shfl_xor_sync(0xffffffffu, m, o);
    const int op = __shfl_xor_sync(0xffffffffu, pos, o);
    if (om < m || (om == m && op < pos)) {
      m = om;
      pos = op;
    }
  }
  if (lane == (pos & 31)) lst[pos] = key;
  __syncwarp();
  return list_min(lst, kpad, lane);
}

// ------------------------------------------------------------------------------------------
// Main kernel
// ------------------------------------------------------------------------------------------
template <typename T, int BT, bool QREG>
__global__ void __launch_bounds__(kStreamThreads, 1) cosine_topk_stream_kernel(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int EPC = Chunk<T>::EPC;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.y * BT;                        // first query of this group
  const int nq = min(BT, p.n_queries - q0);              // live queries in this group
  const uint32_t row_bytes = static_cast<uint32_t>(p.dim) * sizeof(T);
  const int chunks = row_bytes / 16;

  // ---- shared-memory carve-up -------------------------------------------------------------
  unsigned char* tiles = smem_raw;                                                 // stages * tile_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(p.stages) * p.tile_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  float* qs = reinterpret_cast<float*>(bars + 2 * kMaxStages);                     // [BT][dim] fp32
  uint64_t* lists = reinterpret_cast<uint64_t*>(qs + static_cast<size_t>(BT) * p.dim + ((BT * p.dim) & 1));
  // lists: [kConsumerWarps][BT][kpad]

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < BT * p.dim; i += blockDim.x) {
    const int b = i / p.dim;
    qs[i] = (b < nq) ? p.qn[static_cast<size_t>(q0 + b) * p.dim + (i - b * p.dim)] : 0.f;
  }
  for (int i = threadIdx.x; i < kConsumerWarps * BT * p.kpad; i += blockDim.x) {
    const int j = i % p.kpad;
    lists[i] = (j < p.k) ? 0ull : ~0ull;  // 0 = empty slot, ~0 = padding that is never replaced
  }
  __syncthreads();

  if (warp == 0) {
    // ================= producer: one lane drives the bulk-copy ring =================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        mbar_wait(&empty_bar[s], phase ^ 1u);
        const long long row0 = t * p.tile_rows;
        const long long rows = min(static_cast<long long>(p.tile_rows), p.n_rows - row0);
        const uint32_t bytes = static_cast<uint32_t>(rows) * row_bytes;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_copy_g2s(tiles + static_cast<size_t>(s) * p.tile_bytes,
                      static_cast<const unsigned char*>(p.corpus) + static_cast<size_t>(row0) * row_bytes, bytes,
                      &full_bar[s]);
        if (++s == p.stages) {
          s = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ================= consumers: one corpus row per warp at a time =================
    const int cw = warp - 1;
    uint64_t* my_lists = lists + static_cast<size_t>(cw) * BT * p.kpad;
    uint64_t thr[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) thr[b] = 0ull;

    // B=1 fast path: this lane's slice of the query lives in registers.
    float qreg[QREG ? kMaxQReg : 1][EPC];
    if (QREG) {
#pragma unroll
      for (int j = 0; j < kMaxQReg; ++j) {
        const int c = lane + 32 * j;
#pragma unroll
        for (int e = 0; e < EPC; ++e) qreg[j][e] = (c < chunks) ? qs[c * EPC + e] : 0.f;
      }
    }

    int s = 0;
    uint32_t phase = 0;
    for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      mbar_wait(&full_bar[s], phase);
      const long long row0 = t * p.tile_rows;
      const int rows = static_cast<int>(min(static_cast<long long>(p.tile_rows), p.n_rows - row0));
      const unsigned char* tile = tiles + static_cast<size_t>(s) * p.tile_bytes;
      for (int r = cw; r < rows; r += kConsumerWarps) {
        const uint4* row = reinterpret_cast<const uint4*>(tile + static_cast<size_t>(r) * row_bytes);
        float dot[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) dot[b] = 0.f;
        float ssq = 0.f;
        if (QREG) {
#pragma unroll
          for (int j = 0; j < kMaxQReg; ++j) {
            const int c = lane + 32 * j;
            if (c < chunks) {
              float x[EPC];
              Chunk<T>::unpack(row[c], x);
#pragma unroll
              for (int e = 0; e < EPC; ++e) {
                ssq = fmaf(x[e], x[e], ssq);
                dot[0] = fmaf(x[e], qreg[j][e], dot[0]);
              }
            }
          }
        } else {
          for (int c = lane; c < chunks; c += 32) {
            float x[EPC];
            Chunk<T>::unpack(row[c], x);
#pragma unroll
            for (int e = 0; e < EPC; ++e) ssq = fmaf(x[e], x[e], ssq);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
              const float4* qv = reinterpret_cast<const float4*>(qs + static_cast<size_t>(b) * p.dim + c * EPC);
#pragma unroll
              for (int h = 0; h < EPC / 4; ++h) {
                const float4 qq = qv[h];
                dot[b] = fmaf(x[4 * h + 0], qq.x, dot[b]);
                dot[b] = fmaf(x[4 * h + 1], qq.y, dot[b]);
                dot[b] = fmaf(x[4 * h + 2], qq.z, dot[b]);
                dot[b] = fmaf(x[4 * h + 3], qq.w, dot[b]);
              }
            }
          }
        }
        ssq = warp_sum(ssq);
        // sklearn's zero rule: a zero row is divided by 1 and scores 0.
        const float inv = ssq > 0.f ? 1.0f / sqrtf(ssq) : 1.0f;
        const uint32_t grow = p.index_base + static_cast<uint32_t>(row0 + r);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float d = warp_sum(dot[b]);
          const uint64_t key = make_key(d * inv, grow);
          if (b < nq && key > thr[b]) thr[b] = list_insert(my_lists + b * p.kpad, p.kpad, key, lane);
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
  __syncthreads();

  // ---- CTA merge: rank every candidate by counting; rank < k goes to its sorted slot ----------
  const int M = kConsumerWarps * p.kpad;  // per query: candidate slots incl. padding
  for (int b = 0; b < nq; ++b) {
    uint64_t* out = p.partial + (static_cast<size_t>(q0 + b) * gridDim.x + blockIdx.x) * p.k;
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) out[j] = 0ull;
  }
  __syncthreads();
  for (int b = 0; b < nq; ++b) {
    uint64_t* out = p.partial + (static_cast<size_t>(q0 + b) * gridDim.x + blockIdx.x) * p.k;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      const int w = i / p.kpad, j = i - w * p.kpad;
      if (j >= p.k) continue;
      const uint64_t key = lists[(static_cast<size_t>(w) * BT + b) * p.kpad + j];
      if (key == 0ull) continue;
      int rank = 0;
      for (int w2 = 0; w2 < kConsumerWarps && rank < p.k; ++w2) {
        const uint64_t* l2 = lists + (static_cast<size_t>(w2) * BT + b) * p.kpad;
        for (int j2 = 0; j2 < p.k; ++j2) rank += (l2[j2] > key) ? 1 : 0;
      }
      if (rank < p.k) out[rank] = key;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Generic fallback for rows that are not 16-byte multiples / unaligned: plain loads, same epilogue
// ------------------------------------------------------------------------------------------
template <typename T, int BT>
__global__ void __launch_bounds__(kStreamThreads, 1) cosine_topk_generic_kernel(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int q0 = blockIdx.y * BT;
  const int nq = min(BT, p.n_queries - q0);
  float* qs = reinterpret_cast<float*>(smem_raw);
  uint64_t* lists = reinterpret_cast<uint64_t*>(qs + static_cast<size_t>(BT) * p.dim + ((BT * p.dim) & 1));
  for (int i = threadIdx.x; i < BT * p.dim; i += blockDim.x) {
    const int b = i / p.dim;
    qs[i] = (b < nq) ? p.qn[static_cast<size_t>(q0 + b) * p.dim + (i - b * p.dim)] : 0.f;
  }
  for (int i = threadIdx.x; i < nwarps * BT * p.kpad; i += blockDim.x) lists[i] = ((i % p.kpad) < p.k) ? 0ull : ~0ull;
  __syncthreads();
  uint64_t* my_lists = lists + static_cast<size_t>(warp) * BT * p.kpad;
  uint64_t thr[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) thr[b] = 0ull;
  const T* corpus = static_cast<const T*>(p.corpus);
  for (long long r = static_cast<long long>(blockIdx.x) * nwarps + warp; r < p.n_rows;
       r += static_cast<long long>(gridDim.x) * nwarps) {
    const T* row = corpus + static_cast<size_t>(r) * p.dim;
    float dot[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) dot[b] = 0.f;
    float ssq = 0.f;
    for (int c = lane; c < p.dim; c += 32) {
      const float x = to_f32<T>(row[c]);
      ssq = fmaf(x, x, ssq);
#pragma unroll
      for (int b = 0; b < BT; ++b) dot[b] = fmaf(x, qs[b * p.dim + c], dot[b]);
    }
    ssq = warp_sum(ssq);
    const float inv = ssq > 0.f ? 1.0f / sqrtf(ssq) : 1.0f;
    const uint32_t grow = p.index_base + static_cast<uint32_t>(r);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const float d = warp_sum(dot[b]);
      const uint64_t key = make_key(d * inv, grow);
      if (b < nq && key > thr[b]) thr[b] = list_insert(my_lists + b * p.kpad, p.kpad, key, lane);
    }
  }
  __syncthreads();
  const int M = nwarps * p.kpad;
  for (int b = 0; b < nq; ++b) {
    uint64_t* out = p.partial + (static_cast<size_t>(q0 + b) * gridDim.x + blockIdx.x) * p.k;
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) out[j] = 0ull;
  }
  __syncthreads();
  for (int b = 0; b < nq; ++b) {
    uint64_t* out = p.partial + (static_cast<size_t>(q0 + b) * gridDim.x + blockIdx.x) * p.k;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      const int w = i / p.kpad, j = i - w * p.kpad;
      if (j >= p.k) continue;
      const uint64_t key = lists[(static_cast<size_t>(w) * BT + b) * p.kpad + j];
      if (key == 0ull) continue;
      int rank = 0;
      for (int w2 = 0; w2 < nwarps && rank < p.k; ++w2) {
        const uint64_t* l2 = lists + (static_cast<size_t>(w2) * BT + b) * p.kpad;
        for (int j2 = 0; j2 < p.k; ++j2) rank += (l2[j2] > key) ? 1 : 0;
      }
      if (rank < p.k) out[rank] = key;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Host side: configuration + launch
// ------------------------------------------------------------------------------------------
struct StreamConfig {
  bool bulk;  // TMA-ring kernel usable (16-byte rows, aligned base)
  int bt;
  int stages;
  int tile_rows;
  uint32_t tile_bytes;
  long long n_tiles;
  int grid_x, grid_y;
  int threads;
  size_t smem;
  int kpad;
};

static size_t list_bytes(int warps, int bt, int kpad) { return static_cast<size_t>(warps) * bt * kpad * 8; }

static bool make_config(const void* corpus, long long n_rows, int dim, int dtype, int n_queries, int k,
                        StreamConfig* cfg) {
  const size_t es = dtype_size(dtype);
  const size_t row_bytes = static_cast<size_t>(dim) * es;
  const size_t smem_cap = smem_optin();
  const int sms = sm_count();
  cfg->kpad = static_cast<int>(align_up(k, 32));
  cfg->bulk = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(corpus) % 16 == 0) && row_bytes <= 48 * 1024;
  const int bt_opts[4] = {8, 4, 2, 1};
  if (cfg->bulk) {
    const int tile_rows = static_cast<int>(std::max<size_t>(1, (32 * 1024) / row_bytes));
    const uint32_t tile_bytes = static_cast<uint32_t>(align_up(tile_rows * row_bytes, 128));
    for (int bi = 0; bi < 4; ++bi) {
      const int bt = bt_opts[bi];
      if (bt > 1 && bt / 2 >= n_queries) continue;  // do not over-provision query slots
      const size_t fixed = 2 * kMaxStages * 8 + align_up(static_cast<size_t>(bt) * dim * 4, 8) +
                           list_bytes(kConsumerWarps, bt, cfg->kpad) + 128;
      if (fixed + 2 * static_cast<size_t>(tile_bytes) > smem_cap) continue;
      int stages = static_cast<int>((smem_cap - fixed) / tile_bytes);
      stages = std::min(stages, 6);
      if (stages < 2) continue;
      cfg->bt = bt;
      cfg->stages = stages;
      cfg->tile_rows = tile_rows;
      cfg->tile_bytes = tile_bytes;
      cfg->n_tiles = (n_rows + tile_rows - 1) / tile_rows;
      cfg->grid_y = (n_queries + bt - 1) / bt;
      cfg->grid_x = static_cast<int>(std::max<long long>(1, std::min<long long>(sms, cfg->n_tiles)));
      cfg->threads = kStreamThreads;
      cfg->smem = fixed + static_cast<size_t>(stages) * tile_bytes;
      return true;
    }
    cfg->bulk = false;
  }
  for (int bi = 0; bi < 4; ++bi) {
    const int bt = bt_opts[bi];
    if (bt > 1 && bt / 2 >= n_queries) continue;
    const size_t need = align_up(static_cast<size_t>(bt) * dim * 4, 8) + list_bytes(kStreamThreads / 32, bt, cfg->kpad) + 128;
    if (need > smem_cap) continue;
    cfg->bt = bt;
    cfg->stages = 0;
    cfg->tile_rows = 0;
    cfg->tile_bytes = 0;
    cfg->n_tiles = 0;
    cfg->grid_y = (n_queries + bt - 1) / bt;
    const long long want = (n_rows + (kStreamThreads / 32) - 1) / (kStreamThreads / 32);
    cfg->grid_x = static_cast<int>(std::max<long long>(1, std::min<long long>(sms, want)));
    cfg->threads = kStreamThreads;
    cfg->smem = need;
    return true;
  }
  return false;
}

template <typename T, int BT>
static cudaError_t launch_stream(const StreamConfig& cfg, const StreamParams& p, int chunks, cudaStream_t st) {
  const dim3 grid(cfg.grid_x, cfg.grid_y);
  if (cfg.bulk) {
    if (BT == 1 && chunks <= 32 * kMaxQReg) {
      auto kern = cosine_topk_stream_kernel<T, 1, true>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cfg.smem));
      if (e != cudaSuccess) return e;
      kern<<<grid, cfg.threads, cfg.smem, st>>>(p);
    } else {
      auto kern = cosine_topk_stream_kernel<T, BT, false>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cfg.smem));
      if (e != cudaSuccess) return e;
      kern<<<grid, cfg.threads, cfg.smem, st>>>(p);
    }
  } else {
    auto kern = cosine_topk_generic_kernel<T, BT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cfg.smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, cfg.threads, cfg.smem, st>>>(p);
  }
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_stream_bt(const StreamConfig& cfg, const StreamParams& p, int chunks, cudaStream_t st) {
  switch (cfg.bt) {
    case 8: return launch_stream<T, 8>(cfg, p, chunks, st);
    case 4: return launch_stream<T, 4>(cfg, p, chunks, st);
    case 2: return launch_stream<T, 2>(cfg, p, chunks, st);
    default: return launch_stream<T, 1>(cfg, p, chunks, st);
  }
}

}  // namespace ss

using namespace ss;

extern "C" size_t ss_cosine_topk_stream_workspace_bytes(int64_t n_rows, int dim, int corpus_dtype, int n_queries, int k) {
  (void)n_rows;
  (void)corpus_dtype;
  if (dim <= 0 || n_queries <= 0 || k <= 0) return 0;
  const size_t qn = align_up(static_cast<size_t>(n_queries) * dim * 4, 256);
  const size_t partial = align_up(static_cast<size_t>(n_queries) * sm_count() * k * 8, 256);
  return qn + partial + 256;
}

extern "C" int ss_cosine_topk_stream(const void* corpus, int64_t n_rows, int dim, int corpus_dtype, const void* queries,
                                     int n_queries, int query_dtype, int k, uint32_t index_base, void* workspace,
                                     size_t workspace_bytes, uint64_t* out_keys, float* out_scores, int64_t* out_indices,
                                     void* stream) {
  if (!corpus || !queries || !workspace) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_stream: null pointer");
  if (n_rows <= 0 || dim <= 0 || n_queries <= 0 || k <= 0)
    return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_stream: n_rows, dim, n_queries and k must be positive");
  if (!dtype_ok(corpus_dtype) || !dtype_ok(query_dtype)) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_stream: bad dtype");
  if (k > 1024) return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_stream: k > 1024 is not supported");
  if (static_cast<uint64_t>(index_base) + static_cast<uint64_t>(n_rows) > 0xFFFFFFFFull)
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_stream: global row index must fit in 32 bits");
  const size_t need = ss_cosine_topk_stream_workspace_bytes(n_rows, dim, corpus_dtype, n_queries, k);
  if (workspace_bytes < need) return fail(SS_ERR_WORKSPACE, "ss_cosine_topk_stream: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  unsigned char* ws = static_cast<unsigned char*>(workspace);
  ws = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));
  float* qn = reinterpret_cast<float*>(ws);
  uint64_t* partial = reinterpret_cast<uint64_t*>(ws + align_up(static_cast<size_t>(n_queries) * dim * 4, 256));

  switch (query_dtype) {
    case SS_F32: prep_queries_kernel<float><<<n_queries, 256, 0, st>>>(static_cast<const float*>(queries), n_queries, dim, qn); break;
    case SS_BF16: prep_queries_kernel<__nv_bfloat16><<<n_queries, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(queries), n_queries, dim, qn); break;
    default: prep_queries_kernel<__half><<<n_queries, 256, 0, st>>>(static_cast<const __half*>(queries), n_queries, dim, qn); break;
  }
  SS_CUDA_CHECK(cudaGetLastError());

  StreamConfig cfg;
  if (!make_config(corpus, n_rows, dim, corpus_dtype, n_queries, k, &cfg))
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_stream: dim / k combination does not fit in shared memory");
  StreamParams p;
  p.corpus = corpus;
  p.n_rows = n_rows;
  p.dim = dim;
  p.qn = qn;
  p.n_queries = n_queries;
  p.k = k;
  p.kpad = cfg.kpad;
  p.index_base = index_base;
  p.tile_rows = cfg.tile_rows;
  p.n_tiles = cfg.n_tiles;
  p.stages = cfg.stages;
  p.tile_bytes = cfg.tile_bytes;
  p.partial = partial;
  const int chunks = static_cast<int>(static_cast<size_t>(dim) * dtype_size(corpus_dtype) / 16);
  cudaError_t e;
  {
    ProfileScope prof(st);
    switch (corpus_dtype) {
      case SS_F32: e = launch_stream_bt<float>(cfg, p, chunks, st); break;
      case SS_BF16: e = launch_stream_bt<__nv_bfloat16>(cfg, p, chunks, st); break;
      default: e = launch_stream_bt<__half>(cfg, p, chunks, st); break;
    }
  }
  if (e != cudaSuccess) return cuda_fail(e, "cosine_topk_stream launch");
  return ss_topk_merge(partial, cfg.grid_x, n_queries, k, static_cast<int64_t>(cfg.grid_x) * k, k, k, out_keys, out_scores,
                       out_indices, stream);
}
