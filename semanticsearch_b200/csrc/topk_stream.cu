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
#include <cstdlib>
#include <cstring>

#include "ss_common.cuh"
#include "topk_merge.cuh"

namespace ss {

constexpr int kConsumerWarps = 16;                        // warps 0..15 consume, warp 16 produces
constexpr int kStreamThreads = (kConsumerWarps + 1) * 32;  // 544
constexpr int kGenericThreads = 512;
constexpr int kMaxStages = 8;
constexpr int kMaxQReg = 6;  // B=1 fast path keeps the query in registers for <= 6 chunks/lane

struct StreamParams {
  const void* corpus;
  long long n_rows;
  int dim;
  const void* queries;  // [n_queries][dim], raw (un-normalised)
  int query_dtype;
  int n_queries;
  int k;
  int kpad;  // k rounded up to 32
  uint32_t index_base;
  int tile_rows;
  long long n_tiles;
  int stages;
  uint32_t tile_bytes;  // smem bytes reserved per stage (multiple of 128)
  uint64_t* partial;    // [n_queries][gridDim.x][k]
  unsigned int* tickets;  // [gridDim.y], zeroed before launch: last CTA of a query group merges
  uint64_t* out_keys;     // final outputs, each [n_queries][k] or nullptr
  float* out_scores;
  long long* out_indices;
  float* all_scores;      // optional [n_queries][n_rows] full score matrix (ranker drop-in), else nullptr
};

// ------------------------------------------------------------------------------------------
// Query preparation (fused prologue): Qn = q / ||q|| in fp32, zero norm -> divide by 1 (sklearn)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_query_elem(const void* q, int dtype, size_t i) {
  if (dtype == SS_F32) return static_cast<const float*>(q)[i];
  if (dtype == SS_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(q)[i]);
  return __half2float(static_cast<const __half*>(q)[i]);
}

// All threads call.  qs: shared [BT][dim]; qnorm: shared [BT].
template <int BT>
__device__ __forceinline__ void prepare_queries(const StreamParams& p, int q0, int nq, float* qs, float* qnorm) {
  for (int i = threadIdx.x; i < BT * p.dim; i += blockDim.x) {
    const int b = i / p.dim;
    qs[i] = (b < nq) ? load_query_elem(p.queries, p.query_dtype, static_cast<size_t>(q0 + b) * p.dim + (i - b * p.dim)) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < BT) {
    float ss = 0.f;
    for (int c = lane; c < p.dim; c += 32) {
      const float v = qs[warp * p.dim + c];
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    float norm = sqrtf(ss);
    if (norm == 0.f) norm = 1.f;
    if (lane == 0) qnorm[warp] = norm;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BT * p.dim; i += blockDim.x) qs[i] = qs[i] / qnorm[i / p.dim];
  __syncthreads();
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
// CTA epilogue shared by both kernels: fold the warps' lists into one sorted list per query,
// publish it, and let the last CTA of the query group produce the final top-k.
// ------------------------------------------------------------------------------------------
template <int BT>
__device__ __forceinline__ void cta_epilogue(const StreamParams& p, const uint64_t* lists, int nwarps, int q0, int nq,
                                             uint64_t* stage_area, size_t stage_bytes) {
  __shared__ uint64_t scratch[2];
  __shared__ unsigned int s_ticket;
  const int M = nwarps * p.kpad;
  for (int b = 0; b < nq; ++b) {
    uint64_t* out = p.partial + (static_cast<size_t>(q0 + b) * gridDim.x + blockIdx.x) * p.k;
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) out[j] = 0ull;
  }
  __syncthreads();
  // rank every candidate by counting better keys; rank < k goes straight to its sorted slot
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
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.tickets[blockIdx.y], 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();  // acquire: every other CTA's list is visible now
  const int P = gridDim.x;
  const bool stage = stage_area != nullptr && (static_cast<size_t>(P) * p.k + kMergeSurvivorCap) * 8 <= stage_bytes;
  uint64_t* surv = stage ? stage_area + static_cast<size_t>(P) * p.k : nullptr;
  for (int b = 0; b < nq; ++b) {
    const uint64_t* src = p.partial + static_cast<size_t>(q0 + b) * P * p.k;
    if (stage) {
      for (int c = threadIdx.x; c < P * p.k; c += blockDim.x) stage_area[c] = __ldcg(src + c);
      __syncthreads();
      src = stage_area;
    }
    MergeOut out;
    const size_t o = static_cast<size_t>(q0 + b) * p.k;
    out.keys = p.out_keys ? p.out_keys + o : nullptr;
    out.scores = p.out_scores ? p.out_scores + o : nullptr;
    out.indices = p.out_indices ? p.out_indices + o : nullptr;
    block_merge_lists(src, P, p.k, p.k, p.k, out, scratch, surv);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Main kernel
// ------------------------------------------------------------------------------------------
// R corpus rows per warp iteration, BT queries; QREG keeps the (single) query in registers.
template <typename T, int BT, int R, bool QREG>
__global__ void __launch_bounds__(kStreamThreads, 1) cosine_topk_stream_kernel(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NP = Pairs<T>::NP;           // fp32 pairs per 16-byte chunk
  constexpr int NV = R * (BT + 1);           // reduced values per iteration: BT dots + 1 sum-of-squares per row
  constexpr int NVP = next_pow2(NV);
  constexpr int SH = 5 - ilog2(NVP);         // lanes per value after the transposing reduce = 1 << SH
  static_assert(NV <= 32, "too many values per warp iteration");
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.y * BT;
  const int nq = min(BT, p.n_queries - q0);
  const uint32_t row_bytes = static_cast<uint32_t>(p.dim) * sizeof(T);
  const int chunks = row_bytes / 16;

  // ---- shared-memory carve-up -------------------------------------------------------------
  unsigned char* tiles = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(p.stages) * p.tile_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* thr_s = bars + 2 * kMaxStages;                                          // [kConsumerWarps][8]
  float* qs = reinterpret_cast<float*>(thr_s + kConsumerWarps * 8);                 // [BT][dim] fp32
  uint64_t* lists = reinterpret_cast<uint64_t*>(qs + static_cast<size_t>(BT) * p.dim + ((BT * p.dim) & 1));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < kConsumerWarps * 8; i += blockDim.x) thr_s[i] = 0ull;
  for (int i = threadIdx.x; i < kConsumerWarps * BT * p.kpad; i += blockDim.x) {
    const int j = i % p.kpad;
    lists[i] = (j < p.k) ? 0ull : ~0ull;  // 0 = empty slot, ~0 = padding that is never replaced
  }
  __shared__ float qnorm_s[8];
  prepare_queries<BT>(p, q0, nq, qs, qnorm_s);  // ends with __syncthreads()

  if (warp == kConsumerWarps) {
    // ================= producer warp: one lane drives the bulk-copy ring =================
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
    // ================= consumer warps =================
    const int cw = warp;
    uint64_t* my_lists = lists + static_cast<size_t>(cw) * BT * p.kpad;
    uint64_t* my_thr = thr_s + cw * 8;
    // which reduced value this lane ends up holding
    const int vidx = lane >> SH;
    const int v_row = vidx / (BT + 1);
    const int v_q = vidx - v_row * (BT + 1);               // == BT -> the row's sum of squares
    const bool v_live = (vidx < NV) && (v_q < BT) && (v_q < nq);
    const int ssq_lane = (v_row * (BT + 1) + BT) << SH;

    unsigned long long qreg[QREG ? kMaxQReg : 1][NP];
    if (QREG) {
#pragma unroll
      for (int j = 0; j < kMaxQReg; ++j) {
        const int c = lane + 32 * j;
#pragma unroll
        for (int e = 0; e < NP; ++e)
          qreg[j][e] = (c < chunks) ? pack2(qs[c * 2 * NP + 2 * e], qs[c * 2 * NP + 2 * e + 1]) : 0ull;
      }
    }

    const int groups_per_tile = (p.tile_rows + R - 1) / R;
    int s = 0;
    uint32_t phase = 0;
    long long it = 0;  // tiles this CTA has consumed so far (rotates the group -> warp deal)
    for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      mbar_wait(&full_bar[s], phase);
      const long long row0 = t * p.tile_rows;
      const int rows = static_cast<int>(min(static_cast<long long>(p.tile_rows), p.n_rows - row0));
      const unsigned char* tile = tiles + static_cast<size_t>(s) * p.tile_bytes;
      const int ngroups = (rows + R - 1) / R;
      // deal groups round-robin over warps, continuing where the previous tile stopped
      int g = static_cast<int>((cw + kConsumerWarps - (it * groups_per_tile) % kConsumerWarps) % kConsumerWarps);
      for (; g < ngroups; g += kConsumerWarps) {
        const int r0 = g * R;
        const uint4* rowp[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
          rowp[rr] = reinterpret_cast<const uint4*>(tile + static_cast<size_t>(min(r0 + rr, rows - 1)) * row_bytes);
        unsigned long long acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0ull;

        if (QREG) {
#pragma unroll
          for (int j = 0; j < kMaxQReg; ++j) {
            const int c = lane + 32 * j;
            if (c < chunks) {
              uint4 raw[R];
#pragma unroll
              for (int rr = 0; rr < R; ++rr) raw[rr] = rowp[rr][c];
#pragma unroll
              for (int rr = 0; rr < R; ++rr) {
                unsigned long long x[NP];
                Pairs<T>::unpack(raw[rr], x);
#pragma unroll
                for (int e = 0; e < NP; ++e) {
                  acc[rr * (BT + 1) + BT] = ffma2(x[e], x[e], acc[rr * (BT + 1) + BT]);
                  acc[rr * (BT + 1)] = ffma2(x[e], qreg[j][e], acc[rr * (BT + 1)]);
                }
              }
            }
          }
        } else {
          for (int c = lane; c < chunks; c += 32) {
            uint4 raw[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) raw[rr] = rowp[rr][c];
            unsigned long long x[R][NP];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
              Pairs<T>::unpack(raw[rr], x[rr]);
#pragma unroll
              for (int e = 0; e < NP; ++e) acc[rr * (BT + 1) + BT] = ffma2(x[rr][e], x[rr][e], acc[rr * (BT + 1) + BT]);
            }
#pragma unroll
            for (int h = 0; h < NP / 2; ++h) {
#pragma unroll
              for (int b = 0; b < BT; ++b) {
                const ulonglong2 qq =
                    *reinterpret_cast<const ulonglong2*>(qs + static_cast<size_t>(b) * p.dim + c * 2 * NP + 4 * h);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                  acc[rr * (BT + 1) + b] = ffma2(x[rr][2 * h], qq.x, acc[rr * (BT + 1) + b]);
                  acc[rr * (BT + 1) + b] = ffma2(x[rr][2 * h + 1], qq.y, acc[rr * (BT + 1) + b]);
                }
              }
            }
          }
        }

        float vals[NVP];
#pragma unroll
        for (int i = 0; i < NVP; ++i) vals[i] = (i < NV) ? sum2(acc[i]) : 0.f;
        const float mine = transpose_reduce<NVP>(vals, lane);
        const float ssq = __shfl_sync(0xffffffffu, mine, ssq_lane);
        // sklearn's zero rule: a zero row is divided by 1 and scores 0.
        const float inv = ssq > 0.f ? 1.0f / sqrtf(ssq) : 1.0f;
        const bool row_ok = (r0 + v_row) < rows;
        if (p.all_scores != nullptr && v_live && row_ok && (lane & ((1 << SH) - 1)) == 0)
          p.all_scores[static_cast<size_t>(q0 + v_q) * p.n_rows + (row0 + r0 + v_row)] = mine * inv;
        const uint64_t key = make_key(mine * inv, p.index_base + static_cast<uint32_t>(row0 + r0 + v_row));
        const bool pass = v_live && row_ok && key > my_thr[v_live ? v_q : 0];
        uint32_t ballot = __ballot_sync(0xffffffffu, pass);
        while (ballot) {  // rare: a candidate beats this warp's current k-th best
          const int src = __ffs(ballot) - 1;
          ballot &= ~(((SH == 5) ? 0xffffffffu : ((1u << (1 << SH)) - 1u)) << (src & ~((1 << SH) - 1)));
          const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
          const int bb = __shfl_sync(0xffffffffu, v_q, src);
          if (kk > my_thr[bb]) {
            const uint64_t nt = list_insert(my_lists + bb * p.kpad, p.kpad, kk, lane);
            __syncwarp();
            if (lane == 0) my_thr[bb] = nt;
            __syncwarp();
          }
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

  // the tile ring is idle now: reuse it to stage the final merge
  cta_epilogue<BT>(p, lists, kConsumerWarps, q0, nq, reinterpret_cast<uint64_t*>(tiles),
                   static_cast<size_t>(p.stages) * p.tile_bytes);
}

// ------------------------------------------------------------------------------------------
// Generic fallback for rows that are not 16-byte multiples / unaligned: plain loads, same epilogue
// ------------------------------------------------------------------------------------------
template <typename T, int BT>
__global__ void __launch_bounds__(kGenericThreads, 1) cosine_topk_generic_kernel(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int q0 = blockIdx.y * BT;
  const int nq = min(BT, p.n_queries - q0);
  float* qs = reinterpret_cast<float*>(smem_raw);
  uint64_t* lists = reinterpret_cast<uint64_t*>(qs + static_cast<size_t>(BT) * p.dim + ((BT * p.dim) & 1));
  for (int i = threadIdx.x; i < nwarps * BT * p.kpad; i += blockDim.x) lists[i] = ((i % p.kpad) < p.k) ? 0ull : ~0ull;
  __shared__ float qnorm_s[8];
  prepare_queries<BT>(p, q0, nq, qs, qnorm_s);  // ends with __syncthreads()
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
      if (p.all_scores != nullptr && b < nq && lane == 0) p.all_scores[static_cast<size_t>(q0 + b) * p.n_rows + r] = d * inv;
      const uint64_t key = make_key(d * inv, grow);
      if (b < nq && key > thr[b]) thr[b] = list_insert(my_lists + b * p.kpad, p.kpad, key, lane);
    }
  }
  __syncthreads();
  cta_epilogue<BT>(p, lists, nwarps, q0, nq, nullptr, 0);
}

// ------------------------------------------------------------------------------------------
// Host side: configuration + launch
// ------------------------------------------------------------------------------------------
struct StreamConfig {
  bool bulk;  // TMA-ring kernel usable (16-byte rows, aligned base)
  int bt;
  int rows_per_iter;
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
static int rows_per_iter_for(int bt) { return bt <= 2 ? 4 : (bt == 4 ? 3 : 2); }

static int small_corpus_ctas_per_sm() {
  static const int v = getenv("SS_STREAM_SMALL_CTAS") ? std::max(1, atoi(getenv("SS_STREAM_SMALL_CTAS"))) : 1;
  return v;
}

static bool make_config(const void* corpus, long long n_rows, int dim, int dtype, int n_queries, int k,
                        StreamConfig* cfg) {
  const size_t es = dtype_size(dtype);
  const size_t row_bytes = static_cast<size_t>(dim) * es;
  const size_t smem_cap = smem_optin();
  const int sms = sm_count();
  cfg->kpad = static_cast<int>(align_up(k, 32));
  cfg->bulk = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(corpus) % 16 == 0) && row_bytes <= 24 * 1024;
  const int bt_opts[4] = {8, 4, 2, 1};
  if (cfg->bulk) {
    for (int bi = 0; bi < 4; ++bi) {
      const int bt = bt_opts[bi];
      if (bt > 1 && bt / 2 >= n_queries) continue;  // do not over-provision query slots
      const int R = rows_per_iter_for(bt);
      const int tile_rows = static_cast<int>(std::max<size_t>(R, (48 * 1024) / row_bytes / R * R));
      const uint32_t tile_bytes = static_cast<uint32_t>(align_up(tile_rows * row_bytes, 128));
      const size_t fixed = 2 * kMaxStages * 8 + kConsumerWarps * 8 * 8 + align_up(static_cast<size_t>(bt) * dim * 4, 8) +
                           list_bytes(kConsumerWarps, bt, cfg->kpad) + 128;
      if (fixed + 2 * static_cast<size_t>(tile_bytes) > smem_cap) continue;
      int stages = static_cast<int>((smem_cap - fixed) / tile_bytes);
      stages = std::min(stages, 4);
      cfg->bt = bt;
      cfg->rows_per_iter = R;
      cfg->stages = stages;
      cfg->tile_rows = tile_rows;
      cfg->tile_bytes = tile_bytes;
      cfg->n_tiles = (n_rows + tile_rows - 1) / tile_rows;
      cfg->grid_y = (n_queries + bt - 1) / bt;
      // Several query groups (grid.y) share the SMs: with a small corpus a full-width grid per group
      // only multiplies prologue / epilogue / merge work (measured on config 1: 0.94 ms full width, 0.43 ms at one CTA per SM in total).
      long long gx = std::min<long long>(sms, cfg->n_tiles);
      if (cfg->grid_y > 1 && cfg->n_tiles < 8LL * sms)
        gx = std::min<long long>(gx, std::max<long long>(1, (static_cast<long long>(small_corpus_ctas_per_sm()) * sms + cfg->grid_y - 1) / cfg->grid_y));
      cfg->grid_x = static_cast<int>(std::max<long long>(1, gx));
      cfg->threads = kStreamThreads;
      cfg->smem = fixed + static_cast<size_t>(stages) * tile_bytes;
      return true;
    }
    cfg->bulk = false;
  }
  for (int bi = 0; bi < 4; ++bi) {
    const int bt = bt_opts[bi];
    if (bt > 1 && bt / 2 >= n_queries) continue;
    const size_t need = align_up(static_cast<size_t>(bt) * dim * 4, 8) + list_bytes(kGenericThreads / 32, bt, cfg->kpad) + 128;
    if (need > smem_cap) continue;
    cfg->bt = bt;
    cfg->rows_per_iter = 1;
    cfg->stages = 0;
    cfg->tile_rows = 0;
    cfg->tile_bytes = 0;
    cfg->n_tiles = 0;
    cfg->grid_y = (n_queries + bt - 1) / bt;
    const long long want = (n_rows + (kGenericThreads / 32) - 1) / (kGenericThreads / 32);
    cfg->grid_x = static_cast<int>(std::max<long long>(1, std::min<long long>(sms, want)));
    cfg->threads = kGenericThreads;
    cfg->smem = need;
    return true;
  }
  return false;
}

template <typename K>
static cudaError_t launch_kernel(K kern, const StreamConfig& cfg, const StreamParams& p, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cfg.smem));
  if (e != cudaSuccess) return e;
  kern<<<dim3(cfg.grid_x, cfg.grid_y), cfg.threads, cfg.smem, st>>>(p);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_stream_bt(const StreamConfig& cfg, const StreamParams& p, int chunks, cudaStream_t st) {
  if (!cfg.bulk) {
    switch (cfg.bt) {
      case 8: return launch_kernel(cosine_topk_generic_kernel<T, 8>, cfg, p, st);
      case 4: return launch_kernel(cosine_topk_generic_kernel<T, 4>, cfg, p, st);
      case 2: return launch_kernel(cosine_topk_generic_kernel<T, 2>, cfg, p, st);
      default: return launch_kernel(cosine_topk_generic_kernel<T, 1>, cfg, p, st);
    }
  }
  switch (cfg.bt) {
    case 8: return launch_kernel(cosine_topk_stream_kernel<T, 8, 2, false>, cfg, p, st);
    case 4: return launch_kernel(cosine_topk_stream_kernel<T, 4, 3, false>, cfg, p, st);
    case 2: return launch_kernel(cosine_topk_stream_kernel<T, 2, 4, false>, cfg, p, st);
    default:
      if (chunks <= 32 * kMaxQReg) return launch_kernel(cosine_topk_stream_kernel<T, 1, 4, true>, cfg, p, st);
      return launch_kernel(cosine_topk_stream_kernel<T, 1, 4, false>, cfg, p, st);
  }
}

}  // namespace ss

using namespace ss;

extern "C" size_t ss_cosine_topk_stream_workspace_bytes(int64_t n_rows, int dim, int corpus_dtype, int n_queries, int k) {
  (void)n_rows;
  (void)corpus_dtype;
  if (dim <= 0 || n_queries <= 0 || k <= 0) return 0;
  const size_t tickets = align_up(static_cast<size_t>(n_queries) * 4, 256);
  const size_t partial = align_up(static_cast<size_t>(n_queries) * sm_count() * k * 8, 256);
  return tickets + partial + 256;
}

static int run_stream(const void* corpus, int64_t n_rows, int dim, int corpus_dtype, const void* queries, int n_queries,
                      int query_dtype, int k, uint32_t index_base, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                      float* out_scores, int64_t* out_indices, float* all_scores, void* stream) {
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
  unsigned int* tickets = reinterpret_cast<unsigned int*>(ws);
  const size_t tickets_bytes = align_up(static_cast<size_t>(n_queries) * 4, 256);
  uint64_t* partial = reinterpret_cast<uint64_t*>(ws + tickets_bytes);

  StreamConfig cfg;
  if (!make_config(corpus, n_rows, dim, corpus_dtype, n_queries, k, &cfg))
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_stream: dim / k combination does not fit in shared memory");
  SS_CUDA_CHECK(cudaMemsetAsync(tickets, 0, static_cast<size_t>(cfg.grid_y) * 4, st));
  StreamParams p;
  p.corpus = corpus;
  p.n_rows = n_rows;
  p.dim = dim;
  p.queries = queries;
  p.query_dtype = query_dtype;
  p.n_queries = n_queries;
  p.k = k;
  p.kpad = cfg.kpad;
  p.index_base = index_base;
  p.tile_rows = cfg.tile_rows;
  p.n_tiles = cfg.n_tiles;
  p.stages = cfg.stages;
  p.tile_bytes = cfg.tile_bytes;
  p.partial = partial;
  p.tickets = tickets;
  p.out_keys = out_keys;
  p.out_scores = out_scores;
  p.out_indices = reinterpret_cast<long long*>(out_indices);
  p.all_scores = all_scores;
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
  return SS_OK;
}

extern "C" int ss_cosine_topk_stream(const void* corpus, int64_t n_rows, int dim, int corpus_dtype, const void* queries,
                                     int n_queries, int query_dtype, int k, uint32_t index_base, void* workspace,
                                     size_t workspace_bytes, uint64_t* out_keys, float* out_scores, int64_t* out_indices,
                                     void* stream) {
  return run_stream(corpus, n_rows, dim, corpus_dtype, queries, n_queries, query_dtype, k, index_base, workspace, workspace_bytes,
                    out_keys, out_scores, out_indices, nullptr, stream);
}

extern "C" int ss_cosine_scores(const void* corpus, int64_t n_rows, int dim, int corpus_dtype, const void* queries, int n_queries,
                                int query_dtype, void* workspace, size_t workspace_bytes, float* out_all_scores, void* stream) {
  if (!out_all_scores) return fail(SS_ERR_INVALID_ARG, "ss_cosine_scores: null output");
  return run_stream(corpus, n_rows, dim, corpus_dtype, queries, n_queries, query_dtype, 1, 0, workspace, workspace_bytes, nullptr,
                    nullptr, nullptr, out_all_scores, stream);
}
