// K7 — HBM-bound cosine top-k for medium query batches (2..64 queries per corpus pass, any k up
// to 1024) with the dot products on the 5th-gen tensor cores.
//
// Replaces sklearn.cosine_similarity(Q, C) + np.argsort(-s)[:k] (Tool/rank_chunks_optimized.py:
// 215-216,225-235) where the CUDA-core streaming kernel (K1) runs out of FMA throughput (more than
// a few queries per pass) but the batch is far too small to fill K2's 128-query MMA tiles: BASELINE
// config 5 (16 queries, top-100, fp16) and every batch between K1 and K2.
//
// The roles of the operands are swapped with respect to K2: a 128-row CORPUS tile is the MMA "A"
// operand (TMEM lane = corpus row) and the whole query group (padded to 16/32/48/64 rows, resident
// in shared memory for the lifetime of the CTA) is the "B" operand, so one tcgen05.mma of shape
// M=128, N=Npad, K=16 costs 128*Npad/256 cycles — far below the time HBM needs to deliver the
// 4 KB of corpus it consumes.  The kernel therefore reads the corpus exactly once at HBM speed:
//
//   warp 0     TMA producer : cp.async.bulk.tensor of 128 rows x 128 bytes (one K block) per stage
//                             into a deep shared-memory ring (up to 12 x 16 KB in flight)
//   warp 1     MMA issuer   : tcgen05.mma into a double-buffered TMEM accumulator (2 x Npad columns)
//   warps 2-5  epilogue     : one corpus row per thread: tcgen05.ld the row's Npad dot products,
//                             scale by 1/|c| and 1/|q|, compare against the query's current
//                             threshold (shared memory) and append survivors to a small per-query
//                             candidate pool (>= 2k slots).  When a pool overflows, one warp
//                             bitonic-sorts it back to its k best (which raises the query's
//                             threshold) and the candidates that did not fit are re-read from the
//                             still-resident TMEM accumulator.
//   warps 6-9  row norms    : one corpus row per thread: read the same shared-memory stage the MMA
//                             reads (swizzle-agnostic: a row's eight 16-byte chunks in any order),
//                             accumulate sum(c^2) in fp32 and hand 1/|c| to the epilogue — corpus
//                             norms cost no HBM traffic and no separate pass.
//   warp 10    thresholds   : cross-CTA threshold sharing (see tcs_share_thresholds).
//
// Each CTA finally writes k sorted keys per query; ss_topk_merge folds the per-CTA lists.
#include <algorithm>

#include "tc_common.cuh"

namespace ss {

constexpr int T_BM = 128;                      // corpus rows per tile (MMA M, TMEM lanes)
constexpr int T_STAGE_BYTES = T_BM * 128;      // one K block: 128 rows x 128 bytes
constexpr int T_THREADS = 11 * 32;
constexpr int T_MAX_STAGES = 12;
constexpr int T_MAX_GROUP = 64;                // queries per corpus pass (MMA N <= 64)

struct TcStreamParams {
  long long n_rows;
  int dim;
  int n_queries;
  int k;
  int cap;         // candidate-pool capacity per query (power of two, >= 2k)
  int group;       // queries per corpus pass
  int npad;        // group rounded up to a multiple of 16 (MMA N)
  int nkb;         // K blocks per row (ceil(dim / 64))
  int stages;
  uint32_t index_base;
  uint32_t tmem_cols;
  long long n_tiles;
  const void* queries;  // raw [n_queries][dim], same 16-bit dtype as the corpus
  uint64_t* partial;    // [n_queries][gridDim.x][k]
  uint32_t* gbest;      // [n_queries][gridDim.x] best score seen by each CTA (order-preserving bits), zeroed before launch
  int share_rank;       // k when gridDim.x >= k (cross-CTA threshold sharing on), else 0
};

__device__ __forceinline__ void epi4_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// One warp sorts a[0..n) (n a power of two, shared memory) in descending order.
__device__ __forceinline__ void warp_bitonic_desc(uint64_t* a, int n, int lane) {
  for (int k2 = 2; k2 <= n; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const uint64_t x = a[i], y = a[p];
        const bool desc = (i & k2) == 0;
        if (desc ? (x < y) : (x > y)) {
          a[i] = y;
          a[p] = x;
        }
      }
      __syncwarp();
    }
  }
}

// Append make_key(sc, gidx) to a query's pool if it beats the query's current k-th best.  Returns
// true when the pool was full: the candidate stays pending and the caller retries after compaction.
__device__ __noinline__ bool tcs_try_append(uint64_t* pool, int cap, int* cnt_q, uint64_t thr_q, float sc, uint32_t gidx,
                                            int* ovf, int tok, uint32_t* best_q) {
  const uint64_t key = make_key(sc, gidx);
  if (!(key > thr_q)) return false;
  if (static_cast<uint32_t>(key >> 32) > *best_q) atomicMax(best_q, static_cast<uint32_t>(key >> 32));
  const int pos = atomicAdd(cnt_q, 1);
  if (pos < cap) {
    pool[pos] = key;
    return false;
  }
  *ovf = tok;
  return true;
}

// One warp: sort a full pool, keep its k best entries, publish the new threshold.
__device__ __noinline__ void tcs_compact(uint64_t* pool, int cap, int k, int* cnt_q, uint64_t* thr_q, float* thrf_q, int lane) {
  warp_bitonic_desc(pool, cap, lane);  // called only when all cap slots hold candidates
  if (lane == 0) {
    const uint64_t kth = pool[k - 1];
    *cnt_q = k;
    *thr_q = kth;
    const float f = key_score(kth);
    if (f > *thrf_q) *thrf_q = f;  // never below a bound shared by the other CTAs
  }
  __syncwarp();
}

// Cross-CTA threshold sharing (warp 10).  Every CTA scans a disjoint slice of the corpus for the
// same queries; a CTA's own k-th best is a weak filter (it has seen only 1/gridDim.x of the rows).
// Each CTA publishes the best score it has seen per query; if k CTAs have each seen a row scoring
// >= t, then at least k rows score >= t, so rows scoring below the k-th largest published maximum
// can never enter the global top-k.  Published values only grow, so a stale read is merely a
// weaker (still sound) bound.  Runs off the critical path until the epilogue raises `done`.
__device__ __forceinline__ void tcs_share_thresholds(uint32_t* gbest, int n_ctas, int cta, int rank, int nq,
                                                     const volatile uint32_t* best, volatile float* thrf,
                                                     const volatile int* done, int lane) {
  while (*done == 0) {
    if (lane < nq) *reinterpret_cast<volatile uint32_t*>(gbest + static_cast<size_t>(lane) * n_ctas + cta) = best[lane];
    if (lane + 32 < nq) *reinterpret_cast<volatile uint32_t*>(gbest + static_cast<size_t>(lane + 32) * n_ctas + cta) = best[lane + 32];
    for (int q = 0; q < nq; ++q) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = lane + 32 * j;
        v[j] = 0u;
        if (i < n_ctas) v[j] = (i == cta) ? best[q] : *reinterpret_cast<const volatile uint32_t*>(gbest + static_cast<size_t>(q) * n_ctas + i);
      }
      uint32_t t = 0u;  // largest t with at least `rank` published values >= t
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = t | (1u << bit);
        int c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) c += __popc(__ballot_sync(0xffffffffu, v[j] >= cand));
        if (c >= rank) t = cand;
      }
      if (lane == 0 && t != 0u) {
        const float f = ordered_to_float(t);
        if (f > thrf[q]) thrf[q] = f;
      }
    }
    __nanosleep(2000);
  }
}

template <typename T>
__global__ void __launch_bounds__(T_THREADS, 1)
cosine_topk_tcstream_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                            const TcStreamParams p, const uint32_t idesc) {
  extern __shared__ unsigned char tcs_smem_raw[];
  unsigned char* smem = tcs_smem_raw + ((1024u - (smem_u32(tcs_smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = smem;                                                        // [stages][16 KB]
  unsigned char* qtiles = ring + static_cast<size_t>(p.stages) * T_STAGE_BYTES;      // [nkb][npad][128 B]
  uint64_t* pools = reinterpret_cast<uint64_t*>(qtiles + static_cast<size_t>(p.nkb) * p.npad * 128);  // [group][cap]
  uint64_t* thr = pools + static_cast<size_t>(p.group) * p.cap;  // [T_MAX_GROUP] k-th best key per query (16-byte aligned)
  float* thrf = reinterpret_cast<float*>(thr + T_MAX_GROUP);    // [T_MAX_GROUP] score threshold of the pre-filter
  float* inv_q = thrf + T_MAX_GROUP;                            // [T_MAX_GROUP]
  int* cnt = reinterpret_cast<int*>(inv_q + T_MAX_GROUP);       // [T_MAX_GROUP] pool fill
  uint32_t* best = reinterpret_cast<uint32_t*>(cnt + T_MAX_GROUP);  // [T_MAX_GROUP] best score this CTA has seen
  float* inv_c_s = reinterpret_cast<float*>(best + T_MAX_GROUP);    // [2][T_BM] corpus inverse norms per accumulator
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(inv_c_s + 2 * T_BM);
  uint64_t* empty_bar = full_bar + T_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + T_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* norm_full = tmem_empty + 2;
  uint64_t* norm_empty = norm_full + 2;
  uint64_t* q_bar = norm_empty + 2;
  int* ovf = reinterpret_cast<int*>(q_bar + 1);                 // token of the last scan round in which a pool overflowed
  int* done = ovf + 1;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.y * p.group;
  const int nq = min(p.group, p.n_queries - q0);

  if (threadIdx.x == 0) {
    tmap_prefetch(&tmap_q);
    tmap_prefetch(&tmap_c);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1 + 4);  // MMA commit + the four norm warps
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
      mbar_init(&norm_full[a], 4);
      mbar_init(&norm_empty[a], 4);
    }
    mbar_init(q_bar, 1);
    *ovf = 0;
    *done = 0;
    mbar_fence_init();
  }
  // query inverse norms (sklearn: zero norm -> 1) and empty pools, by all warps
  for (int q = warp; q < T_MAX_GROUP; q += T_THREADS / 32) {
    float ss = 0.f;
    if (q < nq) {
      const T* qrow = static_cast<const T*>(p.queries) + static_cast<size_t>(q0 + q) * p.dim;
      for (int c = lane; c < p.dim; c += 32) {
        const float v = to_f32<T>(qrow[c]);
        ss = fmaf(v, v, ss);
      }
      ss = warp_sum(ss);
    }
    if (lane == 0) {
      inv_q[q] = (q < nq) ? (ss > 0.f ? 1.0f / sqrtf(ss) : 1.0f) : 0.f;
      thr[q] = 0ull;
      thrf[q] = (q < nq) ? -INFINITY : INFINITY;  // padding columns never pass the pre-filter
      cnt[q] = 0;
      best[q] = 0u;
    }
  }
  if (warp == 1) tmem_alloc(tmem_ptr_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_bar, static_cast<uint32_t>(p.nkb) * p.npad * 128u);
      for (int kb = 0; kb < p.nkb; ++kb)
        tma_load_2d(qtiles + static_cast<size_t>(kb) * p.npad * 128, &tmap_q, q_bar, kb * (128 / static_cast<int>(sizeof(T))), q0);
      int s = 0;
      uint32_t ph = 0;
      for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], T_STAGE_BYTES);
          tma_load_2d(ring + static_cast<size_t>(s) * T_STAGE_BYTES, &tmap_c, &full_bar[s], kb * (128 / static_cast<int>(sizeof(T))),
                      static_cast<int>(t * T_BM));
          if (++s == p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(q_bar, 0);
    tc_fence_after();
    int s = 0, acc = 0;
    uint32_t ph = 0, acc_ph = 0;
    const uint32_t q_addr = smem_u32(qtiles);
    for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.npad);
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc(smem_u32(ring + static_cast<size_t>(s) * T_STAGE_BYTES));
          const uint64_t db = make_smem_desc(q_addr + static_cast<uint32_t>(kb) * p.npad * 128u);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // four K=16 steps of 32 bytes inside the 128-byte swizzle atom
            umma_f16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == p.nkb - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  } else if (warp >= 6 && warp < 10) {
    // ===================== row norms: sum(c^2) of one corpus row per thread =====================
    constexpr int NP = Pairs<T>::NP;
    const int row = (warp - 6) * 32 + lane;
    const unsigned char* my_row_base = ring + static_cast<size_t>(row) * 128;
    int s = 0, acc = 0;
    uint32_t ph = 0, acc_ph = 0;
    for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      unsigned long long ssq2[4] = {0ull, 0ull, 0ull, 0ull};
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        const uint4* src = reinterpret_cast<const uint4*>(my_row_base + static_cast<size_t>(s) * T_STAGE_BYTES);
        uint4 raw[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) raw[c] = src[c ^ (row & 7)];  // conflict-free: 8 consecutive rows hit 8 distinct chunks
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);  // the data is in registers: release the stage before the math
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          unsigned long long x[NP];
          Pairs<T>::unpack(raw[c], x);
#pragma unroll
          for (int e = 0; e < NP; ++e) ssq2[e & 3] = ffma2(x[e], x[e], ssq2[e & 3]);
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      const float ssq = (sum2(ssq2[0]) + sum2(ssq2[1])) + (sum2(ssq2[2]) + sum2(ssq2[3]));
      mbar_wait(&norm_empty[acc], acc_ph ^ 1u);  // the epilogue has consumed the previous use of this slot
      inv_c_s[acc * T_BM + row] = ssq > 0.f ? 1.0f / sqrtf(ssq) : 1.0f;
      __syncwarp();
      if (lane == 0) mbar_arrive(&norm_full[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  } else if (warp == 10) {
    if (p.share_rank > 0)
      tcs_share_thresholds(p.gbest + static_cast<size_t>(q0) * gridDim.x, static_cast<int>(gridDim.x), static_cast<int>(blockIdx.x),
                           p.share_rank, nq, best, thrf, done, lane);
  } else {
    // ===================== epilogue: one corpus row (TMEM lane) per thread =====================
    const int ew = warp - 2;                   // 0..3
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // corpus row inside the tile == TMEM lane
    int acc = 0;
    uint32_t acc_ph = 0;
    int tok = 1;  // uniform across the 128 epilogue threads: one value per scan round
    for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const long long grow = t * T_BM + row;
      const bool row_ok = grow < p.n_rows;
      const uint32_t gidx = p.index_base + static_cast<uint32_t>(grow);
      mbar_wait(&norm_full[acc], acc_ph);
      const float inv_c = inv_c_s[acc * T_BM + row];
      __syncwarp();
      if (lane == 0) mbar_arrive(&norm_empty[acc]);
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * p.npad);
      unsigned long long pend = 0ull;  // queries whose candidate from this row did not fit in the pool yet
      for (int c0 = 0; c0 < p.npad; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c0), r);
        float iq[16], tf[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<float4*>(&iq[j]) = *reinterpret_cast<const float4*>(inv_q + c0 + j);
          *reinterpret_cast<float4*>(&tf[j]) = *reinterpret_cast<const float4*>(thrf + c0 + j);
        }
        tmem_ld_wait();
        uint32_t m = 0;
        float sc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sc[j] = (__uint_as_float(r[j]) * inv_c) * iq[j];
          m |= (!(sc[j] < tf[j]) ? 1u : 0u) << j;  // NaN passes (ranks last, like np.argsort(-x))
        }
        if (row_ok && m != 0u) {  // rare once the thresholds have warmed up
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if ((m >> j) & 1u) {
              const int q = c0 + j;
              if (q < nq && tcs_try_append(pools + static_cast<size_t>(q) * p.cap, p.cap, &cnt[q], thr[q], sc[j], gidx, ovf, tok, &best[q]))
                pend |= 1ull << q;
            }
          }
        }
      }
      // overflow rounds: sort the full pools back to k entries (which raises their thresholds), then
      // re-read the still-resident accumulator for the candidates that did not fit
      for (;;) {
        epi4_bar_sync();  // appends and overflow marks of this round are visible
        if (*reinterpret_cast<volatile int*>(ovf) != tok) break;
        for (int q = ew; q < nq; q += 4)
          if (cnt[q] >= p.cap) tcs_compact(pools + static_cast<size_t>(q) * p.cap, p.cap, p.k, &cnt[q], &thr[q], &thrf[q], lane);
        epi4_bar_sync();
        ++tok;
        unsigned long long wp = static_cast<unsigned long long>(__reduce_or_sync(0xffffffffu, static_cast<uint32_t>(pend))) |
                                (static_cast<unsigned long long>(__reduce_or_sync(0xffffffffu, static_cast<uint32_t>(pend >> 32))) << 32);
        while (wp) {
          const int q = __ffsll(static_cast<long long>(wp)) - 1;
          wp &= wp - 1;
          uint32_t v;
          tmem_ld1(taddr + static_cast<uint32_t>(q), v);
          tmem_ld_wait();
          if ((pend >> q) & 1ull) {
            pend &= ~(1ull << q);
            const float scq = (__uint_as_float(v) * inv_c) * inv_q[q];
            if (tcs_try_append(pools + static_cast<size_t>(q) * p.cap, p.cap, &cnt[q], thr[q], scq, gidx, ovf, tok, &best[q])) pend |= 1ull << q;
          }
        }
      }
      ++tok;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
    // final: stop the sharing warp, sort every pool and publish the CTA's k best keys per query
    if (threadIdx.x == 64) *reinterpret_cast<volatile int*>(done) = 1;
    epi4_bar_sync();
    for (int q = ew; q < nq; q += 4) {
      uint64_t* pool = pools + static_cast<size_t>(q) * p.cap;
      const int n = min(cnt[q], p.cap);
      for (int i = n + lane; i < p.cap; i += 32) pool[i] = 0ull;
      __syncwarp();
      warp_bitonic_desc(pool, p.cap, lane);
      uint64_t* out = p.partial + (static_cast<size_t>(q0 + q) * gridDim.x + blockIdx.x) * p.k;
      for (int j = lane; j < p.k; j += 32) out[j] = pool[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

struct TcStreamConfig {
  int group, npad, nkb, stages, cap, grid_x, grid_y;
  uint32_t tmem_cols;
  size_t smem;
  long long n_tiles;
};

static size_t tcs_fixed_smem(int nkb, int npad, int group, int cap) {
  return 1024 /*alignment slack*/ + static_cast<size_t>(nkb) * npad * 128 + static_cast<size_t>(group) * cap * 8 +
         T_MAX_GROUP * (8 + 4 + 4 + 4 + 4) + 2 * T_BM * 4 + (2 * T_MAX_STAGES + 9) * 8 + 64;
}

static bool make_tcs_config(long long n_rows, int dim, int n_queries, int k, TcStreamConfig* cfg) {
  const size_t cap_smem = smem_optin();
  const int nkb = (dim + 63) / 64;
  int cap = 32;
  while (cap < 2 * k) cap <<= 1;
  for (int group = std::min(T_MAX_GROUP, (n_queries + 15) / 16 * 16); group >= 1; group = (group > 16 ? group - 16 : group / 2)) {
    const int g = std::min(group, n_queries);
    const int npad = (g + 15) / 16 * 16;
    const size_t fixed = tcs_fixed_smem(nkb, npad, g, cap);
    if (fixed + 4 * static_cast<size_t>(T_STAGE_BYTES) > cap_smem) continue;
    cfg->group = g;
    cfg->npad = npad;
    cfg->nkb = nkb;
    cfg->cap = cap;
    cfg->stages = static_cast<int>(std::min<size_t>(T_MAX_STAGES, (cap_smem - fixed) / T_STAGE_BYTES));
    cfg->smem = fixed + static_cast<size_t>(cfg->stages) * T_STAGE_BYTES;
    cfg->n_tiles = (n_rows + T_BM - 1) / T_BM;
    cfg->grid_x = static_cast<int>(std::max<long long>(1, std::min<long long>(sm_count(), cfg->n_tiles)));
    cfg->grid_y = (n_queries + g - 1) / g;
    uint32_t cols = 32;
    while (cols < static_cast<uint32_t>(2 * npad)) cols <<= 1;
    cfg->tmem_cols = cols;
    return true;
  }
  return false;
}

static size_t tcs_partial_bytes(int n_queries, int k) { return align_up(static_cast<size_t>(n_queries) * sm_count() * k * 8, 256); }

}  // namespace ss

using namespace ss;

extern "C" size_t ss_cosine_topk_tcstream_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k) {
  (void)n_rows;
  (void)dim;
  if (n_queries <= 0 || k <= 0) return 0;
  return tcs_partial_bytes(n_queries, k) + align_up(static_cast<size_t>(n_queries) * sm_count() * 4, 256) + 256;
}

extern "C" int ss_cosine_topk_tcstream(const void* corpus, int64_t n_rows, int dim, int dtype, const void* queries, int n_queries,
                                       int k, uint32_t index_base, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                                       float* out_scores, int64_t* out_indices, void* stream) {
  if (!corpus || !queries || !workspace) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_tcstream: null pointer");
  if (n_rows <= 0 || dim <= 0 || n_queries <= 0 || k <= 0)
    return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_tcstream: sizes must be positive");
  if (dtype != SS_BF16 && dtype != SS_F16)
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_tcstream: corpus and queries must be bf16 or fp16");
  if (k > 1024) return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_tcstream: k > 1024 is not supported");
  if (dim % 8 != 0 || (reinterpret_cast<uintptr_t>(corpus) & 15) || (reinterpret_cast<uintptr_t>(queries) & 15))
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_tcstream: rows must be 16-byte multiples and 16-byte aligned");
  if (n_rows > 0x7FFFFFFFll - T_BM || static_cast<uint64_t>(index_base) + static_cast<uint64_t>(n_rows) > 0xFFFFFFFFull)
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_tcstream: row indices must fit in 32 bits");
  if (workspace_bytes < ss_cosine_topk_tcstream_workspace_bytes(n_rows, dim, n_queries, k))
    return fail(SS_ERR_WORKSPACE, "ss_cosine_topk_tcstream: workspace too small");
  TcStreamConfig cfg;
  if (!make_tcs_config(n_rows, dim, n_queries, k, &cfg))
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_tcstream: dim / k combination does not fit in shared memory");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint64_t* partial = reinterpret_cast<uint64_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));

  CUtensorMap tmap_q, tmap_c;
  if (!make_tmap_rows(&tmap_q, queries, dtype, n_queries, dim, cfg.npad) || !make_tmap_rows(&tmap_c, corpus, dtype, n_rows, dim, T_BM))
    return fail(SS_ERR_CUDA, "ss_cosine_topk_tcstream: cuTensorMapEncodeTiled failed");

  TcStreamParams p;
  p.n_rows = n_rows;
  p.dim = dim;
  p.n_queries = n_queries;
  p.k = k;
  p.cap = cfg.cap;
  p.group = cfg.group;
  p.npad = cfg.npad;
  p.nkb = cfg.nkb;
  p.stages = cfg.stages;
  p.index_base = index_base;
  p.tmem_cols = cfg.tmem_cols;
  p.n_tiles = cfg.n_tiles;
  p.queries = queries;
  p.partial = partial;
  p.gbest = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(partial) + tcs_partial_bytes(n_queries, k));
  // sharing needs k CTAs that each hold at least one row, and the 8-values-per-lane read of the sharing warp
  p.share_rank = (cfg.grid_x >= k && cfg.grid_x <= 256 && cfg.grid_x > 1) ? k : 0;
  if (p.share_rank > 0) SS_CUDA_CHECK(cudaMemsetAsync(p.gbest, 0, static_cast<size_t>(n_queries) * cfg.grid_x * 4, st));
  const uint32_t idesc = make_idesc(dtype == SS_BF16 ? 1 : 0, T_BM, cfg.npad);
  cudaError_t e;
  {
    ProfileScope prof(st);
    if (dtype == SS_BF16) {
      e = cudaFuncSetAttribute(cosine_topk_tcstream_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(cfg.smem));
      if (e == cudaSuccess)
        cosine_topk_tcstream_kernel<__nv_bfloat16><<<dim3(cfg.grid_x, cfg.grid_y), T_THREADS, cfg.smem, st>>>(tmap_q, tmap_c, p, idesc);
    } else {
      e = cudaFuncSetAttribute(cosine_topk_tcstream_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(cfg.smem));
      if (e == cudaSuccess)
        cosine_topk_tcstream_kernel<__half><<<dim3(cfg.grid_x, cfg.grid_y), T_THREADS, cfg.smem, st>>>(tmap_q, tmap_c, p, idesc);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (e != cudaSuccess) return cuda_fail(e, "cosine_topk_tcstream launch");
  return ss_topk_merge(partial, cfg.grid_x, n_queries, k, static_cast<int64_t>(cfg.grid_x) * k, k, k, out_keys, out_scores,
                       out_indices, stream);
}
