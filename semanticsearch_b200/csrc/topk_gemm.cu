// K2 — query-batch cosine similarity on the 5th-gen tensor cores with a fused top-k epilogue.
//
// Replaces sklearn.cosine_similarity(Q, C) + np.argsort(-s) (Tool/rank_chunks_optimized.py:215-216,
// 225-235) when the query batch makes the work a real GEMM.  D[query, corpus_row] = Q . C^T with
// bf16/fp16 operands and fp32 accumulation:
//
//   warp 0   TMA producer : cp.async.bulk.tensor (SWIZZLE_128B) of a 128 x 64 query tile and a
//                           256 x 64 corpus tile per K block into a 4-stage shared-memory ring
//   warp 1   MMA issuer   : one elected thread issues tcgen05.mma (M=128, N=256, K=16) on shared
//                           memory descriptors; accumulators live in TMEM, double-buffered
//                           (2 x 256 columns) so the next tile's MMAs overlap this tile's epilogue
//   warps 2-5 epilogue    : tcgen05.ld the 128 x 256 accumulator tile (one query row per thread),
//                           scale by 1/|c| (shared-memory broadcast) and 1/|q|, threshold against
//                           the thread's current k-th best and keep a sorted top-k in registers.
//
// The B x N score matrix is never written.  Work is split into (query block, corpus chunk) units
// so that all query blocks of one chunk run concurrently and share corpus tiles through L2; each
// unit publishes its finds into one global best-first key list per query (lock-free, see publish_key).
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

namespace ss {

constexpr int G_BM = 128;
constexpr int G_BN = 256;
constexpr int G_BK = 64;
constexpr int G_MAX_STAGES = 6;
constexpr int G_A_BYTES = G_BM * G_BK * 2;
// CG = 1: one CTA per 128 x 256 tile.  CG = 2: a CTA pair (two SMs of one TPC) computes a 256 x 256
// tile with tcgen05.mma.cta_group::2 — each CTA stages its own 128 query rows and only HALF of the
// corpus tile, so the L2 -> shared-memory traffic and the shared-memory bandwidth per MMA drop by a
// third and the ring gets deeper.
template <int CG> struct GemmCfg {
  static constexpr int kBRows = G_BN / CG;                 // corpus rows this CTA stages per tile
  static constexpr int kBBytes = kBRows * G_BK * 2;
  static constexpr int kStageBytes = G_A_BYTES + kBBytes;  // 48 KB / 32 KB
  static constexpr int kMaxStages = CG == 1 ? 4 : G_MAX_STAGES;
};
constexpr int G_EPI_WARPS = 8;  // two warps per TMEM lane quadrant, each takes half of the tile's columns
constexpr int G_THREADS = (2 + G_EPI_WARPS) * 32;
constexpr int G_EPI_THREADS = G_EPI_WARPS * 32;
constexpr int G_EPI_COLS = G_BN / (G_EPI_WARPS / 4);  // columns per epilogue thread and tile
constexpr int G_TMEM_COLS = 512;

struct GemmParams {
  int n_queries;
  long long n_rows;
  int dim;
  int k;
  uint32_t index_base;
  int n_qb;  // query blocks of 128 * CG rows
  long long n_tiles;
  int tiles_per_chunk;
  int n_chunks;
  int warm_chunks;  // the first warm_chunks chunks hold ONE tile each: the launch's cold thresholds cost one tile per CTA, not a chunk
  long long n_units;
  int stages;
  const float* inv_c;  // padded to a multiple of G_BN entries, NaN past n_rows (never selected)
  const float* inv_gmax;  // [n_fill / 32] largest 1/|c| of each 32-row group (NaN rows ignored)
  const float* inv_q;
  uint64_t* gtop;  // [n_queries][k] global best-first key list of every query, zeroed before launch (see publish_list)
  int* progress;   // [gridDim.x] units started by every CTA's producer, zeroed before launch (drift throttle); may be null
  int max_lead;    // a producer starts its j-th unit only when the slowest CTA has started its (j - max_lead)-th
};

// Per-thread top-k list in shared memory, entry j of epilogue thread e at list[j * 256 + e] (conflict-free across a
// warp).  The list is UNSORTED: the thread remembers where its weakest entry sits (`min_pos`), a candidate that beats the
// k-th best replaces that entry with one store, and one pass of k independent loads finds the new weakest entry
// (score ascending, then row index descending, so that among equal scores the higher row leaves first and strict '>'
// keeps the earlier, lower-index entry).  A sorted insertion costs a dependent load / compare / store chain per slot,
// ~500 cycles per event in the cold phase of a launch, when every column passes in some lane of the warp; the global
// lists take keys in any order, so nothing needs the sort.  Precondition: sc > current k-th best.  Returns the new one.
struct ScoreIdx {
  float v;
  int ix;
};
__device__ __noinline__ float epi_list_insert(ScoreIdx* list, int k, float sc, int c, int& min_pos) {
  ScoreIdx e;
  e.v = sc;
  e.ix = c;
  list[min_pos * G_EPI_THREADS] = e;
  float mv = INFINITY;
  int mi = -1, mp = 0;
#pragma unroll 4
  for (int j = 0; j < k; ++j) {
    const ScoreIdx x = list[j * G_EPI_THREADS];
    const bool weaker = x.v < mv || (x.v == mv && x.ix > mi);
    mv = weaker ? x.v : mv;
    mi = weaker ? x.ix : mi;
    mp = weaker ? j : mp;
  }
  min_pos = mp;
  return mv;
}

// Tile range of a chunk.  The first `warm_chunks` chunks are single tiles: every unit of the first waves is short, so the
// global lists hold useful thresholds (the k-th best of ~10 k rows per query) after ~0.1 ms instead of after the ~0.55 ms that
// 8-tile units with cold thresholds take (profiles/r02_k2_trace_experiment.txt); the remaining chunks have tiles_per_chunk tiles.
__device__ __forceinline__ void chunk_tiles(int chunk, int warm_chunks, int tiles_per_chunk, long long n_tiles, long long& t0, long long& t1) {
  if (chunk < warm_chunks) {
    t0 = chunk;
    t1 = chunk + 1;
  } else {
    t0 = warm_chunks + static_cast<long long>(chunk - warm_chunks) * tiles_per_chunk;
    t1 = t0 + tiles_per_chunk;
  }
  t1 = t1 < n_tiles ? t1 : n_tiles;
}

// Global per-query result list, shared by every unit of the launch: gtop[query][0..k) holds packed keys, best first.
// A unit publishes a key with a cascade of 64-bit atomicMax operations — slot s keeps the larger of (its key, the
// incoming key) and hands the smaller one to slot s + 1.  Slots only grow, what a slot hands down never exceeds what
// it keeps, and every key lives in exactly one place (a slot or the register of one cascade in flight), so at ANY moment
// the array is sorted and its non-empty slots are distinct real rows: gtop[query][k-1] != 0 proves that k rows score at
// least that much, i.e. it is a sound lower bound of the final k-th best, and after the last cascade the array is the
// exact top-k by (score desc, row asc) whatever the interleaving.  Units read that bound when they start, so only the
// first wave of units runs with cold thresholds, and no per-unit partial lists or merge pass exist.
__device__ __forceinline__ void publish_key(unsigned long long* g, int k, unsigned long long x) {
  for (int s = 0; s < k && x != 0ull; ++s) {
    const unsigned long long old = atomicMax(g + s, x);
    x = old < x ? old : x;
  }
}

// One 32-column group of one query row.  Cheap bound first: with a non-negative threshold no column can
// beat it unless max(raw dot) * max(1/|c| of the group) * 1/|q| does, which needs no per-column multiply;
// the exact scores (raw * 1/|c| * 1/|q|) are formed only for the rare group that passes.
__device__ __forceinline__ void epi_group(const uint32_t (&r)[32], const float* inv_grp, float gmax, float inv_q, float& thr,
                                          float thr_floor, ScoreIdx* my_list, int& min_pos, int k, int col0) {
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    m = fmaxf(fmaxf(fmaxf(m, __uint_as_float(r[j])), fmaxf(__uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]))), __uint_as_float(r[j + 3]));
  if (thr < 0.f || fmaxf(m, 0.f) * gmax * inv_q > thr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 iv = *reinterpret_cast<const float4*>(inv_grp + j);
      const float ivs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float sc = (__uint_as_float(r[j + e]) * ivs[e]) * inv_q;
        if (sc > thr) thr = fmaxf(thr_floor, epi_list_insert(my_list, k, sc, col0 + j + e, min_pos));
      }
    }
  }
}

template <int CG>
__global__ void __launch_bounds__(G_THREADS, 1)
cosine_topk_gemm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                        const GemmParams p, const uint32_t idesc) {
  constexpr int G_STAGE_BYTES = GemmCfg<CG>::kStageBytes;
  const int G_STAGES = p.stages;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;  // rank 0 of a pair issues the MMAs
  const int n_workers = static_cast<int>(gridDim.x) / CG;       // CTAs (CG = 1) or CTA pairs (CG = 2)
  const int worker = static_cast<int>(blockIdx.x) / CG;
  extern __shared__ unsigned char gemm_smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  unsigned char* smem = gemm_smem_raw + ((1024u - (smem_u32(gemm_smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = smem;
  float* sinv = reinterpret_cast<float*>(tiles + static_cast<size_t>(G_STAGES) * G_STAGE_BYTES);  // [2][G_BN] corpus inverse norms
  float* sgmax = sinv + 2 * G_BN;                                                                  // [2][G_BN / 32] their 32-row group maxima
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sgmax + 2 * (G_BN / 32));
  uint64_t* empty_bar = full_bar + G_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + G_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* inv_full = tmem_empty + 2;
  uint64_t* inv_empty = inv_full + 2;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(inv_empty + 2);
  ScoreIdx* lists = reinterpret_cast<ScoreIdx*>(tmem_ptr_s + 4);  // [k][G_EPI_THREADS] per-thread top-k lists

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.dim + G_BK - 1) / G_BK;

  if (threadIdx.x == 0) {
    tmap_prefetch(&tmap_q);
    tmap_prefetch(&tmap_c);
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], G_EPI_WARPS * CG);  // the epilogue warps of every CTA that shares the accumulator
      mbar_init(&inv_full[a], 1);
      mbar_init(&inv_empty[a], G_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_pair(tmem_ptr_s, G_TMEM_COLS); else tmem_alloc(tmem_ptr_s, G_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {
      int s = 0, ia = 0, j_unit = 0;
      uint32_t ph = 0, ia_ph = 0;
      for (long long u = worker; u < p.n_units; u += n_workers) {
        // Drift throttle.  The query blocks that share a corpus chunk run on different CTAs at about the same time and
        // meet in L2; with a static unit list nothing keeps them together for thousands of units, and a CTA that runs
        // ahead re-reads from DRAM what the others will fetch again later.  Every producer counts the units it has
        // started; nobody starts unit j before the slowest CTA has started unit j - max_lead (the slowest never waits).
        if (p.progress != nullptr) {
          ++j_unit;
          if (lane == 0) *reinterpret_cast<volatile int*>(p.progress + blockIdx.x) = j_unit;
          // The wait is a performance hint, never a dependency: it is bounded (~0.25 ms, several units), so a CTA whose
          // siblings are not resident yet (another kernel holds their SMs) carries on instead of waiting for them.
          for (int spins = 0; spins < 512; ++spins) {
            int mn = 0x7fffffff;
            for (int i = lane; i < static_cast<int>(gridDim.x); i += 32) mn = min(mn, *reinterpret_cast<volatile int*>(p.progress + i));
            mn = __reduce_min_sync(0xffffffffu, mn);
            if (j_unit - mn <= p.max_lead) break;
            __nanosleep(500);
          }
        }
        if (lane != 0) continue;
        const int chunk = static_cast<int>(u / p.n_qb), qb = static_cast<int>(u % p.n_qb) * CG + static_cast<int>(cta_rank);
        long long t0, t1;
        chunk_tiles(chunk, p.warm_chunks, p.tiles_per_chunk, p.n_tiles, t0, t1);
        for (long long t = t0; t < t1; ++t) {
          // this tile's 256 corpus inverse norms (1 KB) ride the TMA engine too
          mbar_wait(&inv_empty[ia], ia_ph ^ 1u);
          mbar_arrive_expect_tx(&inv_full[ia], G_BN * 4 + (G_BN / 32) * 4);
          bulk_copy_g2s(sinv + ia * G_BN, p.inv_c + t * G_BN, G_BN * 4, &inv_full[ia]);
          bulk_copy_g2s(sgmax + ia * (G_BN / 32), p.inv_gmax + t * (G_BN / 32), (G_BN / 32) * 4, &inv_full[ia]);
          if (++ia == 2) {
            ia = 0;
            ia_ph ^= 1u;
          }
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1u);
            unsigned char* a_dst = tiles + static_cast<size_t>(s) * G_STAGE_BYTES;
            if (CG == 1) {
              mbar_arrive_expect_tx(&full_bar[s], G_STAGE_BYTES);
              tma_load_2d(a_dst, &tmap_q, &full_bar[s], kb * G_BK, qb * G_BM);
              tma_load_2d(a_dst + G_A_BYTES, &tmap_c, &full_bar[s], kb * G_BK, static_cast<int>(t * G_BN));
            } else {
              // both CTAs credit the LEADER's barrier, which expects the bytes of the whole pair
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * G_STAGE_BYTES);
              const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[s]), 0);
              tma_load_2d_pair(a_dst, &tmap_q, lead_bar, kb * G_BK, qb * G_BM);
              tma_load_2d_pair(a_dst + G_A_BYTES, &tmap_c, lead_bar, kb * G_BK,
                               static_cast<int>(t * G_BN) + static_cast<int>(cta_rank) * GemmCfg<CG>::kBRows);
            }
            if (++s == G_STAGES) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
      // a CTA that has issued its last load holds nobody back
      if (p.progress != nullptr && lane == 0) *reinterpret_cast<volatile int*>(p.progress + blockIdx.x) = 0x7fffffff;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader only when CG = 2) =====================
    int s = 0, acc = 0;
    uint32_t ph = 0, acc_ph = 0;
    const uint32_t tiles_lo = smem_desc_lo(smem_u32(tiles));
    for (long long u = worker; cta_rank == 0 && u < p.n_units; u += n_workers) {
      const int chunk = static_cast<int>(u / p.n_qb);
      long long t0, t1;
      chunk_tiles(chunk, p.warm_chunks, p.tiles_per_chunk, p.n_tiles, t0, t1);
      for (long long t = t0; t < t1; ++t) {
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * G_BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);  // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_lo = tiles_lo + static_cast<uint32_t>(s) * (G_STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + (G_A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k)  // +32 bytes (2 x 16-byte units) per K=16 step inside the swizzle atom
            umma_f16_lohi<CG>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_elect<CG>(&empty_bar[s]);                      // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if (kb == nkb - 1) umma_commit_elect<CG>(&tmem_full[acc]);  // accumulator complete -> epilogue
          if (++s == G_STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_ph ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue: one query row x half of the tile's columns per thread =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // which 128 columns of the 256-column tile
    const int row = quad * 32 + lane;          // query row inside the 128-row block
    const int et = (warp - 2) * 32 + lane;     // 0..255 among epilogue threads
    ScoreIdx* my_list = lists + et;
    int acc = 0;
    uint32_t acc_ph = 0;
    const uint32_t lead_tmem_empty0 = CG == 2 ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u;
    const uint32_t lead_tmem_empty1 = CG == 2 ? mapa_u32(smem_u32(&tmem_empty[1]), 0) : 0u;
    for (long long u = worker; u < p.n_units; u += n_workers) {
      const int chunk = static_cast<int>(u / p.n_qb), qb = static_cast<int>(u % p.n_qb) * CG + static_cast<int>(cta_rank);
      long long t0, t1;
      chunk_tiles(chunk, p.warm_chunks, p.tiles_per_chunk, p.n_tiles, t0, t1);
      const int query = qb * G_BM + row;
      const float inv_q = query < p.n_queries ? p.inv_q[query] : 0.f;
      for (int j = 0; j < p.k; ++j) {
        ScoreIdx e;
        e.v = -INFINITY;
        e.ix = -1;
        my_list[j * G_EPI_THREADS] = e;
      }
      int min_pos = 0;  // every slot is empty: any of them is the weakest
      // Seed the threshold with the k-th best score published so far for this query (all units, all chunks): rows
      // scoring below it cannot reach the global top-k.  One step below: equal scores still compete on the row index.
      float thr = -INFINITY;
      unsigned long long* gq = reinterpret_cast<unsigned long long*>(p.gtop) + static_cast<size_t>(query < p.n_queries ? query : 0) * p.k;
      if (query < p.n_queries) {
        const uint32_t g = static_cast<uint32_t>(*reinterpret_cast<volatile unsigned long long*>(gq + p.k - 1) >> 32);
        if (g > 1u) thr = ordered_to_float(g - 1u);
      }
      const float thr_floor = thr;
      for (long long t = t0; t < t1; ++t) {
        const float* inv_tile = sinv + acc * G_BN + half * G_EPI_COLS;
        const float* gmax_tile = sgmax + acc * (G_BN / 32) + half * (G_EPI_COLS / 32);
        mbar_wait(&inv_full[acc], acc_ph);
        mbar_wait(&tmem_full[acc], acc_ph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * G_BN + half * G_EPI_COLS);
        const int col_base = static_cast<int>(t * G_BN) + half * G_EPI_COLS;
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr, ra);
#pragma unroll 1
        for (int c0 = 0; c0 < G_EPI_COLS; c0 += 64) {
          tmem_ld_wait();                                   // ra = columns c0 .. c0+31
          tmem_ld32(taddr + static_cast<uint32_t>(c0 + 32), rb);  // in flight while ra is processed
          epi_group(ra, inv_tile + c0, gmax_tile[c0 >> 5], inv_q, thr, thr_floor, my_list, min_pos, p.k, col_base + c0);
          tmem_ld_wait();                                   // rb = columns c0+32 .. c0+63
          if (c0 + 64 < G_EPI_COLS) tmem_ld32(taddr + static_cast<uint32_t>(c0 + 64), ra);
          epi_group(rb, inv_tile + c0 + 32, gmax_tile[(c0 >> 5) + 1], inv_q, thr, thr_floor, my_list, min_pos, p.k, col_base + c0 + 32);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&inv_empty[acc]);
          if (CG == 1) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_cluster(acc ? lead_tmem_empty1 : lead_tmem_empty0);
        }
        if (++acc == 2) {
          acc = 0;
          acc_ph ^= 1u;
        }
      }
      if (query < p.n_queries) {
        // publish what this unit found (any order): only keys above the current global k-th best can enter
        unsigned long long gk = *reinterpret_cast<volatile unsigned long long*>(gq + p.k - 1);
        for (int j = 0; j < p.k; ++j) {
          const ScoreIdx e = my_list[j * G_EPI_THREADS];
          if (e.ix < 0) continue;
          const unsigned long long key = make_key(e.v, p.index_base + static_cast<uint32_t>(e.ix));
          if (key > gk) {
            publish_key(gq, p.k, key);
            gk = *reinterpret_cast<volatile unsigned long long*>(gq + p.k - 1);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, G_TMEM_COLS); else tmem_dealloc(tmem_base, G_TMEM_COLS);
  }
}

// Final form of the global lists: packed keys -> (keys, scores, indices); empty slots read -inf / -1.
__global__ void __launch_bounds__(256) decode_keys_kernel(const uint64_t* __restrict__ keys, long long n, uint64_t* __restrict__ out_keys,
                                                          float* __restrict__ out_scores, long long* __restrict__ out_indices) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  if (out_keys) out_keys[i] = key;
  if (out_scores) out_scores[i] = key ? key_score(key) : -INFINITY;
  if (out_indices) out_indices[i] = key ? key_index(key) : -1;
}

// ---- vectorised row inverse norms (pre-pass over the corpus, HBM-bound) -------------------------
template <typename T>
__global__ void __launch_bounds__(256) row_inv_norms_vec_kernel(const T* __restrict__ rows, long long n_rows, int dim,
                                                                float zero_value, float* __restrict__ out, long long n_fill) {
  constexpr int NP = Pairs<T>::NP;
  // entries past the last row (up to the padded length) are NaN: such columns are never selected
  for (long long i = n_rows + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_fill;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __int_as_float(0x7fc00000);
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int chunks = dim * static_cast<int>(sizeof(T)) / 16;
  for (long long r = warp * 2; r < n_rows; r += nwarps * 2) {
    const uint4* ra = reinterpret_cast<const uint4*>(rows + static_cast<size_t>(r) * dim);
    const bool two = r + 1 < n_rows;
    const uint4* rb = two ? reinterpret_cast<const uint4*>(rows + static_cast<size_t>(r + 1) * dim) : ra;
    unsigned long long sa = 0ull, sb = 0ull;
    for (int c = lane; c < chunks; c += 32) {
      const uint4 xa = __ldg(ra + c), xb = __ldg(rb + c);
      unsigned long long pa[NP], pb[NP];
      Pairs<T>::unpack(xa, pa);
      Pairs<T>::unpack(xb, pb);
#pragma unroll
      for (int e = 0; e < NP; ++e) {
        sa = ffma2(pa[e], pa[e], sa);
        sb = ffma2(pb[e], pb[e], sb);
      }
    }
    const float fa = warp_sum(sum2(sa)), fb = warp_sum(sum2(sb));
    if (lane == 0) {
      out[r] = fa > 0.f ? 1.0f / sqrtf(fa) : zero_value;
      if (two) out[r + 1] = fb > 0.f ? 1.0f / sqrtf(fb) : zero_value;
    }
  }
}

// largest finite inverse norm of every 32-row group (feeds the epilogue's multiply-free bound)
__global__ void __launch_bounds__(256) inv_group_max_kernel(const float* __restrict__ inv, long long n_groups, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long g = warp; g < n_groups; g += nwarps) {
    float v = inv[g * 32 + lane];
    v = (v == v) ? v : 0.f;  // NaN marks rows past the end
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) out[g] = v;
  }
}

template <typename T>
static cudaError_t launch_inv_norms_vec(const void* rows, long long n_rows, int dim, float zero_value, float* out, long long n_fill,
                                        cudaStream_t st) {
  const long long want = (n_rows + 15) / 16;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(want, static_cast<long long>(sm_count()) * 16)));
  row_inv_norms_vec_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(rows), n_rows, dim, zero_value, out, n_fill);
  return cudaGetLastError();
}

struct GemmPlan {
  int n_qb;
  long long n_tiles;
  int tiles_per_chunk;
  int n_chunks;
  int warm_chunks;
};

// CTA pairs need cluster launches of 2 CTAs with ~212 KB of shared memory each to be schedulable on this
// device (they are on a full B200; MIG slices or an odd SM count fall back to single-CTA tiles).
static bool gemm_pairs_supported() {
  static int cached[64];  // 0 = unknown, 1 = yes, 2 = no
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (cached[dev] == 0) {
    bool ok = sm_count() % 2 == 0;
    if (ok) {
      const size_t smem = smem_optin() - 1024;
      ok = cudaFuncSetAttribute(cosine_topk_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) == cudaSuccess;
      if (ok) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(2);
        lc.blockDim = dim3(G_THREADS);
        lc.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        int n_clusters = 0;
        ok = cudaOccupancyMaxActiveClusters(&n_clusters, cosine_topk_gemm_kernel<2>, &lc) == cudaSuccess && n_clusters >= 1;
      }
      cudaGetLastError();  // a failed probe must not poison later launches
    }
    cached[dev] = ok ? 1 : 2;
  }
  return cached[dev] == 1;
}
static int gemm_cta_group(int n_queries) { return (n_queries > G_BM && gemm_pairs_supported()) ? 2 : 1; }

static GemmPlan make_gemm_plan(long long n_rows, int n_queries) {
  GemmPlan g;
  const int cg = gemm_cta_group(n_queries);
  g.n_qb = (n_queries + G_BM * cg - 1) / (G_BM * cg);
  g.n_tiles = (n_rows + G_BN - 1) / G_BN;
  // Short units keep the query blocks that share a corpus chunk inside an L2-sized window (their drift is bounded by the
  // unit length); the global result lists make a unit start cheap (one threshold read per query).  Measured on B200
  // (profiles/r02_k2_units_sweep.txt): chunks of ~8 tiles (3 MB of corpus) are the optimum at 1.25 M and at 10 M rows.
  static const int units_per_worker = getenv("SS_GEMM_UNITS_PER_WORKER") ? std::max(1, atoi(getenv("SS_GEMM_UNITS_PER_WORKER"))) : 1024;
  const long long target_units = static_cast<long long>(sm_count() / cg) * units_per_worker;
  long long n_chunks = std::max<long long>(1, (target_units + g.n_qb - 1) / g.n_qb);
  static const int min_tiles = getenv("SS_GEMM_MIN_TILES") ? std::max(1, atoi(getenv("SS_GEMM_MIN_TILES"))) : 8;
  n_chunks = std::min<long long>(n_chunks, std::max<long long>(1, g.n_tiles / min_tiles));  // at least ~min_tiles tiles per chunk
  g.tiles_per_chunk = static_cast<int>((g.n_tiles + n_chunks - 1) / n_chunks);
  // warm-up prefix of single-tile chunks (16 measured best on 0.3 M and 1.25 M-row shards, profiles/r02_k2_warm_sweep.txt)
  static const int warm = getenv("SS_GEMM_WARM_TILES") ? std::max(0, atoi(getenv("SS_GEMM_WARM_TILES"))) : 16;
  g.warm_chunks = g.tiles_per_chunk > 1 ? static_cast<int>(std::min<long long>(warm, g.n_tiles / 4)) : 0;
  const long long rest = g.n_tiles - g.warm_chunks;
  g.n_chunks = g.warm_chunks + static_cast<int>((rest + g.tiles_per_chunk - 1) / g.tiles_per_chunk);
  return g;
}

}  // namespace ss

using namespace ss;

extern "C" size_t ss_cosine_topk_gemm_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k) {
  (void)dim;
  if (n_rows <= 0 || n_queries <= 0 || k <= 0) return 0;
  const GemmPlan g = make_gemm_plan(n_rows, n_queries);
  (void)g;
  return align_up(static_cast<size_t>(n_rows), G_BN) * 4 + align_up(align_up(static_cast<size_t>(n_rows), G_BN) / 32 * 4, 256) +
         align_up(static_cast<size_t>(n_queries) * 4, 256) + align_up(static_cast<size_t>(n_queries) * k * 8, 256) + 4096 + 256;
}

// corpus_norms_valid: the head of `workspace` (inverse norms of the corpus rows and their 32-row group maxima, whose
// offsets depend on n_rows only) still holds what an earlier call wrote for THIS corpus — a resident index skips the
// norm pre-pass (one read of the whole corpus, 2.3 ms of a 57 ms step at 10 M x 768).
static int cosine_topk_gemm_impl(const void* corpus, int64_t n_rows, int dim, int dtype, const void* queries, int n_queries,
                                 int k, uint32_t index_base, void* workspace, size_t workspace_bytes, bool corpus_norms_valid,
                                 uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream) {
  if (!corpus || !queries || !workspace) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_gemm: null pointer");
  if (n_rows <= 0 || dim <= 0 || n_queries <= 0 || k <= 0) return fail(SS_ERR_INVALID_ARG, "ss_cosine_topk_gemm: sizes must be positive");
  if (dtype != SS_BF16 && dtype != SS_F16) return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_gemm: corpus and queries must be bf16 or fp16");
  if (k > 16) return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_gemm: k > 16 is not supported (use ss_cosine_topk_stream)");
  if (dim % 8 != 0 || (reinterpret_cast<uintptr_t>(corpus) & 15) || (reinterpret_cast<uintptr_t>(queries) & 15))
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_gemm: rows must be 16-byte multiples and 16-byte aligned");
  if (n_rows > 0x7FFFFFFFll || static_cast<uint64_t>(index_base) + static_cast<uint64_t>(n_rows) > 0xFFFFFFFFull)
    return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_gemm: row indices must fit in 32 bits");
  if (workspace_bytes < ss_cosine_topk_gemm_workspace_bytes(n_rows, dim, n_queries, k))
    return fail(SS_ERR_WORKSPACE, "ss_cosine_topk_gemm: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  unsigned char* ws = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  float* inv_c = reinterpret_cast<float*>(ws);
  const long long n_fill = static_cast<long long>(align_up(static_cast<size_t>(n_rows), G_BN));
  ws += static_cast<size_t>(n_fill) * 4;
  float* inv_gmax = reinterpret_cast<float*>(ws);
  ws += align_up(static_cast<size_t>(n_fill) / 32 * 4, 256);
  float* inv_q = reinterpret_cast<float*>(ws);
  ws += align_up(static_cast<size_t>(n_queries) * 4, 256);
  uint64_t* gtop = reinterpret_cast<uint64_t*>(ws);
  const size_t gtop_bytes = align_up(static_cast<size_t>(n_queries) * k * 8, 256);
  int* progress = reinterpret_cast<int*>(ws + gtop_bytes);  // 1024 ints: one per CTA of the grid (<= 2 * workers)
  SS_CUDA_CHECK(cudaMemsetAsync(gtop, 0, gtop_bytes + 4096, st));

  cudaError_t e = cudaSuccess;
  if (dtype == SS_BF16) {
    if (!corpus_norms_valid) e = launch_inv_norms_vec<__nv_bfloat16>(corpus, n_rows, dim, 1.0f, inv_c, n_fill, st);
    if (e == cudaSuccess) e = launch_inv_norms_vec<__nv_bfloat16>(queries, n_queries, dim, 1.0f, inv_q, 0, st);
  } else {
    if (!corpus_norms_valid) e = launch_inv_norms_vec<__half>(corpus, n_rows, dim, 1.0f, inv_c, n_fill, st);
    if (e == cudaSuccess) e = launch_inv_norms_vec<__half>(queries, n_queries, dim, 1.0f, inv_q, 0, st);
  }
  if (e == cudaSuccess && !corpus_norms_valid) {
    const long long n_groups = n_fill / 32;
    const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((n_groups + 7) / 8, static_cast<long long>(sm_count()) * 8)));
    inv_group_max_kernel<<<blocks, 256, 0, st>>>(inv_c, n_groups, inv_gmax);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return cuda_fail(e, "row_inv_norms_vec launch");

  const int cg = gemm_cta_group(n_queries);
  CUtensorMap tmap_q, tmap_c;
  if (!make_tmap_rows(&tmap_q, queries, dtype, n_queries, dim, G_BM) || !make_tmap_rows(&tmap_c, corpus, dtype, n_rows, dim, G_BN / cg))
    return fail(SS_ERR_CUDA, "ss_cosine_topk_gemm: cuTensorMapEncodeTiled failed");

  const GemmPlan g = make_gemm_plan(n_rows, n_queries);
  GemmParams p;
  p.n_queries = n_queries;
  p.n_rows = n_rows;
  p.dim = dim;
  p.k = k;
  p.index_base = index_base;
  p.n_qb = g.n_qb;
  p.n_tiles = g.n_tiles;
  p.tiles_per_chunk = g.tiles_per_chunk;
  p.n_chunks = g.n_chunks;
  p.warm_chunks = g.warm_chunks;
  p.n_units = static_cast<long long>(g.n_chunks) * g.n_qb;
  p.inv_c = inv_c;
  p.inv_gmax = inv_gmax;
  p.inv_q = inv_q;
  p.gtop = gtop;
  static const int max_lead = getenv("SS_GEMM_MAX_LEAD") ? atoi(getenv("SS_GEMM_MAX_LEAD")) : 2;
  p.max_lead = max_lead;
  p.progress = (max_lead > 0 && sm_count() <= 1024) ? progress : nullptr;
  const uint32_t idesc = make_idesc(dtype == SS_BF16 ? 1 : 0, G_BM * cg, G_BN);
  const size_t per_stage = cg == 2 ? GemmCfg<2>::kStageBytes : GemmCfg<1>::kStageBytes;
  const size_t fixed = 1024 /*alignment slack*/ + 256 /*barriers*/ + 2 * G_BN * 4 + 2 * (G_BN / 32) * 4 + static_cast<size_t>(k) * G_EPI_THREADS * sizeof(ScoreIdx);
  p.stages = static_cast<int>(std::min<size_t>(cg == 2 ? GemmCfg<2>::kMaxStages : GemmCfg<1>::kMaxStages, (smem_optin() - fixed) / per_stage));
  if (p.stages < 2) return fail(SS_ERR_UNSUPPORTED, "ss_cosine_topk_gemm: not enough shared memory");
  const size_t smem = fixed + static_cast<size_t>(p.stages) * per_stage;
  const int workers = static_cast<int>(std::max<long long>(1, std::min<long long>(sm_count() / cg, p.n_units)));
  if (cg == 1) {
    e = cudaFuncSetAttribute(cosine_topk_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) {
      ProfileScope prof(st);
      cosine_topk_gemm_kernel<1><<<workers, G_THREADS, smem, st>>>(tmap_q, tmap_c, p, idesc);
      e = cudaGetLastError();
    }
  } else {
    e = cudaFuncSetAttribute(cosine_topk_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) {
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(static_cast<unsigned int>(workers * 2));
      lc.blockDim = dim3(G_THREADS);
      lc.dynamicSmemBytes = smem;
      lc.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;  // CTA pair: two SMs of one TPC
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      lc.attrs = attr;
      lc.numAttrs = 1;
      ProfileScope prof(st);
      e = cudaLaunchKernelEx(&lc, cosine_topk_gemm_kernel<2>, tmap_q, tmap_c, p, idesc);
    }
  }
  if (e != cudaSuccess) return cuda_fail(e, "cosine_topk_gemm launch");
  const long long n_keys = static_cast<long long>(n_queries) * k;
  decode_keys_kernel<<<static_cast<unsigned int>((n_keys + 255) / 256), 256, 0, st>>>(gtop, n_keys, out_keys, out_scores,
                                                                                     reinterpret_cast<long long*>(out_indices));
  SS_CUDA_CHECK(cudaGetLastError());
  return SS_OK;
}

extern "C" int ss_cosine_topk_gemm(const void* corpus, int64_t n_rows, int dim, int dtype, const void* queries, int n_queries,
                                   int k, uint32_t index_base, void* workspace, size_t workspace_bytes, uint64_t* out_keys,
                                   float* out_scores, int64_t* out_indices, void* stream) {
  return cosine_topk_gemm_impl(corpus, n_rows, dim, dtype, queries, n_queries, k, index_base, workspace, workspace_bytes, false,
                               out_keys, out_scores, out_indices, stream);
}

extern "C" int ss_cosine_topk_gemm_resident(const void* corpus, int64_t n_rows, int dim, int dtype, const void* queries,
                                            int n_queries, int k, uint32_t index_base, void* workspace, size_t workspace_bytes,
                                            int corpus_norms_valid, uint64_t* out_keys, float* out_scores, int64_t* out_indices,
                                            void* stream) {
  return cosine_topk_gemm_impl(corpus, n_rows, dim, dtype, queries, n_queries, k, index_base, workspace, workspace_bytes,
                               corpus_norms_valid != 0, out_keys, out_scores, out_indices, stream);
}
