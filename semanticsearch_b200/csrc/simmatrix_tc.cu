// K3 (tensor-core form) — segmented sentence x sentence similarity matrices S_d = En_d . En_d^T on
// the 5th-gen tensor cores with fp32-class accuracy (3xTF32 error-compensated products).
//
// Replaces the arithmetic of create_similarity_matrix (Method/semantic_common.py:158-164,186-191:
// row L2 normalisation, then torch.mm / numpy matmul in fp32), the C99 similarity at
// Method/Semantic_Splitter_Optimized.py:169 and the controller's diagnostic recomputations
// (data_process/simple_chunk_controller.py:614,682,743) for a whole ragged batch in one launch.
//
// Parity needs |S - S_ref| <= 1e-5, which rules out plain TF32/BF16 products (1e-3).  Every fp32
// operand x is split on the fly into hi = x with the low 13 mantissa bits dropped (what the tensor core
// reads of a kind::tf32 operand anyway: the raw plane IS the hi operand) and lo = x - hi (exact), and the
// product is hi.hi + hi.lo + lo.hi.  The main term runs as kind::tf32 MMAs on the raw planes; the two
// cross terms are 2^-10 of it, so fp16 copies of hi and lo (11-bit significands) carry them with a relative
// error of 2^-21 per product, 2^-19 = 1.9e-6 for a whole dot product in the worst case: they run as
// kind::f16 MMAs with K = 16 — half the tensor time and half the shared-memory operand bytes of tf32 cross
// terms (bf16 copies would be range-safe but their 8-bit significand puts the worst case at 1.5e-5, past the
// parity bound).  fp16's range is handled per row: the splitter of a row takes the exponent of the largest
// magnitude among the 32 elements of the row's first K block that is not all zero (every CTA and both operand
// roles see the same blocks, so the choice is consistent), stores hi * 2^-e and lo * 2^(12-e), and the epilogue multiplies the cross
// accumulator by 2^(e_a + e_b - 12) — exact powers of two.  A later element more than 2^15 times that
// magnitude saturates (cvt.satfinite) and raises `range_flag`; callers that validate recompute such a batch
// with the fp32 FFMA kernel.  fp32 accumulation happens in TMEM.
//
// Work unit = one 128 x 128 upper-triangular tile (ti <= tj) of one document, listed in a host-built
// unit table.  Persistent CTAs, 14 warps:
//   warp 0      TMA producer : raw fp32 row tiles (128 rows x 32 floats, SWIZZLE_128B) of the tile's
//                              row block (A) and column block (B) into a 3-stage ring
//   warps 6-13  splitters    : one tile row per thread: rewrite the raw plane in place as `hi`, write
//                              `lo` to a second plane (same swizzled offsets), accumulate the row's
//                              sum of squares, fence.proxy.async and hand the stage to the MMA warp
//   warp 1      MMA issuer   : 12 tcgen05.mma (M=128, N=16..128, K=8) per K block into a double-
//                              buffered TMEM accumulator
//   warps 2-5   epilogue     : tcgen05.ld, scale by 1/|e_i| 1/|e_j| (zero rows stay zero), write the
//                              tile and, for off-diagonal tiles, its transpose; both writes are
//                              coalesced (the transpose directly from the TMEM lane layout, the
//                              direct tile through a padded shared-memory transpose).
#include <algorithm>

#include "tc_common.cuh"

namespace ss {

constexpr int S_BM = 128;
constexpr int S_BK = 32;                    // fp32 elements per K block = one 128-byte swizzle atom
constexpr int S_PLANE = S_BM * 128;         // 16 KB: 128 rows x 128 bytes
constexpr int S_STAGE_BYTES = 4 * S_PLANE;  // A raw | B raw | A fp16 {hi | lo} | B fp16 {hi | lo}
constexpr int S_STAGES = 3;
constexpr int S_THREADS = 14 * 32;
constexpr int S_TMEM_COLS = 512;  // 2 buffers x (main accumulator | correction accumulator) x 128 columns

struct SimTcParams {
  const int* offsets;          // [n_docs + 1] row offsets
  const long long* s_offsets;  // [n_docs + 1] element offsets into out
  const int4* units;           // [n_units] {doc, ti, tj, 0}
  long long n_units;
  int dim;
  int nkb;
  float* out;
  int* range_flag;  // nullable: set to 1 when an element saturated its row's fp16 scale (see the header comment)
};

// Unit table entry {a, b, c, kind}.  kind 0: tile (ti = b, tj = c) of document a.  kind 1: a PACKED window —
// documents a .. a + b - 1 (each <= 64 rows, c rows in total, <= 128) share one diagonal tile: their rows
// are contiguous in memory, the MMA computes the whole window's Gram matrix and the epilogue keeps only
// each document's own block.  Three quarters of the reference corpus are documents of <= 32 sentences.
struct SimUnit {
  int row_base, n, ti, tj;  // packed: n = rows of the window, ti = tj = 0
  long long s_off;
  int packed, first_doc, n_docs;
};

__device__ __forceinline__ SimUnit load_unit(const SimTcParams& p, long long u) {
  const int4 e = __ldg(p.units + u);
  SimUnit r;
  r.row_base = __ldg(p.offsets + e.x);
  r.packed = e.w;
  r.first_doc = e.x;
  if (e.w) {
    r.n = e.z;
    r.n_docs = e.y;
    r.ti = 0;
    r.tj = 0;
    r.s_off = 0;
  } else {
    r.n = __ldg(p.offsets + e.x + 1) - r.row_base;
    r.n_docs = 1;
    r.s_off = __ldg(p.s_offsets + e.x);
    r.ti = e.y;
    r.tj = e.z;
  }
  return r;
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ void umma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(kSmemDescHi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

__global__ void __launch_bounds__(S_THREADS, 1)
segmented_simmatrix_tc_kernel(const __grid_constant__ CUtensorMap tmap_rows, const SimTcParams p) {
  extern __shared__ unsigned char simtc_smem_raw[];
  unsigned char* smem = simtc_smem_raw + ((1024u - (smem_u32(simtc_smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = smem;                                                       // [S_STAGES][64 KB]
  float* scratch = reinterpret_cast<float*>(tiles + S_STAGES * S_STAGE_BYTES);      // [4 warps][32][33] transpose staging
  float* inv_s = scratch + 4 * 32 * 33;                                              // [2][256]: 1/|row| of the A rows, then the B rows
  float* scale_s = inv_s + 2 * 256;                                                  // [2][256]: 2^(e - 6) of the A rows, then the B rows
  long long* row_base_s = reinterpret_cast<long long*>(scale_s + 2 * 256);           // [128] per tile row: offset of (row, window col 0)
  int* row_lo = reinterpret_cast<int*>(row_base_s + 128);                             // [128] first / one-past-last window column it may write
  int* row_hi = row_lo + 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(row_hi + 128);
  uint64_t* ready_bar = full_bar + S_STAGES;
  uint64_t* empty_bar = ready_bar + S_STAGES;
  uint64_t* tmem_full = empty_bar + S_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* norm_full = tmem_empty + 2;
  uint64_t* norm_empty = norm_full + 2;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(norm_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tmap_prefetch(&tmap_rows);
    for (int s = 0; s < S_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 8);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
      mbar_init(&norm_full[a], 8);
      mbar_init(&norm_empty[a], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_s, S_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const SimUnit un = load_unit(p, u);
        const bool diag = un.ti == un.tj;
        const int ra = un.row_base + un.ti * S_BM, rb = un.row_base + un.tj * S_BM;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          unsigned char* st = tiles + s * S_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], diag ? S_PLANE : 2 * S_PLANE);
          tma_load_2d(st, &tmap_rows, &full_bar[s], kb * S_BK, ra);
          if (!diag) tma_load_2d(st + S_PLANE, &tmap_rows, &full_bar[s], kb * S_BK, rb);
          if (++s == S_STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (all lanes run the loop, one elected lane issues) =====================
    int s = 0, acc = 0;
    uint32_t ph = 0, acc_ph = 0;
    const uint32_t tiles_lo = smem_desc_lo(smem_u32(tiles));
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const SimUnit un = load_unit(p, u);
      const bool diag = un.ti == un.tj;
      const int ncols = min(S_BM, un.n - un.tj * S_BM);
      const uint32_t idesc = make_idesc(2 /*tf32*/, S_BM, (ncols + 15) & ~15);
      const uint32_t idesc_h = make_idesc(0 /*fp16*/, S_BM, (ncols + 15) & ~15);
      mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
      tc_fence_after();
      // The tensor core truncates when it adds into the fp32 accumulator, a bias that grows with the
      // number of accumulations at full magnitude.  The hi.hi products (96 MMAs at d = 768) therefore get
      // an accumulator of their own; the small cross terms go to a second one (their truncation
      // error is ~1e-3 smaller) and the epilogue adds the two in fp32.
      const uint32_t d_main = tmem_base + static_cast<uint32_t>(acc * 2 * S_BM);
      const uint32_t d_corr = d_main + S_BM;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&ready_bar[s], ph);  // fp16 planes written and fenced by the splitters
        tc_fence_after();
        const uint32_t a_hi = tiles_lo + static_cast<uint32_t>(s) * (S_STAGE_BYTES >> 4);
        const uint32_t b_hi = diag ? a_hi : a_hi + (S_PLANE >> 4);
        const uint32_t a_cb = a_hi + (2 * S_PLANE >> 4);                 // row r: fp16 hi[0..32) | fp16 lo[0..32), both scaled
        const uint32_t b_cb = diag ? a_cb : a_hi + (3 * S_PLANE >> 4);
#pragma unroll
        for (int k = 0; k < S_BK / 8; ++k)  // K = 8 tf32 per MMA = 32 bytes = 2 descriptor units
          umma_tf32_lohi(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < S_BK / 16; ++k) {  // K = 16 fp16 per MMA = 32 bytes; the lo half of a row starts 64 bytes in
          umma_f16_lohi<1>(d_corr, a_cb + 2 * k, b_cb + 4 + 2 * k, idesc_h, (kb | k) != 0 ? 1u : 0u);  // hi . lo
          umma_f16_lohi<1>(d_corr, a_cb + 4 + 2 * k, b_cb + 2 * k, idesc_h, 1u);                        // lo . hi
        }
        umma_commit_elect<1>(&empty_bar[s]);
        if (kb == p.nkb - 1) umma_commit_elect<1>(&tmem_full[acc]);
        if (++s == S_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  } else if (warp >= 6) {
    // ===================== splitters: one tile row per thread =====================
    const int srow = (warp - 6) * 32 + lane;  // 0..127: A rows, 128..255: B rows
    const int r = srow & 127;
    const uint32_t plane_off = srow < 128 ? 0u : static_cast<uint32_t>(S_PLANE);
    int s = 0, acc = 0;
    uint32_t ph = 0, acc_ph = 0;
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int4 e = __ldg(p.units + u);
      // diagonal tiles have no separate B rows; rows past the end of the document / packed window (the tile
      // is padded with whatever follows in memory) feed only outputs that are never written, so they are
      // not split either
      bool active;
      if (e.w) {
        active = srow < 128 && r < e.z;
      } else {
        const int n_doc = __ldg(p.offsets + e.x + 1) - __ldg(p.offsets + e.x);
        const int blk = srow < 128 ? e.y : e.z;
        active = (srow < 128 || e.y != e.z) && (blk * S_BM + r < n_doc);
      }
      float ssq0 = 0.f, ssq1 = 0.f;
      float sc_hi = 1.f, sc_lo = 4096.f;  // 2^-e and 2^(12 - e) of this row, fixed at its first K block
      int row_e = 0;
      bool sat = false, have_scale = false;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        if (active) {
          unsigned char* base = tiles + s * S_STAGE_BYTES + plane_off + static_cast<uint32_t>(r) * 128u;
          unsigned char* comb = base + 2 * S_PLANE;
          const int sw = r & 7;  // SWIZZLE_128B: logical 16-byte chunk c of row r sits at chunk c ^ (r & 7)
          if (!have_scale) {
            // the scale is fixed by the first K block that holds a normal number: the all-zero blocks before it convert
            // to zeros under any scale, and every CTA that handles this row sees the same blocks in the same order
            uint32_t mx = 0u;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 w = *reinterpret_cast<const uint4*>(base + (c ^ sw) * 16);
              mx = max(max(mx, w.x & 0x7FFFFFFFu), max(max(w.y & 0x7FFFFFFFu, w.z & 0x7FFFFFFFu), w.w & 0x7FFFFFFFu));
            }
            const int be = static_cast<int>(mx >> 23);  // biased exponent of the largest magnitude (0: zeros / denormals)
            if (be != 0 && be != 255) {
              have_scale = true;
              row_e = min(max(be - 127, -100), 100);
              sc_hi = __uint_as_float(static_cast<uint32_t>(127 - row_e) << 23);
              sc_lo = __uint_as_float(static_cast<uint32_t>(127 + 12 - row_e) << 23);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // elements 8q .. 8q+7 of the row: raw chunks 2q, 2q+1 -> fp16 hi chunk q, lo chunk q + 4
            // conflict-free: 8 consecutive rows touch 8 distinct 16-byte chunks in every access
            const float4 v0 = *reinterpret_cast<const float4*>(base + ((2 * q) ^ sw) * 16);
            const float4 v1 = *reinterpret_cast<const float4*>(base + ((2 * q + 1) ^ sw) * 16);
            ssq0 = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, ssq0))));
            ssq1 = fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, ssq1))));
            const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = x[2 * e] * sc_hi, b = x[2 * e + 1] * sc_hi;
              sat |= fmaxf(fabsf(a), fabsf(b)) > 16000.f;  // lo * 2^12 reaches 4x the scaled hi: both must stay below 65504
              const float la = (x[2 * e] - tf32_trunc(x[2 * e])) * sc_lo, lb = (x[2 * e + 1] - tf32_trunc(x[2 * e + 1])) * sc_lo;
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[e]) : "f"(b), "f"(a));    // cvt packs its FIRST source into the upper half: element 2e lands in the lower one
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l[e]) : "f"(lb), "f"(la));
            }
            *reinterpret_cast<uint4*>(comb + (q ^ sw) * 16) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(comb + ((q + 4) ^ sw) * 16) = make_uint4(l[0], l[1], l[2], l[3]);
          }
          fence_proxy_async();  // generic-proxy writes above -> visible to the tensor core's async-proxy reads
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[s]);
        if (++s == S_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      const float ssq = ssq0 + ssq1;
      mbar_wait(&norm_empty[acc], acc_ph ^ 1u);
      // zero rows stay zero (reference: norm 0 -> 1e-9, and 0 / 1e-9 == 0)
      if (active) {
        inv_s[acc * 256 + srow] = ssq > 0.f ? 1.0f / sqrtf(ssq) : 0.f;
        scale_s[acc * 256 + srow] = __uint_as_float(static_cast<uint32_t>(127 + row_e - 6) << 23);  // 2^(e - 6): a * b gives 2^(e_a + e_b - 12)
        if (sat && p.range_flag != nullptr) *p.range_flag = 1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&norm_full[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  } else {
    // ===================== epilogue: one tile row (TMEM lane) per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    float* scr = scratch + (warp - 2) * 32 * 33;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const SimUnit un = load_unit(p, u);
      const bool diag = un.ti == un.tj;
      const int ncols = min(S_BM, un.n - un.tj * S_BM);
      // Per row of the tile: the window-local column range [c_lo, c_hi) it may write, the element offset of
      // (row, window column 0) for the direct tile, and offset + leading dimension for the transposed tile.
      int c_lo = 0, c_hi = 0, ld = 1;
      long long rbase = 0, mbase = 0;
      if (!un.packed) {
        const int gr = un.ti * S_BM + row;  // this thread's row inside the document
        if (gr < un.n) {
          c_hi = ncols;
          ld = un.n;
          rbase = un.s_off + static_cast<long long>(gr) * un.n + un.tj * S_BM;
          mbase = un.s_off + static_cast<long long>(un.tj) * S_BM * un.n + gr;
        }
      } else if (row < un.n) {
        int lo_d = un.first_doc, hi_d = un.first_doc + un.n_docs;  // last document with offsets[d] <= global row
        const int grow = un.row_base + row;
        while (hi_d - lo_d > 1) {
          const int mid = (lo_d + hi_d) >> 1;
          if (__ldg(p.offsets + mid) <= grow) lo_d = mid; else hi_d = mid;
        }
        const int d_lo = __ldg(p.offsets + lo_d) - un.row_base;
        const int d_n = __ldg(p.offsets + lo_d + 1) - __ldg(p.offsets + lo_d);
        const long long d_off = __ldg(p.s_offsets + lo_d);
        c_lo = d_lo;
        c_hi = d_lo + d_n;
        ld = d_n;
        rbase = d_off + static_cast<long long>(row - d_lo) * d_n - d_lo;
        mbase = d_off - static_cast<long long>(d_lo) * d_n + (row - d_lo);
      }
      row_lo[(warp - 2) * 32 + lane] = c_lo;
      row_hi[(warp - 2) * 32 + lane] = c_hi;
      row_base_s[(warp - 2) * 32 + lane] = rbase;
      __syncwarp();
      mbar_wait(&norm_full[acc], acc_ph);
      const float* inv_a = inv_s + acc * 256;
      const float* inv_b = diag ? inv_a : inv_a + 128;
      const float my_inv = inv_a[row];
      const float* scale_b = scale_s + acc * 256 + (diag ? 0 : 128);
      const float my_scale = scale_s[acc * 256 + row];
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * 2 * S_BM);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        float x[32];
        {
          uint32_t raw[32], cor[32];
          tmem_ld32(taddr + static_cast<uint32_t>(c0), raw);
          tmem_ld32(taddr + static_cast<uint32_t>(S_BM + c0), cor);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 ib = *reinterpret_cast<const float4*>(inv_b + c0 + j);
            const float4 sb = *reinterpret_cast<const float4*>(scale_b + c0 + j);
            x[j] = fmaf(__uint_as_float(cor[j]), my_scale * sb.x, __uint_as_float(raw[j])) * (my_inv * ib.x);
            x[j + 1] = fmaf(__uint_as_float(cor[j + 1]), my_scale * sb.y, __uint_as_float(raw[j + 1])) * (my_inv * ib.y);
            x[j + 2] = fmaf(__uint_as_float(cor[j + 2]), my_scale * sb.z, __uint_as_float(raw[j + 2])) * (my_inv * ib.z);
            x[j + 3] = fmaf(__uint_as_float(cor[j + 3]), my_scale * sb.w, __uint_as_float(raw[j + 3])) * (my_inv * ib.w);
          }
        }
        // transposed tile S[col][row]: lanes hold consecutive rows -> consecutive addresses.  Diagonal
        // tiles mirror their strict upper triangle so that S is bit-for-bit symmetric.
        if (c_hi > c_lo) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            if (c >= c_lo && c < c_hi && (!diag || c > row)) p.out[mbase + static_cast<long long>(c) * ld] = x[j];
          }
        }
        // direct tile S[row][col]: transpose 32 x 32 through padded shared memory so that lanes
        // hold consecutive columns
#pragma unroll
        for (int j = 0; j < 32; ++j) scr[lane * 33 + j] = x[j];
        __syncwarp();
        const int c = c0 + lane;
        if (!un.packed) {
          // one document per tile: row j of this warp starts at an affine address, no per-row table look-ups
          const int r0 = un.ti * S_BM + quad * 32;
          const int rows_valid = un.n - r0;  // rows j < rows_valid lie inside the document
          float* dst = p.out + un.s_off + static_cast<long long>(r0) * un.n + un.tj * S_BM + c;
          const bool col_ok = c < ncols;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float y = scr[j * 33 + lane];
            if (col_ok && j < rows_valid && (!diag || c >= quad * 32 + j)) dst[static_cast<long long>(j) * un.n] = y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int rj = (warp - 2) * 32 + j;
            const float y = scr[j * 33 + lane];
            if (c >= row_lo[rj] && c < row_hi[rj] && (!diag || c >= quad * 32 + j)) p.out[row_base_s[rj] + c] = y;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tmem_empty[acc]);
        mbar_arrive(&norm_empty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S_TMEM_COLS);
  }
}

}  // namespace ss

using namespace ss;

extern "C" int ss_segmented_plan128_host(const int32_t* offsets_host, int n_docs, int32_t* units_host, int64_t capacity_units,
                                         int64_t* total_units) {
  if (!offsets_host || n_docs < 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_plan128_host: bad arguments");
  constexpr int kPackMaxDoc = 64;  // documents up to this many rows are packed into shared 128-row windows
  int64_t t = 0;
  auto emit = [&](int a, int b, int c, int kind) -> bool {
    if (units_host) {
      if (t >= capacity_units) return false;
      units_host[4 * t + 0] = a;
      units_host[4 * t + 1] = b;
      units_host[4 * t + 2] = c;
      units_host[4 * t + 3] = kind;
    }
    ++t;
    return true;
  };
  int d = 0;
  while (d < n_docs) {
    const int64_t n = static_cast<int64_t>(offsets_host[d + 1]) - offsets_host[d];
    if (n < 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_plan128_host: offsets must be non-decreasing");
    if (n <= kPackMaxDoc) {
      // greedy window of consecutive small documents (their rows are contiguous)
      int d1 = d;
      int64_t rows = 0;
      while (d1 < n_docs) {
        const int64_t m = static_cast<int64_t>(offsets_host[d1 + 1]) - offsets_host[d1];
        if (m < 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_plan128_host: offsets must be non-decreasing");
        if (m > kPackMaxDoc || rows + m > S_BM) break;
        rows += m;
        ++d1;
      }
      if (rows > 0) {
        if (d1 - d == 1) {
          if (!emit(d, 0, 0, 0)) return fail(SS_ERR_WORKSPACE, "ss_segmented_plan128_host: unit table too small");
        } else if (!emit(d, d1 - d, static_cast<int>(rows), 1)) {
          return fail(SS_ERR_WORKSPACE, "ss_segmented_plan128_host: unit table too small");
        }
      }
      d = d1;
      continue;
    }
    const int T = static_cast<int>((n + S_BM - 1) / S_BM);
    for (int ti = 0; ti < T; ++ti)
      for (int tj = ti; tj < T; ++tj)
        if (!emit(d, ti, tj, 0)) return fail(SS_ERR_WORKSPACE, "ss_segmented_plan128_host: unit table too small");
    ++d;
  }
  if (total_units) *total_units = t;
  return SS_OK;
}

extern "C" int ss_segmented_simmatrix_tc(const float* rows, int64_t total_rows, int dim, const int32_t* offsets,
                                         const int64_t* s_offsets, const int32_t* units, int64_t n_units, float* out_S,
                                         int32_t* out_range_flag, void* stream) {
  if (!rows || !offsets || !s_offsets || !units || !out_S) return fail(SS_ERR_INVALID_ARG, "ss_segmented_simmatrix_tc: null pointer");
  if (dim <= 0 || total_rows <= 0 || n_units < 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_simmatrix_tc: bad sizes");
  if (n_units == 0) return SS_OK;
  if ((dim & 3) != 0 || (reinterpret_cast<uintptr_t>(rows) & 15) != 0)
    return fail(SS_ERR_UNSUPPORTED, "ss_segmented_simmatrix_tc: rows must be 16-byte multiples and 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(units) & 15) != 0) return fail(SS_ERR_INVALID_ARG, "ss_segmented_simmatrix_tc: unit table must be 16-byte aligned");
  if (total_rows > 0x7FFFFFFFll - S_BM) return fail(SS_ERR_UNSUPPORTED, "ss_segmented_simmatrix_tc: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmap;
  if (!make_tmap_rows(&tmap, rows, SS_F32, total_rows, dim, S_BM))
    return fail(SS_ERR_CUDA, "ss_segmented_simmatrix_tc: cuTensorMapEncodeTiled failed");
  SimTcParams p;
  p.offsets = offsets;
  p.s_offsets = reinterpret_cast<const long long*>(s_offsets);
  p.units = reinterpret_cast<const int4*>(units);
  p.n_units = n_units;
  p.dim = dim;
  p.nkb = (dim + S_BK - 1) / S_BK;
  p.out = out_S;
  p.range_flag = out_range_flag;
  const size_t smem = 1024 + static_cast<size_t>(S_STAGES) * S_STAGE_BYTES + 4 * 32 * 33 * 4 + 4 * 256 * 4 + 128 * (8 + 4 + 4) + 256;
  cudaError_t e = cudaFuncSetAttribute(segmented_simmatrix_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e == cudaSuccess) {
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(sm_count(), n_units)));
    ProfileScope prof(st);
    segmented_simmatrix_tc_kernel<<<grid, S_THREADS, smem, st>>>(tmap, p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return cuda_fail(e, "segmented_simmatrix_tc launch");
  return SS_OK;
}
