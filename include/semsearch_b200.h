/*
 * semsearch_b200.h — C ABI of libsemsearch_b200.so (sm_100a CUDA kernels for the dense
 * embedding-similarity hot path of Mineooo1405/SemanticSearch).
 *
 * The reference has no FFI: its seam is Python module attributes.  Each entry
 * point below replaces one reference arithmetic site; the Python host layer
 * (semanticsearch_b200/) binds them with ctypes and mirrors the reference's call signatures.
 * Reference citations are relative to the reference repository root.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream); no entry point synchronises the host;
 *   - functions return 0 on success, a negative ss_status_t otherwise; ss_last_error() returns
 *     a thread-local, human-readable message for the last failure;
 *   - no hidden device allocation: scratch memory is a caller-provided workspace whose size
 *     comes from the matching *_workspace_bytes() call;
 *   - there is no CPU fallback anywhere.
 *
 * Top-k results are exchanged as packed 64-bit "keys": high 32 bits = the fp32 score mapped
 * to an order-preserving unsigned integer, low 32 bits = 0xFFFFFFFF - global_row_index.
 * Larger key == better candidate; ties in score resolve to the LOWER row index, on any
 * number of GPUs.  Key 0 denotes an empty slot.
 */
#ifndef SEMSEARCH_B200_H_
#define SEMSEARCH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { SS_F32 = 0, SS_BF16 = 1, SS_F16 = 2 } ss_dtype_t;

typedef enum {
  SS_OK = 0,
  SS_ERR_INVALID_ARG = -1,
  SS_ERR_CUDA = -2,
  SS_ERR_WORKSPACE = -3,
  SS_ERR_UNSUPPORTED = -4
} ss_status_t;

/* Library / device introspection (no reference counterpart). */
int ss_version(void);
const char* ss_last_error(void);
int ss_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_per_block_optin);

/* Dominant-kernel timing for bench.py's roofline figure: between begin and end each launch of a
 * path's dominant kernel is bracketed by a CUDA event pair on the launching stream (no host
 * synchronisation until ss_profile_end, which returns one duration in ms per launch). */
int ss_profile_begin(int max_records);
int ss_profile_end(float* ms_out_host, int capacity, int* n_recorded, int* n_dropped);

/* ---- K1: streaming cosine + fused top-k, small query batches --------------------------------
 * Replaces sklearn cosine_similarity(query 1xd, chunks Nxd)[0] followed by np.argsort(-s)
 * (Tool/rank_chunks_optimized.py:215-216, :225-235).  Corpus rows are L2-normalised on the fly
 * (norm 0 -> 1, sklearn's rule), the corpus is read exactly once per group of up to 8 queries,
 * and the score matrix is never written to memory.
 * corpus:  n_rows x dim, row-major, dtype corpus_dtype.   queries: n_queries x dim, query_dtype.
 * outputs (each may be NULL): out_keys [n_queries][k] best-first; out_scores fp32 [n_queries][k];
 * out_indices int64 [n_queries][k] (index_base + local row; -1 / -inf for empty slots when
 * n_rows < k). */
size_t ss_cosine_topk_stream_workspace_bytes(int64_t n_rows, int dim, int corpus_dtype, int n_queries, int k);
int ss_cosine_topk_stream(const void* corpus, int64_t n_rows, int dim, int corpus_dtype,
                          const void* queries, int n_queries, int query_dtype,
                          int k, uint32_t index_base,
                          void* workspace, size_t workspace_bytes,
                          uint64_t* out_keys, float* out_scores, int64_t* out_indices,
                          void* stream);

/* Full score matrix and full ordering (the ranker needs every chunk's score and rank, not only the
 * top k): ss_cosine_scores writes out_all_scores[q][row] = cos(query q, corpus row) with the K1
 * streaming kernel (Tool/rank_chunks_optimized.py:215-216); ss_rank_order is np.argsort(-scores) and
 * the 1-based rank lookup of :225-235 (out_order[q][r] = row at rank r, out_rank1[q][row] = rank + 1;
 * equal scores -> lower row first).  Workspace for ss_cosine_scores: the ss_cosine_topk_stream size with k = 1. */
int ss_cosine_scores(const void* corpus, int64_t n_rows, int dim, int corpus_dtype, const void* queries, int n_queries,
                     int query_dtype, void* workspace, size_t workspace_bytes, float* out_all_scores, void* stream);
size_t ss_rank_order_workspace_bytes(int n_queries, int64_t n);
int ss_rank_order(const float* scores, int n_queries, int64_t n, void* workspace, size_t workspace_bytes,
                  int32_t* out_order, int32_t* out_rank1, void* stream);

/* ---- K9: cosine + top-k for many queries over a small (L2-resident) corpus ------------------------
 * Same contract as ss_cosine_topk_stream (any dtype, any dim, k <= 4096) for the reference's own scale
 * (BASELINE.json config 1: 100 queries x 10 000 chunks x 384 fp32): a 64 x 64-tile fp32 score kernel writes
 * the n_queries x n_rows matrix into the workspace, then one CTA per query selects its top k exactly
 * (3-pass radix select, lowest-index ties, bitonic sort of the k keys).  n_rows <= 2^24. */
size_t ss_cosine_topk_small_workspace_bytes(int64_t n_rows, int n_queries);
int ss_cosine_topk_small(const void* corpus, int64_t n_rows, int dim, int corpus_dtype,
                         const void* queries, int n_queries, int query_dtype, int k, uint32_t index_base,
                         void* workspace, size_t workspace_bytes,
                         uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);

/* ---- K8: segmented ranking of query groups (cosine, ranks, reciprocal-rank fusion, percentiles) ----
 * For every group g (chunk rows offsets[g]..offsets[g+1] of the fp32 matrix `chunks`, query g of `queries`), in
 * one launch: out_cos = cosine_similarity(q, chunks)[0] (Tool/rank_chunks_optimized.py:215-216);
 * out_rank_cos / out_rank_bm25 = the 1-based rank lookups of np.argsort(-scores) (:225-235; bm25_scores are
 * computed by the caller, NULL skips the lexical term); out_rrf = 1/(k_rrf + rank_cos) + 1/(k_rrf + rank_bm25)
 * in fp64 (:238-239); out_order = the group's local row indices by fused score, best first (:250);
 * out_thr[g] = {np.percentile(rrf, upper), np.percentile(rrf, lower)} (:518-519).  Equal scores rank
 * lower-row-first.  max_group_rows <= 8192. */
int ss_segmented_rank_rrf(const float* chunks, int dim, const int32_t* offsets, int n_groups, int max_group_rows,
                          const float* queries, const float* bm25_scores, double k_rrf,
                          double upper_percentile, double lower_percentile,
                          float* out_cos, int32_t* out_rank_cos, int32_t* out_rank_bm25, double* out_rrf,
                          int32_t* out_order, double* out_thr, void* stream);

/* ---- K2: tensor-core cosine + fused top-k, large query batches ---------------------------------
 * Same contract as ss_cosine_topk_stream for bf16/fp16 corpora and queries of the same dtype, k <= 16,
 * dim % 8 == 0: D = Q C^T on tcgen05 tensor cores (TMA-fed shared-memory operands, fp32
 * accumulators in TMEM), corpus/query inverse norms applied and the top-k selected in the
 * accumulator epilogue; the B x N score matrix is never materialised.  Replaces the same reference
 * lines (Tool/rank_chunks_optimized.py:215-216,225-235) when many queries are ranked at once. */
size_t ss_cosine_topk_gemm_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k);
int ss_cosine_topk_gemm(const void* corpus, int64_t n_rows, int dim, int dtype,
                        const void* queries, int n_queries, int k, uint32_t index_base,
                        void* workspace, size_t workspace_bytes,
                        uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);

/* ss_cosine_topk_gemm for a RESIDENT corpus: with corpus_norms_valid != 0 the head of `workspace` still holds the
 * inverse row norms an earlier call (valid or not) computed for this same corpus into this same workspace, and the
 * norm pre-pass over the corpus is skipped.  The caller owns that guarantee (same corpus bytes, same workspace
 * pointer, no other call used the workspace in between). */
int ss_cosine_topk_gemm_resident(const void* corpus, int64_t n_rows, int dim, int dtype, const void* queries, int n_queries,
                                 int k, uint32_t index_base, void* workspace, size_t workspace_bytes, int corpus_norms_valid,
                                 uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);

/* ---- K7: tensor-core streaming cosine + fused top-k, medium query batches -----------------------
 * Same contract as ss_cosine_topk_gemm (bf16/fp16, one dtype for corpus and queries, dim % 8 == 0)
 * with k <= 1024: the corpus tile is the MMA M operand and up to 64 queries stay resident in shared
 * memory as the N operand, corpus norms are accumulated from the shared-memory tiles the MMA reads, so
 * the corpus streams from HBM exactly once per group of 64 queries; survivors of a per-query
 * threshold go to shared-memory candidate pools that are bitonic-sorted back to k when nearly full.
 * Replaces Tool/rank_chunks_optimized.py:215-216,225-235 for batches too large for K1's FMA budget
 * and too small for K2's 128-query tiles (BASELINE.json config 5: 16 queries, top-100, fp16). */
size_t ss_cosine_topk_tcstream_workspace_bytes(int64_t n_rows, int dim, int n_queries, int k);
int ss_cosine_topk_tcstream(const void* corpus, int64_t n_rows, int dim, int dtype,
                            const void* queries, int n_queries, int k, uint32_t index_base,
                            void* workspace, size_t workspace_bytes,
                            uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);

/* ---- K6: k-way merge of best-first key lists ------------------------------------------------
 * Merges n_lists sorted lists per query (per-CTA partials, or per-GPU results after an NCCL
 * all-gather) into the global top k_out.  key(q, p, j) = keys_in[q*query_stride + p*list_stride + j].
 * No reference counterpart (the reference is single-process); it is the distributed form of
 * np.argsort(-scores)[:k] at Tool/rank_chunks_optimized.py:225. */
int ss_topk_merge(const uint64_t* keys_in, int n_lists, int n_queries, int k_in,
                  int64_t query_stride, int64_t list_stride, int k_out,
                  uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);

/* ---- K6p: top-k key exchange over NVLink peer memory, fused with the merge -----------------------
 * One box, one process per GPU.  Every rank allocates an exchange buffer (ss_peer_alloc: cudaMalloc +
 * CUDA IPC handle), the 64-byte handles travel through the host (e.g. torch.distributed
 * all_gather_object), and every rank maps its peers' buffers (ss_peer_open).  ss_topk_peer_exchange_merge
 * then replaces "NCCL all-gather of B x k keys + ss_topk_merge": a push kernel stores this rank's keys
 * into slot [rank] of every rank's buffer through the peer mappings and raises a per-rank flag, a merge
 * kernel (one CTA per query) waits on its own buffer's flags and merges the `world` lists.
 * peer_bases_device: DEVICE array of `world` pointers (entry r = rank r's buffer as mapped in this
 * process, entry `rank` = the local allocation).  seq: 1, 2, 3, ... identical on every rank for the same
 * search.  All ranks must call in lock-step (the wait is bounded; see ss_peer_status). */
size_t ss_peer_buffer_bytes(int world, int max_queries, int k);
int ss_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle_out64);
int ss_peer_open(const unsigned char* handle64, void** dev_ptr_out);
int ss_peer_close(void* dev_ptr);
int ss_peer_free(void* dev_ptr);
int ss_topk_peer_exchange_merge(const uint64_t* local_keys, int n_queries, int k, int rank, int world,
                                void* const* peer_bases_device, int max_queries, uint32_t seq,
                                uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);
/* The same exchange with the sequence number kept on the device (in `local_buffer`, this rank's own allocation): the
 * launch parameters never change from one search to the next, so the two kernels can be captured in a CUDA graph
 * together with the local search and replayed (every rank replays in lock-step).  A merge that gives up waiting for a
 * peer (after many seconds) does not trap: it records 1 + the missing rank in the buffer's status word, which
 * ss_peer_status copies to the host (0 = healthy; synchronises). */
int ss_topk_peer_exchange_merge_auto(const uint64_t* local_keys, int n_queries, int k, int rank, int world,
                                     void* const* peer_bases_device, int max_queries, void* local_buffer,
                                     uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);
int ss_peer_status(const void* local_buffer, int* status_out_host);

/* Row inverse L2 norms, 1/sqrt(sum x^2) with zero rows -> zero_value (1.0 reproduces sklearn,
 * Tool/rank_chunks_optimized.py:216; 1e9 reproduces norms[norms==0]=1e-9 at
 * Method/semantic_common.py:158-160). */
int ss_row_inv_norms(const void* rows, int64_t n_rows, int dim, int dtype, float zero_value,
                     float* out_inv_norms, void* stream);

/* ---- K3: segmented sentence x sentence similarity matrices ------------------------------------
 * For every document d (rows offsets[d]..offsets[d+1] of the fp32 matrix `rows`), S_d = En En^T with
 * rows L2-normalised on the fly (zero rows -> zero row/column), written as a dense n_d x n_d fp32
 * block at out_S + s_offsets[d].  Replaces the arithmetic of create_similarity_matrix
 * (Method/semantic_common.py:158-164,186-191), Method/Semantic_Splitter_Optimized.py:169 and
 * data_process/simple_chunk_controller.py:614,682,743.  ss_segmented_plan_host is pure host index
 * arithmetic (HOST pointers): s_offsets = prefix sums of n^2, tile_prefix = prefix sums of
 * T(T+1)/2 with T = ceil(n/64); the caller uploads both. */
int ss_segmented_plan_host(const int32_t* offsets_host, int n_docs, int64_t* s_offsets_host,
                           int32_t* tile_prefix_host, int64_t* total_tiles, int32_t* max_doc_rows);
int ss_segmented_simmatrix(const float* rows, int dim, const int32_t* offsets, const int64_t* s_offsets,
                           const int32_t* tile_prefix, int n_docs, int64_t total_tiles, float* out_S, void* stream);

/* K3 on the tensor cores (same outputs, same reference lines as ss_segmented_simmatrix): 128 x 128
 * upper-triangular tiles, every fp32 operand split on the fly into tf32 hi + lo and multiplied as
 * hi.hi + hi.lo + lo.hi with tcgen05.mma kind::tf32 (fp32 accumulation in TMEM), so |S - S_fp32| stays
 * ~1e-6, inside the 1e-5 parity bound that plain TF32 would miss by two orders of magnitude.  Needs
 * dim % 4 == 0 and 16-byte aligned rows.  The two cross terms run as fp16 MMAs on per-row power-of-two scaled copies (the
 * scale comes from the row's first 32 elements); out_range_flag (device int32, nullable, caller-zeroed) is set to 1 when
 * a later element of some row exceeded that scale by more than 2^14 and saturated — such a batch should be recomputed
 * with ss_segmented_simmatrix (embeddings never do this; rows scaled over 12 orders of magnitude are fine).
 * ss_segmented_plan128_host (HOST pointers) lists the work units
 * {doc, tile row, tile column, 0} (int32 x 4 each); call it with units_host = NULL to size the table. */
int ss_segmented_plan128_host(const int32_t* offsets_host, int n_docs, int32_t* units_host, int64_t capacity_units,
                              int64_t* total_units);
int ss_segmented_simmatrix_tc(const float* rows, int64_t total_rows, int dim, const int32_t* offsets,
                              const int64_t* s_offsets, const int32_t* units, int64_t n_units, float* out_S,
                              int32_t* out_range_flag, void* stream);

/* ---- K4: semantic-grouping threshold pass ------------------------------------------------------
 * Per document, from S (layout of K3): sim_sharp = sigmoid(((S-mu)/sigma)/tau) in fp32 with zero
 * diagonal (Method/Semantic_Grouping_Optimized.py:100-113), centrality (:115), the quantile
 * thresholds of the positive entries q80/q65/q60 and 0.1*std (:351-355,458-462,534-537,559-561)
 * and each row's top-(k_eff+1) neighbours, value-descending/index-ascending (:270-283; the >= floor
 * filter and max-symmetrisation are applied by the caller).
 * out_doc_stats[d] = {mu, sigma, q80, q65, q60, 0.1*std(pos), count(pos), k}; out_knn_* are
 * [total_rows][33] (-1 / 0 padded).  knn_mode: 0 = auto k (:347), >0 = explicit knn_k (<= 32),
 * -1 = max(5, min(20, n-1)) (:349).  s_is_symmetric != 0 promises S == S^T bit for bit (true for the output of
 * ss_segmented_simmatrix*): the order statistics and moments of the positive values are then taken over the strict
 * upper triangle, which holds every value of the full multiset exactly once more. */
int ss_group_threshold_pass(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, float tau,
                            int knn_mode, int s_is_symmetric, float* out_sharp, double* out_centrality, double* out_doc_stats,
                            int32_t* out_knn_idx, float* out_knn_val, void* stream);

/* ---- §8f-2: block sums of sim_sharp over cluster member lists, co-association of a label sweep ---------
 * Device form of the reference's `_mean_between` / `_mean_within` double loops and of the per-(sentence, cluster)
 * means of the reassignment pass (Method/Semantic_Grouping_Optimized.py:118-130,566-588): for every document d of a
 * batch (sim_sharp packed like K4's output), with groups group_prefix[d]..group_prefix[d+1] and the members
 * (document-local row indices, repeats allowed — the reference's lists are multisets) of group g at
 * members[member_prefix[g]..member_prefix[g+1]]:
 *   out_rowsum[rowsum_offsets[d] + x*G + g] = sum_{y in g} sim_sharp[x][y]         (float64, list order)
 *   out_block [block_offsets[d]  + a*G + b] = sum_{x in a} rowsum[x][b]            (float64, list order)
 * so mean_between(A,B) = block[A][B]/(|A||B|), mean_within(A) = block[A][A]/(|A|(|A|-1)) (sim_sharp is symmetric with
 * a zero diagonal) and the reassignment mean of x against g = rowsum[x][g]/|g|.  row_doc[total_rows] = document of
 * each row.  Summation order is fixed, results are reproducible.
 * ss_group_coassociation: C[i][j] = #{l : labels[l][i] == labels[l][j]} / n_labelings off the diagonal, 0 on it
 * (the consensus matrix of the Louvain resolution sweep, :231-241); labels = int32[n_labelings][n]. */
int ss_group_block_sums(const float* sharp, const int32_t* offsets, const int64_t* s_offsets, int n_docs, int total_rows,
                        const int32_t* row_doc, const int32_t* group_prefix, const int32_t* member_prefix,
                        const int32_t* members, const int64_t* rowsum_offsets, const int64_t* block_offsets,
                        double* out_rowsum, double* out_block, void* stream);
int ss_group_coassociation(const int32_t* labels, int n_labelings, int n, double* out_C, void* stream);

/* Similarity-distribution statistics of each document's strict upper triangle of S (layout of K3):
 * values >= 1 - eps are dropped, then out_stats[d] = {count, min, max, mean, std, p10, p25, p50, p75,
 * p80, p85, p90, p95} with numpy's float32 percentile arithmetic.  Device form of
 * analyze_similarity_distribution (Method/semantic_common.py:250-270): count == 0 means every value was
 * filtered and all fields hold max(sims) (:257-260); count == -1 means fewer than 2 rows (None). */
int ss_similarity_distribution(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, float eps,
                               double* out_stats, void* stream);

/* a12 — diameter-bounded splitting of each document's sentences, the controller's optional stage on sim = emb @ emb.T
 * (data_process/simple_chunk_controller.py:571-594,614): a span [a, b) whose diameter 1 - min_{i != j} S[i][j] exceeds
 * `threshold` is cut after its lowest adjacent similarity S[i][i+1] (first minimum) and both halves are treated the
 * same way.  out_span_ends[offsets[d] + s] = end (exclusive, document-local) of the s-th final span of document d,
 * ascending; out_n_spans[d] = their number (1 = not split); out_diameter[d] = diameter of the whole document
 * (0 for a single row).  S has the layout of K3; max_doc_rows <= 4096. */
int ss_diameter_split(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, int max_doc_rows,
                      double threshold, int32_t* out_span_ends, int32_t* out_n_spans, double* out_diameter, void* stream);

/* C99 rank transform of each document's S (layout of K3): global row+column rank
 * (Method/Semantic_Splitter_Optimized.py:189-192) or the clipped mask_size x mask_size local rank
 * (:171-186) when bit 0 of use_local_rank is set.  Bit 1 (value 2, global mode only) promises that every S equals its
 * transpose bit for bit (true for the output of ss_segmented_simmatrix*): the column ranks are then taken as the
 * transposed row ranks.  workspace_rows: int32[total_rows + total_rows / 16 + 2 * n_docs + 8] scratch (row -> document map, tile list
 * of the local mode). */
int ss_c99_rank_matrix(const float* S, const int32_t* offsets, const int64_t* s_offsets, int n_docs, int total_rows,
                       int max_doc_rows, int use_local_rank, int mask_size, int32_t* workspace_rows, float* out_R,
                       void* stream);

/* K10 — C99 divisive cut search on each document's rank matrix R (output layout of ss_c99_rank_matrix):
 * device form of the greedy loop at Method/Semantic_Splitter_Optimized.py:194-238.  Every round takes the cut
 * with the largest gain = 0.5 * (mean R[a:c,a:c] + mean R[c:b,c:b]) - mean R[a:b,a:b] over all segments [a, b)
 * and a+m <= c <= b-m (first best in segment-list order, ascending c), until no candidate is left, max_cuts is
 * reached, or (stop_by_gain != 0) the gain is below max(min_gain, 0.1 * |mean of that segment|).  Block means
 * come from a float64 summed-area table (sat_workspace: sum over documents of (n+1)^2 doubles; sat_offsets =
 * int64[n_docs] start of each table, device).  min_chunk / max_cuts: per-document int32 arrays (device) or
 * null to use the *_all scalars (max_cuts < 0 = unlimited).  out_cuts[offsets[d] + i] = i-th picked cut of
 * document d, out_n_cuts[d] = count (0 when n < 2 * min_chunk, :165-166; -1 = min_chunk < 1), out_profile
 * (nullable) [offsets[d] + i] = inside density after i cuts (D_series, :206,234).  max_doc_rows <= 4096. */
int ss_c99_divisive_cuts(const float* R, const int32_t* offsets, const int64_t* s_offsets, const int64_t* sat_offsets,
                         int n_docs, int max_doc_rows, const int32_t* min_chunk, int min_chunk_all, const int32_t* max_cuts,
                         int max_cuts_all, double min_gain, int stop_by_gain, double* sat_workspace, int32_t* out_cuts,
                         int32_t* out_n_cuts, double* out_profile, void* stream);

/* ---- K5: semantic-splitter passes over ragged documents ---------------------------------------
 * Documents are concatenated: rows = [total_rows x dim]; offsets = int32[n_docs+1] (device) CSR
 * row offsets.
 * ss_segmented_adjacent_cosine: out_adj[r] = cos(rows[r], rows[r+1]) (rows L2-normalised on the
 * fly, zero rows -> 0), one streaming pass; replaces `_embed` + the adjacent dot loop at
 * Method/Semantic_Splitter_Optimized.py:140-152,412.  The slot of each document's last row is
 * zeroed by ss_segmented_percentile.
 * ss_segmented_percentile: per document d = 1 - adj, out_thr[doc] = np.percentile(d, pct)
 * (linear interpolation, fp64), out_flags[r] = d[r] > thr (BASELINE.json config 3 breakpoints);
 * optional out_smooth[r] = median-of-3 smoothed adj (Splitter:340-356) and out_stats[doc] =
 * {median, MAD + 1e-9, P25, P75} of the smoothed series (Splitter:417-437).  Documents with a
 * single row get thr = NaN.  max_doc_rows bounds the per-document shared-memory sort (<= 8193). */
int ss_segmented_adjacent_cosine(const void* rows, int64_t n_rows, int dim, int dtype, float* out_adj, void* stream);
int ss_segmented_percentile(float* adj, const int32_t* offsets, int n_docs, int max_doc_rows, double pct,
                            double* out_thr, uint8_t* out_flags, double* out_stats, float* out_smooth, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMSEARCH_B200_H_ */
