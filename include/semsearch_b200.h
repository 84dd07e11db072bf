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

/* ---- K6: k-way merge of best-first key lists ------------------------------------------------
 * Merges n_lists sorted lists per query (per-CTA partials, or per-GPU results after an NCCL
 * all-gather) into the global top k_out.  key(q, p, j) = keys_in[q*query_stride + p*list_stride + j].
 * No reference counterpart (the reference is single-process); it is the distributed form of
 * np.argsort(-scores)[:k] at Tool/rank_chunks_optimized.py:225. */
int ss_topk_merge(const uint64_t* keys_in, int n_lists, int n_queries, int k_in,
                  int64_t query_stride, int64_t list_stride, int k_out,
                  uint64_t* out_keys, float* out_scores, int64_t* out_indices, void* stream);

/* Row inverse L2 norms, 1/sqrt(sum x^2) with zero rows -> zero_value (1.0 reproduces sklearn,
 * Tool/rank_chunks_optimized.py:216; 1e9 reproduces norms[norms==0]=1e-9 at
 * Method/semantic_common.py:158-160). */
int ss_row_inv_norms(const void* rows, int64_t n_rows, int dim, int dtype, float zero_value,
                     float* out_inv_norms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMSEARCH_B200_H_ */
