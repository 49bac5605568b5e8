/*
 * mmr_b200.h -- C ABI of the B200-native retrieval hot path.
 *
 * Drop-in boundary for ONE path of ppddddpp/multi-modal-retrieval-predict-project:
 *   gallery similarity search -> top-K -> label/KG rerank -> P@K / Recall@K / nDCG / mAP / MRR.
 * Every entry point below names the reference interface (file:line under /root/reference)
 * it replaces.  The reference is pure Python (numpy / sklearn / heapq / pandas); its
 * "FFI" is therefore a ctypes binding, loaded next to PyTorch, which
 * supplies device memory and streams.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++ / torch types.
 *  - Every function returns an int status: 0 = MMR_OK, otherwise an MMR_E* code;
 *    mmr_last_error() returns a thread-local message.  No exception crosses the boundary.
 *  - Data pointers may be HOST or DEVICE pointers (detected with
 *    cudaPointerGetAttributes).  Host inputs are staged to the device inside the call;
 *    if any OUTPUT pointer is a host pointer the call synchronises `stream` before it
 *    returns, otherwise it is asynchronous on `stream`.
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *  - Row ids are int64 and GLOBAL: local row + the index's `row_offset` (row shards).
 *  - Ordering rule everywhere: score descending, then row id ascending (the reference's
 *    np.argsort(...)[::-1] leaves tie order unspecified).
 *  - There is NO CPU fallback: with no usable sm_100 device every call fails with
 *    MMR_ENODEV / MMR_ECUDA.
 */
#ifndef MMR_B200_H
#define MMR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMR_ABI_VERSION 2

/* status codes */
#define MMR_OK        0
#define MMR_EINVAL    1   /* bad argument (maps to ValueError / RuntimeError in the Python mirror) */
#define MMR_ECUDA     2   /* CUDA runtime / driver error */
#define MMR_ENOMEM    3   /* device allocation failed */
#define MMR_ENODEV    4   /* no sm_100 device */
#define MMR_EUNSUP    5   /* valid request the kernels do not cover (e.g. K > MMR_MAX_K) */

/* element types */
#define MMR_F32   0
#define MMR_BF16  1

/* search algorithm selection */
#define MMR_ALGO_AUTO  0   /* scan for small query batches, GEMM otherwise */
#define MMR_ALGO_SCAN  1   /* HBM-streaming scan, CUDA-core fp32 FMA (any storage dtype) */
#define MMR_ALGO_GEMM  2   /* tcgen05/TMEM bf16 GEMM with fused top-K epilogue (bf16 storage) */

/* mmr_index_create flags */
#define MMR_FLAG_BORROW 1  /* emb is a device pointer already in storage layout
                              (dtype_in == dtype_store, row stride == d_pad): use it in place */

#define MMR_MAX_K 1024

/* mmr_index_tune knobs (per handle; they select among kernel instantiations that all return the same
 * results -- the parity tests pin each of them explicitly, production code leaves them on AUTO) */
#define MMR_TUNE_GEMM_VARIANT 0  /* value: MMR_GEMM_VARIANT_* */
#define MMR_TUNE_GEMM_PARTS   1  /* value: 0 = automatic, > 0 = number of gallery parts per query tile */
#define MMR_TUNE_GEMM_PAIR    2  /* value: MMR_GEMM_PAIR_* */
#define MMR_GEMM_VARIANT_AUTO  0 /* = SHORT whenever every candidate list of the launch gets a full gallery tile */
#define MMR_GEMM_VARIANT_LONG  1 /* lists fill and compact, lockstep start-up, no pacing (round-1 form for long launches) */
#define MMR_GEMM_VARIANT_SHORT 2 /* probe pass on the first tile + paced sharers + bound-filtered final pass */
#define MMR_GEMM_PAIR_AUTO 0     /* cta_group::2 CTA pairs whenever the batch has >= 2 query tiles */
#define MMR_GEMM_PAIR_OFF  1     /* single-CTA (cta_group::1) instantiation */

typedef struct mmr_index mmr_index;                 /* one gallery row-shard resident in HBM */
typedef struct mmr_rerank_tables mmr_rerank_tables; /* label bitmasks + KG vectors in HBM   */

int         mmr_abi_version(void);
const char* mmr_last_error(void);
/* Number of kernels this library has launched in this process (all handles, all streams); the
 * benchmark reports the difference across its timed region as `gpu_launches`. */
int64_t     mmr_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Gallery.  Replaces RetrievalEngine.__init__ (Retrieval/retrieval.py:24-32): the gallery is
 * loaded once and stays resident.  `emb` is (n, d) row-major of dtype_in (host or device).
 * Stored as dtype_store with rows zero-padded to d_pad (multiple of 64 elements) plus one
 * fp32 inverse L2 norm per row (0 for a zero row, so that zero rows score 0 like sklearn's
 * normalize() zero-norm rule).  bf16 storage rounds to nearest even; the norm is taken over
 * the ROUNDED values.  `row_offset` is added to every row id this shard reports.
 * ------------------------------------------------------------------------------------- */
int mmr_index_create(mmr_index** out, const void* emb, int64_t n, int32_t d, int32_t dtype_in,
                     int32_t dtype_store, int64_t row_offset, int32_t device, int32_t flags,
                     void* stream);
int mmr_index_destroy(mmr_index* index);
int mmr_index_info(const mmr_index* index, int64_t* n, int32_t* d, int32_t* d_pad,
                   int32_t* dtype_store, int32_t* device, int64_t* row_offset, int64_t* hbm_bytes);
/* device pointers of the stored gallery (n x d_pad) and the inverse norms (n) -- for tests and
 * roofline accounting; owned by the index. */
int mmr_index_device_ptrs(const mmr_index* index, const void** emb, const float** inv_norm);
/* RetrievalEngine.get_embeddings_for_ids (Retrieval/retrieval.py:41-50): gather rows by GLOBAL
 * row id into `out` (m x d fp32); a row id outside the shard (e.g. -1 = unknown id) yields zeros. */
int mmr_index_get_rows(const mmr_index* index, const int64_t* rows, int64_t m, float* out, void* stream);

/* Live kernel timing for roofline accounting: while enabled, mmr_search brackets its dominant
 * kernel (the GEMM or the scan) with CUDA events on the caller's stream.  Calling with enable = 0
 * (or 1 again) synchronises, returns the summed duration in milliseconds and the number of timed
 * launches since the previous call (either pointer may be NULL), and resets the counters. */
int mmr_index_profile(mmr_index* index, int32_t enable, double* kernel_ms_sum, int32_t* kernel_launches);

/* Kernel selection of this handle's searches (MMR_TUNE_*); takes effect from the next mmr_search. */
int mmr_index_tune(mmr_index* index, int32_t knob, int32_t value);
/* What the most recent mmr_search on this handle ran: algo = MMR_ALGO_SCAN / MMR_ALGO_GEMM (0 before the
 * first search); for the GEMM additionally variant = MMR_GEMM_VARIANT_LONG / _SHORT, pair = 1 for
 * cta_group::2, the number of gallery parts and the gallery tiles (256 rows) per part.  Any pointer may be
 * NULL.  The benchmark labels its roofline with this instead of guessing the dispatch. */
int mmr_index_last_plan(const mmr_index* index, int32_t* algo, int32_t* variant, int32_t* pair, int32_t* n_parts,
                        int32_t* tiles_per_part);

/* ---------------------------------------------------------------------------------------
 * Exact cosine search + top-K.  Replaces the exact form cosine_similarity(Q, G) +
 * np.argsort(row)[::-1][:k] (Evaluate/retrieval_overlap.py:85,90; Retrieval/retrieval.py:128,134)
 * and stands in for DLSRetrievalEngine.retrieve (Retrieval/retrieval.py:140-244), whose greedy
 * walk is approximate.  q is (b, d) of q_dtype.  With a bf16 index the queries are rounded to
 * bf16 first.  score = (dot(q, g) * inv_norm(g)) * inv_norm(q) accumulated in fp32.
 * Outputs (b, k): best first; if the shard has fewer than k rows the tail is row = -1,
 * score = -inf.  `exclude_rows` (b global row ids, may be NULL, -1 = none) removes one row per
 * query (the np.fill_diagonal(sim, -1) of Retrieval/retrieval.py:129 when building a link graph).
 * ------------------------------------------------------------------------------------- */
int mmr_search(mmr_index* index, const void* q, int32_t b, int32_t q_dtype, int32_t k, int32_t algo,
               const int64_t* exclude_rows, float* out_scores, int64_t* out_rows, void* stream);

/* K-way merge of per-shard top-K lists (the step after the NCCL all-gather): inputs are
 * (n_lists, b, k_in) list-major -- the layout all_gather_into_tensor produces -- with row = -1
 * padding allowed; outputs (b, k_out) best first.  out_src (may be NULL) receives for every
 * output slot the flat source position list*k_in + j (or -1), so that callers can carry
 * payload (rerank features) through the merge.  Row ids must be below 2^32 - 1 (the merge packs them into
 * the low word of its ordering keys; mmr_index_create enforces the same bound on row_offset + n). */
int mmr_merge_topk(const float* scores, const int64_t* rows, int32_t n_lists, int32_t b, int32_t k_in,
                   int32_t k_out, float* out_scores, int64_t* out_rows, int32_t* out_src,
                   int32_t device, void* stream);

/* Same merge for lists that are not packed back to back: list l starts l * stride ELEMENTS after
 * the base pointer (device pointers).  Lets each rank all-gather ONE blob holding its scores, rows
 * and rerank payload, and merge straight out of the gathered buffer. */
int mmr_merge_topk_strided(const float* scores, const int64_t* rows, int32_t n_lists, int32_t b, int32_t k_in,
                           int64_t scores_list_stride, int64_t rows_list_stride, int32_t k_out,
                           float* out_scores, int64_t* out_rows, int32_t* out_src, int32_t device,
                           void* stream);

/* Payload carried through a merge (device pointers): out[q][i] = payload[list * list_stride +
 * q * k_in + j] for the source position out_src[q][i] = list * k_in + j of mmr_merge_topk*, 0 where
 * the slot is empty.  Used for the per-candidate embedding cosine each rank computed before the
 * exchange (Retrieval/reranker.py:298 evaluated by the rank that owns the row). */
int mmr_gather_payload(const float* payload, int64_t list_stride, const int32_t* src, int32_t b, int32_t k_in,
                       int32_t k_out, float* out, int32_t device, void* stream);

/* The (ids, scores) pair DLSRetrievalEngine.retrieve returns after a rerank
 * (Retrieval/retrieval.py:257-269): rows (b, k) candidate ids, order (b, keep) and scores4
 * (b, keep, 4) as written by mmr_rerank* -> out_rows (b, keep) ids in reranked order (-1 where
 * order < 0) and out_final (b, keep) combined scores.  Device pointers. */
int mmr_apply_order(const int64_t* rows, const int32_t* order, const double* scores4, int32_t b, int32_t k,
                    int32_t keep, int64_t* out_rows, double* out_final, int32_t device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU exchange over NVLink peer memory (gallery row-sharded, one process per GPU; the
 * replacement for "concatenate every shard's candidates and argsort", i.e. the distributed form
 * of Evaluate/retrieval_overlap.py:85,90, followed by Reranker.rerank, Retrieval/reranker.py:240-333).
 * Queries are split across the ranks; the owner of a query merges the G local top-K lists and reranks
 * them.  No collective library call is on the data path: lists and results move with peer stores issued
 * by the kernels on either side of the exchange.
 *   create          allocate this rank's region (sized for b_max queries, k_max <= 128 results)
 *   handle          CUDA IPC handle of the region (mmr_exchange_handle_bytes() bytes) -- exchange the
 *                   handles of all ranks out of band (torch.distributed all_gather) and pass them,
 *                   rank-major, to open
 *   search_scatter  mmr_search whose selection kernel stores every query's top-k {score, global row}
 *                   straight into the OWNER rank's region (owner of q = q / ceil(b / G)) + signal
 *   rerank          ONE kernel per owner, one CTA per owned query: wait for all ranks' lists of `step`, merge,
 *                   label / KG features + min-max + combine + order (the embedding feature, reranker.py:298,
 *                   is the search score the candidate was found with; a candidate's record index is its
 *                   GLOBAL row id; q_rec (b) device array of the queries' record indices), store the
 *                   query's (ids, combined scores) (keep = topk or k) into EVERY rank's result buffer,
 *                   signal; then wait for all ranks' slices.  *ids / *fin point at the full (b, keep) result
 *                   in this rank's region (valid until step + 2 is scattered).
 *   close_peers     unmap the peers' regions; put a barrier between this and destroy (which frees the
 *                   region the peers had mapped)
 * `step` must increase by one per round on every rank (buffers alternate by its parity), and all ranks
 * must run a step with the same (b, k).  Every device-side wait is bounded (set_timeout, default 20 s) and
 * gives up at once when a peer calls abort; a wait that gives up, or a (b, k) mismatch between ranks, makes
 * the next call on the handle (and status) return MMR_ECUDA instead of hanging the GPU.
 * ------------------------------------------------------------------------------------- */
typedef struct mmr_exchange mmr_exchange;
int mmr_exchange_create(mmr_exchange** out, int32_t rank, int32_t world, int32_t b_max, int32_t k_max,
                        int32_t device);
int mmr_exchange_handle_bytes(void);
int mmr_exchange_handle(mmr_exchange* ex, void* handle_out);
int mmr_exchange_open(mmr_exchange* ex, const void* handles);
int mmr_exchange_set_timeout(mmr_exchange* ex, int32_t milliseconds);
/* *code (may be NULL): 0 ok, 1 / 2 timed out waiting for lists / results, 3 (b, k) mismatch, 4 aborted */
int mmr_exchange_status(mmr_exchange* ex, int32_t* code);
int mmr_exchange_abort(mmr_exchange* ex);
int mmr_exchange_close_peers(mmr_exchange* ex);
int mmr_exchange_destroy(mmr_exchange* ex);
int mmr_search_scatter(mmr_index* index, mmr_exchange* ex, const void* q, int32_t b, int32_t q_dtype, int32_t k,
                       int32_t algo, uint32_t step, void* stream);
int mmr_exchange_rerank(mmr_exchange* ex, const mmr_rerank_tables* tables, const int64_t* q_rec, int32_t b,
                        int32_t k, double alpha, double beta, double gamma, int32_t topk, uint32_t step,
                        const int64_t** ids, const double** fin, void* stream);

/* ---------------------------------------------------------------------------------------
 * Rerank.  Replaces Reranker.rerank (Retrieval/reranker.py:240-333).
 * Tables (built once, Reranker.__init__/_load_kg :29-129): per record `label_words` uint64
 * words of label bits (get_record_label_set :161-179) and a d_kg fp32 KG vector
 * (get_record_kg_vec :181-220; rows already L2-normalised on the host exactly as :120 does).
 * Record index -1 = unknown record: empty label set, zero KG vector.
 * ------------------------------------------------------------------------------------- */
int mmr_rerank_tables_create(mmr_rerank_tables** out, const uint64_t* label_masks, int32_t label_words,
                             const float* kg_vecs, int32_t d_kg, int64_t n_rec, int32_t device,
                             void* stream);
int mmr_rerank_tables_destroy(mmr_rerank_tables* tables);

/* Raw feature scores of b x k candidates (reranker.py:298-319), fp64 out (b, k, 3):
 *   [0] cosine(q_emb, cand_emb)  -- safe_cos :135-142 (0 when a norm is 0), fp32 arithmetic
 *   [1] Jaccard(labels(q), labels(c)) -- jaccard_sets :145-149 (0 when both empty)
 *   [2] cosine(kg(q), kg(c))
 * q_emb: (b, d) fp32.  Candidate embeddings: `cand_emb` (b, k, d) fp32, or NULL to gather them
 * from `index` by GLOBAL row id `cand_rows` (rows outside the shard contribute cosine 0 and are
 * flagged in `owned` (b,k) uint8, may be NULL -- used for the sharded path).
 * q_rec (b) / cand_rec (b, k): record indices into the tables (-1 = unknown).
 * cand_count (b, may be NULL = k everywhere): valid candidates per query. */
int mmr_rerank_features(const mmr_index* index, const mmr_rerank_tables* tables, const float* q_emb,
                        const float* cand_emb, const int64_t* cand_rows, const int64_t* q_rec,
                        const int64_t* cand_rec, const int32_t* cand_count, int32_t b, int32_t k,
                        int32_t d, double* out_raw, uint8_t* owned, void* stream);
/* Sharded path: the fp32 cosine(q_emb, gallery row) of b x k candidates given by GLOBAL row id
 * (safe_cos, reranker.py:135-142) for the rows this shard owns (0 and owned = 0 elsewhere); each
 * rank evaluates it for its LOCAL top-K before the exchange so it can travel with the lists. */
int mmr_candidate_cosine(const mmr_index* index, const float* q_emb, const int64_t* cand_rows, int32_t b,
                         int32_t k, int32_t d, float* out_cos, uint8_t* owned, void* stream);
/* mmr_rerank with the embedding cosines supplied (b, k) fp32 instead of recomputed: label Jaccard
 * and KG cosine come from the (replicated) tables, then the same combine. */
int mmr_rerank_with_cos(const mmr_rerank_tables* tables, const float* emb_cos, const int64_t* q_rec,
                        const int64_t* cand_rec, const int32_t* cand_count, int32_t b, int32_t k,
                        double alpha, double beta, double gamma, int32_t topk, int32_t* out_order,
                        double* out_scores, int32_t device, void* stream);
/* min-max scale each feature over the query's candidates in fp64 (minmax_scale_list :152-159,
 * all-zeros when max == min), final = alpha*emb_n + beta*lab_n + gamma*kg_n (:325), order by final
 * descending then candidate position ascending (:327), keep topk (0 = all).  out_order (b, topk)
 * candidate positions (-1 pad), out_scores (b, topk, 4) = final, emb_n, lab_n, kg_n. */
int mmr_rerank_combine(const double* raw, const int32_t* cand_count, int32_t b, int32_t k, double alpha,
                       double beta, double gamma, int32_t topk, int32_t* out_order, double* out_scores,
                       int32_t device, void* stream);
/* The fused tail of a search step (single shard): rerank the search result itself.  rows / scores (b, k) as
 * written by mmr_search (k <= 128; -1 rows = padding); the embedding feature (reranker.py:298) is the search
 * score, a candidate's record index is its GLOBAL row id, q_rec (b) the queries' record indices (-1 =
 * unknown).  One kernel: label / KG features, min-max, combine, order.  Outputs what retrieve(...,
 * reranker=...) returns (Retrieval/retrieval.py:257-269): out_ids / out_final (b, keep), keep = topk or k,
 * ids in reranked order (-1 pad) + combined scores; out_scores4 (b, keep, 4), may be NULL, = final, emb_n,
 * lab_n, kg_n of Reranker.rerank's tuples.  MMR_EUNSUP for k > 128 or a KG dimension the kernel does not
 * cover (> 512 or not a multiple of 4): use mmr_rerank then. */
int mmr_rerank_scored(const mmr_rerank_tables* tables, const int64_t* rows, const float* scores,
                      const int64_t* q_rec, int32_t b, int32_t k, double alpha, double beta, double gamma,
                      int32_t topk, int64_t* out_ids, double* out_final, double* out_scores4, int32_t device,
                      void* stream);
/* features + combine in one call (single-shard path). */
int mmr_rerank(const mmr_index* index, const mmr_rerank_tables* tables, const float* q_emb,
               const float* cand_emb, const int64_t* cand_rows, const int64_t* q_rec,
               const int64_t* cand_rec, const int32_t* cand_count, int32_t b, int32_t k, int32_t d,
               double alpha, double beta, double gamma, int32_t topk, int32_t* out_order,
               double* out_scores, void* stream);

/* ---------------------------------------------------------------------------------------
 * Metrics.  Replaces Helpers/retrieval_metrics.py (precision_at_k :4-11, recall_at_k :74-79,
 * average_precision :24-38, the reciprocal rank of mean_reciprocal_rank :56-72, ndcg_at_k :81-89),
 * called per query by Evaluate/retrieval_eval.py:147-160.
 * retrieved: (q, k_ret) int64 item ids (-1 = padding past ret_count); ret_count (q, may be NULL).
 * Relevance: CSR -- rel_indptr (q+1), rel_sorted = per query sorted UNIQUE item ids;
 * rel_list_len (q, may be NULL = unique count) = len(relevant) AS PASSED (AP denominator, a JSON
 * list keeps duplicates).  k >= 1.  log2_tbl[i] = log2(i + 2) for i < max(k, 1) computed by the
 * host's libm so nDCG is bit-identical to numpy's.  out: (q, 5) fp64 =
 * [P@k, Recall@k, AP@k, RR (whole list), nDCG@k]; the mean over queries stays with the caller
 * (np.mean's pairwise summation). */
int mmr_metrics(const int64_t* retrieved, const int32_t* ret_count, int32_t q, int32_t k_ret,
                const int64_t* rel_indptr, const int64_t* rel_sorted, const int64_t* rel_list_len,
                int32_t k, const double* log2_tbl, double* out, int32_t device, void* stream);

/* Label-overlap relevance + full-ranking metrics on 64-bit label masks ("next" rows:
 * Helpers/contructGT.py:68-81 and Evaluate/retrieval_overlap.py:84-115).
 * out_relevant (nq, ng) uint8 = ((qmask & gmask) != 0) and not (exclude_self and i == j). */
int mmr_label_relevance(const uint64_t* q_masks, int64_t nq, const uint64_t* g_masks, int64_t ng,
                        int32_t label_words, int32_t exclude_self, uint8_t* out_relevant,
                        int32_t device, void* stream);

/* Full-ranking reduction of compute_ranking_metrics (Evaluate/retrieval_overlap.py:89-100): for every
 * query the 1-based rank of the best-scoring RELEVANT gallery row of this shard (relevant = label
 * masks overlap; 0 if none) and the number of relevant rows (:103-112) -- two sweeps over the gallery,
 * no (Q, N) matrix, no sort.  g_masks is (n, label_words) for this shard's rows. */
int mmr_first_relevant_rank(mmr_index* index, const void* q, int32_t b, int32_t q_dtype, const uint64_t* q_masks,
                            const uint64_t* g_masks, int32_t label_words, int64_t* out_rank,
                            int64_t* out_total, void* stream);

/* evaluate_label_attention's retrieval metrics (Trainner/train_label_attention.py:106-125) over n record
 * embeddings (n, d) fp32 with their L2 norms (np.linalg.norm, :108) and label bit masks: per record i,
 * over the full ranking of all n items by cosine dot / (norm_i * norm_j) (self included with label 0;
 * relevant = labels share a positive), out[i][0] = sklearn average_precision_score and out[i][1 + t] =
 * mean relevance of the top topk[t] items ("recall@k" in the reference).  n_topk <= 8.  The caller
 * averages over i (np.mean, :123-124). */
int mmr_label_ranking_eval(const float* emb, const float* norms, int32_t n, int32_t d, const uint64_t* label_masks,
                           int32_t label_words, const int32_t* topk, int32_t n_topk, double* out, int32_t device,
                           void* stream);

/* Result-set diversity (Evaluate/retrieval_diversity_compute.py:171-194): emb (b, k, d) fp32 ->
 * 1 - mean pairwise cosine (compute_embedding_diversity; 0 for fewer than 2 items) and label masks
 * (b, k, label_words) -> |union of labels| / mean label count over the items that have labels
 * (compute_label_diversity_from_labels; 0 if none).  Either half may be omitted (NULL). */
int mmr_result_diversity(const float* emb, const uint64_t* label_masks, const int32_t* counts, int32_t b,
                         int32_t k, int32_t d, int32_t label_words, double* out_emb_div,
                         double* out_label_div, int32_t device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMR_B200_H */
