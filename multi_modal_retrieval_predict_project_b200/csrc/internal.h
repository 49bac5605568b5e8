// Internal (non-ABI) declarations shared by the translation units of libmmr_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "common.cuh"

namespace mmr {

// Grow-only device buffer owned by a handle (never shrinks; freed with the handle).
struct DeviceBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  void release();
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

// Multi-GPU exchange (exchange.cu): base address of every rank's region as mapped into THIS process, and the
// description of where a select kernel puts a query's top-k when the result goes straight to the rank that
// owns the query (list-major layout [source rank][local query][kp], see exchange.cu).
constexpr int kMaxWorld = 16;
struct PeerTable {
  uint8_t* base[kMaxWorld];
};
struct PeerSink {
  PeerTable peers;
  size_t off_scores, off_rows;  // byte offsets of the fp32 score / int64 row lists inside every region
  int per;                      // queries per owner rank (owner of q = q / per)
  int kp;                       // padded list length (multiple of 4, >= k)
  int my_rank;                  // this rank = the list index at the owner
};

bool is_device_ptr(const void* p);
int elem_size(int dtype);

// RAII device guard
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev);
  ~DeviceGuard();
};

// Stages a possibly-host input on the device (async copy on `stream`) using `buf`.
int stage_in(const void* src, size_t bytes, DeviceBuf& buf, cudaStream_t stream, const void** dev_out);

// ------------------------------------------------------------------------------------------
// kernel launchers (one per .cu)
// ------------------------------------------------------------------------------------------

// ingest.cu: rows (n, d) of dtype_in -> (n, d_pad) of dtype_store zero-padded + fp32 inverse norms.
// src/dst are device pointers.  dst may be nullptr (norms only; src already in storage layout).
int launch_ingest(const void* src, int dtype_in, int64_t n, int d, int64_t src_ld, void* dst,
                  int dtype_store, int d_pad, float* inv_norm, cudaStream_t stream);
// ingest.cu: gather rows by global id -> fp32 (m, d); outside [row_offset, row_offset+n) -> zeros.
int launch_gather_rows(const void* emb, int dtype_store, int64_t n, int d, int d_pad, int64_t row_offset,
                       const int64_t* rows, int64_t m, float* out, cudaStream_t stream);

// scan_topk.cu: HBM-streaming scan.  q_f32 (b, d_pad) fp32 raw queries, q_inv (b) inverse norms.
// Writes per-(query, CTA) sorted partial lists of kp = `*kp_out` keys into `partial`
// ((b, n_parts, kp) uint64, padding key 0).  n_parts/kp are returned for the select step.
struct ScanPlan {
  int n_parts;   // CTAs along the gallery
  int kp;        // keys per partial list (>= k)
  int warps;     // warps per CTA
  int qb;        // queries per CTA pass
  size_t smem;   // dynamic shared memory per CTA
  size_t partial_bytes;
};
int plan_scan(int64_t n, int d_pad, int dtype_store, int b, int k, int num_sms, ScanPlan* plan);
int launch_scan(const void* emb, int dtype_store, const float* inv_norm, int64_t n, int d_pad,
                const float* q_f32, const float* q_inv, int b, int k, const int64_t* exclude_local,
                const ScanPlan& plan, uint64_t* partial, cudaStream_t stream);

// select.cu: per query, top `k_out` of `n_parts * kp` keys -> (score, global row), best first.
// `sink` != nullptr: the lists are stored into the owner ranks' exchange regions instead of out_scores /
// out_rows (which then serve as staging when the fused kernel does not cover the shape).
int launch_select_keys(const uint64_t* keys, int b, int64_t keys_per_query, int k_out, int64_t row_offset,
                       float* out_scores, int64_t* out_rows, cudaStream_t stream, const PeerSink* sink = nullptr);
// select.cu: merge (n_lists, b, k_in) (score,row) lists -> (b, k_out); out_src optional.
int launch_merge_lists(const float* scores, const int64_t* rows, int n_lists, int b, int k_in,
                       int64_t scores_list_stride, int64_t rows_list_stride, int k_out, float* out_scores,
                       int64_t* out_rows, int32_t* out_src, cudaStream_t stream);

// select.cu: payload gather through a merge's source positions; reranked ids + combined score.
int launch_gather_payload(const float* payload, int64_t list_stride, const int32_t* src, int b, int k_in, int k_out,
                          float* out, cudaStream_t stream);
int launch_apply_order(const int64_t* rows, const int32_t* order, const double* scores4, int b, int k, int keep,
                       int64_t* out_rows, double* out_final, cudaStream_t stream);

// gemm_topk.cu: tcgen05 GEMM + fused top-K (bf16 storage).
// Per-handle tuning (mmr_index_tune): the defaults are the production choice; tests pin each kernel
// instantiation explicitly instead of relying on the size heuristics.
struct GemmTune {
  int variant = MMR_GEMM_VARIANT_AUTO;  // MMR_GEMM_VARIANT_*: which start-up / pruning variant of the kernel runs
  int parts = 0;                        // > 0: force this many gallery parts (clamped to the tile count)
  int pair = MMR_GEMM_PAIR_AUTO;        // MMR_GEMM_PAIR_*: cta_group::2 CTA pairs or single CTAs
};
struct GemmPlan {
  int m_tiles;          // query tiles of 128
  int pair;             // 1 = cta_group::2 (a 2-CTA cluster works on two query tiles x one gallery tile)
  int probe;            // 1 = the short-launch (kProbe) instantiation runs
  int m_group;          // query tiles (pairs in pair mode) scheduled together: one wave = m_group x n_parts units
  int n_parts;          // gallery parts (CTAs per query tile)
  int n_lists;          // candidate lists per query (parts x epilogue warpgroups)
  int tiles_per_part;   // gallery tiles of 256 rows per part
  int cap;              // candidate capacity per list
  size_t cand_bytes;    // candidate buffer bytes
  size_t count_bytes;   // per-(query, list) counts
  size_t pub_bytes;     // per-(list, query) published pruning thresholds
};
int plan_gemm(int64_t n, int d_pad, int b, int k, int num_sms, const GemmTune& tune, GemmPlan* plan);
int launch_gemm_topk(const void* emb_bf16, const float* inv_norm, int64_t n, int d_pad, const void* q_bf16,
                     const float* q_inv, int b, int k, const int64_t* exclude_local, const GemmPlan& plan,
                     uint64_t* cand, int32_t* counts, uint32_t* tau_pub, cudaStream_t stream);
int launch_select_var(const uint64_t* cand, const int32_t* counts, int b, int n_parts, int cap, int per_part,
                      int k_out, int64_t row_offset, const int64_t* exclude_local, const uint32_t* tau_pub, int b_pad,
                      float* out_scores, int64_t* out_rows, cudaStream_t stream, const PeerSink* sink = nullptr);
// select.cu: (b, k) dense lists -> the owners' regions (fallback of the fused select + scatter)
int launch_scatter_lists(const float* scores, const int64_t* rows, int b, int k, const PeerSink& sink, cudaStream_t stream);

// rerank.cu
int launch_rerank_features(const void* emb, int dtype_store, int64_t n, int d_pad, int64_t row_offset,
                           const uint64_t* label_masks, int label_words, const float* kg, int d_kg,
                           int64_t n_rec, const float* q_emb, const float* cand_emb,
                           const int64_t* cand_rows, const int64_t* q_rec, const int64_t* cand_rec,
                           const int32_t* cand_count, int b, int k, int d, double* out_raw, uint8_t* owned,
                           const float* emb_cos_in, float* cos_out, cudaStream_t stream);
int launch_rerank_combine(const double* raw, const int32_t* cand_count, int b, int k, double alpha,
                          double beta, double gamma, int topk, int32_t* out_order, double* out_scores,
                          cudaStream_t stream);

// rerank.cu: the fused tail (features + min-max + combine + order) of a search result whose embedding
// feature is the search score; candidates' record index == their GLOBAL row id.
bool tail_supported(int k, const void* kg, int d_kg);
int launch_rerank_scored(const int64_t* rows, const float* scores, const int64_t* q_rec, const uint64_t* masks,
                         int label_words, const float* kg, int d_kg, int64_t n_rec, int b, int k, double alpha,
                         double beta, double gamma, int topk, int64_t* out_ids, double* out_fin,
                         double* out_scores4, cudaStream_t stream);

// metrics.cu
int launch_metrics(const int64_t* retrieved, const int32_t* ret_count, int q, int k_ret,
                   const int64_t* rel_indptr, const int64_t* rel_sorted, const int64_t* rel_list_len, int k,
                   const double* log2_tbl, double* out, cudaStream_t stream);
int launch_label_relevance(const uint64_t* q_masks, int64_t nq, const uint64_t* g_masks, int64_t ng,
                           int label_words, int exclude_self, uint8_t* out, cudaStream_t stream);

// eval.cu
int launch_first_relevant_rank(const void* emb, int dtype_store, const float* inv_norm, int64_t n, int d_pad,
                               const float* q_f32, const float* q_inv, int b, const uint64_t* q_masks,
                               const uint64_t* g_masks, int words, int64_t* out_rank, int64_t* out_total,
                               cudaStream_t stream);
int launch_label_ranking(const float* emb, const float* norms, int n, int d, const uint64_t* masks, int words,
                         const int32_t* topk, int n_topk, double* out, cudaStream_t stream);
int launch_diversity(const float* emb, const uint64_t* masks, const int32_t* counts, int b, int k, int d, int words,
                     double* out_emb_div, double* out_label_div, cudaStream_t stream);

}  // namespace mmr

// ------------------------------------------------------------------------------------------------
// handles (opaque in the C ABI; shared by the translation units of the library)
// ------------------------------------------------------------------------------------------------
#include <vector>

struct mmr_index {
  int device = 0;
  int num_sms = 148;
  int64_t n = 0;
  int d = 0, d_pad = 0;
  int dtype = MMR_BF16;
  int64_t row_offset = 0;
  void* emb = nullptr;
  bool owns_emb = true;
  float* inv_norm = nullptr;
  // Re-entrancy: the grow-only workspaces below are one set per handle.  `mu` serialises the host side of
  // a call; `last_done` (recorded at the end of every call on its stream) makes the NEXT call's stream
  // wait for the previous call's kernels, so two threads / streams sharing one handle never overlap on
  // the device either (calls with host outputs synchronise anyway).
  std::mutex mu;
  cudaEvent_t last_done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_last = false;
  mmr::GemmTune tune;
  int last_algo = 0, last_variant = 0, last_pair = 0, last_parts = 0, last_tiles_per_part = 0;
  mmr::DeviceBuf q_in, q_store, q_f32, q_inv, scratch, excl_in, excl_local, partial, counts, tau_pub, out_scores, out_rows;
  // live kernel timing (mmr_index_profile)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_start, ev_stop;
  size_t ev_used = 0;
};

struct mmr_rerank_tables {
  int device = 0;
  int64_t n_rec = 0;
  int label_words = 0;
  int d_kg = 0;
  uint64_t* masks = nullptr;
  float* kg = nullptr;
};

namespace mmr {
// api.cu: the body of mmr_search; with `sink` the lists go to the owner ranks' exchange regions.
int search_impl(mmr_index* ix, const void* q, int32_t b, int32_t q_dtype, int32_t k, int32_t algo,
                const int64_t* exclude_rows, float* out_scores, int64_t* out_rows, const PeerSink* sink,
                void* stream);
int check_device(int device, int* num_sms);
}  // namespace mmr

