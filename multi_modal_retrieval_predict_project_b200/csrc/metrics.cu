// K6: retrieval metrics on device.
//
// Replaces the per-query Python loops of reference Helpers/retrieval_metrics.py
// (precision_at_k :4-11, recall_at_k :74-79 [the second definition, which shadows :13-22],
// average_precision :24-38, the reciprocal rank inside mean_reciprocal_rank :65-71, ndcg_at_k :81-89)
// as called by Evaluate/retrieval_eval.py:147-160.  One warp per query: lanes test membership of
// the retrieved ids in the query's sorted relevance list (binary search), then lane 0 replays the
// reference's sequential fp64 arithmetic in the same order, so every per-query value is
// bit-identical to the Python result.  Integer/latency-bound; reported as time only.
//
// Also the "next" rows: label-overlap relevance (Helpers/contructGT.py:68-81).
#include <math_constants.h>

#include "internal.h"

namespace mmr {
namespace {

constexpr int kMaxRet = 4096;  // retrieved list length supported per query (bitmask in smem)

__device__ __forceinline__ bool contains_sorted(const int64_t* a, int64_t n, int64_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = a[mid];
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && a[lo] == x;
}

// 4 warps per CTA, one query per warp.
__global__ void __launch_bounds__(128)
metrics_kernel(const int64_t* __restrict__ retrieved, const int32_t* __restrict__ ret_count, int nq, int k_ret,
               const int64_t* __restrict__ rel_indptr, const int64_t* __restrict__ rel_sorted,
               const int64_t* __restrict__ rel_list_len, int k, const double* __restrict__ log2_tbl,
               double* __restrict__ out) {
  __shared__ uint32_t hitmask[4][kMaxRet / 32];
  __shared__ uint32_t dupmask[4][kMaxRet / 32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 4 + warp;
  if (q >= nq) return;
  const int64_t* ret = retrieved + static_cast<int64_t>(q) * k_ret;
  int count = ret_count != nullptr ? min(ret_count[q], k_ret) : k_ret;
  const int64_t r0 = rel_indptr[q], r1 = rel_indptr[q + 1];
  const int64_t n_unique = r1 - r0;
  const int64_t list_len = rel_list_len != nullptr ? rel_list_len[q] : n_unique;
  const int kk = min(k, count);  // retrieved[:k]

  // membership + "same id appeared earlier within the first k" (recall de-duplicates, :78)
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    bool hit = false, dup = false;
    if (i < count) {
      const int64_t id = ret[i];
      hit = id >= 0 && contains_sorted(rel_sorted + r0, n_unique, id);
      if (hit && i < kk) {
        for (int j = 0; j < i; ++j) dup |= (ret[j] == id);
      }
    }
    const uint32_t hm = __ballot_sync(0xffffffffu, hit);
    const uint32_t dm = __ballot_sync(0xffffffffu, dup);
    if (lane == 0) {
      hitmask[warp][base >> 5] = hm;
      dupmask[warp][base >> 5] = dm;
    }
  }
  __syncwarp();
  if (lane != 0) return;

  // sequential replay in the reference's order (fp64)
  int hits_k = 0, distinct_hits_k = 0, first_hit = -1;
  double ap_sum = 0.0, dcg = 0.0;
  for (int i = 0; i < count; ++i) {
    const bool hit = (hitmask[warp][i >> 5] >> (i & 31)) & 1u;
    if (!hit) continue;
    if (first_hit < 0) first_hit = i;
    if (i < kk) {
      ++hits_k;
      if (!((dupmask[warp][i >> 5] >> (i & 31)) & 1u)) ++distinct_hits_k;
      ap_sum += static_cast<double>(hits_k) / static_cast<double>(i + 1);  // score += hits / i   (:37)
      dcg += 1.0 / log2_tbl[i];                                           // score / np.log2(idx + 2) (:83)
    }
  }
  double idcg = 0.0;
  for (int i = 0; i < hits_k; ++i) idcg += 1.0 / log2_tbl[i];             // ideal = hits sorted first (:86)
  double* o = out + static_cast<int64_t>(q) * 5;
  o[0] = static_cast<double>(hits_k) / static_cast<double>(k);                             // :11
  o[1] = list_len == 0 ? 0.0 : static_cast<double>(distinct_hits_k) / static_cast<double>(n_unique);  // :75-79
  o[2] = list_len > 0 ? ap_sum / static_cast<double>(list_len) : 0.0;                        // :38
  o[3] = first_hit >= 0 ? 1.0 / static_cast<double>(first_hit + 1) : 0.0;                   // :69
  o[4] = idcg > 0.0 ? dcg / idcg : 0.0;                                                     // :89
}

__global__ void label_relevance_kernel(const uint64_t* __restrict__ qm, int64_t nq, const uint64_t* __restrict__ gm,
                                       int64_t ng, int words, int exclude_self, uint8_t* __restrict__ out) {
  const int64_t total = nq * ng;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = t / ng, j = t - i * ng;
    uint64_t any = 0;
    for (int w = 0; w < words; ++w) any |= qm[i * words + w] & gm[j * words + w];
    out[t] = (any != 0 && !(exclude_self && i == j)) ? 1 : 0;
  }
}

}  // namespace

int launch_metrics(const int64_t* retrieved, const int32_t* ret_count, int q, int k_ret, const int64_t* rel_indptr,
                   const int64_t* rel_sorted, const int64_t* rel_list_len, int k, const double* log2_tbl,
                   double* out, cudaStream_t stream) {
  if (q == 0) return MMR_OK;
  if (k_ret > kMaxRet) return fail(MMR_EUNSUP, "metrics: retrieved lists longer than 4096 are not supported");
  if (k < 1) return fail(MMR_EINVAL, "metrics: k must be >= 1 (the reference divides by k)");
  metrics_kernel<<<(q + 3) / 4, 128, 0, stream>>>(retrieved, ret_count, q, k_ret, rel_indptr, rel_sorted,
                                                  rel_list_len, k, log2_tbl, out);
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_label_relevance(const uint64_t* q_masks, int64_t nq, const uint64_t* g_masks, int64_t ng,
                           int label_words, int exclude_self, uint8_t* out, cudaStream_t stream) {
  if (nq == 0 || ng == 0) return MMR_OK;
  const int64_t total = nq * ng;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  label_relevance_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(q_masks, nq, g_masks, ng, label_words,
                                                                       exclude_self, out);
  MMR_LAUNCHED();
  return MMR_OK;
}

}  // namespace mmr
