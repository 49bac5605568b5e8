// "Next" rows of the hot path: evaluation reductions over the full ranking
// and result-set diversity, on the device.
//
//  * first_relevant_rank_kernel -- compute_ranking_metrics (reference Evaluate/retrieval_overlap.py:84-115):
//    the reference materialises cosine_similarity(Q, G), argsorts every row and walks the ranking with
//    Python sets until the first gallery item sharing a label with the query.  Here: one CTA per query,
//    two sweeps over the gallery -- (A) best ordering key among the RELEVANT rows (label masks overlap)
//    and the number of relevant rows, (B) number of rows whose key beats it -- so
//    rank = 1 + #better, without the (Q, N) matrix or any sort.  Keys use the library-wide rule
//    (score desc, row asc); scores are (dot * inv_norm[g]) * inv_norm[q] in fp32.
//  * diversity_kernel -- compute_embedding_diversity / compute_label_diversity_from_labels
//    (reference Evaluate/retrieval_diversity_compute.py:171-194): 1 - mean pairwise cosine of a result
//    set, and |union of labels| / mean label count over the items that have labels.
// Latency-bound helpers for evaluation-sized inputs; reported as time only.
#include <math_constants.h>

#include "internal.h"

namespace mmr {
namespace {

template <typename T>
__device__ __forceinline__ float elem(const T* p, int64_t i);
template <>
__device__ __forceinline__ float elem<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float elem<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other > v ? other : v;
  }
  return v;
}

// one CTA (8 warps) per query; dynamic smem: d_pad floats (the query)
template <typename T>
__global__ void __launch_bounds__(256)
first_relevant_rank_kernel(const T* __restrict__ emb, const float* __restrict__ inv_norm, int64_t n, int d_pad,
                           const float* __restrict__ q, const float* __restrict__ q_inv,
                           const uint64_t* __restrict__ q_masks, const uint64_t* __restrict__ g_masks, int words,
                           int64_t* __restrict__ out_rank, int64_t* __restrict__ out_total) {
  extern __shared__ __align__(16) float qs[];
  __shared__ uint64_t s_best[8];
  __shared__ unsigned long long s_cnt[8];
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < d_pad; i += blockDim.x) qs[i] = q[static_cast<int64_t>(qi) * d_pad + i];
  __syncthreads();
  const float qinv = q_inv[qi];
  const uint64_t* qm = q_masks + static_cast<int64_t>(qi) * words;

  auto row_key = [&](int64_t r) -> uint64_t {
    const T* g = emb + r * d_pad;
    float acc = 0.f;
    for (int i = lane; i < d_pad; i += 32) acc = fmaf(elem<T>(g, i), qs[i], acc);
    acc = warp_sum(acc);
    return make_key((acc * inv_norm[r]) * qinv, static_cast<uint32_t>(r));
  };

  // sweep A: best key among relevant rows + number of relevant rows
  uint64_t best = 0ull;
  unsigned long long total = 0;
  for (int64_t r = warp; r < n; r += 8) {
    uint64_t any = 0;
    for (int w = 0; w < words; ++w) any |= qm[w] & g_masks[r * words + w];
    if (any != 0) {  // warp-uniform
      ++total;
      const uint64_t k = row_key(r);
      best = k > best ? k : best;
    }
  }
  if (lane == 0) {
    s_best[warp] = best;
    s_cnt[warp] = total;
  }
  __syncthreads();
  best = 0ull;
  total = 0;
  for (int w = 0; w < 8; ++w) {
    best = s_best[w] > best ? s_best[w] : best;
    total += s_cnt[w];
  }
  __syncthreads();
  // sweep B: rows ranked ahead of the best relevant row
  unsigned long long ahead = 0;
  if (best != 0ull) {
    for (int64_t r = warp; r < n; r += 8) ahead += row_key(r) > best ? 1ull : 0ull;
  }
  if (lane == 0) s_cnt[warp] = ahead;
  __syncthreads();
  if (threadIdx.x == 0) {
    ahead = 0;
    for (int w = 0; w < 8; ++w) ahead += s_cnt[w];
    out_rank[qi] = best != 0ull ? static_cast<int64_t>(ahead) + 1 : 0;  // 0 = no relevant item
    out_total[qi] = static_cast<int64_t>(total);
  }
}

// one CTA (4 warps) per result set: emb (b, k, d) fp32; optional label masks (b, k, words)
__global__ void __launch_bounds__(128)
diversity_kernel(const float* __restrict__ emb, const uint64_t* __restrict__ masks, const int32_t* __restrict__ counts,
                 int k, int d, int words, double* __restrict__ out_emb_div, double* __restrict__ out_label_div) {
  __shared__ double s_sum[4];
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = counts != nullptr ? min(counts[qi], k) : k;
  if (emb != nullptr && out_emb_div != nullptr) {
    const float* e = emb + static_cast<int64_t>(qi) * k * d;
    double sum = 0.0;
    // pairs (i < j) dealt round-robin to warps; cos = dot / (max(|a|, 1e-8) * max(|b|, 1e-8))  (:176-178)
    int pair = 0;
    for (int i = 0; i < cnt; ++i) {
      for (int j = i + 1; j < cnt; ++j, ++pair) {
        if ((pair & 3) != warp) continue;
        float dot = 0.f, sa = 0.f, sb = 0.f;
        for (int t = lane; t < d; t += 32) {
          const float a = e[i * d + t], b = e[j * d + t];
          dot = fmaf(a, b, dot);
          sa = fmaf(a, a, sa);
          sb = fmaf(b, b, sb);
        }
        dot = warp_sum(dot);
        sa = warp_sum(sa);
        sb = warp_sum(sb);
        sum += static_cast<double>(dot / (fmaxf(sqrtf(sa), 1e-8f) * fmaxf(sqrtf(sb), 1e-8f)));
      }
    }
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      const double pairs = 0.5 * cnt * (cnt - 1);
      out_emb_div[qi] = cnt < 2 ? 0.0 : 1.0 - (s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3]) / pairs;  // :172-173,181-182
    }
  }
  if (masks != nullptr && out_label_div != nullptr && threadIdx.x == 0) {
    const uint64_t* m = masks + static_cast<int64_t>(qi) * k * words;
    int uni = 0, nonzero = 0;
    long long sizes = 0;
    for (int w = 0; w < words; ++w) {
      uint64_t u = 0;
      for (int i = 0; i < cnt; ++i) u |= m[i * words + w];
      uni += __popcll(u);
    }
    for (int i = 0; i < cnt; ++i) {
      int s = 0;
      for (int w = 0; w < words; ++w) s += __popcll(m[i * words + w]);
      if (s > 0) {
        ++nonzero;
        sizes += s;
      }
    }
    // unique labels / mean label count over items that have labels (:186-194)
    out_label_div[qi] = nonzero == 0 ? 0.0 : static_cast<double>(uni) / (static_cast<double>(sizes) / nonzero);
  }
}

// label_ranking_kernel -- evaluate_label_attention (reference Trainner/train_label_attention.py:106-125):
// all-pairs cosine of N record embeddings, and per query i over the FULL ranking of all N items
// (self included, with label 0): mean of the relevance flags among the top-k ("recall@k" in the
// reference's naming) and sklearn's average_precision_score.  No (N, N) matrix, no argsort: with
//   tp(s) = #relevant with score >= s,  all(s) = #items with score >= s
// average precision = (1/|rel|) * sum over relevant j of tp(s_j) / all(s_j)  (ties share one threshold,
// exactly sklearn's step-wise definition), and item j is in the top-k iff fewer than k items are ranked
// ahead of it (score desc, row asc).  One CTA per query; dynamic smem: n scores + n flags.
__global__ void __launch_bounds__(256)
label_ranking_kernel(const float* __restrict__ emb, const float* __restrict__ norms, int n, int d,
                     const uint64_t* __restrict__ masks, int words, const int32_t* __restrict__ topk, int n_topk,
                     double* __restrict__ out) {
  extern __shared__ __align__(16) float sm_scores[];
  uint8_t* const rel = reinterpret_cast<uint8_t*>(sm_scores + n);
  __shared__ double s_ap[8];
  __shared__ int s_hits[8][8];
  __shared__ int s_nrel[8];
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* q = emb + static_cast<int64_t>(qi) * d;
  const float qn = norms[qi];
  const uint64_t* qm = masks + static_cast<int64_t>(qi) * words;
  // phase 1: sims[i][j] = dot / (norm_i * norm_j)  (:108-109), relevance = labels share a positive, self excluded (:114-115)
  for (int j = warp; j < n; j += 8) {
    const float* g = emb + static_cast<int64_t>(j) * d;
    float acc = 0.f;
    for (int t = lane; t < d; t += 32) acc = fmaf(g[t], q[t], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      sm_scores[j] = acc / (qn * norms[j]);
      uint64_t any = 0;
      for (int w = 0; w < words; ++w) any |= qm[w] & masks[static_cast<int64_t>(j) * words + w];
      rel[j] = (any != 0 && j != qi) ? 1 : 0;
    }
  }
  __syncthreads();
  // phase 2: every relevant item looks at the whole ranking
  double ap = 0.0;
  int hits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int nrel = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    if (rel[j] == 0) continue;
    ++nrel;
    const float sj = sm_scores[j];
    int ge_all = 0, ge_rel = 0, ahead = 0;
    for (int m = 0; m < n; ++m) {
      const float s = sm_scores[m];
      const bool ge = s >= sj;
      ge_all += ge ? 1 : 0;
      ge_rel += (ge && rel[m] != 0) ? 1 : 0;
      ahead += (s > sj || (s == sj && m < j)) ? 1 : 0;
    }
    ap += static_cast<double>(ge_rel) / static_cast<double>(ge_all);
    for (int t = 0; t < n_topk; ++t) hits[t] += ahead < topk[t] ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ap += __shfl_xor_sync(0xffffffffu, ap, o);
    nrel += __shfl_xor_sync(0xffffffffu, nrel, o);
    for (int t = 0; t < 8; ++t) hits[t] += __shfl_xor_sync(0xffffffffu, hits[t], o);
  }
  if (lane == 0) {
    s_ap[warp] = ap;
    s_nrel[warp] = nrel;
    for (int t = 0; t < 8; ++t) s_hits[warp][t] = hits[t];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    int r = 0;
    for (int w = 0; w < 8; ++w) {
      a += s_ap[w];
      r += s_nrel[w];
    }
    double* o = out + static_cast<int64_t>(qi) * (1 + n_topk);
    o[0] = r == 0 ? 0.0 : a / static_cast<double>(r);  // sklearn: no positive => 0
    for (int t = 0; t < n_topk; ++t) {
      int h = 0;
      for (int w = 0; w < 8; ++w) h += s_hits[w][t];
      o[1 + t] = static_cast<double>(h) / static_cast<double>(topk[t]);  // sorted_labels[:k].mean()  (:121)
    }
  }
}

}  // namespace

int launch_label_ranking(const float* emb, const float* norms, int n, int d, const uint64_t* masks, int words,
                         const int32_t* topk, int n_topk, double* out, cudaStream_t stream) {
  if (n == 0) return MMR_OK;
  const size_t smem = static_cast<size_t>(n) * (sizeof(float) + 1) + 16;
  if (smem > 200 * 1024) return fail(MMR_EUNSUP, "label ranking evaluation: more than ~40k records");
  if (smem > 48 * 1024)
    MMR_CUDA_TRY(cudaFuncSetAttribute(label_ranking_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  label_ranking_kernel<<<n, 256, smem, stream>>>(emb, norms, n, d, masks, words, topk, n_topk, out);
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_first_relevant_rank(const void* emb, int dtype_store, const float* inv_norm, int64_t n, int d_pad,
                               const float* q_f32, const float* q_inv, int b, const uint64_t* q_masks,
                               const uint64_t* g_masks, int words, int64_t* out_rank, int64_t* out_total,
                               cudaStream_t stream) {
  if (b == 0) return MMR_OK;
  const size_t smem = static_cast<size_t>(d_pad) * sizeof(float);
  if (dtype_store == MMR_BF16) {
    first_relevant_rank_kernel<__nv_bfloat16><<<b, 256, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(emb), inv_norm, n, d_pad, q_f32, q_inv, q_masks, g_masks, words, out_rank,
        out_total);
  } else {
    first_relevant_rank_kernel<float><<<b, 256, smem, stream>>>(static_cast<const float*>(emb), inv_norm, n, d_pad,
                                                                q_f32, q_inv, q_masks, g_masks, words, out_rank,
                                                                out_total);
  }
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_diversity(const float* emb, const uint64_t* masks, const int32_t* counts, int b, int k, int d, int words,
                     double* out_emb_div, double* out_label_div, cudaStream_t stream) {
  if (b == 0) return MMR_OK;
  diversity_kernel<<<b, 128, 0, stream>>>(emb, masks, counts, k, d, words, out_emb_div, out_label_div);
  MMR_LAUNCHED();
  return MMR_OK;
}

}  // namespace mmr
