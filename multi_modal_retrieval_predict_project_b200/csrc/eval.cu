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

}  // namespace

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
