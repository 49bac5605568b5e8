// K5: label / knowledge-graph rerank.
//
// Replaces Reranker.rerank (reference Retrieval/reranker.py:240-333):
//   emb_scores[i] = safe_cos(q_emb, cand_emb[i])                                  (:298, safe_cos :135-142)
//   lab_scores[i] = jaccard(labels(q), labels(c_i))                                (:301-304, :145-149)
//   kg_scores[i]  = safe_cos(kg(q), kg(c_i))                                       (:307-319)
//   x_n = minmax_scale_list(x) in fp64, all zeros when max == min                  (:152-159, :322-324)
//   final = alpha*emb_n + beta*lab_n + gamma*kg_n                                  (:325)
//   order = argsort(final)[::-1][:topk]                                            (:327-329)
// The pandas .loc label lookups that dominate the reference (25 ms per 100 candidates) become one
// popcount on 64-bit label masks; KG vectors are rows of a device table.
//
// features kernel: one CTA per query, one warp per candidate (fp32 dot products with shuffle
// reductions, exactly safe_cos's formula dot/(||a||*||b||)).  combine kernel: one CTA per query,
// fp64 with explicit non-fused mul/add so the result matches numpy's unfused arithmetic.
// Latency-bound (K <= 1024 candidates per query); reported as time only.
#include <math_constants.h>

#include "internal.h"
#include "rerank_tail.cuh"

namespace mmr {
namespace {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ float safe_cos_finish(float dot, float ssa, float ssb) {
  const float na = sqrtf(ssa), nb = sqrtf(ssb);
  if (na == 0.f || nb == 0.f) return 0.f;
  return dot / (na * nb);
}

template <typename T>
__global__ void __launch_bounds__(128)
rerank_features_kernel(const T* __restrict__ emb, int64_t n, int d_pad, int64_t row_offset,
                       const uint64_t* __restrict__ label_masks, int label_words, const float* __restrict__ kg,
                       int d_kg, int64_t n_rec, const float* __restrict__ q_emb, const float* __restrict__ cand_emb,
                       const int64_t* __restrict__ cand_rows, const int64_t* __restrict__ q_rec,
                       const int64_t* __restrict__ cand_rec, const int32_t* __restrict__ cand_count, int k, int d,
                       double* __restrict__ out_raw, uint8_t* __restrict__ owned,
                       const float* __restrict__ emb_cos_in, float* __restrict__ cos_out) {
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int count = cand_count != nullptr ? min(cand_count[qi], k) : k;
  const float* qv = q_emb + static_cast<int64_t>(qi) * d;
  const int64_t qr = q_rec != nullptr ? q_rec[qi] : -1;
  const bool q_known = qr >= 0 && qr < n_rec;
  const float* qkg = (q_known && kg != nullptr) ? kg + qr * d_kg : nullptr;
  const uint64_t* qmask = (q_known && label_masks != nullptr) ? label_masks + qr * label_words : nullptr;

  float qss = 0.f;
  if (emb_cos_in == nullptr) {
    for (int i = lane; i < d; i += 32) qss = fmaf(qv[i], qv[i], qss);
    qss = warp_sum(qss);
  }
  float qkss = 0.f;
  if (qkg != nullptr) {
    for (int i = lane; i < d_kg; i += 32) qkss = fmaf(qkg[i], qkg[i], qkss);
    qkss = warp_sum(qkss);
  }

  for (int j = warp; j < k; j += nwarps) {
    double* o = out_raw != nullptr ? out_raw + (static_cast<int64_t>(qi) * k + j) * 3 : nullptr;
    if (j >= count) {
      if (lane == 0) {
        if (o != nullptr) { o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; }
        if (cos_out != nullptr) cos_out[static_cast<int64_t>(qi) * k + j] = 0.f;
        if (owned != nullptr) owned[static_cast<int64_t>(qi) * k + j] = 0;
      }
      continue;
    }
    // --- embedding cosine
    float dot = 0.f, css = 0.f;
    bool have = true;
    if (emb_cos_in != nullptr) {
      // precomputed by the rank that owns the candidate row (sharded path)
    } else if (cand_emb != nullptr) {
      const float* c = cand_emb + (static_cast<int64_t>(qi) * k + j) * d;
      for (int i = lane; i < d; i += 32) {
        const float x = c[i];
        dot = fmaf(x, qv[i], dot);
        css = fmaf(x, x, css);
      }
    } else {
      const int64_t local = cand_rows[static_cast<int64_t>(qi) * k + j] - row_offset;
      have = local >= 0 && local < n;
      if (have) {
        const T* c = emb + local * d_pad;
        for (int i = lane; i < d; i += 32) {
          const float x = to_f32<T>(c[i]);
          dot = fmaf(x, qv[i], dot);
          css = fmaf(x, x, css);
        }
      }
    }
    dot = warp_sum(dot);
    css = warp_sum(css);
    const float e = emb_cos_in != nullptr ? emb_cos_in[static_cast<int64_t>(qi) * k + j]
                                          : (have ? safe_cos_finish(dot, qss, css) : 0.f);
    if (cos_out != nullptr) {  // cosine-only mode
      if (lane == 0) {
        cos_out[static_cast<int64_t>(qi) * k + j] = e;
        if (owned != nullptr) owned[static_cast<int64_t>(qi) * k + j] = have ? 1 : 0;
      }
      continue;
    }
    // --- label Jaccard + KG cosine
    const int64_t cr = cand_rec != nullptr ? cand_rec[static_cast<int64_t>(qi) * k + j] : -1;
    const bool c_known = cr >= 0 && cr < n_rec;
    int inter = 0, uni = 0;
    if (label_masks != nullptr) {
      for (int w = lane; w < label_words; w += 32) {
        const uint64_t a = qmask != nullptr ? qmask[w] : 0ull;
        const uint64_t bb = c_known ? label_masks[cr * label_words + w] : 0ull;
        inter += __popcll(a & bb);
        uni += __popcll(a | bb);
      }
      inter = __reduce_add_sync(0xffffffffu, inter);
      uni = __reduce_add_sync(0xffffffffu, uni);
    }
    float kdot = 0.f, kss = 0.f;
    if (qkg != nullptr && c_known) {
      const float* c = kg + cr * d_kg;
      for (int i = lane; i < d_kg; i += 32) {
        const float x = c[i];
        kdot = fmaf(x, qkg[i], kdot);
        kss = fmaf(x, x, kss);
      }
    }
    kdot = warp_sum(kdot);
    kss = warp_sum(kss);
    if (lane == 0) {
      o[0] = static_cast<double>(e);
      o[1] = uni == 0 ? 0.0 : static_cast<double>(inter) / static_cast<double>(uni);
      o[2] = static_cast<double>((qkg != nullptr && c_known) ? safe_cos_finish(kdot, qkss, kss) : 0.f);
      if (owned != nullptr) owned[static_cast<int64_t>(qi) * k + j] = have ? 1 : 0;
    }
  }
}

// Vectorised variant for the hot configuration (bf16 gallery rows gathered by candidate row, or a
// precomputed cosine): 128-bit loads of the candidate's embedding row and KG vector (a 512-d bf16
// row is two fully coalesced 512 B requests per warp instead of sixteen 64 B ones).  The kernel is
// latency-bound (a chain row id -> row address -> gather -> shuffle reduction per candidate), so:
//   * the candidate row / record ids of the whole query are staged in shared memory up front,
//   * every warp works on kCand candidates at once -- all their gathers are issued before the first
//     reduction, and the shuffle reductions of the candidates interleave,
//   * the query's embedding / KG vector live in shared memory in the lane-sliced order the gathers
//     use (lane l owns elements (it*32 + l)*8 .. +8), which keeps registers free for loads in flight.
// kIts = ceil(d_pad / 256) embedding slices per lane, kKIts = ceil(d_kg / 128) KG slices per lane.
// Same formulas as the generic kernel (safe_cos: dot / (||a|| * ||b||)).
constexpr int kCand = 2;
constexpr int kVecMaxK = 1024;

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

template <int kIts, int kKIts>
__global__ void __launch_bounds__(256)
rerank_features_vec_kernel(const __nv_bfloat16* __restrict__ emb, int64_t n, int d_pad, int64_t row_offset,
                           const uint64_t* __restrict__ label_masks, int label_words, const float* __restrict__ kg,
                           int d_kg, int64_t n_rec, const float* __restrict__ q_emb,
                           const int64_t* __restrict__ cand_rows, const int64_t* __restrict__ q_rec,
                           const int64_t* __restrict__ cand_rec, const int32_t* __restrict__ cand_count, int k, int d,
                           double* __restrict__ out_raw, uint8_t* __restrict__ owned,
                           const float* __restrict__ emb_cos_in, float* __restrict__ cos_out) {
  __shared__ __align__(16) float qs[kIts * 256];
  __shared__ __align__(16) float qks[kKIts * 128];
  extern __shared__ __align__(16) int64_t s_ids[];  // [k] local rows (-1 = not in this shard) | [k] record ids
  int64_t* const s_row = s_ids;
  int64_t* const s_rec = s_ids + k;
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int count = cand_count != nullptr ? min(cand_count[qi], k) : k;
  const int64_t qr = q_rec != nullptr ? q_rec[qi] : -1;
  const bool q_known = qr >= 0 && qr < n_rec;
  const float* qkg = (q_known && kg != nullptr) ? kg + qr * d_kg : nullptr;
  const uint64_t* qmask = (q_known && label_masks != nullptr) ? label_masks + qr * label_words : nullptr;
  const bool need_emb = emb_cos_in == nullptr;
  const bool feats = cos_out == nullptr;
  const int nv = d_pad >> 3;  // 16-byte vectors per embedding row
  const int nk = d_kg >> 2;   // float4 per KG row
  const int64_t base = static_cast<int64_t>(qi) * k;

  for (int i = threadIdx.x; i < kIts * 256; i += blockDim.x)
    qs[i] = (need_emb && i < d) ? q_emb[static_cast<int64_t>(qi) * d + i] : 0.f;
  for (int i = threadIdx.x; i < kKIts * 128; i += blockDim.x) qks[i] = (qkg != nullptr && i < d_kg) ? qkg[i] : 0.f;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    int64_t local = -1;
    if (need_emb && j < count) {
      local = cand_rows[base + j] - row_offset;
      if (local < 0 || local >= n) local = -1;
    }
    s_row[j] = local;
    const int64_t cr = (feats && cand_rec != nullptr && j < count) ? cand_rec[base + j] : -1;
    s_rec[j] = (cr >= 0 && cr < n_rec) ? cr : -1;
  }
  __syncthreads();
  float qss = 0.f, qkss = 0.f;
#pragma unroll
  for (int it = 0; it < kIts; ++it) {
    const float4 a = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8);
    const float4 c = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8 + 4);
    qss = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, qss))));
    qss = fmaf(c.x, c.x, fmaf(c.y, c.y, fmaf(c.z, c.z, fmaf(c.w, c.w, qss))));
  }
#pragma unroll
  for (int it = 0; it < kKIts; ++it) {
    const float4 a = *reinterpret_cast<const float4*>(qks + (it * 32 + lane) * 4);
    qkss = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, qkss))));
  }
  qss = warp_sum(qss);
  qkss = warp_sum(qkss);

  for (int j0 = warp * kCand; j0 < k; j0 += nwarps * kCand) {
    // ---- issue every gather of this warp's kCand candidates before the first reduction ----
    uint4 x[kCand][kIts];
    float4 y[kCand][kKIts];
    uint64_t cmask[kCand];
    bool have[kCand], do_kg[kCand], known[kCand];
#pragma unroll
    for (int c = 0; c < kCand; ++c) {
      const int j = j0 + c;
      const bool live = j < count;
      const int64_t local = live ? s_row[j] : -1;
      const int64_t cr = live ? s_rec[j] : -1;
      have[c] = local >= 0;
      known[c] = cr >= 0;
      do_kg[c] = feats && qkg != nullptr && known[c];
      const uint4* ce = reinterpret_cast<const uint4*>(emb + (have[c] ? local : 0) * d_pad);
#pragma unroll
      for (int it = 0; it < kIts; ++it) {
        const int u = it * 32 + lane;
        x[c][it] = (need_emb && have[c] && u < nv) ? __ldg(ce + u) : make_uint4(0u, 0u, 0u, 0u);
      }
      const float4* ck = reinterpret_cast<const float4*>(kg + (do_kg[c] ? cr : 0) * d_kg);
#pragma unroll
      for (int it = 0; it < kKIts; ++it) {
        const int u = it * 32 + lane;
        y[c][it] = (do_kg[c] && u < nk) ? __ldg(ck + u) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cmask[c] = (feats && label_masks != nullptr && known[c] && lane < label_words)
                     ? label_masks[cr * label_words + lane] : 0ull;
    }
    // ---- reduce ----
    float dot[kCand], css[kCand], kdot[kCand], kss[kCand];
    int inter[kCand], uni[kCand];
#pragma unroll
    for (int c = 0; c < kCand; ++c) {
      dot[c] = css[c] = kdot[c] = kss[c] = 0.f;
      inter[c] = uni[c] = 0;
    }
#pragma unroll
    for (int it = 0; it < kIts; ++it) {
      const float4 qa = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8);
      const float4 qb = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8 + 4);
      const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
      for (int c = 0; c < kCand; ++c) {
        const uint32_t w[4] = {x[c][it].x, x[c][it].y, x[c][it].z, x[c][it].w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float lo = bf16_lo(w[h]), hi = bf16_hi(w[h]);
          dot[c] = fmaf(lo, qv[2 * h], dot[c]);
          css[c] = fmaf(lo, lo, css[c]);
          dot[c] = fmaf(hi, qv[2 * h + 1], dot[c]);
          css[c] = fmaf(hi, hi, css[c]);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < kKIts; ++it) {
      const float4 qa = *reinterpret_cast<const float4*>(qks + (it * 32 + lane) * 4);
#pragma unroll
      for (int c = 0; c < kCand; ++c) {
        const float4 v = y[c][it];
        kdot[c] = fmaf(v.x, qa.x, fmaf(v.y, qa.y, fmaf(v.z, qa.z, fmaf(v.w, qa.w, kdot[c]))));
        kss[c] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, kss[c]))));
      }
    }
    if (feats && label_masks != nullptr) {
      // label words beyond the first 32 (never the case for the 43-label table) go through the loop
      const uint64_t a0 = (qmask != nullptr && lane < label_words) ? qmask[lane] : 0ull;
#pragma unroll
      for (int c = 0; c < kCand; ++c) {
        inter[c] = __popcll(a0 & cmask[c]);
        uni[c] = __popcll(a0 | cmask[c]);
        for (int w = lane + 32; w < label_words; w += 32) {
          const uint64_t a = qmask != nullptr ? qmask[w] : 0ull;
          const uint64_t bb = known[c] ? label_masks[s_rec[min(j0 + c, k - 1)] * label_words + w] : 0ull;
          inter[c] += __popcll(a & bb);
          uni[c] += __popcll(a | bb);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < kCand; ++c) {
        dot[c] += __shfl_xor_sync(0xffffffffu, dot[c], o);
        css[c] += __shfl_xor_sync(0xffffffffu, css[c], o);
        kdot[c] += __shfl_xor_sync(0xffffffffu, kdot[c], o);
        kss[c] += __shfl_xor_sync(0xffffffffu, kss[c], o);
      }
    }
#pragma unroll
    for (int c = 0; c < kCand; ++c) {
      const int j = j0 + c;
      if (j >= k) continue;
      const int64_t at = base + j;
      if (feats && label_masks != nullptr) {
        inter[c] = __reduce_add_sync(0xffffffffu, inter[c]);
        uni[c] = __reduce_add_sync(0xffffffffu, uni[c]);
      }
      if (lane != 0) continue;
      const bool live = j < count;
      const float e = !live ? 0.f : (need_emb ? (have[c] ? safe_cos_finish(dot[c], qss, css[c]) : 0.f) : emb_cos_in[at]);
      if (!feats) {
        cos_out[at] = e;
      } else if (!live) {
        out_raw[at * 3 + 0] = 0.0; out_raw[at * 3 + 1] = 0.0; out_raw[at * 3 + 2] = 0.0;
      } else {
        out_raw[at * 3 + 0] = static_cast<double>(e);
        out_raw[at * 3 + 1] = uni[c] == 0 ? 0.0 : static_cast<double>(inter[c]) / static_cast<double>(uni[c]);
        out_raw[at * 3 + 2] = static_cast<double>(do_kg[c] ? safe_cos_finish(kdot[c], qkss, kss[c]) : 0.f);
      }
      if (owned != nullptr) owned[at] = (live && have[c]) ? 1 : 0;
    }
  }
}

// One CTA per query.  dynamic smem: 4 * k doubles (final, emb_n, lab_n, kg_n).
__global__ void __launch_bounds__(256)
rerank_combine_kernel(const double* __restrict__ raw, const int32_t* __restrict__ cand_count, int k, double alpha,
                      double beta, double gamma, int topk, int32_t* __restrict__ out_order,
                      double* __restrict__ out_scores) {
  extern __shared__ __align__(16) double sm[];
  double* fin = sm;
  double* nrm = sm + k;  // [3][k]
  __shared__ double s_lo[3], s_hi[3];
  const int qi = blockIdx.x;
  const int count = cand_count != nullptr ? min(cand_count[qi], k) : k;
  const double* r = raw + static_cast<int64_t>(qi) * k * 3;
  const int keep = (topk > 0 && topk < k) ? topk : k;

  // NaN-aware min / max per feature (np.nanmin / np.nanmax): fmin/fmax drop NaNs
  if (threadIdx.x < 96) {
    const int f = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double lo = CUDART_INF, hi = -CUDART_INF;
    for (int j = lane; j < count; j += 32) {
      const double x = r[j * 3 + f];
      lo = fmin(lo, x);
      hi = fmax(hi, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      s_lo[f] = lo;
      s_hi[f] = hi;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < count; j += blockDim.x) {
    double v[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const double range = __dsub_rn(s_hi[f], s_lo[f]);
      v[f] = (range == 0.0) ? 0.0 : __ddiv_rn(__dsub_rn(r[j * 3 + f], s_lo[f]), range);
      nrm[f * k + j] = v[f];
    }
    // (alpha*e + beta*l) + gamma*g with every product and sum rounded separately (numpy does not fuse)
    fin[j] = __dadd_rn(__dadd_rn(__dmul_rn(alpha, v[0]), __dmul_rn(beta, v[1])), __dmul_rn(gamma, v[2]));
  }
  __syncthreads();
  // rank by counting: final descending, candidate position ascending
  for (int j = threadIdx.x; j < count; j += blockDim.x) {
    const double fj = fin[j];
    int rank = 0;
    for (int i = 0; i < count; ++i) {
      const double fi = fin[i];
      rank += (fi > fj || (fi == fj && i < j)) ? 1 : 0;
    }
    if (rank < keep) {
      const int64_t o = static_cast<int64_t>(qi) * keep + rank;
      out_order[o] = j;
      out_scores[o * 4 + 0] = fj;
      out_scores[o * 4 + 1] = nrm[0 * k + j];
      out_scores[o * 4 + 2] = nrm[1 * k + j];
      out_scores[o * 4 + 3] = nrm[2 * k + j];
    }
  }
  // padding when fewer valid candidates than `keep`
  for (int j = count + threadIdx.x; j < keep; j += blockDim.x) {
    const int64_t o = static_cast<int64_t>(qi) * keep + j;
    out_order[o] = -1;
    out_scores[o * 4 + 0] = 0.0; out_scores[o * 4 + 1] = 0.0;
    out_scores[o * 4 + 2] = 0.0; out_scores[o * 4 + 3] = 0.0;
  }
}

// Single-shard fused tail: one CTA per query; candidates = the search result (rows, scores) (b, k), best first,
// -1 rows = padding.  Writes what retrieve(..., reranker=...) returns (Retrieval/retrieval.py:257-269): ids in
// reranked order + combined scores (b, keep), optionally the four score columns of Reranker.rerank's tuples.
template <int kKIts>
__global__ void __launch_bounds__(kTailThreads)
rerank_scored_kernel(const int64_t* __restrict__ rows, const float* __restrict__ scores,
                     const int64_t* __restrict__ q_rec, TailTables t, int k, double alpha, double beta, double gamma,
                     int keep, int64_t* __restrict__ out_ids, double* __restrict__ out_fin,
                     double* __restrict__ out_scores4) {
  __shared__ TailSmem sm;
  __shared__ int s_count;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < k) {
    const int64_t r = rows[static_cast<int64_t>(q) * k + threadIdx.x];
    sm.cand_row[threadIdx.x] = r;
    sm.cand_score[threadIdx.x] = scores[static_cast<int64_t>(q) * k + threadIdx.x];
    if (r >= 0) atomicAdd(&s_count, 1);   // valid candidates form a prefix (search pads at the end)
  }
  __syncthreads();
  const int count = s_count;
  const int64_t base = static_cast<int64_t>(q) * keep;
  rerank_tail<kKIts>(sm, count, q_rec != nullptr ? q_rec[q] : -1, t, alpha, beta, gamma, keep,
                     [&](int rank, int j, double fin, double e, double l, double g) {
                       out_ids[base + rank] = sm.cand_row[j];
                       out_fin[base + rank] = fin;
                       if (out_scores4 != nullptr) {
                         double* o = out_scores4 + (base + rank) * 4;
                         o[0] = fin; o[1] = e; o[2] = l; o[3] = g;
                       }
                     });
  for (int r = count + threadIdx.x; r < keep; r += kTailThreads) {  // fewer valid candidates than `keep`
    out_ids[base + r] = -1;
    out_fin[base + r] = 0.0;
    if (out_scores4 != nullptr) {
      double* o = out_scores4 + (base + r) * 4;
      o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0;
    }
  }
}

}  // namespace

// KG table shapes the fused tail covers (else callers use the unfused kernels)
bool tail_supported(int k, const void* kg, int d_kg) {
  return k >= 1 && k <= kTailMaxK &&
         (kg == nullptr || (d_kg % 4 == 0 && d_kg <= 512 && (reinterpret_cast<uintptr_t>(kg) & 15u) == 0));
}

int launch_rerank_scored(const int64_t* rows, const float* scores, const int64_t* q_rec, const uint64_t* masks,
                         int label_words, const float* kg, int d_kg, int64_t n_rec, int b, int k, double alpha,
                         double beta, double gamma, int topk, int64_t* out_ids, double* out_fin,
                         double* out_scores4, cudaStream_t stream) {
  if (b == 0 || k == 0) return MMR_OK;
  if (!tail_supported(k, kg, d_kg)) return fail(MMR_EUNSUP, "rerank_scored: needs k <= 128 and a KG dimension <= 512 that is a multiple of 4");
  const int keep = (topk > 0 && topk < k) ? topk : k;
  const TailTables t{masks, label_words, kg, d_kg, n_rec};
  if (kg == nullptr || d_kg <= 384)
    rerank_scored_kernel<3><<<b, kTailThreads, 0, stream>>>(rows, scores, q_rec, t, k, alpha, beta, gamma, keep, out_ids,
                                                            out_fin, out_scores4);
  else
    rerank_scored_kernel<4><<<b, kTailThreads, 0, stream>>>(rows, scores, q_rec, t, k, alpha, beta, gamma, keep, out_ids,
                                                            out_fin, out_scores4);
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_rerank_features(const void* emb, int dtype_store, int64_t n, int d_pad, int64_t row_offset,
                           const uint64_t* label_masks, int label_words, const float* kg, int d_kg, int64_t n_rec,
                           const float* q_emb, const float* cand_emb, const int64_t* cand_rows,
                           const int64_t* q_rec, const int64_t* cand_rec, const int32_t* cand_count, int b, int k,
                           int d, double* out_raw, uint8_t* owned, const float* emb_cos_in, float* cos_out,
                           cudaStream_t stream) {
  if (b == 0 || k == 0) return MMR_OK;
  if (emb_cos_in == nullptr && cand_emb == nullptr && (emb == nullptr || cand_rows == nullptr))
    return fail(MMR_EINVAL, "Please provide candidate_embs or an index with candidate rows.");
  // vectorised path: gathered bf16 rows (or a precomputed cosine), 16-byte aligned tables
  const bool emb_vec_ok = emb_cos_in != nullptr ||
                          (cand_emb == nullptr && dtype_store == MMR_BF16 && d_pad <= 1024 &&
                           (reinterpret_cast<uintptr_t>(emb) & 15u) == 0);
  const bool kg_vec_ok = kg == nullptr || (d_kg % 4 == 0 && d_kg <= 512 && (reinterpret_cast<uintptr_t>(kg) & 15u) == 0);
  if (emb_vec_ok && kg_vec_ok && k <= kVecMaxK) {
    const __nv_bfloat16* e16 = static_cast<const __nv_bfloat16*>(emb);
    const int its = emb_cos_in != nullptr ? 1 : (d_pad + 255) / 256;
    const int kits = kg == nullptr ? 1 : (d_kg + 127) / 128;
    const size_t smem = static_cast<size_t>(2) * k * sizeof(int64_t);
#define MMR_RERANK_VEC(ITS, KITS)                                                                                  \
  rerank_features_vec_kernel<ITS, KITS><<<b, 256, smem, stream>>>(                                                 \
      e16, n, d_pad, row_offset, label_masks, label_words, kg, d_kg, n_rec, q_emb, cand_rows, q_rec, cand_rec,     \
      cand_count, k, d, out_raw, owned, emb_cos_in, cos_out)
    if (its <= 1 && kits <= 3) {
      MMR_RERANK_VEC(1, 3);
    } else if (its <= 2 && kits <= 3) {
      MMR_RERANK_VEC(2, 3);
    } else {
      MMR_RERANK_VEC(4, 4);
    }
#undef MMR_RERANK_VEC
    MMR_LAUNCHED();
    return MMR_OK;
  }
  if (dtype_store == MMR_BF16 && cand_emb == nullptr) {
    rerank_features_kernel<__nv_bfloat16><<<b, 128, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(emb), n, d_pad, row_offset, label_masks, label_words, kg, d_kg, n_rec,
        q_emb, cand_emb, cand_rows, q_rec, cand_rec, cand_count, k, d, out_raw, owned, emb_cos_in, cos_out);
  } else {
    rerank_features_kernel<float><<<b, 128, 0, stream>>>(
        static_cast<const float*>(emb), n, d_pad, row_offset, label_masks, label_words, kg, d_kg, n_rec, q_emb,
        cand_emb, cand_rows, q_rec, cand_rec, cand_count, k, d, out_raw, owned, emb_cos_in, cos_out);
  }
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_rerank_combine(const double* raw, const int32_t* cand_count, int b, int k, double alpha, double beta,
                          double gamma, int topk, int32_t* out_order, double* out_scores, cudaStream_t stream) {
  if (b == 0 || k == 0) return MMR_OK;
  const size_t smem = static_cast<size_t>(4) * k * sizeof(double);
  if (smem > 48 * 1024) {
    MMR_CUDA_TRY(cudaFuncSetAttribute(rerank_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  }
  rerank_combine_kernel<<<b, 256, smem, stream>>>(raw, cand_count, k, alpha, beta, gamma, topk, out_order,
                                                  out_scores);
  MMR_LAUNCHED();
  return MMR_OK;
}

}  // namespace mmr
