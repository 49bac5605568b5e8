// Fused tail of a search step for ONE query: label / KG features of its (<= 128) candidates, min-max scaling,
// weighted sum and final order -- Reranker.rerank (reference Retrieval/reranker.py:298-329) without the three
// round trips through global memory of the unfused kernels (rerank_features -> rerank_combine -> apply_order).
//
// The embedding feature emb_scores[i] (reranker.py:298, safe_cos(q_emb, cand_emb[i])) is the search score the
// candidate was selected with: the same quantity, so the K x D gather of candidate
// rows -- as many bytes as the KG gather -- is not repeated.  Label Jaccard (:301-304), KG cosine (:307-319),
// minmax_scale_list (:152-159) and the combine (:325-327) use exactly the arithmetic of rerank.cu, in the same
// order, so that the fused and the unfused kernels agree bit for bit when fed the same cosines.
#pragma once

#include <math_constants.h>

#include "internal.h"

namespace mmr {

constexpr int kTailMaxK = 128;     // candidates per query the fused tail covers
constexpr int kTailThreads = 256;
#ifndef MMR_TAIL_CAND
#define MMR_TAIL_CAND 2
#endif
constexpr int kTailCand = MMR_TAIL_CAND;   // candidates whose KG gathers a warp keeps in flight

__device__ __forceinline__ float tail_safe_cos(float dot, float ssa, float ssb) {  // safe_cos, reranker.py:135-142
  const float na = sqrtf(ssa), nb = sqrtf(ssb);
  if (na == 0.f || nb == 0.f) return 0.f;
  return dot / (na * nb);
}

struct TailTables {  // device view of mmr_rerank_tables
  const uint64_t* masks;
  int label_words;
  const float* kg;
  int d_kg;
  int64_t n_rec;
};

struct TailSmem {
  int64_t cand_row[kTailMaxK];   // GLOBAL row == record index, best first (search order)
  float cand_score[kTailMaxK];
  double raw[3][kTailMaxK];      // emb, label, kg
  double fin[kTailMaxK];
  double nrm[3][kTailMaxK];
  __align__(16) float qks[512];  // the query's KG vector, lane-sliced like the gathers
  double lo[3], hi[3];
};

// All kTailThreads threads of the CTA call this with the query's `count` candidates in sm.cand_row / cand_score.
// emit(rank, j, final, emb_n, lab_n, kg_n) is called once for every rank < min(count, keep) by one thread.
// kKIts = ceil(d_kg / 128) slices of 4 floats per lane (d_kg a multiple of 4, table 16-byte aligned).
template <int kKIts, typename Emit>
__device__ __forceinline__ void rerank_tail(TailSmem& sm, int count, int64_t qr, const TailTables& t, double alpha,
                                            double beta, double gamma, int keep, Emit emit) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kTailThreads >> 5;
  const bool q_known = qr >= 0 && qr < t.n_rec;
  const float* qkg = (q_known && t.kg != nullptr) ? t.kg + qr * t.d_kg : nullptr;
  const uint64_t* qmask = (q_known && t.masks != nullptr) ? t.masks + qr * t.label_words : nullptr;
  const int nk = t.d_kg >> 2;  // float4 per KG row
  for (int i = tid; i < kKIts * 128; i += kTailThreads) sm.qks[i] = (qkg != nullptr && i < t.d_kg) ? qkg[i] : 0.f;
  __syncthreads();
  float qkss = 0.f;
#pragma unroll
  for (int it = 0; it < kKIts; ++it) {
    const float4 a = *reinterpret_cast<const float4*>(sm.qks + (it * 32 + lane) * 4);
    qkss = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, qkss))));
  }
  qkss = warp_sum(qkss);
  // ---- KG cosine: kTailCand candidates per warp in flight (all gathers issued before the first reduction) ----
  for (int j0 = warp * kTailCand; j0 < count; j0 += nwarps * kTailCand) {
    float4 y[kTailCand][kKIts];
    bool do_kg[kTailCand];
#pragma unroll
    for (int c = 0; c < kTailCand; ++c) {
      const int j = j0 + c;
      const int64_t cr = j < count ? sm.cand_row[j] : -1;
      do_kg[c] = qkg != nullptr && cr >= 0 && cr < t.n_rec;
      const float4* ck = reinterpret_cast<const float4*>(t.kg + (do_kg[c] ? cr : 0) * t.d_kg);
#pragma unroll
      for (int it = 0; it < kKIts; ++it) {
        const int u = it * 32 + lane;
        y[c][it] = (do_kg[c] && u < nk) ? __ldg(ck + u) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float kdot[kTailCand], kss[kTailCand];
#pragma unroll
    for (int c = 0; c < kTailCand; ++c) kdot[c] = kss[c] = 0.f;
#pragma unroll
    for (int it = 0; it < kKIts; ++it) {
      const float4 qa = *reinterpret_cast<const float4*>(sm.qks + (it * 32 + lane) * 4);
#pragma unroll
      for (int c = 0; c < kTailCand; ++c) {
        const float4 v = y[c][it];
        kdot[c] = fmaf(v.x, qa.x, fmaf(v.y, qa.y, fmaf(v.z, qa.z, fmaf(v.w, qa.w, kdot[c]))));
        kss[c] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, kss[c]))));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < kTailCand; ++c) {
        kdot[c] += __shfl_xor_sync(0xffffffffu, kdot[c], o);
        kss[c] += __shfl_xor_sync(0xffffffffu, kss[c], o);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < kTailCand; ++c)
        if (j0 + c < count) sm.raw[2][j0 + c] = static_cast<double>(do_kg[c] ? tail_safe_cos(kdot[c], qkss, kss[c]) : 0.f);
    }
  }
  // ---- embedding feature (the search score) and label Jaccard: one thread per candidate ----
  if (tid < count) {
    sm.raw[0][tid] = static_cast<double>(sm.cand_score[tid]);
    const int64_t cr = sm.cand_row[tid];
    const bool known = cr >= 0 && cr < t.n_rec;
    int inter = 0, uni = 0;
    if (t.masks != nullptr) {
      for (int w = 0; w < t.label_words; ++w) {
        const uint64_t a = qmask != nullptr ? qmask[w] : 0ull;
        const uint64_t bb = known ? t.masks[cr * t.label_words + w] : 0ull;
        inter += __popcll(a & bb);
        uni += __popcll(a | bb);
      }
    }
    sm.raw[1][tid] = uni == 0 ? 0.0 : static_cast<double>(inter) / static_cast<double>(uni);
  }
  __syncthreads();
  // ---- min-max per feature (np.nanmin / np.nanmax: fmin / fmax drop NaNs), combine, order ----
  if (tid < 96) {
    const int f = tid >> 5;
    double lo = CUDART_INF, hi = -CUDART_INF;
    for (int j = lane; j < count; j += 32) {
      const double x = sm.raw[f][j];
      lo = fmin(lo, x);
      hi = fmax(hi, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      sm.lo[f] = lo;
      sm.hi[f] = hi;
    }
  }
  __syncthreads();
  if (tid < count) {
    double v[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const double range = __dsub_rn(sm.hi[f], sm.lo[f]);
      v[f] = (range == 0.0) ? 0.0 : __ddiv_rn(__dsub_rn(sm.raw[f][tid], sm.lo[f]), range);
      sm.nrm[f][tid] = v[f];
    }
    // (alpha*e + beta*l) + gamma*g with every product and sum rounded separately (numpy does not fuse)
    sm.fin[tid] = __dadd_rn(__dadd_rn(__dmul_rn(alpha, v[0]), __dmul_rn(beta, v[1])), __dmul_rn(gamma, v[2]));
  }
  __syncthreads();
  if (tid < count) {  // rank by counting: final descending, candidate position ascending
    const double fj = sm.fin[tid];
    int rank = 0;
    for (int i = 0; i < count; ++i) {
      const double fi = sm.fin[i];
      rank += (fi > fj || (fi == fj && i < tid)) ? 1 : 0;
    }
    if (rank < keep) emit(rank, tid, fj, sm.nrm[0][tid], sm.nrm[1][tid], sm.nrm[2][tid]);
  }
}

}  // namespace mmr
