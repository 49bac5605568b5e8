// K2: HBM-streaming exact cosine scan with fused per-warp top-K (small query batches, and the
// fp32 parity path).
//
// Replaces, for a handful of queries at a time, cosine_similarity(q, G) + np.argsort(row)[::-1][:k]
// (reference Evaluate/retrieval_overlap.py:85,90) and the per-candidate cosine of
// DLSRetrievalEngine.retrieve (Retrieval/retrieval.py:203-207,224-226), without ever writing the
// (Q, N) score matrix.
//
// Layout / roofline: the gallery (n, d_pad) is streamed exactly once per query group with 128-bit
// `ld.global.nc.L1::no_allocate` loads, one warp per row block (a row of 512 bf16 = 1 KiB = two fully
// coalesced 512-byte warp requests).  Queries live in registers as fp32; dot products are fp32 FMAs
// reduced with warp shuffles.  Algorithmic bytes per query group = n * d_pad * sizeof(T) + 4 n
// (inverse norms); the kernel is HBM-bound (B = 1: 1 KiB per 512 FMAs).
//
// Top-K: every warp keeps, per query, an unsorted candidate buffer of 2*kp keys in shared memory
// and a strict threshold tau (score of its current k-th best).  A score enters only if
// s > tau -- a warp-uniform test, because after the shuffle reduction every lane holds s.  Rows are
// visited in increasing order by a warp, so on an exact tie the earlier row (smaller id) is already
// in the buffer and the strict test implements "score desc, row asc".  When the buffer fills, the
// warp bitonic-sorts it and keeps k.  At the end the CTA bitonic-sorts its warps' buffers and
// writes one sorted partial list per (query, CTA); select.cu merges the partial lists.
#include "internal.h"

namespace mmr {
namespace {

template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<__nv_bfloat16> {
  static constexpr int V = 8;  // elements per 16-byte load
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xFFFF0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xFFFF0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xFFFF0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xFFFF0000u);
  }
};
template <>
struct ElemTraits<float> {
  static constexpr int V = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
};

__device__ __forceinline__ uint4 ld_stream_16B(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// CH = 16-byte chunks per lane per row, QB = queries per pass, R = rows in flight per warp.
template <typename T, int CH, int QB, int R>
__global__ void __launch_bounds__(512, 1)
scan_topk_kernel(const T* __restrict__ emb, const float* __restrict__ inv_norm, int64_t n, int d_pad,
                 const float* __restrict__ q, const float* __restrict__ q_inv, int b, int k, int kp,
                 const int64_t* __restrict__ exclude_local, uint64_t* __restrict__ partial) {
  constexpr int V = ElemTraits<T>::V;
  extern __shared__ __align__(16) uint64_t smem_keys[];  // [QB][nwarps][cap]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int cap = 2 * kp;
  const int qbase = blockIdx.y * QB;

  // queries -> registers (fp32), zero beyond d_pad / beyond b
  float qr[QB][CH][V];
  float qinv[QB];
  int64_t excl[QB];
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) {
    const bool qok = qbase + qi < b;
    qinv[qi] = qok ? q_inv[qbase + qi] : 0.f;
    excl[qi] = (qok && exclude_local != nullptr) ? exclude_local[qbase + qi] : -1;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int off = c * 32 * V + lane * V;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        qr[qi][c][e] = (qok && off < d_pad) ? q[static_cast<int64_t>(qbase + qi) * d_pad + off + e] : 0.f;
      }
    }
  }

  uint64_t* buf[QB];
  int cnt[QB];
  float tau[QB];
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) {
    buf[qi] = smem_keys + (static_cast<size_t>(qi) * nwarps + warp) * cap;
    cnt[qi] = 0;
    tau[qi] = -INFINITY;
  }

  const int64_t gw = static_cast<int64_t>(blockIdx.x) * nwarps + warp;
  const int64_t tw = static_cast<int64_t>(gridDim.x) * nwarps;

  for (int64_t r0 = gw * R; r0 < n; r0 += tw * R) {
    uint4 v[R][CH];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = r0 + r;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int off = c * 32 * V + lane * V;
        if (row < n && off < d_pad) {
          v[r][c] = ld_stream_16B(emb + row * d_pad + off);
        } else {
          v[r][c] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    const float ginv_lane = (lane < R && r0 + lane < n) ? __ldg(inv_norm + r0 + lane) : 0.f;

    float acc[R][QB];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) acc[r][qi] = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float f[V];
        ElemTraits<T>::unpack(v[r][c], f);
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) {
#pragma unroll
          for (int e = 0; e < V; ++e) acc[r][qi] = fmaf(f[e], qr[qi][c][e], acc[r][qi]);
        }
      }
    }

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = r0 + r;
      const float gi = __shfl_sync(0xffffffffu, ginv_lane, r);
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) {
        const float dot = warp_sum(acc[r][qi]);
        const float s = (dot * gi) * qinv[qi];
        if (row < n && s > tau[qi] && row != excl[qi] && qbase + qi < b) {  // warp-uniform
          if (lane == 0) buf[qi][cnt[qi]] = make_key(s, static_cast<uint32_t>(row));
          ++cnt[qi];
          if (cnt[qi] == cap) {
            __syncwarp();
            warp_bitonic_sort_desc(buf[qi], cap, lane);
            cnt[qi] = k;
            tau[qi] = key_score(buf[qi][k - 1]);
            __syncwarp();
          }
        }
      }
    }
  }

  // pad the unused tail with the sentinel, then one block-wide sort per query
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) {
    __syncwarp();
    for (int i = cnt[qi] + lane; i < cap; i += 32) buf[qi][i] = 0ull;
  }
  __syncthreads();
  const int per_q = nwarps * cap;
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) {
    if (qbase + qi >= b) break;
    uint64_t* s = smem_keys + static_cast<size_t>(qi) * per_q;
    block_bitonic_sort_desc(s, per_q);
    uint64_t* out = partial + (static_cast<int64_t>(qbase + qi) * gridDim.x + blockIdx.x) * kp;
    for (int i = threadIdx.x; i < kp; i += blockDim.x) out[i] = (i < k) ? s[i] : 0ull;
  }
}

template <typename T, int CH, int QB>
int launch_one(const T* emb, const float* inv_norm, int64_t n, int d_pad, const float* q, const float* q_inv,
               int b, int k, const int64_t* excl, const ScanPlan& plan, uint64_t* partial, cudaStream_t stream) {
  // rows in flight per warp: 8 x 1 KiB for the batch-1 bf16 d=512 case (16 warps => 128 KiB per SM in
  // flight), fewer when the query registers (QB x CH x V) leave less room
  constexpr int kRowRegs = (QB == 1) ? 16 : 8;  // uint4 registers spent on in-flight rows
  constexpr int R = (kRowRegs / CH) < 1 ? 1 : (kRowRegs / CH > 8 ? 8 : kRowRegs / CH);
  auto kern = scan_topk_kernel<T, CH, QB, R>;
  MMR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.smem)));
  dim3 grid(plan.n_parts, (b + QB - 1) / QB);
  kern<<<grid, plan.warps * 32, plan.smem, stream>>>(emb, inv_norm, n, d_pad, q, q_inv, b, k, plan.kp, excl,
                                                     partial);
  MMR_LAUNCHED();
  return MMR_OK;
}

template <typename T, int CH>
int dispatch_qb(const T* emb, const float* inv_norm, int64_t n, int d_pad, const float* q, const float* q_inv,
                int b, int k, const int64_t* excl, const ScanPlan& plan, uint64_t* partial, cudaStream_t stream) {
  constexpr int V = ElemTraits<T>::V;
  constexpr int kMaxQB = (64 / (CH * V)) >= 4 ? 4 : ((64 / (CH * V)) >= 2 ? 2 : 1);
  if (plan.qb == 4) {
    if constexpr (kMaxQB >= 4) return launch_one<T, CH, 4>(emb, inv_norm, n, d_pad, q, q_inv, b, k, excl, plan, partial, stream);
  }
  if (plan.qb == 2) {
    if constexpr (kMaxQB >= 2) return launch_one<T, CH, 2>(emb, inv_norm, n, d_pad, q, q_inv, b, k, excl, plan, partial, stream);
  }
  if (plan.qb == 1) return launch_one<T, CH, 1>(emb, inv_norm, n, d_pad, q, q_inv, b, k, excl, plan, partial, stream);
  return fail(MMR_EINVAL, "scan: inconsistent plan (qb)");
}

template <typename T>
int dispatch_ch(const T* emb, const float* inv_norm, int64_t n, int d_pad, const float* q, const float* q_inv,
                int b, int k, const int64_t* excl, const ScanPlan& plan, uint64_t* partial, cudaStream_t stream) {
  constexpr int V = ElemTraits<T>::V;
  const int ch = (d_pad + 32 * V - 1) / (32 * V);
  if (ch <= 1) return dispatch_qb<T, 1>(emb, inv_norm, n, d_pad, q, q_inv, b, k, excl, plan, partial, stream);
  if (ch <= 2) return dispatch_qb<T, 2>(emb, inv_norm, n, d_pad, q, q_inv, b, k, excl, plan, partial, stream);
  if (ch <= 4) return dispatch_qb<T, 4>(emb, inv_norm, n, d_pad, q, q_inv, b, k, excl, plan, partial, stream);
  if (ch <= 8) return dispatch_qb<T, 8>(emb, inv_norm, n, d_pad, q, q_inv, b, k, excl, plan, partial, stream);
  return fail(MMR_EUNSUP, "scan: embedding dimension too large for the scan kernel (bf16: d <= 2048, fp32: d <= 1024)");
}

}  // namespace

int plan_scan(int64_t n, int d_pad, int dtype_store, int b, int k, int num_sms, ScanPlan* plan) {
  if (k < 1 || k > MMR_MAX_K) return fail(MMR_EUNSUP, "scan: k must be in [1, " + std::to_string(MMR_MAX_K) + "]");
  const int V = dtype_store == MMR_BF16 ? 8 : 4;
  const int ch_raw = (d_pad + 32 * V - 1) / (32 * V);
  if (ch_raw > 8) return fail(MMR_EUNSUP, "scan: embedding dimension too large for the scan kernel");
  const int ch = ch_raw <= 1 ? 1 : (ch_raw <= 2 ? 2 : (ch_raw <= 4 ? 4 : 8));
  const int kp = next_pow2(k) < 16 ? 16 : next_pow2(k);
  const int cap = 2 * kp;
  int max_qb = 64 / (ch * V);
  max_qb = max_qb >= 4 ? 4 : (max_qb >= 2 ? 2 : 1);
  // shared-memory budget: qb * warps * cap * 8 bytes <= 128 KiB
  int warps = 16;
  int qb = max_qb;
  if (b < qb) qb = b >= 2 ? 2 : 1;
  if (qb > max_qb) qb = max_qb;
  const size_t budget = 128 * 1024;
  while (static_cast<size_t>(qb) * warps * cap * 8 > budget && qb > 1) qb >>= 1;
  while (static_cast<size_t>(qb) * warps * cap * 8 > budget && warps > 4) warps >>= 1;
  if (static_cast<size_t>(qb) * warps * cap * 8 > 200 * 1024) return fail(MMR_EUNSUP, "scan: k too large");
  // enough rows per warp to amortise the final per-CTA sort; at most one CTA per SM
  int64_t parts = (n + static_cast<int64_t>(warps) * 64 - 1) / (static_cast<int64_t>(warps) * 64);
  if (parts < 1) parts = 1;
  if (parts > num_sms) parts = num_sms;
  plan->n_parts = static_cast<int>(parts);
  plan->kp = kp;
  plan->warps = warps;
  plan->qb = qb;
  plan->smem = static_cast<size_t>(qb) * warps * cap * 8;
  plan->partial_bytes = static_cast<size_t>(b) * plan->n_parts * kp * sizeof(uint64_t);
  return MMR_OK;
}

int launch_scan(const void* emb, int dtype_store, const float* inv_norm, int64_t n, int d_pad, const float* q_f32,
                const float* q_inv, int b, int k, const int64_t* exclude_local, const ScanPlan& plan,
                uint64_t* partial, cudaStream_t stream) {
  if (dtype_store == MMR_BF16) {
    return dispatch_ch<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(emb), inv_norm, n, d_pad, q_f32, q_inv, b,
                                      k, exclude_local, plan, partial, stream);
  }
  return dispatch_ch<float>(static_cast<const float*>(emb), inv_norm, n, d_pad, q_f32, q_inv, b, k, exclude_local,
                            plan, partial, stream);
}

}  // namespace mmr
