// Gallery / query ingest: dtype conversion, zero padding to d_pad, inverse L2 norms.
//
// Replaces the load step of RetrievalEngine.__init__ (reference Retrieval/retrieval.py:24-32) plus
// the row normalisation that sklearn's cosine_similarity redoes on EVERY call
// (normalize(): row_norms = sqrt(einsum('ij,ij->i')), zero norms -> 1).  Here the norm is taken
// once; a zero row gets inverse norm 0 so it scores exactly 0 like in sklearn.
//
// HBM-bound streaming kernel: one warp per row, 128-bit loads/stores, shuffle reduction.
#include "internal.h"

namespace mmr {
namespace {

__device__ __forceinline__ float load_elem(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float load_elem(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float store_elem(float* p, int64_t i, float v) {
  p[i] = v;
  return v;
}
__device__ __forceinline__ float store_elem(__nv_bfloat16* p, int64_t i, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  p[i] = h;
  return __bfloat162float(h);
}
__device__ __forceinline__ float round_to(float v, const float*) { return v; }
__device__ __forceinline__ float round_to(float v, const __nv_bfloat16*) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// Generic (scalar) path: any d, any alignment.
template <typename TIn, typename TOut>
__global__ void ingest_rows_kernel(const TIn* __restrict__ src, int64_t n, int d, int64_t src_ld,
                                   TOut* __restrict__ dst, int d_pad, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    const TIn* s = src + r * src_ld;
    float ss = 0.f;
    for (int i = lane; i < d_pad; i += 32) {
      float v = (i < d) ? load_elem(s, i) : 0.f;
      if (dst != nullptr) {
        v = store_elem(dst + r * static_cast<int64_t>(d_pad), i, v);
      } else {
        v = round_to(v, dst);
      }
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) inv_norm[r] = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
  }
}

// Vector path fp32 -> bf16 (the common ingest): d % 8 == 0 and 16-byte aligned rows.
__global__ void ingest_f32_to_bf16_vec_kernel(const float* __restrict__ src, int64_t n, int d, int64_t src_ld,
                                              __nv_bfloat16* __restrict__ dst, int d_pad,
                                              float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    const float4* s = reinterpret_cast<const float4*>(src + r * src_ld);
    uint4* o = reinterpret_cast<uint4*>(dst + r * static_cast<int64_t>(d_pad));
    float ss = 0.f;
    for (int c = lane; c < d_pad / 8; c += 32) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (c * 8 < d) {
        a = __ldg(s + 2 * c);
        b = __ldg(s + 2 * c + 1);
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      float2 f0 = __bfloat1622float2(p0), f1 = __bfloat1622float2(p1), f2 = __bfloat1622float2(p2),
             f3 = __bfloat1622float2(p3);
      ss = fmaf(f0.x, f0.x, ss); ss = fmaf(f0.y, f0.y, ss);
      ss = fmaf(f1.x, f1.x, ss); ss = fmaf(f1.y, f1.y, ss);
      ss = fmaf(f2.x, f2.x, ss); ss = fmaf(f2.y, f2.y, ss);
      ss = fmaf(f3.x, f3.x, ss); ss = fmaf(f3.y, f3.y, ss);
      uint4 w;
      w.x = *reinterpret_cast<uint32_t*>(&p0);
      w.y = *reinterpret_cast<uint32_t*>(&p1);
      w.z = *reinterpret_cast<uint32_t*>(&p2);
      w.w = *reinterpret_cast<uint32_t*>(&p3);
      o[c] = w;
    }
    ss = warp_sum(ss);
    if (lane == 0) inv_norm[r] = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
  }
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ emb, int64_t n, int d, int d_pad, int64_t row_offset,
                                   const int64_t* __restrict__ rows, int64_t m, float* __restrict__ out) {
  for (int64_t r = blockIdx.x; r < m; r += gridDim.x) {
    int64_t local = rows[r] - row_offset;
    bool ok = local >= 0 && local < n;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      out[r * d + i] = ok ? load_elem(emb + local * d_pad, i) : 0.f;
    }
  }
}

}  // namespace

int launch_ingest(const void* src, int dtype_in, int64_t n, int d, int64_t src_ld, void* dst, int dtype_store,
                  int d_pad, float* inv_norm, cudaStream_t stream) {
  if (n == 0) return MMR_OK;
  const int threads = 256;
  int64_t want = (n + (threads / 32) - 1) / (threads / 32);
  int blocks = static_cast<int>(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
  bool vec_ok = dtype_in == MMR_F32 && dtype_store == MMR_BF16 && dst != nullptr && d % 8 == 0 &&
                src_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
  if (vec_ok) {
    ingest_f32_to_bf16_vec_kernel<<<blocks, threads, 0, stream>>>(
        static_cast<const float*>(src), n, d, src_ld, static_cast<__nv_bfloat16*>(dst), d_pad, inv_norm);
  } else if (dtype_in == MMR_F32 && dtype_store == MMR_F32) {
    ingest_rows_kernel<float, float><<<blocks, threads, 0, stream>>>(
        static_cast<const float*>(src), n, d, src_ld, static_cast<float*>(dst), d_pad, inv_norm);
  } else if (dtype_in == MMR_F32 && dtype_store == MMR_BF16) {
    ingest_rows_kernel<float, __nv_bfloat16><<<blocks, threads, 0, stream>>>(
        static_cast<const float*>(src), n, d, src_ld, static_cast<__nv_bfloat16*>(dst), d_pad, inv_norm);
  } else if (dtype_in == MMR_BF16 && dtype_store == MMR_BF16) {
    ingest_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(src), n, d, src_ld, static_cast<__nv_bfloat16*>(dst), d_pad, inv_norm);
  } else if (dtype_in == MMR_BF16 && dtype_store == MMR_F32) {
    ingest_rows_kernel<__nv_bfloat16, float><<<blocks, threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(src), n, d, src_ld, static_cast<float*>(dst), d_pad, inv_norm);
  } else {
    return fail(MMR_EINVAL, "ingest: unsupported dtype combination");
  }
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_gather_rows(const void* emb, int dtype_store, int64_t n, int d, int d_pad, int64_t row_offset,
                       const int64_t* rows, int64_t m, float* out, cudaStream_t stream) {
  if (m == 0) return MMR_OK;
  int blocks = static_cast<int>(m < 148 * 16 ? m : 148 * 16);
  if (dtype_store == MMR_F32) {
    gather_rows_kernel<float><<<blocks, 128, 0, stream>>>(static_cast<const float*>(emb), n, d, d_pad,
                                                          row_offset, rows, m, out);
  } else {
    gather_rows_kernel<__nv_bfloat16><<<blocks, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(emb), n, d,
                                                                  d_pad, row_offset, rows, m, out);
  }
  MMR_LAUNCHED();
  return MMR_OK;
}

}  // namespace mmr
