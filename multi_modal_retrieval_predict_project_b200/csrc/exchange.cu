// Multi-GPU exchange over NVLink peer memory (one process per GPU, gallery row-sharded).
//
// The sharded search has one exchange step: every rank holds a local top-K
// per query and the global top-K is the merge of the G local lists.  Instead of an NCCL all-gather
// followed by a merge that every rank repeats for every query, the query batch is split across the
// ranks ("owner" of query q = q / ceil(b / G)) and the exchange is fused into the kernels:
//
//   scatter   ONE kernel per rank computes the fp32 embedding cosine of its K local candidates
//             (the rerank feature of Retrieval/reranker.py:298, evaluated by the rank that owns
//             the gallery row) and stores {score, cosine, global row} for query q straight into
//             the OWNER's exchange buffer with 128-bit NVLink stores -- the all-to-all is the
//             kernel's epilogue.  A one-warp signal kernel then raises this rank's flag on every peer.
//   merge     the owner waits for the G flags (a spinning one-warp kernel, stream-ordered), merges
//             its slice of the queries out of the exchange buffer and reranks it.
//   publish   the owner stores its slice of the final (ids, scores) into every rank's result buffer
//             (peer stores again) and signals; collect waits for the G result flags.
//
// Buffers are double buffered by step parity: a rank can only start step s + 2 after it has seen
// every peer's step s + 1 signal, which each peer raises (in stream order) after it finished
// reading step s.  Peer mappings come from CUDA IPC handles exchanged once through
// torch.distributed (plumbing); no NCCL call is on the data path.
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.h"

namespace mmr {
constexpr int kMaxWorld = 16;
struct PeerTable {  // base address of every rank's region as mapped into THIS process
  uint8_t* base[kMaxWorld];
};
}  // namespace mmr
using mmr::kMaxWorld;
using mmr::PeerTable;

struct mmr_exchange {
  int device = 0;
  int rank = 0, world = 1;
  int b_max = 0, k_max = 0;
  size_t blob_bytes = 0;    // per parity: world * per_max * kp_max * 16
  size_t result_bytes = 0;  // per parity: world * per_max * k_max * 16
  size_t flags_off = 0, region_bytes = 0;
  uint8_t* local = nullptr;
  PeerTable peers{};
  std::vector<void*> opened;
  bool open = false;
};

namespace mmr {
namespace {

__device__ __forceinline__ float xbf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float xbf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// flags: uint32 [2 kinds][world]; kind 0 = blob of step arrived from src, kind 1 = result arrived from src
__global__ void exchange_signal_kernel(PeerTable peers, size_t flags_off, int kind, int world, int my_rank,
                                       uint32_t step) {
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();  // everything this stream wrote to the peers before this kernel is ordered first
    uint32_t* flag = reinterpret_cast<uint32_t*>(peers.base[r] + flags_off) + kind * world + my_rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
  }
}

__global__ void exchange_wait_kernel(const uint32_t* __restrict__ flags, int world, uint32_t step) {
  const int r = threadIdx.x;
  if (r < world) {
    uint32_t v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
      if (static_cast<int32_t>(v - step) >= 0) break;  // flags only grow (wrap-safe compare)
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

// One CTA per query.  cosine(q, gallery row) of the K local candidates (fp32, safe_cos formula
// dot / (||a|| * ||b||), 128-bit gathers, two candidates per warp in flight), staged in shared
// memory, then the query's {scores, cosines, rows} lists go to the owner rank's buffer.
template <int kIts>
__global__ void __launch_bounds__(256)
exchange_cos_scatter_kernel(const __nv_bfloat16* __restrict__ emb, int64_t n, int d_pad, int64_t row_offset,
                            const float* __restrict__ q_emb, int d, const int64_t* __restrict__ rows,
                            const float* __restrict__ scores, int k, int kp, int per, int my_rank, PeerTable peers,
                            size_t off_scores, size_t off_cos, size_t off_rows) {
  __shared__ __align__(16) float qs[kIts * 256];
  __shared__ __align__(16) float s_cos[MMR_MAX_K];
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nv = d_pad >> 3;
  for (int i = threadIdx.x; i < kIts * 256; i += blockDim.x) qs[i] = i < d ? q_emb[static_cast<int64_t>(q) * d + i] : 0.f;
  __syncthreads();
  float qss = 0.f;
#pragma unroll
  for (int it = 0; it < kIts; ++it) {
    const float4 a = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8);
    const float4 c = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8 + 4);
    qss = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, qss))));
    qss = fmaf(c.x, c.x, fmaf(c.y, c.y, fmaf(c.z, c.z, fmaf(c.w, c.w, qss))));
  }
  qss = warp_sum(qss);
  const int64_t base = static_cast<int64_t>(q) * k;
  for (int j0 = warp * 2; j0 < k; j0 += nwarps * 2) {
    uint4 x[2][kIts];
    bool have[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = j0 + c;
      int64_t local = j < k ? rows[base + j] - row_offset : -1;
      have[c] = j < k && local >= 0 && local < n;
      const uint4* ce = reinterpret_cast<const uint4*>(emb + (have[c] ? local : 0) * d_pad);
#pragma unroll
      for (int it = 0; it < kIts; ++it) {
        const int u = it * 32 + lane;
        x[c][it] = (have[c] && u < nv) ? __ldg(ce + u) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float dot[2] = {0.f, 0.f}, css[2] = {0.f, 0.f};
#pragma unroll
    for (int it = 0; it < kIts; ++it) {
      const float4 qa = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8);
      const float4 qb = *reinterpret_cast<const float4*>(qs + (it * 32 + lane) * 8 + 4);
      const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t w[4] = {x[c][it].x, x[c][it].y, x[c][it].z, x[c][it].w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float lo = xbf16_lo(w[h]), hi = xbf16_hi(w[h]);
          dot[c] = fmaf(lo, qv[2 * h], dot[c]);
          css[c] = fmaf(lo, lo, css[c]);
          dot[c] = fmaf(hi, qv[2 * h + 1], dot[c]);
          css[c] = fmaf(hi, hi, css[c]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        dot[c] += __shfl_xor_sync(0xffffffffu, dot[c], o);
        css[c] += __shfl_xor_sync(0xffffffffu, css[c], o);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (j0 + c < k) {
          const float na = sqrtf(qss), nb = sqrtf(css[c]);
          s_cos[j0 + c] = (have[c] && na != 0.f && nb != 0.f) ? dot[c] / (na * nb) : 0.f;
        }
      }
    }
  }
  __syncthreads();
  // ---- the exchange: this query's lists -> the owner rank's buffer (16-byte NVLink stores) ----
  const int dest = q / per;
  const int64_t slot = (static_cast<int64_t>(my_rank) * per + (q - dest * per)) * kp;  // element offset
  uint8_t* const dbase = peers.base[dest];
  float4* const d_sc = reinterpret_cast<float4*>(dbase + off_scores + slot * 4);
  float4* const d_co = reinterpret_cast<float4*>(dbase + off_cos + slot * 4);
  longlong2* const d_ro = reinterpret_cast<longlong2*>(dbase + off_rows + slot * 8);
  for (int v = threadIdx.x; v < kp / 4; v += blockDim.x) {
    float sc[4], co[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = v * 4 + e;
      sc[e] = j < k ? scores[base + j] : -INFINITY;
      co[e] = j < k ? s_cos[j] : 0.f;
    }
    d_sc[v] = make_float4(sc[0], sc[1], sc[2], sc[3]);
    d_co[v] = make_float4(co[0], co[1], co[2], co[3]);
  }
  for (int v = threadIdx.x; v < kp / 2; v += blockDim.x) {
    const int j = v * 2;
    d_ro[v] = make_longlong2(j < k ? rows[base + j] : -1, j + 1 < k ? rows[base + j + 1] : -1);
  }
  __threadfence_system();
}

// my slice of the final (ids, combined scores) -> every rank's result buffer
__global__ void exchange_publish_kernel(const int64_t* __restrict__ ids, const double* __restrict__ fin, int64_t count,
                                        int64_t dst_elem_off, int world, PeerTable peers, size_t off_ids,
                                        size_t off_fin) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) {
    const int64_t id = ids[i];
    const double f = fin[i];
    for (int r = 0; r < world; ++r) {
      reinterpret_cast<int64_t*>(peers.base[r] + off_ids)[dst_elem_off + i] = id;
      reinterpret_cast<double*>(peers.base[r] + off_fin)[dst_elem_off + i] = f;
    }
  }
  __threadfence_system();
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace
}  // namespace mmr

using namespace mmr;

namespace {
struct Layout {
  int per, kp;
  size_t off_scores, off_cos, off_rows;  // within the region, for the given parity
  size_t off_ids, off_fin;
};
Layout layout_for(const mmr_exchange* ex, int b, int k, uint32_t step) {
  Layout L;
  L.per = (b + ex->world - 1) / ex->world;
  L.kp = (k + 3) / 4 * 4;
  const size_t parity = step & 1u;
  const size_t lists = static_cast<size_t>(ex->world) * L.per * L.kp;
  const size_t blob0 = parity * ex->blob_bytes;
  L.off_scores = blob0;
  L.off_cos = blob0 + align_up(lists * 4, 256);
  L.off_rows = L.off_cos + align_up(lists * 4, 256);
  const size_t res0 = 2 * ex->blob_bytes + parity * ex->result_bytes;
  L.off_ids = res0;
  L.off_fin = res0 + align_up(static_cast<size_t>(ex->world) * L.per * k * 8, 256);
  return L;
}
int check_sizes(const mmr_exchange* ex, int b, int k, const char* who) {
  if (ex == nullptr || !ex->open) return fail(MMR_EINVAL, std::string(who) + ": exchange is not open");
  if (b < 1 || k < 1 || b > ex->b_max || k > ex->k_max)
    return fail(MMR_EINVAL, std::string(who) + ": batch / k exceed the sizes the exchange was created for");
  return MMR_OK;
}
}  // namespace

extern "C" {

int mmr_exchange_create(mmr_exchange** out, int32_t rank, int32_t world, int32_t b_max, int32_t k_max, int32_t device) {
  MMR_REQUIRE(out != nullptr, "mmr_exchange_create: out is NULL");
  MMR_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "mmr_exchange_create: bad rank / world");
  MMR_REQUIRE(b_max >= 1 && k_max >= 1 && k_max <= MMR_MAX_K, "mmr_exchange_create: bad sizes");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(MMR_ENODEV, "mmr_exchange_create: cannot select the device (no CPU fallback)");
  mmr_exchange* ex = new mmr_exchange();
  ex->device = device;
  ex->rank = rank;
  ex->world = world;
  ex->b_max = b_max;
  ex->k_max = k_max;
  const size_t per = (static_cast<size_t>(b_max) + world - 1) / world;
  const size_t kp = (static_cast<size_t>(k_max) + 3) / 4 * 4;
  ex->blob_bytes = align_up(3 * 256 + static_cast<size_t>(world) * per * kp * 16, 256);
  ex->result_bytes = align_up(2 * 256 + static_cast<size_t>(world) * per * k_max * 16, 256);
  ex->flags_off = 2 * ex->blob_bytes + 2 * ex->result_bytes;
  ex->region_bytes = ex->flags_off + 256;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->local), ex->region_bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete ex;
    return fail(MMR_ENOMEM, std::string("mmr_exchange_create: cudaMalloc: ") + cudaGetErrorString(e));
  }
  e = cudaMemset(ex->local, 0, ex->region_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(ex->local);
    delete ex;
    return fail(MMR_ECUDA, std::string("mmr_exchange_create: ") + cudaGetErrorString(e));
  }
  for (int r = 0; r < kMaxWorld; ++r) ex->peers.base[r] = nullptr;
  ex->peers.base[rank] = ex->local;
  ex->open = world == 1;
  *out = ex;
  return MMR_OK;
}

int mmr_exchange_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

int mmr_exchange_handle(mmr_exchange* ex, void* handle_out) {
  MMR_REQUIRE(ex != nullptr && handle_out != nullptr, "mmr_exchange_handle: NULL argument");
  DeviceGuard guard(ex->device);
  cudaIpcMemHandle_t h;
  MMR_CUDA_TRY(cudaIpcGetMemHandle(&h, ex->local));
  memcpy(handle_out, &h, sizeof(h));
  return MMR_OK;
}

int mmr_exchange_open(mmr_exchange* ex, const void* handles) {
  MMR_REQUIRE(ex != nullptr && handles != nullptr, "mmr_exchange_open: NULL argument");
  if (ex->open) return MMR_OK;
  DeviceGuard guard(ex->device);
  const uint8_t* hb = static_cast<const uint8_t*>(handles);
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + static_cast<size_t>(r) * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(MMR_ECUDA, "mmr_exchange_open: cudaIpcOpenMemHandle for rank " + std::to_string(r) + ": " +
                                 cudaGetErrorString(e) + " (peer access over NVLink is required)");
    }
    ex->opened.push_back(p);
    ex->peers.base[r] = static_cast<uint8_t*>(p);
  }
  ex->open = true;
  return MMR_OK;
}

int mmr_exchange_destroy(mmr_exchange* ex) {
  if (ex == nullptr) return MMR_OK;
  DeviceGuard guard(ex->device);
  cudaDeviceSynchronize();
  for (void* p : ex->opened) cudaIpcCloseMemHandle(p);
  if (ex->local) cudaFree(ex->local);
  cudaGetLastError();
  delete ex;
  return MMR_OK;
}

int mmr_exchange_scatter(mmr_exchange* ex, const mmr_index* ix, const float* q_emb, const int64_t* rows,
                         const float* scores, int32_t b, int32_t k, uint32_t step, void* stream_v) {
  MMR_TRY(check_sizes(ex, b, k, "mmr_exchange_scatter"));
  MMR_REQUIRE(ix && q_emb && rows && scores, "mmr_exchange_scatter: NULL argument");
  if (!(is_device_ptr(q_emb) && is_device_ptr(rows) && is_device_ptr(scores)))
    return fail(MMR_EINVAL, "mmr_exchange_scatter: device pointers only");
  int64_t n = 0, row_offset = 0, bytes = 0;
  int32_t d = 0, d_pad = 0, dtype = 0, device = 0;
  MMR_TRY(mmr_index_info(ix, &n, &d, &d_pad, &dtype, &device, &row_offset, &bytes));
  const void* emb = nullptr;
  const float* inv = nullptr;
  MMR_TRY(mmr_index_device_ptrs(ix, &emb, &inv));
  if (dtype != MMR_BF16 || d_pad > 1024 || (reinterpret_cast<uintptr_t>(emb) & 15u) != 0)
    return fail(MMR_EUNSUP, "mmr_exchange_scatter: needs a bf16 index with d <= 1024");
  if (device != ex->device) return fail(MMR_EINVAL, "mmr_exchange_scatter: index and exchange live on different devices");
  DeviceGuard guard(ex->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const Layout L = layout_for(ex, b, k, step);
  const __nv_bfloat16* e16 = static_cast<const __nv_bfloat16*>(emb);
  const int its = (d_pad + 255) / 256;
#define MMR_XCHG(ITS)                                                                                              \
  exchange_cos_scatter_kernel<ITS><<<b, 256, 0, stream>>>(e16, n, d_pad, row_offset, q_emb, d, rows, scores, k, L.kp, \
                                                          L.per, ex->rank, ex->peers, L.off_scores, L.off_cos,       \
                                                          L.off_rows)
  if (its <= 1) {
    MMR_XCHG(1);
  } else if (its == 2) {
    MMR_XCHG(2);
  } else {
    MMR_XCHG(4);
  }
#undef MMR_XCHG
  MMR_LAUNCHED();
  exchange_signal_kernel<<<1, 32, 0, stream>>>(ex->peers, ex->flags_off, 0, ex->world, ex->rank, step);
  MMR_LAUNCHED();
  return MMR_OK;
}

int mmr_exchange_merge(mmr_exchange* ex, int32_t b, int32_t k, uint32_t step, float* out_scores, int64_t* out_rows,
                       float* out_cos, void* stream_v) {
  MMR_TRY(check_sizes(ex, b, k, "mmr_exchange_merge"));
  DeviceGuard guard(ex->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const Layout L = layout_for(ex, b, k, step);
  exchange_wait_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const uint32_t*>(ex->local + ex->flags_off), ex->world,
                                             step);
  MMR_LAUNCHED();
  const int q_lo = std::min(b, ex->rank * L.per), q_hi = std::min(b, (ex->rank + 1) * L.per);
  const int nloc = q_hi - q_lo;
  if (nloc <= 0) return MMR_OK;
  MMR_REQUIRE(out_scores && out_rows && out_cos, "mmr_exchange_merge: NULL output");
  if (!(is_device_ptr(out_scores) && is_device_ptr(out_rows) && is_device_ptr(out_cos)))
    return fail(MMR_EINVAL, "mmr_exchange_merge: device pointers only");
  const float* sc = reinterpret_cast<const float*>(ex->local + L.off_scores);
  const float* co = reinterpret_cast<const float*>(ex->local + L.off_cos);
  const int64_t* ro = reinterpret_cast<const int64_t*>(ex->local + L.off_rows);
  int32_t* src = nullptr;
  MMR_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&src), static_cast<size_t>(nloc) * k * sizeof(int32_t), stream));
  const int64_t stride = static_cast<int64_t>(L.per) * L.kp;
  int st = launch_merge_lists(sc, ro, ex->world, nloc, L.kp, stride, stride, k, out_scores, out_rows, src, stream);
  if (st == MMR_OK) st = launch_gather_payload(co, stride, src, nloc, L.kp, k, out_cos, stream);
  cudaFreeAsync(src, stream);
  return st;
}

int mmr_exchange_publish(mmr_exchange* ex, const int64_t* ids, const double* fin, int32_t b, int32_t keep,
                         uint32_t step, void* stream_v) {
  MMR_TRY(check_sizes(ex, b, keep, "mmr_exchange_publish"));
  DeviceGuard guard(ex->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const Layout L = layout_for(ex, b, keep, step);
  const int q_lo = std::min(b, ex->rank * L.per), q_hi = std::min(b, (ex->rank + 1) * L.per);
  const int64_t count = static_cast<int64_t>(q_hi - q_lo) * keep;
  if (count > 0) {
    MMR_REQUIRE(ids && fin, "mmr_exchange_publish: NULL argument");
    if (!(is_device_ptr(ids) && is_device_ptr(fin))) return fail(MMR_EINVAL, "mmr_exchange_publish: device pointers only");
    exchange_publish_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, stream>>>(
        ids, fin, count, static_cast<int64_t>(q_lo) * keep, ex->world, ex->peers, L.off_ids, L.off_fin);
    MMR_LAUNCHED();
  }
  exchange_signal_kernel<<<1, 32, 0, stream>>>(ex->peers, ex->flags_off, 1, ex->world, ex->rank, step);
  MMR_LAUNCHED();
  return MMR_OK;
}

int mmr_exchange_collect(mmr_exchange* ex, int32_t b, int32_t keep, uint32_t step, const int64_t** ids,
                         const double** fin, void* stream_v) {
  MMR_TRY(check_sizes(ex, b, keep, "mmr_exchange_collect"));
  MMR_REQUIRE(ids && fin, "mmr_exchange_collect: NULL argument");
  DeviceGuard guard(ex->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const Layout L = layout_for(ex, b, keep, step);
  exchange_wait_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const uint32_t*>(ex->local + ex->flags_off) + ex->world,
                                             ex->world, step);
  MMR_LAUNCHED();
  *ids = reinterpret_cast<const int64_t*>(ex->local + L.off_ids);
  *fin = reinterpret_cast<const double*>(ex->local + L.off_fin);
  return MMR_OK;
}

}  // extern "C"
