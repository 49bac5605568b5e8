// Multi-GPU exchange over NVLink peer memory (one process per GPU, gallery row-sharded).
//
// The sharded search has one exchange step: every rank holds a local top-K per query
// and the global top-K is the merge of the G local lists.  Instead of an NCCL all-gather followed by a merge
// and a rerank that every rank repeats for every query, the query batch is split across the ranks ("owner" of
// query q = q / ceil(b / G)) and the exchange is the epilogue / prologue of the kernels on either side of it:
//
//   scatter   mmr_search_scatter: the search's selection kernel (select.cu, RemoteSink) stores every query's
//             top-K {score, global row} straight into the OWNER's region with NVLink stores -- the all-to-all
//             is that kernel's output -- and a one-warp kernel raises this rank's flag on every peer.
//   rerank    mmr_exchange_rerank: ONE kernel per owner, one CTA per owned query: wait for the G flags
//             (bounded spin), merge the G lists, label / KG features, min-max, combine, order
//             (rerank_tail.cuh = Reranker.rerank, reference Retrieval/reranker.py:298-329), store the query's
//             final (ids, scores) into EVERY rank's result buffer; the last CTA raises the result flags.
//   collect   a one-warp kernel waits for the G result flags; the full (b, keep) result then sits in this
//             rank's region.
//
// Regions are double buffered by step parity: a rank can only start step s + 2 after it has seen every peer's
// step s + 1 result flag, which each peer raises (in stream order) after it finished reading step s.  Peer
// mappings come from CUDA IPC handles exchanged once through torch.distributed (plumbing); no collective
// library call is on the data path.
//
// Failure handling: every wait is bounded (mmr_exchange_set_timeout, default 20 s) and checks an abort word
// that mmr_exchange_abort raises on all peers; a wait that gives up -- or finds that a peer ran the step
// with a different (b, k) -- writes a code into a host-mapped error word, which the next call on the handle
// (and mmr_exchange_status) reports as MMR_ECUDA instead of hanging the GPU.
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.h"
#include "rerank_tail.cuh"

using mmr::kMaxWorld;
using mmr::PeerSink;
using mmr::PeerTable;

namespace mmr {
// control block at ctl_off of every region
struct Ctl {
  uint32_t list_flag[kMaxWorld];  // step of the last list block received from rank r
  uint32_t res_flag[kMaxWorld];   // step of the last result slice received from rank r
  uint32_t meta[kMaxWorld];       // (b, k) fingerprint rank r ran that step with
  uint32_t abort;                 // != 0: give up waiting
  uint32_t done[2];               // per parity: CTAs of the rerank kernel that have published
};
enum : uint32_t { kErrNone = 0, kErrTimeoutLists = 1, kErrTimeoutResults = 2, kErrMismatch = 3, kErrAborted = 4 };
}  // namespace mmr
using mmr::Ctl;

struct mmr_exchange {
  int device = 0;
  int rank = 0, world = 1;
  int b_max = 0, k_max = 0;
  size_t list_bytes = 0;    // per parity: scores + rows, world * per_max * kp_max entries each
  size_t result_bytes = 0;  // per parity: ids + scores, world * per_max * k_max entries each
  size_t ctl_off = 0, region_bytes = 0;
  uint8_t* local = nullptr;
  PeerTable peers{};
  std::vector<void*> opened;
  bool open = false;
  uint32_t* err_host = nullptr;  // host-mapped error word (written by a kernel that gives up)
  uint32_t* err_dev = nullptr;
  uint32_t timeout_ms = 20000;
};

namespace mmr {
namespace {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Lanes r < world of ONE warp wait until flags[r] >= step (wrap-safe), at most timeout_ns, giving up at once
// when the abort word is raised.  With `meta` the fingerprint every source rank published must equal `mine`.
// On failure the code goes to the host-mapped error word.  Returns true (warp-uniform) when everything arrived.
__device__ __forceinline__ bool wait_flags(const uint32_t* flags, const uint32_t* meta, uint32_t mine,
                                           const uint32_t* abort_word, int world, uint32_t step, uint64_t timeout_ns,
                                           uint32_t* err, uint32_t timeout_code) {
  const int r = threadIdx.x & 31;
  uint32_t code = kErrNone;
  if (r < world) {
    const uint64_t t0 = global_ns();
    for (;;) {
      if (static_cast<int32_t>(ld_acquire_sys(flags + r) - step) >= 0) break;  // flags only grow
      if (*reinterpret_cast<const volatile uint32_t*>(abort_word) != 0u) {
        code = kErrAborted;
        break;
      }
      if (global_ns() - t0 > timeout_ns) {
        code = timeout_code;
        break;
      }
      __nanosleep(200);
    }
    if (code == kErrNone && meta != nullptr && *reinterpret_cast<const volatile uint32_t*>(meta + r) != mine)
      code = kErrMismatch;
    if (code != kErrNone) {
      *reinterpret_cast<volatile uint32_t*>(err) = code;
      __threadfence_system();
    }
  }
  return __all_sync(0xffffffffu, code == kErrNone);
}

inline uint32_t fingerprint(int b, int k) {  // host: the (b, k) a rank ran a step with
  return (static_cast<uint32_t>(b) * 2654435761u) ^ (static_cast<uint32_t>(k) * 40503u) ^ 0x9E3779B9u;
}

// kind 0: "my lists of `step` are in your region" (+ the (b, k) fingerprint), kind 1: "my result slice is"
__global__ void exchange_signal_kernel(PeerTable peers, size_t ctl_off, int kind, int world, int my_rank, uint32_t step,
                                       uint32_t meta) {
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();  // everything this stream stored into the peers before this kernel is ordered first
    Ctl* ctl = reinterpret_cast<Ctl*>(peers.base[r] + ctl_off);
    if (kind == 0) {
      *reinterpret_cast<volatile uint32_t*>(&ctl->meta[my_rank]) = meta;
      st_release_sys(&ctl->list_flag[my_rank], step);
    } else {
      st_release_sys(&ctl->res_flag[my_rank], step);
    }
  }
}

__global__ void exchange_wait_kernel(const Ctl* ctl, int kind, int world, uint32_t step, uint64_t timeout_ns,
                                     uint32_t* err) {
  (void)wait_flags(kind == 0 ? ctl->list_flag : ctl->res_flag, nullptr, 0u, &ctl->abort, world, step, timeout_ns, err,
                   kind == 0 ? kErrTimeoutLists : kErrTimeoutResults);
  __threadfence_system();
}

// One CTA per OWNED query: wait -> merge G lists -> rerank tail -> publish to every rank -> (last CTA) signal.
struct RerankArgs {
  const float* l_scores;   // this rank's list block of the step's parity: [world][per][kp]
  const int64_t* l_rows;
  Ctl* ctl;                // this rank's control block
  TailTables t;
  const int64_t* q_rec;    // (b) record index of every query of the batch
  int q_lo, nloc, b, k, kp, per, world, my_rank, keep;
  double alpha, beta, gamma;
  PeerTable peers;
  size_t off_ids, off_fin, ctl_off;
  uint32_t step, meta;
  uint64_t timeout_ns;
  uint32_t* err;
};

template <int kKIts>
__global__ void __launch_bounds__(kTailThreads)
exchange_rerank_kernel(RerankArgs a) {
  __shared__ TailSmem sm;
  __shared__ __align__(16) uint64_t keys[kMaxWorld * kTailMaxK];
  __shared__ uint32_t s_bound;
  __shared__ int s_m, s_ok, s_last;
  const int tid = threadIdx.x;
  const int ql = blockIdx.x;          // query within this rank's slice
  const int q = a.q_lo + ql;          // query within the batch
  if (tid == 0) {
    s_bound = 0u;
    s_m = 0;
  }
  if (tid < 32) {
    const bool ok = wait_flags(a.ctl->list_flag, a.ctl->meta, a.meta, &a.ctl->abort, a.world, a.step, a.timeout_ns,
                               a.err, kErrTimeoutLists);
    if (tid == 0) s_ok = ok ? 1 : 0;
  }
  __syncthreads();
  int count = 0;
  if (s_ok) {
    // ---- merge: a list whose k-th entry exists holds k candidates >= that entry, so nothing below the best such
    // entry can be in the global top-k; the survivors are ranked by counting (keys are unique) ----
    if (tid < a.world) {
      const int64_t at = (static_cast<int64_t>(tid) * a.per + ql) * a.kp + (a.k - 1);
      if (a.l_rows[at] >= 0) atomicMax(&s_bound, f32_to_ordered(a.l_scores[at]));
    }
    __syncthreads();
    const uint32_t bound = s_bound;
    const int total = a.world * a.kp;
    for (int i = tid; i < total; i += kTailThreads) {
      const int l = i / a.kp, j = i - l * a.kp;
      const int64_t at = (static_cast<int64_t>(l) * a.per + ql) * a.kp + j;
      const int64_t row = a.l_rows[at];
      if (row >= 0) {
        const uint64_t key = make_key(a.l_scores[at], static_cast<uint32_t>(row));
        if (static_cast<uint32_t>(key >> 32) >= bound) {
          const int at_key = atomicAdd(&s_m, 1);
#ifdef MMR_DIAG
          if (at_key >= kMaxWorld * kTailMaxK) __trap();  // bounds-checked build
#endif
          keys[at_key] = key;
        }
      }
    }
    __syncthreads();
    const int m = s_m;
    for (int i = tid; i < m; i += kTailThreads) {
      const uint64_t mine = keys[i];
      int r = 0;
      for (int j = 0; j < m; ++j) r += keys[j] > mine ? 1 : 0;
      if (r < a.k) {
        sm.cand_row[r] = static_cast<int64_t>(key_row(mine));
        sm.cand_score[r] = key_score(mine);
      }
    }
    count = m < a.k ? m : a.k;
    __syncthreads();
  }
  // ---- rerank + publish: the query's final (ids, scores) go to every rank's result buffer ----
  const int64_t base = static_cast<int64_t>(q) * a.keep;
  rerank_tail<kKIts>(sm, count, a.q_rec[q], a.t, a.alpha, a.beta, a.gamma, a.keep,
                     [&](int rank, int j, double fin, double, double, double) {
                       const int64_t id = sm.cand_row[j];
                       for (int r = 0; r < a.world; ++r) {
                         reinterpret_cast<int64_t*>(a.peers.base[r] + a.off_ids)[base + rank] = id;
                         reinterpret_cast<double*>(a.peers.base[r] + a.off_fin)[base + rank] = fin;
                       }
                     });
  for (int rank = count + tid; rank < a.keep; rank += kTailThreads) {  // fewer candidates than `keep` (or a failed wait)
    for (int r = 0; r < a.world; ++r) {
      reinterpret_cast<int64_t*>(a.peers.base[r] + a.off_ids)[base + rank] = -1;
      reinterpret_cast<double*>(a.peers.base[r] + a.off_fin)[base + rank] = 0.0;
    }
  }
  // ---- the last CTA of the launch raises this rank's result flag on every peer ----
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(&a.ctl->done[a.step & 1u], 1u);
    s_last = prev + 1u == gridDim.x ? 1 : 0;
    if (s_last) a.ctl->done[a.step & 1u] = 0u;  // ready for step + 2
  }
  __syncthreads();
  if (s_last && tid < a.world) {
    __threadfence_system();
    st_release_sys(&reinterpret_cast<Ctl*>(a.peers.base[tid] + a.ctl_off)->res_flag[a.my_rank], a.step);
  }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Layout {
  int per, kp;
  size_t off_scores, off_rows;  // lists, for the step's parity
  size_t off_ids, off_fin;      // results, for the step's parity
};
Layout layout_for(const mmr_exchange* ex, int b, int k, int keep, uint32_t step) {
  Layout L;
  L.per = (b + ex->world - 1) / ex->world;
  L.kp = (k + 3) / 4 * 4;
  const size_t parity = step & 1u;
  const size_t lists = static_cast<size_t>(ex->world) * L.per * L.kp;
  L.off_scores = parity * ex->list_bytes;
  L.off_rows = L.off_scores + align_up(lists * 4, 256);
  L.off_ids = 2 * ex->list_bytes + parity * ex->result_bytes;
  L.off_fin = L.off_ids + align_up(static_cast<size_t>(ex->world) * L.per * keep * 8, 256);
  return L;
}

int check_call(mmr_exchange* ex, int b, int k, const char* who) {
  if (ex == nullptr || !ex->open) return fail(MMR_EINVAL, std::string(who) + ": exchange is not open");
  if (b < 1 || k < 1 || b > ex->b_max || k > ex->k_max)
    return fail(MMR_EINVAL, std::string(who) + ": batch / k exceed the sizes the exchange was created for");
  const uint32_t code = *reinterpret_cast<volatile uint32_t*>(ex->err_host);
  if (code != kErrNone) {
    static const char* what[] = {"", "timed out waiting for a peer's candidate lists", "timed out waiting for a peer's results",
                                 "a peer ran the step with a different batch size or k", "aborted"};
    return fail(MMR_ECUDA, std::string(who) + ": the exchange failed in an earlier step (" + what[code < 5 ? code : 0] +
                               "); destroy it and create a new one");
  }
  return MMR_OK;
}

}  // namespace
}  // namespace mmr

using namespace mmr;

extern "C" {

int mmr_exchange_create(mmr_exchange** out, int32_t rank, int32_t world, int32_t b_max, int32_t k_max, int32_t device) {
  MMR_REQUIRE(out != nullptr, "mmr_exchange_create: out is NULL");
  *out = nullptr;
  MMR_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "mmr_exchange_create: bad rank / world");
  MMR_REQUIRE(b_max >= 1 && k_max >= 1, "mmr_exchange_create: bad sizes");
  if (k_max > kTailMaxK)
    return fail(MMR_EUNSUP, "mmr_exchange_create: the fused exchange covers k <= 128 (use the all-gather transport)");
  MMR_TRY(check_device(device, nullptr));
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
  ex->list_bytes = align_up(2 * 256 + static_cast<size_t>(world) * per * kp * 12, 256);
  ex->result_bytes = align_up(2 * 256 + static_cast<size_t>(world) * per * k_max * 16, 256);
  ex->ctl_off = 2 * ex->list_bytes + 2 * ex->result_bytes;
  ex->region_bytes = ex->ctl_off + align_up(sizeof(Ctl), 256);
  auto cleanup = [&](int code) {
    if (ex->local) cudaFree(ex->local);
    if (ex->err_host) cudaFreeHost(ex->err_host);
    cudaGetLastError();
    delete ex;
    return code;
  };
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->local), ex->region_bytes);
  if (e != cudaSuccess) return cleanup(fail(MMR_ENOMEM, std::string("mmr_exchange_create: cudaMalloc: ") + cudaGetErrorString(e)));
  e = cudaHostAlloc(reinterpret_cast<void**>(&ex->err_host), sizeof(uint32_t), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *ex->err_host = kErrNone;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&ex->err_dev), ex->err_host, 0);
  }
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->region_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cleanup(fail(MMR_ECUDA, std::string("mmr_exchange_create: ") + cudaGetErrorString(e)));
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

int mmr_exchange_set_timeout(mmr_exchange* ex, int32_t milliseconds) {
  MMR_REQUIRE(ex != nullptr && milliseconds >= 1, "mmr_exchange_set_timeout: bad argument");
  ex->timeout_ms = static_cast<uint32_t>(milliseconds);
  return MMR_OK;
}

int mmr_exchange_status(mmr_exchange* ex, int32_t* code) {
  MMR_REQUIRE(ex != nullptr, "mmr_exchange_status: exchange is NULL");
  const uint32_t c = *reinterpret_cast<volatile uint32_t*>(ex->err_host);
  if (code) *code = static_cast<int32_t>(c);
  return c == kErrNone ? MMR_OK : check_call(ex, 1, 1, "mmr_exchange_status");
}

int mmr_exchange_abort(mmr_exchange* ex) {
  // raise the abort word in every mapped region (ours included): kernels waiting on this exchange give up
  MMR_REQUIRE(ex != nullptr, "mmr_exchange_abort: exchange is NULL");
  DeviceGuard guard(ex->device);
  cudaStream_t s = nullptr;
  MMR_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));  // must not queue behind a spinning kernel
  const uint32_t one = 1u;
  cudaError_t e = cudaSuccess;
  for (int r = 0; r < ex->world && e == cudaSuccess; ++r) {
    if (ex->peers.base[r] == nullptr) continue;
    e = cudaMemcpyAsync(ex->peers.base[r] + ex->ctl_off + offsetof(Ctl, abort), &one, sizeof(one), cudaMemcpyHostToDevice, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaStreamDestroy(s);
  if (e != cudaSuccess) return fail(MMR_ECUDA, std::string("mmr_exchange_abort: ") + cudaGetErrorString(e));
  return MMR_OK;
}

int mmr_exchange_close_peers(mmr_exchange* ex) {
  // phase 1 of a shutdown: unmap every peer's region.  Callers put a barrier between this and
  // mmr_exchange_destroy, which frees the region the peers had mapped (CUDA requires importers to close first).
  if (ex == nullptr) return MMR_OK;
  DeviceGuard guard(ex->device);
  cudaDeviceSynchronize();
  for (void* p : ex->opened) cudaIpcCloseMemHandle(p);
  ex->opened.clear();
  for (int r = 0; r < kMaxWorld; ++r)
    if (r != ex->rank) ex->peers.base[r] = nullptr;
  ex->open = false;
  cudaGetLastError();
  return MMR_OK;
}

int mmr_exchange_destroy(mmr_exchange* ex) {
  if (ex == nullptr) return MMR_OK;
  mmr_exchange_close_peers(ex);
  DeviceGuard guard(ex->device);
  if (ex->local) cudaFree(ex->local);
  if (ex->err_host) cudaFreeHost(ex->err_host);
  cudaGetLastError();
  delete ex;
  return MMR_OK;
}

int mmr_search_scatter(mmr_index* ix, mmr_exchange* ex, const void* q, int32_t b, int32_t q_dtype, int32_t k,
                       int32_t algo, uint32_t step, void* stream_v) {
  MMR_TRY(check_call(ex, b, k, "mmr_search_scatter"));
  MMR_REQUIRE(ix != nullptr && q != nullptr, "mmr_search_scatter: NULL argument");
  if (ix->device != ex->device) return fail(MMR_EINVAL, "mmr_search_scatter: index and exchange live on different devices");
  const Layout L = layout_for(ex, b, k, k, step);
  PeerSink sink;
  sink.peers = ex->peers;
  sink.off_scores = L.off_scores;
  sink.off_rows = L.off_rows;
  sink.per = L.per;
  sink.kp = L.kp;
  sink.my_rank = ex->rank;
  MMR_TRY(search_impl(ix, q, b, q_dtype, k, algo, nullptr, nullptr, nullptr, &sink, stream_v));
  DeviceGuard guard(ex->device);
  exchange_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_v)>>>(ex->peers, ex->ctl_off, 0, ex->world, ex->rank,
                                                                            step, fingerprint(b, k));
  MMR_LAUNCHED();
  return MMR_OK;
}

int mmr_exchange_rerank(mmr_exchange* ex, const mmr_rerank_tables* t, const int64_t* q_rec, int32_t b, int32_t k,
                        double alpha, double beta, double gamma, int32_t topk, uint32_t step, const int64_t** ids,
                        const double** fin, void* stream_v) {
  MMR_TRY(check_call(ex, b, k, "mmr_exchange_rerank"));
  MMR_REQUIRE(q_rec != nullptr && ids != nullptr && fin != nullptr && topk >= 0, "mmr_exchange_rerank: bad argument");
  if (!is_device_ptr(q_rec)) return fail(MMR_EINVAL, "mmr_exchange_rerank: q_rec must be a device pointer");
  if (t != nullptr && t->device != ex->device)
    return fail(MMR_EINVAL, "mmr_exchange_rerank: tables and exchange live on different devices");
  const TailTables tt{t ? t->masks : nullptr, t ? t->label_words : 0, t ? t->kg : nullptr, t ? t->d_kg : 0,
                      t ? t->n_rec : 0};
  if (!tail_supported(k, tt.kg, tt.d_kg) || ex->world * ((k + 3) / 4 * 4) > kMaxWorld * kTailMaxK)
    return fail(MMR_EUNSUP, "mmr_exchange_rerank: needs k <= 128 and a KG dimension <= 512 that is a multiple of 4");
  DeviceGuard guard(ex->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int keep = (topk > 0 && topk < k) ? topk : k;
  const Layout L = layout_for(ex, b, k, keep, step);
  const int q_lo = std::min(b, ex->rank * L.per), q_hi = std::min(b, (ex->rank + 1) * L.per);
  const uint64_t timeout_ns = static_cast<uint64_t>(ex->timeout_ms) * 1000000ull;
  Ctl* ctl = reinterpret_cast<Ctl*>(ex->local + ex->ctl_off);
  if (q_hi > q_lo) {
    RerankArgs a;
    a.l_scores = reinterpret_cast<const float*>(ex->local + L.off_scores);
    a.l_rows = reinterpret_cast<const int64_t*>(ex->local + L.off_rows);
    a.ctl = ctl;
    a.t = tt;
    a.q_rec = q_rec;
    a.q_lo = q_lo;
    a.nloc = q_hi - q_lo;
    a.b = b;
    a.k = k;
    a.kp = L.kp;
    a.per = L.per;
    a.world = ex->world;
    a.my_rank = ex->rank;
    a.keep = keep;
    a.alpha = alpha;
    a.beta = beta;
    a.gamma = gamma;
    a.peers = ex->peers;
    a.off_ids = L.off_ids;
    a.off_fin = L.off_fin;
    a.ctl_off = ex->ctl_off;
    a.step = step;
    a.meta = fingerprint(b, k);
    a.timeout_ns = timeout_ns;
    a.err = ex->err_dev;
    if (tt.kg == nullptr || tt.d_kg <= 384)
      exchange_rerank_kernel<3><<<a.nloc, kTailThreads, 0, stream>>>(a);
    else
      exchange_rerank_kernel<4><<<a.nloc, kTailThreads, 0, stream>>>(a);
    MMR_LAUNCHED();
  } else {  // this rank owns no query of the batch: it still has to wait for the lists (buffer reuse) and signal
    exchange_wait_kernel<<<1, 32, 0, stream>>>(ctl, 0, ex->world, step, timeout_ns, ex->err_dev);
    MMR_LAUNCHED();
    exchange_signal_kernel<<<1, 32, 0, stream>>>(ex->peers, ex->ctl_off, 1, ex->world, ex->rank, step, 0u);
    MMR_LAUNCHED();
  }
  exchange_wait_kernel<<<1, 32, 0, stream>>>(ctl, 1, ex->world, step, timeout_ns, ex->err_dev);
  MMR_LAUNCHED();
  *ids = reinterpret_cast<const int64_t*>(ex->local + L.off_ids);
  *fin = reinterpret_cast<const double*>(ex->local + L.off_fin);
  return MMR_OK;
}

}  // extern "C"
