// C-ABI entry points of libmmr_b200.so (declared in include/mmr_b200.h).
//
// Host-side orchestration only: argument checks, host<->device staging, workspace management,
// kernel sequencing on the caller's stream.  No compute happens on the CPU and there is no CPU
// fallback: every path ends in a kernel launch or an error code.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "internal.h"

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
namespace mmr {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};
#ifdef MMR_DIAG
// diagnostic builds: MMR_B200_NO_TAU_SHARE=1 disables the cross-CTA threshold exchange of the GEMM path
static const bool g_share_tau = std::getenv("MMR_B200_NO_TAU_SHARE") == nullptr;
#else
constexpr bool g_share_tau = true;
#endif
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int DeviceBuf::ensure(size_t bytes) {
  if (bytes <= cap) return MMR_OK;
  if (p != nullptr) {
    cudaError_t e = cudaFree(p);  // implicit device sync: nothing in flight uses the old buffer
    p = nullptr;
    cap = 0;
    if (e != cudaSuccess) return fail(MMR_ECUDA, std::string("cudaFree: ") + cudaGetErrorString(e));
  }
  size_t want = bytes + bytes / 4;
  want = (want + 255) / 256 * 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    p = nullptr;
    return fail(MMR_ENOMEM, "cudaMalloc(" + std::to_string(bytes) + " bytes): " + cudaGetErrorString(e));
  }
  cap = want;
  return MMR_OK;
}
void DeviceBuf::release() {
  if (p != nullptr) cudaFree(p);
  p = nullptr;
  cap = 0;
}

bool is_device_ptr(const void* p) {
  if (p == nullptr) return false;
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int elem_size(int dtype) { return dtype == MMR_BF16 ? 2 : 4; }

DeviceGuard::DeviceGuard(int dev) {
  if (cudaGetDevice(&prev) != cudaSuccess) {
    cudaGetLastError();
    prev = -1;
  }
  ok = cudaSetDevice(dev) == cudaSuccess;
  if (!ok) cudaGetLastError();
}
DeviceGuard::~DeviceGuard() {
  if (prev >= 0) cudaSetDevice(prev);
}

int stage_in(const void* src, size_t bytes, DeviceBuf& buf, cudaStream_t stream, const void** dev_out) {
  if (src == nullptr || bytes == 0) {
    *dev_out = src;
    return MMR_OK;
  }
  if (is_device_ptr(src)) {
    *dev_out = src;
    return MMR_OK;
  }
  MMR_TRY(buf.ensure(bytes));
  MMR_CUDA_TRY(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, stream));
  *dev_out = buf.p;
  return MMR_OK;
}

// Device validation is cached per ordinal: cudaGetDeviceProperties costs milliseconds and the
// handle-less entry points (merge, rerank, metrics) run once per query batch.
int check_device(int device, int* num_sms) {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cached_sms[kMaxDev];  // 0 = unknown, > 0 = validated sm_100 device
  if (device >= 0 && device < kMaxDev) {
    const int c = cached_sms[device].load(std::memory_order_acquire);
    if (c > 0) {
      if (num_sms != nullptr) *num_sms = c;
      return MMR_OK;
    }
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(MMR_ENODEV, "no CUDA device available: libmmr_b200 has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(MMR_EINVAL, "invalid device ordinal " + std::to_string(device));
  int major = 0, minor = 0, sms = 0;
  MMR_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  MMR_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  MMR_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (major != 10) {
    return fail(MMR_ENODEV, "device " + std::to_string(device) + " is sm_" + std::to_string(major) +
                                std::to_string(minor) + "; libmmr_b200 is built for sm_100a only");
  }
  {
    // The handle-less entry points take their temporaries from the device's stream-ordered pool.
    // By default the pool gives memory back to the OS at every synchronisation, which turns each
    // synchronous call (host result buffers) into a multi-millisecond cuMemCreate/map cycle: keep it.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  if (device < kMaxDev) cached_sms[device].store(sms, std::memory_order_release);
  if (num_sms != nullptr) *num_sms = sms;
  return MMR_OK;
}

namespace {

// Per-call temporaries for handle-less entry points (stream-ordered allocations).
struct CallScope {
  cudaStream_t stream;
  std::vector<void*> temps;
  struct Out { void* host; void* dev; size_t bytes; };
  std::vector<Out> outs;
  explicit CallScope(cudaStream_t s) : stream(s) {}
  ~CallScope() {
    for (void* t : temps) cudaFreeAsync(t, stream);
  }
  int alloc(size_t bytes, void** p) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(p, bytes, stream);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(MMR_ENOMEM, "cudaMallocAsync(" + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    temps.push_back(*p);
    return MMR_OK;
  }
  template <typename T>
  int in(const T* src, size_t count, const T** dev) {
    if (src == nullptr || count == 0 || is_device_ptr(src)) {
      *dev = src;
      return MMR_OK;
    }
    void* p = nullptr;
    MMR_TRY(alloc(count * sizeof(T), &p));
    MMR_CUDA_TRY(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    *dev = static_cast<const T*>(p);
    return MMR_OK;
  }
  template <typename T>
  int out(T* dst, size_t count, T** dev) {
    if (dst == nullptr || count == 0 || is_device_ptr(dst)) {
      *dev = dst;
      return MMR_OK;
    }
    void* p = nullptr;
    MMR_TRY(alloc(count * sizeof(T), &p));
    outs.push_back({dst, p, count * sizeof(T)});
    *dev = static_cast<T*>(p);
    return MMR_OK;
  }
  int finish() {
    for (const Out& o : outs) MMR_CUDA_TRY(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, stream));
    if (!outs.empty()) MMR_CUDA_TRY(cudaStreamSynchronize(stream));
    return MMR_OK;
  }
};

__global__ void globalize_exclude_kernel(const int64_t* __restrict__ in, int b, int64_t row_offset, int64_t n,
                                         int64_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b) {
    int64_t l = in[i] - row_offset;
    out[i] = (in[i] >= 0 && l >= 0 && l < n) ? l : -1;
  }
}

// Device-side ordering of consecutive calls on one handle (see mmr_index::mu): call under the mutex.
int enter_call(mmr_index* ix, cudaStream_t stream) {
  if (ix->has_last && ix->last_stream != stream) MMR_CUDA_TRY(cudaStreamWaitEvent(stream, ix->last_done, 0));
  return MMR_OK;
}
int leave_call(mmr_index* ix, cudaStream_t stream) {
  if (ix->last_done == nullptr) MMR_CUDA_TRY(cudaEventCreateWithFlags(&ix->last_done, cudaEventDisableTiming));
  MMR_CUDA_TRY(cudaEventRecord(ix->last_done, stream));
  ix->last_stream = stream;
  ix->has_last = true;
  return MMR_OK;
}

}  // namespace
}  // namespace mmr

using namespace mmr;

extern "C" {

int mmr_abi_version(void) { return MMR_ABI_VERSION; }
const char* mmr_last_error(void) { return g_last_error.c_str(); }
int64_t mmr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mmr_index_profile(mmr_index* ix, int32_t enable, double* kernel_ms_sum, int32_t* kernel_launches) {
  MMR_REQUIRE(ix != nullptr, "mmr_index_profile: index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  DeviceGuard guard(ix->device);
  double sum = 0.0;
  if (ix->ev_used > 0) {
    MMR_CUDA_TRY(cudaEventSynchronize(ix->ev_stop[ix->ev_used - 1]));
    for (size_t i = 0; i < ix->ev_used; ++i) {
      float ms = 0.f;
      MMR_CUDA_TRY(cudaEventElapsedTime(&ms, ix->ev_start[i], ix->ev_stop[i]));
      sum += ms;
    }
  }
  if (kernel_ms_sum) *kernel_ms_sum = sum;
  if (kernel_launches) *kernel_launches = static_cast<int32_t>(ix->ev_used);
  ix->ev_used = 0;
  ix->profiling = enable != 0;
  return MMR_OK;
}

int mmr_index_tune(mmr_index* ix, int32_t knob, int32_t value) {
  MMR_REQUIRE(ix != nullptr, "mmr_index_tune: index is NULL");
  std::lock_guard<std::mutex> lock(ix->mu);
  switch (knob) {
    case MMR_TUNE_GEMM_VARIANT:
      MMR_REQUIRE(value >= MMR_GEMM_VARIANT_AUTO && value <= MMR_GEMM_VARIANT_SHORT, "mmr_index_tune: bad GEMM variant");
      ix->tune.variant = value;
      return MMR_OK;
    case MMR_TUNE_GEMM_PARTS:
      MMR_REQUIRE(value >= 0, "mmr_index_tune: parts must be >= 0");
      ix->tune.parts = value;
      return MMR_OK;
    case MMR_TUNE_GEMM_PAIR:
      MMR_REQUIRE(value == MMR_GEMM_PAIR_AUTO || value == MMR_GEMM_PAIR_OFF, "mmr_index_tune: bad pair mode");
      ix->tune.pair = value;
      return MMR_OK;
    default:
      return fail(MMR_EINVAL, "mmr_index_tune: unknown knob " + std::to_string(knob));
  }
}

int mmr_index_last_plan(const mmr_index* ix, int32_t* algo, int32_t* variant, int32_t* pair, int32_t* n_parts,
                        int32_t* tiles_per_part) {
  MMR_REQUIRE(ix != nullptr, "mmr_index_last_plan: index is NULL");
  if (algo) *algo = ix->last_algo;
  if (variant) *variant = ix->last_variant;
  if (pair) *pair = ix->last_pair;
  if (n_parts) *n_parts = ix->last_parts;
  if (tiles_per_part) *tiles_per_part = ix->last_tiles_per_part;
  return MMR_OK;
}

int mmr_index_create(mmr_index** out, const void* emb, int64_t n, int32_t d, int32_t dtype_in, int32_t dtype_store,
                     int64_t row_offset, int32_t device, int32_t flags, void* stream_v) {
  MMR_REQUIRE(out != nullptr, "mmr_index_create: out is NULL");
  *out = nullptr;
  MMR_REQUIRE(n >= 0 && d >= 1, "mmr_index_create: need n >= 0 and d >= 1");
  MMR_REQUIRE(n < 0xFFFFFFFFll, "mmr_index_create: a shard holds at most 2^32-2 rows");
  // the cross-shard merge packs GLOBAL row ids into the low 32 bits of its ordering keys
  MMR_REQUIRE(row_offset >= 0 && row_offset + n < 0xFFFFFFFFll,
              "mmr_index_create: global row ids (row_offset + n) must stay below 2^32-1");
  MMR_REQUIRE(dtype_in == MMR_F32 || dtype_in == MMR_BF16, "mmr_index_create: dtype_in must be MMR_F32 or MMR_BF16");
  MMR_REQUIRE(dtype_store == MMR_F32 || dtype_store == MMR_BF16, "mmr_index_create: bad dtype_store");
  MMR_REQUIRE(emb != nullptr || n == 0, "mmr_index_create: emb is NULL");
  int num_sms = 148;
  MMR_TRY(check_device(device, &num_sms));
  DeviceGuard guard(device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

  mmr_index* ix = new mmr_index();
  ix->device = device;
  ix->num_sms = num_sms;
  ix->n = n;
  ix->d = d;
  ix->d_pad = static_cast<int>(round_up(d, 64));
  ix->dtype = dtype_store;
  ix->row_offset = row_offset;
  const size_t es = elem_size(dtype_store);
  auto cleanup = [&](int code) {
    if (ix->owns_emb && ix->emb) cudaFree(ix->emb);
    if (ix->inv_norm) cudaFree(ix->inv_norm);
    delete ix;
    return code;
  };
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ix->inv_norm), (n > 0 ? n : 1) * sizeof(float));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cleanup(fail(MMR_ENOMEM, std::string("inv_norm alloc: ") + cudaGetErrorString(e)));
  }
  const bool borrow = (flags & MMR_FLAG_BORROW) != 0;
  if (borrow) {
    if (!(is_device_ptr(emb) && dtype_in == dtype_store && d == ix->d_pad &&
          reinterpret_cast<uintptr_t>(emb) % 16 == 0)) {
      return cleanup(fail(MMR_EINVAL, "MMR_FLAG_BORROW needs a 16-byte aligned device pointer in storage layout "
                                      "(dtype_in == dtype_store, d a multiple of 64)"));
    }
    ix->emb = const_cast<void*>(emb);
    ix->owns_emb = false;
    int s = launch_ingest(emb, dtype_in, n, d, d, nullptr, dtype_store, ix->d_pad, ix->inv_norm, stream);
    if (s != MMR_OK) return cleanup(s);
  } else {
    e = cudaMalloc(&ix->emb, (n > 0 ? n : 1) * static_cast<size_t>(ix->d_pad) * es);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return cleanup(fail(MMR_ENOMEM, std::string("gallery alloc: ") + cudaGetErrorString(e)));
    }
    if (n > 0) {
      if (is_device_ptr(emb)) {
        int s = launch_ingest(emb, dtype_in, n, d, d, ix->emb, dtype_store, ix->d_pad, ix->inv_norm, stream);
        if (s != MMR_OK) return cleanup(s);
      } else {
        // host gallery: stream it through a bounded staging buffer, chunk by chunk
        const size_t in_row = static_cast<size_t>(d) * elem_size(dtype_in);
        int64_t chunk_rows = static_cast<int64_t>((256ull << 20) / in_row);
        if (chunk_rows < 1) chunk_rows = 1;
        if (chunk_rows > n) chunk_rows = n;
        void* staging = nullptr;
        e = cudaMalloc(&staging, static_cast<size_t>(chunk_rows) * in_row);
        if (e != cudaSuccess) {
          cudaGetLastError();
          return cleanup(fail(MMR_ENOMEM, std::string("staging alloc: ") + cudaGetErrorString(e)));
        }
        int s = MMR_OK;
        for (int64_t r = 0; r < n && s == MMR_OK; r += chunk_rows) {
          const int64_t rows = (n - r) < chunk_rows ? (n - r) : chunk_rows;
          e = cudaMemcpyAsync(staging, static_cast<const char*>(emb) + static_cast<size_t>(r) * in_row,
                              static_cast<size_t>(rows) * in_row, cudaMemcpyHostToDevice, stream);
          if (e != cudaSuccess) {
            s = fail(MMR_ECUDA, std::string("gallery H2D: ") + cudaGetErrorString(e));
            break;
          }
          s = launch_ingest(staging, dtype_in, rows, d, d,
                            static_cast<char*>(ix->emb) + static_cast<size_t>(r) * ix->d_pad * es, dtype_store,
                            ix->d_pad, ix->inv_norm + r, stream);
        }
        cudaStreamSynchronize(stream);
        cudaFree(staging);
        if (s != MMR_OK) return cleanup(s);
      }
    }
  }
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return cleanup(fail(MMR_ECUDA, std::string("ingest: ") + cudaGetErrorString(e)));
  *out = ix;
  return MMR_OK;
}

int mmr_index_destroy(mmr_index* ix) {
  if (ix == nullptr) return MMR_OK;
  DeviceGuard guard(ix->device);
  cudaDeviceSynchronize();
  if (ix->owns_emb && ix->emb) cudaFree(ix->emb);
  if (ix->inv_norm) cudaFree(ix->inv_norm);
  for (DeviceBuf* b : {&ix->q_in, &ix->q_store, &ix->q_f32, &ix->q_inv, &ix->scratch, &ix->excl_in, &ix->excl_local, &ix->partial,
                       &ix->counts, &ix->tau_pub, &ix->out_scores, &ix->out_rows})
    b->release();
  for (cudaEvent_t e : ix->ev_start) cudaEventDestroy(e);
  for (cudaEvent_t e : ix->ev_stop) cudaEventDestroy(e);
  if (ix->last_done != nullptr) cudaEventDestroy(ix->last_done);
  delete ix;
  return MMR_OK;
}

int mmr_index_info(const mmr_index* ix, int64_t* n, int32_t* d, int32_t* d_pad, int32_t* dtype_store,
                   int32_t* device, int64_t* row_offset, int64_t* hbm_bytes) {
  MMR_REQUIRE(ix != nullptr, "mmr_index_info: index is NULL");
  if (n) *n = ix->n;
  if (d) *d = ix->d;
  if (d_pad) *d_pad = ix->d_pad;
  if (dtype_store) *dtype_store = ix->dtype;
  if (device) *device = ix->device;
  if (row_offset) *row_offset = ix->row_offset;
  if (hbm_bytes) *hbm_bytes = ix->n * static_cast<int64_t>(ix->d_pad) * elem_size(ix->dtype) + ix->n * 4;
  return MMR_OK;
}

int mmr_index_device_ptrs(const mmr_index* ix, const void** emb, const float** inv_norm) {
  MMR_REQUIRE(ix != nullptr, "mmr_index_device_ptrs: index is NULL");
  if (emb) *emb = ix->emb;
  if (inv_norm) *inv_norm = ix->inv_norm;
  return MMR_OK;
}

int mmr_index_get_rows(const mmr_index* ix, const int64_t* rows, int64_t m, float* out, void* stream_v) {
  MMR_REQUIRE(ix != nullptr, "mmr_index_get_rows: index is NULL");
  MMR_REQUIRE(m >= 0 && (m == 0 || (rows != nullptr && out != nullptr)), "mmr_index_get_rows: NULL argument");
  DeviceGuard guard(ix->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CallScope cs(stream);
  const int64_t* d_rows;
  float* d_out;
  MMR_TRY(cs.in(rows, static_cast<size_t>(m), &d_rows));
  MMR_TRY(cs.out(out, static_cast<size_t>(m) * ix->d, &d_out));
  MMR_TRY(launch_gather_rows(ix->emb, ix->dtype, ix->n, ix->d, ix->d_pad, ix->row_offset, d_rows, m, d_out, stream));
  return cs.finish();
}

}  // extern "C"

namespace mmr {
// mmr_search proper.  `sink` != nullptr (mmr_search_scatter): the per-query lists go straight into the owner
// ranks' exchange regions; out_scores / out_rows are then optional device staging buffers.
int search_impl(mmr_index* ix, const void* q, int32_t b, int32_t q_dtype, int32_t k, int32_t algo,
                const int64_t* exclude_rows, float* out_scores, int64_t* out_rows, const PeerSink* sink,
                void* stream_v) {
  MMR_REQUIRE(ix != nullptr, "mmr_search: index is NULL");
  MMR_REQUIRE(b >= 0 && k >= 1, "mmr_search: need b >= 0 and k >= 1");
  MMR_REQUIRE(q_dtype == MMR_F32 || q_dtype == MMR_BF16, "mmr_search: bad q_dtype");
  MMR_REQUIRE(algo == MMR_ALGO_AUTO || algo == MMR_ALGO_SCAN || algo == MMR_ALGO_GEMM, "mmr_search: bad algo");
  if (b == 0) return MMR_OK;
  MMR_REQUIRE(q != nullptr && (sink != nullptr || (out_scores != nullptr && out_rows != nullptr)),
              "mmr_search: NULL argument");
  if (k > MMR_MAX_K) return fail(MMR_EUNSUP, "mmr_search: k > MMR_MAX_K (1024)");
  if (algo == MMR_ALGO_GEMM && ix->dtype != MMR_BF16)
    return fail(MMR_EUNSUP, "mmr_search: the tcgen05 GEMM path needs a bf16 index");

  std::lock_guard<std::mutex> lock(ix->mu);
  DeviceGuard guard(ix->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  MMR_TRY(enter_call(ix, stream));

  // queries -> storage dtype (bf16 index: round to bf16) padded to d_pad, + inverse norms
  const void* d_q = nullptr;
  MMR_TRY(stage_in(q, static_cast<size_t>(b) * ix->d * elem_size(q_dtype), ix->q_in, stream, &d_q));
  MMR_TRY(ix->q_store.ensure(static_cast<size_t>(b) * ix->d_pad * elem_size(ix->dtype)));
  MMR_TRY(ix->q_inv.ensure(static_cast<size_t>(b) * sizeof(float)));
  MMR_TRY(launch_ingest(d_q, q_dtype, b, ix->d, ix->d, ix->q_store.p, ix->dtype, ix->d_pad, ix->q_inv.as<float>(),
                        stream));

  // exclusions: global row ids -> local (or -1)
  const int64_t* d_excl = nullptr;
  if (exclude_rows != nullptr) {
    const void* d_ex_in = nullptr;
    MMR_TRY(stage_in(exclude_rows, static_cast<size_t>(b) * sizeof(int64_t), ix->excl_in, stream, &d_ex_in));
    MMR_TRY(ix->excl_local.ensure(static_cast<size_t>(b) * sizeof(int64_t)));
    globalize_exclude_kernel<<<(b + 127) / 128, 128, 0, stream>>>(static_cast<const int64_t*>(d_ex_in), b,
                                                                  ix->row_offset, ix->n,
                                                                  ix->excl_local.as<int64_t>());
    MMR_LAUNCHED();
    d_excl = ix->excl_local.as<int64_t>();
  }

  // outputs
  const bool host_scores = sink == nullptr && !is_device_ptr(out_scores);
  const bool host_rows = sink == nullptr && !is_device_ptr(out_rows);
  float* d_scores = out_scores;
  int64_t* d_rows = out_rows;
  if (sink != nullptr && (k > 128 || d_scores == nullptr || d_rows == nullptr)) {
    // shapes outside the fused select + scatter kernel stage the dense lists in the handle's buffers
    MMR_TRY(ix->out_scores.ensure(static_cast<size_t>(b) * k * sizeof(float)));
    MMR_TRY(ix->out_rows.ensure(static_cast<size_t>(b) * k * sizeof(int64_t)));
    d_scores = ix->out_scores.as<float>();
    d_rows = ix->out_rows.as<int64_t>();
  }
  if (host_scores) {
    MMR_TRY(ix->out_scores.ensure(static_cast<size_t>(b) * k * sizeof(float)));
    d_scores = ix->out_scores.as<float>();
  }
  if (host_rows) {
    MMR_TRY(ix->out_rows.ensure(static_cast<size_t>(b) * k * sizeof(int64_t)));
    d_rows = ix->out_rows.as<int64_t>();
  }

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ix->profiling) {
    if (ix->ev_used == ix->ev_start.size()) {
      cudaEvent_t a, c;
      MMR_CUDA_TRY(cudaEventCreate(&a));
      MMR_CUDA_TRY(cudaEventCreate(&c));
      ix->ev_start.push_back(a);
      ix->ev_stop.push_back(c);
    }
    ev0 = ix->ev_start[ix->ev_used];
    ev1 = ix->ev_stop[ix->ev_used];
    ++ix->ev_used;
  }

  int use = algo;
  // measured on 10M x 512 (scripts/sweep_batch.py): the TMA-fed GEMM kernel stays HBM-bound and flat
  // (1.5-1.9 ms) from b = 2 to b = 128 while the scan re-streams the gallery per query group; b = 1 is the
  // scan's latency path
  if (use == MMR_ALGO_AUTO) use = (ix->dtype == MMR_BF16 && b >= 2 && ix->n >= 1024) ? MMR_ALGO_GEMM : MMR_ALGO_SCAN;

  if (use == MMR_ALGO_GEMM) {
    GemmPlan gp;
    const int k_eff = k + (d_excl != nullptr ? 1 : 0);  // the excluded row is dropped by the select step
    MMR_TRY(plan_gemm(ix->n, ix->d_pad, b, k_eff, ix->num_sms, ix->tune, &gp));
    ix->last_algo = MMR_ALGO_GEMM;
    ix->last_variant = gp.probe ? MMR_GEMM_VARIANT_SHORT : MMR_GEMM_VARIANT_LONG;
    ix->last_pair = gp.pair;
    ix->last_parts = gp.n_parts;
    ix->last_tiles_per_part = gp.tiles_per_part;
    MMR_TRY(ix->partial.ensure(gp.cand_bytes));
    MMR_TRY(ix->counts.ensure(gp.count_bytes));
    MMR_TRY(ix->tau_pub.ensure(gp.pub_bytes));
    if (ev0) MMR_CUDA_TRY(cudaEventRecord(ev0, stream));
    MMR_TRY(launch_gemm_topk(ix->emb, ix->inv_norm, ix->n, ix->d_pad, ix->q_store.p, ix->q_inv.as<float>(), b, k_eff,
                             d_excl, gp, ix->partial.as<uint64_t>(), ix->counts.as<int32_t>(),
                             g_share_tau ? ix->tau_pub.as<uint32_t>() : nullptr, stream));
    if (ev1) MMR_CUDA_TRY(cudaEventRecord(ev1, stream));
    MMR_TRY(launch_select_var(ix->partial.as<uint64_t>(), ix->counts.as<int32_t>(), b, gp.n_lists, gp.cap, k_eff, k,
                              ix->row_offset, d_excl, g_share_tau ? ix->tau_pub.as<uint32_t>() : nullptr,
                              gp.m_tiles * 128, d_scores, d_rows, stream, sink));
  } else {
    const float* q_f32 = nullptr;
    if (ix->dtype == MMR_F32) {
      q_f32 = ix->q_store.as<float>();
    } else {
      MMR_TRY(ix->q_f32.ensure(static_cast<size_t>(b) * ix->d_pad * sizeof(float)));
      MMR_TRY(ix->scratch.ensure(static_cast<size_t>(b) * sizeof(float)));
      // exact upcast of the bf16-rounded queries (the norms it recomputes are discarded)
      MMR_TRY(launch_ingest(ix->q_store.p, MMR_BF16, b, ix->d_pad, ix->d_pad, ix->q_f32.p, MMR_F32, ix->d_pad,
                            ix->scratch.as<float>(), stream));
      q_f32 = ix->q_f32.as<float>();
    }
    ScanPlan sp;
    MMR_TRY(plan_scan(ix->n, ix->d_pad, ix->dtype, b, k, ix->num_sms, &sp));
    ix->last_algo = MMR_ALGO_SCAN;
    ix->last_variant = ix->last_pair = ix->last_tiles_per_part = 0;
    ix->last_parts = sp.n_parts;
    MMR_TRY(ix->partial.ensure(sp.partial_bytes));
    if (ev0) MMR_CUDA_TRY(cudaEventRecord(ev0, stream));
    MMR_TRY(launch_scan(ix->emb, ix->dtype, ix->inv_norm, ix->n, ix->d_pad, q_f32, ix->q_inv.as<float>(), b, k, d_excl,
                        sp, ix->partial.as<uint64_t>(), stream));
    if (ev1) MMR_CUDA_TRY(cudaEventRecord(ev1, stream));
    MMR_TRY(launch_select_keys(ix->partial.as<uint64_t>(), b, static_cast<int64_t>(sp.n_parts) * sp.kp, k,
                               ix->row_offset, d_scores, d_rows, stream, sink));
  }

  if (host_scores)
    MMR_CUDA_TRY(cudaMemcpyAsync(out_scores, d_scores, static_cast<size_t>(b) * k * sizeof(float),
                                 cudaMemcpyDeviceToHost, stream));
  if (host_rows)
    MMR_CUDA_TRY(cudaMemcpyAsync(out_rows, d_rows, static_cast<size_t>(b) * k * sizeof(int64_t),
                                 cudaMemcpyDeviceToHost, stream));
  if (host_scores || host_rows) MMR_CUDA_TRY(cudaStreamSynchronize(stream));
  return leave_call(ix, stream);
}
}  // namespace mmr

extern "C" {

int mmr_search(mmr_index* ix, const void* q, int32_t b, int32_t q_dtype, int32_t k, int32_t algo,
               const int64_t* exclude_rows, float* out_scores, int64_t* out_rows, void* stream_v) {
  return search_impl(ix, q, b, q_dtype, k, algo, exclude_rows, out_scores, out_rows, nullptr, stream_v);
}

int mmr_merge_topk_strided(const float* scores, const int64_t* rows, int32_t n_lists, int32_t b, int32_t k_in,
                           int64_t scores_list_stride, int64_t rows_list_stride, int32_t k_out, float* out_scores,
                           int64_t* out_rows, int32_t* out_src, int32_t device, void* stream_v) {
  MMR_REQUIRE(n_lists >= 1 && b >= 0 && k_in >= 1 && k_out >= 1, "mmr_merge_topk: bad sizes");
  if (b == 0) return MMR_OK;
  MMR_REQUIRE(scores && rows && out_scores && out_rows, "mmr_merge_topk: NULL argument");
  if (k_out > MMR_MAX_K) return fail(MMR_EUNSUP, "mmr_merge_topk: k_out > MMR_MAX_K");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CallScope cs(stream);
  const size_t per_list = static_cast<size_t>(b) * k_in, n_out = static_cast<size_t>(b) * k_out;
  const bool contiguous = static_cast<size_t>(scores_list_stride) == per_list &&
                          static_cast<size_t>(rows_list_stride) == per_list;
  if (!contiguous && !(is_device_ptr(scores) && is_device_ptr(rows)))
    return fail(MMR_EINVAL, "mmr_merge_topk_strided: strided lists must be device pointers");
  const float* d_s;
  const int64_t* d_r;
  float* d_os;
  int64_t* d_or;
  int32_t* d_src;
  MMR_TRY(cs.in(scores, per_list * n_lists, &d_s));
  MMR_TRY(cs.in(rows, per_list * n_lists, &d_r));
  MMR_TRY(cs.out(out_scores, n_out, &d_os));
  MMR_TRY(cs.out(out_rows, n_out, &d_or));
  MMR_TRY(cs.out(out_src, n_out, &d_src));
  MMR_TRY(launch_merge_lists(d_s, d_r, n_lists, b, k_in, scores_list_stride, rows_list_stride, k_out, d_os, d_or,
                             d_src, stream));
  return cs.finish();
}

int mmr_merge_topk(const float* scores, const int64_t* rows, int32_t n_lists, int32_t b, int32_t k_in, int32_t k_out,
                   float* out_scores, int64_t* out_rows, int32_t* out_src, int32_t device, void* stream_v) {
  const int64_t per_list = static_cast<int64_t>(b) * k_in;
  return mmr_merge_topk_strided(scores, rows, n_lists, b, k_in, per_list, per_list, k_out, out_scores, out_rows,
                                out_src, device, stream_v);
}

int mmr_gather_payload(const float* payload, int64_t list_stride, const int32_t* src, int32_t b, int32_t k_in,
                       int32_t k_out, float* out, int32_t device, void* stream_v) {
  MMR_REQUIRE(b >= 0 && k_in >= 1 && k_out >= 1, "mmr_gather_payload: bad sizes");
  if (b == 0) return MMR_OK;
  MMR_REQUIRE(payload && src && out, "mmr_gather_payload: NULL argument");
  if (!(is_device_ptr(payload) && is_device_ptr(src) && is_device_ptr(out)))
    return fail(MMR_EINVAL, "mmr_gather_payload: device pointers only");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  return launch_gather_payload(payload, list_stride, src, b, k_in, k_out, out, static_cast<cudaStream_t>(stream_v));
}

int mmr_apply_order(const int64_t* rows, const int32_t* order, const double* scores4, int32_t b, int32_t k,
                    int32_t keep, int64_t* out_rows, double* out_final, int32_t device, void* stream_v) {
  MMR_REQUIRE(b >= 0 && k >= 1 && keep >= 1 && keep <= k, "mmr_apply_order: bad sizes");
  if (b == 0) return MMR_OK;
  MMR_REQUIRE(rows && order && scores4 && out_rows && out_final, "mmr_apply_order: NULL argument");
  if (!(is_device_ptr(rows) && is_device_ptr(order) && is_device_ptr(scores4) && is_device_ptr(out_rows) &&
        is_device_ptr(out_final)))
    return fail(MMR_EINVAL, "mmr_apply_order: device pointers only");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  return launch_apply_order(rows, order, scores4, b, k, keep, out_rows, out_final, static_cast<cudaStream_t>(stream_v));
}

int mmr_rerank_tables_create(mmr_rerank_tables** out, const uint64_t* label_masks, int32_t label_words,
                             const float* kg_vecs, int32_t d_kg, int64_t n_rec, int32_t device, void* stream_v) {
  MMR_REQUIRE(out != nullptr, "mmr_rerank_tables_create: out is NULL");
  *out = nullptr;
  MMR_REQUIRE(n_rec >= 0 && label_words >= 0 && d_kg >= 0, "mmr_rerank_tables_create: bad sizes");
  MMR_REQUIRE(label_words == 0 || label_masks != nullptr || n_rec == 0, "mmr_rerank_tables_create: label_masks NULL");
  MMR_REQUIRE(d_kg == 0 || kg_vecs != nullptr || n_rec == 0, "mmr_rerank_tables_create: kg_vecs NULL");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  mmr_rerank_tables* t = new mmr_rerank_tables();
  t->device = device;
  t->n_rec = n_rec;
  t->label_words = label_words;
  t->d_kg = d_kg;
  auto cleanup = [&](int code) {
    if (t->masks) cudaFree(t->masks);
    if (t->kg) cudaFree(t->kg);
    delete t;
    return code;
  };
  const size_t mb = static_cast<size_t>(n_rec) * label_words * sizeof(uint64_t);
  const size_t kb = static_cast<size_t>(n_rec) * d_kg * sizeof(float);
  if (mb > 0) {
    if (cudaMalloc(reinterpret_cast<void**>(&t->masks), mb) != cudaSuccess) {
      cudaGetLastError();
      return cleanup(fail(MMR_ENOMEM, "rerank tables: label mask allocation failed"));
    }
    if (cudaMemcpyAsync(t->masks, label_masks, mb, cudaMemcpyDefault, stream) != cudaSuccess)
      return cleanup(fail(MMR_ECUDA, "rerank tables: label mask copy failed"));
  }
  if (kb > 0) {
    if (cudaMalloc(reinterpret_cast<void**>(&t->kg), kb) != cudaSuccess) {
      cudaGetLastError();
      return cleanup(fail(MMR_ENOMEM, "rerank tables: KG vector allocation failed"));
    }
    if (cudaMemcpyAsync(t->kg, kg_vecs, kb, cudaMemcpyDefault, stream) != cudaSuccess)
      return cleanup(fail(MMR_ECUDA, "rerank tables: KG vector copy failed"));
  }
  if (cudaStreamSynchronize(stream) != cudaSuccess) return cleanup(fail(MMR_ECUDA, "rerank tables: sync failed"));
  *out = t;
  return MMR_OK;
}

int mmr_rerank_tables_destroy(mmr_rerank_tables* t) {
  if (t == nullptr) return MMR_OK;
  DeviceGuard guard(t->device);
  cudaDeviceSynchronize();
  if (t->masks) cudaFree(t->masks);
  if (t->kg) cudaFree(t->kg);
  delete t;
  return MMR_OK;
}

static int rerank_features_impl(CallScope& cs, const mmr_index* ix, const mmr_rerank_tables* t, const float* q_emb,
                                const float* cand_emb, const int64_t* cand_rows, const int64_t* q_rec,
                                const int64_t* cand_rec, const int32_t* cand_count, int32_t b, int32_t k, int32_t d,
                                double* d_raw, uint8_t* d_owned, const float* emb_cos = nullptr,
                                float* d_cos_out = nullptr) {
  if (emb_cos == nullptr) {
    MMR_REQUIRE(q_emb != nullptr, "Query embedding not found. Provide query_emb.");
    if (cand_emb == nullptr && (ix == nullptr || cand_rows == nullptr))
      return fail(MMR_EINVAL, "Please provide candidate_embs or candidate_emb_lookup.");
    if (cand_emb == nullptr) MMR_REQUIRE(ix->d == d, "candidate_embs rows must match the index dimension");
  }
  const size_t bk = static_cast<size_t>(b) * k;
  const float *d_q, *d_ce;
  const int64_t *d_cr, *d_qr, *d_crec;
  const int32_t* d_cc;
  MMR_TRY(cs.in(q_emb, static_cast<size_t>(b) * d, &d_q));
  MMR_TRY(cs.in(cand_emb, bk * d, &d_ce));
  MMR_TRY(cs.in(cand_rows, bk, &d_cr));
  MMR_TRY(cs.in(q_rec, static_cast<size_t>(b), &d_qr));
  MMR_TRY(cs.in(cand_rec, bk, &d_crec));
  MMR_TRY(cs.in(cand_count, static_cast<size_t>(b), &d_cc));
  const float* d_ec;
  MMR_TRY(cs.in(emb_cos, bk, &d_ec));
  return launch_rerank_features(ix ? ix->emb : nullptr, ix ? ix->dtype : MMR_F32, ix ? ix->n : 0, ix ? ix->d_pad : 0,
                                ix ? ix->row_offset : 0, t ? t->masks : nullptr, t ? t->label_words : 0,
                                t ? t->kg : nullptr, t ? t->d_kg : 0, t ? t->n_rec : 0, d_q, d_ce, d_cr, d_qr, d_crec,
                                d_cc, b, k, d, d_raw, d_owned, d_ec, d_cos_out, cs.stream);
}

int mmr_rerank_features(const mmr_index* ix, const mmr_rerank_tables* t, const float* q_emb, const float* cand_emb,
                        const int64_t* cand_rows, const int64_t* q_rec, const int64_t* cand_rec,
                        const int32_t* cand_count, int32_t b, int32_t k, int32_t d, double* out_raw, uint8_t* owned,
                        void* stream_v) {
  MMR_REQUIRE(b >= 0 && k >= 0 && d >= 1, "mmr_rerank_features: bad sizes");
  if (b == 0 || k == 0) return MMR_OK;
  MMR_REQUIRE(out_raw != nullptr, "mmr_rerank_features: out_raw is NULL");
  const int device = ix ? ix->device : (t ? t->device : 0);
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  double* d_raw;
  uint8_t* d_owned;
  MMR_TRY(cs.out(out_raw, static_cast<size_t>(b) * k * 3, &d_raw));
  MMR_TRY(cs.out(owned, static_cast<size_t>(b) * k, &d_owned));
  MMR_TRY(rerank_features_impl(cs, ix, t, q_emb, cand_emb, cand_rows, q_rec, cand_rec, cand_count, b, k, d, d_raw,
                               d_owned));
  return cs.finish();
}

int mmr_rerank_combine(const double* raw, const int32_t* cand_count, int32_t b, int32_t k, double alpha, double beta,
                       double gamma, int32_t topk, int32_t* out_order, double* out_scores, int32_t device,
                       void* stream_v) {
  MMR_REQUIRE(b >= 0 && k >= 0 && topk >= 0, "mmr_rerank_combine: bad sizes");
  if (b == 0 || k == 0) return MMR_OK;
  MMR_REQUIRE(raw && out_order && out_scores, "mmr_rerank_combine: NULL argument");
  if (k > 4096) return fail(MMR_EUNSUP, "mmr_rerank_combine: more than 4096 candidates per query");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  const int keep = (topk > 0 && topk < k) ? topk : k;
  const double* d_raw;
  const int32_t* d_cc;
  int32_t* d_order;
  double* d_sc;
  MMR_TRY(cs.in(raw, static_cast<size_t>(b) * k * 3, &d_raw));
  MMR_TRY(cs.in(cand_count, static_cast<size_t>(b), &d_cc));
  MMR_TRY(cs.out(out_order, static_cast<size_t>(b) * keep, &d_order));
  MMR_TRY(cs.out(out_scores, static_cast<size_t>(b) * keep * 4, &d_sc));
  MMR_TRY(launch_rerank_combine(d_raw, d_cc, b, k, alpha, beta, gamma, topk, d_order, d_sc, cs.stream));
  return cs.finish();
}

int mmr_rerank(const mmr_index* ix, const mmr_rerank_tables* t, const float* q_emb, const float* cand_emb,
               const int64_t* cand_rows, const int64_t* q_rec, const int64_t* cand_rec, const int32_t* cand_count,
               int32_t b, int32_t k, int32_t d, double alpha, double beta, double gamma, int32_t topk,
               int32_t* out_order, double* out_scores, void* stream_v) {
  MMR_REQUIRE(b >= 0 && k >= 0 && d >= 1 && topk >= 0, "mmr_rerank: bad sizes");
  if (b == 0 || k == 0) return MMR_OK;
  MMR_REQUIRE(out_order && out_scores, "mmr_rerank: NULL output");
  if (k > 4096) return fail(MMR_EUNSUP, "mmr_rerank: more than 4096 candidates per query");
  const int device = ix ? ix->device : (t ? t->device : 0);
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  const int keep = (topk > 0 && topk < k) ? topk : k;
  void* raw_v = nullptr;
  MMR_TRY(cs.alloc(static_cast<size_t>(b) * k * 3 * sizeof(double), &raw_v));
  double* d_raw = static_cast<double*>(raw_v);
  MMR_TRY(rerank_features_impl(cs, ix, t, q_emb, cand_emb, cand_rows, q_rec, cand_rec, cand_count, b, k, d, d_raw,
                               nullptr));
  const int32_t* d_cc;
  MMR_TRY(cs.in(cand_count, static_cast<size_t>(b), &d_cc));
  int32_t* d_order;
  double* d_sc;
  MMR_TRY(cs.out(out_order, static_cast<size_t>(b) * keep, &d_order));
  MMR_TRY(cs.out(out_scores, static_cast<size_t>(b) * keep * 4, &d_sc));
  MMR_TRY(launch_rerank_combine(d_raw, d_cc, b, k, alpha, beta, gamma, topk, d_order, d_sc, cs.stream));
  return cs.finish();
}

int mmr_rerank_scored(const mmr_rerank_tables* t, const int64_t* rows, const float* scores, const int64_t* q_rec,
                      int32_t b, int32_t k, double alpha, double beta, double gamma, int32_t topk, int64_t* out_ids,
                      double* out_final, double* out_scores4, int32_t device, void* stream_v) {
  MMR_REQUIRE(b >= 0 && k >= 0 && topk >= 0, "mmr_rerank_scored: bad sizes");
  if (b == 0 || k == 0) return MMR_OK;
  MMR_REQUIRE(rows && scores && out_ids && out_final, "mmr_rerank_scored: NULL argument");
  if (t != nullptr) device = t->device;
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  const int keep = (topk > 0 && topk < k) ? topk : k;
  const size_t bk = static_cast<size_t>(b) * k, bo = static_cast<size_t>(b) * keep;
  const int64_t *d_rows, *d_qr;
  const float* d_sc;
  int64_t* d_ids;
  double *d_fin, *d_s4;
  MMR_TRY(cs.in(rows, bk, &d_rows));
  MMR_TRY(cs.in(scores, bk, &d_sc));
  MMR_TRY(cs.in(q_rec, static_cast<size_t>(b), &d_qr));
  MMR_TRY(cs.out(out_ids, bo, &d_ids));
  MMR_TRY(cs.out(out_final, bo, &d_fin));
  MMR_TRY(cs.out(out_scores4, bo * 4, &d_s4));
  MMR_TRY(launch_rerank_scored(d_rows, d_sc, d_qr, t ? t->masks : nullptr, t ? t->label_words : 0, t ? t->kg : nullptr,
                               t ? t->d_kg : 0, t ? t->n_rec : 0, b, k, alpha, beta, gamma, topk, d_ids, d_fin, d_s4,
                               cs.stream));
  return cs.finish();
}

int mmr_candidate_cosine(const mmr_index* ix, const float* q_emb, const int64_t* cand_rows, int32_t b, int32_t k,
                         int32_t d, float* out_cos, uint8_t* owned, void* stream_v) {
  MMR_REQUIRE(ix != nullptr, "mmr_candidate_cosine: index is NULL");
  MMR_REQUIRE(b >= 0 && k >= 0 && d >= 1, "mmr_candidate_cosine: bad sizes");
  if (b == 0 || k == 0) return MMR_OK;
  MMR_REQUIRE(q_emb && cand_rows && out_cos, "mmr_candidate_cosine: NULL argument");
  DeviceGuard guard(ix->device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  float* d_cos;
  uint8_t* d_owned;
  MMR_TRY(cs.out(out_cos, static_cast<size_t>(b) * k, &d_cos));
  MMR_TRY(cs.out(owned, static_cast<size_t>(b) * k, &d_owned));
  MMR_TRY(rerank_features_impl(cs, ix, nullptr, q_emb, nullptr, cand_rows, nullptr, nullptr, nullptr, b, k, d, nullptr,
                               d_owned, nullptr, d_cos));
  return cs.finish();
}

int mmr_rerank_with_cos(const mmr_rerank_tables* t, const float* emb_cos, const int64_t* q_rec,
                        const int64_t* cand_rec, const int32_t* cand_count, int32_t b, int32_t k, double alpha,
                        double beta, double gamma, int32_t topk, int32_t* out_order, double* out_scores,
                        int32_t device, void* stream_v) {
  MMR_REQUIRE(b >= 0 && k >= 0 && topk >= 0, "mmr_rerank_with_cos: bad sizes");
  if (b == 0 || k == 0) return MMR_OK;
  MMR_REQUIRE(emb_cos && out_order && out_scores, "mmr_rerank_with_cos: NULL argument");
  if (k > 4096) return fail(MMR_EUNSUP, "mmr_rerank_with_cos: more than 4096 candidates per query");
  if (t != nullptr) device = t->device;
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  const int keep = (topk > 0 && topk < k) ? topk : k;
  void* raw_v = nullptr;
  MMR_TRY(cs.alloc(static_cast<size_t>(b) * k * 3 * sizeof(double), &raw_v));
  double* d_raw = static_cast<double*>(raw_v);
  MMR_TRY(rerank_features_impl(cs, nullptr, t, nullptr, nullptr, nullptr, q_rec, cand_rec, cand_count, b, k, 1, d_raw,
                               nullptr, emb_cos, nullptr));
  const int32_t* d_cc;
  MMR_TRY(cs.in(cand_count, static_cast<size_t>(b), &d_cc));
  int32_t* d_order;
  double* d_sc;
  MMR_TRY(cs.out(out_order, static_cast<size_t>(b) * keep, &d_order));
  MMR_TRY(cs.out(out_scores, static_cast<size_t>(b) * keep * 4, &d_sc));
  MMR_TRY(launch_rerank_combine(d_raw, d_cc, b, k, alpha, beta, gamma, topk, d_order, d_sc, cs.stream));
  return cs.finish();
}

int mmr_metrics(const int64_t* retrieved, const int32_t* ret_count, int32_t q, int32_t k_ret, const int64_t* rel_indptr,
                const int64_t* rel_sorted, const int64_t* rel_list_len, int32_t k, const double* log2_tbl, double* out,
                int32_t device, void* stream_v) {
  MMR_REQUIRE(q >= 0 && k_ret >= 0, "mmr_metrics: bad sizes");
  MMR_REQUIRE(k >= 1, "mmr_metrics: k must be >= 1");
  if (q == 0) return MMR_OK;
  MMR_REQUIRE(rel_indptr && log2_tbl && out && (retrieved || k_ret == 0), "mmr_metrics: NULL argument");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  // the CSR length lives in rel_indptr[q]: read it on the host if we can, else from the device
  int64_t nnz = 0;
  if (is_device_ptr(rel_indptr)) {
    MMR_CUDA_TRY(cudaMemcpyAsync(&nnz, rel_indptr + q, sizeof(int64_t), cudaMemcpyDeviceToHost, cs.stream));
    MMR_CUDA_TRY(cudaStreamSynchronize(cs.stream));
  } else {
    nnz = rel_indptr[q];
  }
  const int64_t* d_ret;
  const int32_t* d_rc;
  const int64_t *d_ip, *d_rs, *d_ll;
  const double* d_tbl;
  double* d_out;
  const int tbl_n = (k > k_ret ? k : k_ret) > 0 ? (k > k_ret ? k : k_ret) : 1;
  MMR_TRY(cs.in(retrieved, static_cast<size_t>(q) * k_ret, &d_ret));
  MMR_TRY(cs.in(ret_count, static_cast<size_t>(q), &d_rc));
  MMR_TRY(cs.in(rel_indptr, static_cast<size_t>(q) + 1, &d_ip));
  MMR_TRY(cs.in(rel_sorted, static_cast<size_t>(nnz), &d_rs));
  MMR_TRY(cs.in(rel_list_len, static_cast<size_t>(q), &d_ll));
  MMR_TRY(cs.in(log2_tbl, static_cast<size_t>(tbl_n), &d_tbl));
  MMR_TRY(cs.out(out, static_cast<size_t>(q) * 5, &d_out));
  MMR_TRY(launch_metrics(d_ret, d_rc, q, k_ret, d_ip, d_rs, d_ll, k, d_tbl, d_out, cs.stream));
  return cs.finish();
}

int mmr_first_relevant_rank(mmr_index* ix, const void* q, int32_t b, int32_t q_dtype, const uint64_t* q_masks,
                            const uint64_t* g_masks, int32_t label_words, int64_t* out_rank, int64_t* out_total,
                            void* stream_v) {
  MMR_REQUIRE(ix != nullptr, "mmr_first_relevant_rank: index is NULL");
  MMR_REQUIRE(b >= 0 && label_words >= 1, "mmr_first_relevant_rank: bad sizes");
  MMR_REQUIRE(q_dtype == MMR_F32 || q_dtype == MMR_BF16, "mmr_first_relevant_rank: bad q_dtype");
  if (b == 0) return MMR_OK;
  MMR_REQUIRE(q && q_masks && g_masks && out_rank && out_total, "mmr_first_relevant_rank: NULL argument");
  std::lock_guard<std::mutex> lock(ix->mu);
  DeviceGuard guard(ix->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  MMR_TRY(enter_call(ix, stream));
  // queries -> storage dtype (rounded like mmr_search) -> fp32 padded + inverse norms
  const void* d_q = nullptr;
  MMR_TRY(stage_in(q, static_cast<size_t>(b) * ix->d * elem_size(q_dtype), ix->q_in, stream, &d_q));
  MMR_TRY(ix->q_store.ensure(static_cast<size_t>(b) * ix->d_pad * elem_size(ix->dtype)));
  MMR_TRY(ix->q_inv.ensure(static_cast<size_t>(b) * sizeof(float)));
  MMR_TRY(launch_ingest(d_q, q_dtype, b, ix->d, ix->d, ix->q_store.p, ix->dtype, ix->d_pad, ix->q_inv.as<float>(),
                        stream));
  const float* q_f32 = ix->q_store.as<float>();
  if (ix->dtype == MMR_BF16) {
    MMR_TRY(ix->q_f32.ensure(static_cast<size_t>(b) * ix->d_pad * sizeof(float)));
    MMR_TRY(ix->scratch.ensure(static_cast<size_t>(b) * sizeof(float)));
    MMR_TRY(launch_ingest(ix->q_store.p, MMR_BF16, b, ix->d_pad, ix->d_pad, ix->q_f32.p, MMR_F32, ix->d_pad,
                          ix->scratch.as<float>(), stream));
    q_f32 = ix->q_f32.as<float>();
  }
  CallScope cs(stream);
  const uint64_t *d_qm, *d_gm;
  int64_t *d_rank, *d_total;
  MMR_TRY(cs.in(q_masks, static_cast<size_t>(b) * label_words, &d_qm));
  MMR_TRY(cs.in(g_masks, static_cast<size_t>(ix->n) * label_words, &d_gm));
  MMR_TRY(cs.out(out_rank, static_cast<size_t>(b), &d_rank));
  MMR_TRY(cs.out(out_total, static_cast<size_t>(b), &d_total));
  MMR_TRY(launch_first_relevant_rank(ix->emb, ix->dtype, ix->inv_norm, ix->n, ix->d_pad, q_f32, ix->q_inv.as<float>(), b,
                                     d_qm, d_gm, label_words, d_rank, d_total, stream));
  MMR_TRY(cs.finish());
  return leave_call(ix, stream);
}

int mmr_label_ranking_eval(const float* emb, const float* norms, int32_t n, int32_t d, const uint64_t* label_masks,
                           int32_t label_words, const int32_t* topk, int32_t n_topk, double* out, int32_t device,
                           void* stream_v) {
  MMR_REQUIRE(n >= 0 && d >= 1 && label_words >= 1 && n_topk >= 0 && n_topk <= 8, "mmr_label_ranking_eval: bad sizes");
  if (n == 0) return MMR_OK;
  MMR_REQUIRE(emb && norms && label_masks && out && (topk || n_topk == 0), "mmr_label_ranking_eval: NULL argument");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  const float *d_e, *d_n;
  const uint64_t* d_m;
  const int32_t* d_k;
  double* d_out;
  MMR_TRY(cs.in(emb, static_cast<size_t>(n) * d, &d_e));
  MMR_TRY(cs.in(norms, static_cast<size_t>(n), &d_n));
  MMR_TRY(cs.in(label_masks, static_cast<size_t>(n) * label_words, &d_m));
  MMR_TRY(cs.in(topk, static_cast<size_t>(n_topk), &d_k));
  MMR_TRY(cs.out(out, static_cast<size_t>(n) * (1 + n_topk), &d_out));
  MMR_TRY(launch_label_ranking(d_e, d_n, n, d, d_m, label_words, d_k, n_topk, d_out, cs.stream));
  return cs.finish();
}

int mmr_result_diversity(const float* emb, const uint64_t* label_masks, const int32_t* counts, int32_t b, int32_t k,
                         int32_t d, int32_t label_words, double* out_emb_div, double* out_label_div, int32_t device,
                         void* stream_v) {
  MMR_REQUIRE(b >= 0 && k >= 0 && d >= 0 && label_words >= 0, "mmr_result_diversity: bad sizes");
  if (b == 0) return MMR_OK;
  MMR_REQUIRE((emb && out_emb_div) || (label_masks && out_label_div), "mmr_result_diversity: nothing to compute");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  const float* d_e;
  const uint64_t* d_m;
  const int32_t* d_c;
  double *d_oe, *d_ol;
  MMR_TRY(cs.in(emb, static_cast<size_t>(b) * k * d, &d_e));
  MMR_TRY(cs.in(label_masks, static_cast<size_t>(b) * k * label_words, &d_m));
  MMR_TRY(cs.in(counts, static_cast<size_t>(b), &d_c));
  MMR_TRY(cs.out(out_emb_div, static_cast<size_t>(b), &d_oe));
  MMR_TRY(cs.out(out_label_div, static_cast<size_t>(b), &d_ol));
  MMR_TRY(launch_diversity(d_e, d_m, d_c, b, k, d, label_words, d_oe, d_ol, cs.stream));
  return cs.finish();
}

int mmr_label_relevance(const uint64_t* q_masks, int64_t nq, const uint64_t* g_masks, int64_t ng, int32_t label_words,
                        int32_t exclude_self, uint8_t* out_relevant, int32_t device, void* stream_v) {
  MMR_REQUIRE(nq >= 0 && ng >= 0 && label_words >= 1, "mmr_label_relevance: bad sizes");
  if (nq == 0 || ng == 0) return MMR_OK;
  MMR_REQUIRE(q_masks && g_masks && out_relevant, "mmr_label_relevance: NULL argument");
  MMR_TRY(check_device(device, nullptr));
  DeviceGuard guard(device);
  CallScope cs(static_cast<cudaStream_t>(stream_v));
  const uint64_t *d_q, *d_g;
  uint8_t* d_o;
  MMR_TRY(cs.in(q_masks, static_cast<size_t>(nq) * label_words, &d_q));
  MMR_TRY(cs.in(g_masks, static_cast<size_t>(ng) * label_words, &d_g));
  MMR_TRY(cs.out(out_relevant, static_cast<size_t>(nq) * ng, &d_o));
  MMR_TRY(launch_label_relevance(d_q, nq, d_g, ng, label_words, exclude_self, d_o, cs.stream));
  return cs.finish();
}

}  // extern "C"
