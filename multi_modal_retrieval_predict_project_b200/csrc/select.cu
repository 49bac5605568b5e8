// K4: on-device K-way merge / final selection.
//
// Replaces the tail of np.argsort(row)[::-1][:k] (reference Evaluate/retrieval_overlap.py:90) and
// heapq.nsmallest(K, heap) (Retrieval/retrieval.py:240): given per-(query, partition) candidate
// lists (from the scan CTAs, the GEMM CTAs, or the per-GPU lists after the NCCL all-gather) produce
// the global top-K per query, best first, ordered by (score desc, row asc).
//
// Three kernels, all on packed 64-bit keys (ordered score, inverted row: score desc, row asc):
//   select_warp_kernel  one WARP per query over the GEMM's per-(query, list) candidate lists, k <= 128 -- the
//                       selection of every batched search step (and, with RemoteSink, its multi-GPU scatter);
//   select_fast_kernel  one CTA per query, <= 4096 candidate slots, k <= 128 (merge of per-GPU / per-scan-CTA
//                       lists; two passes above 4096 slots);
//   select_topk_kernel  one CTA per query, any k <= 1024: shared-memory bitonic network, more candidates than
//                       fit (8192 keys) are consumed in rounds that keep the running top-K.
// Latency/launch-bound, tiny next to the search kernels (reported as time only).
#include "internal.h"

namespace mmr {
namespace {

constexpr int kMaxSortKeys = 8192;  // 64 KiB of shared memory

struct KeySourceFlat {  // keys contiguous per query
  const uint64_t* keys;
  int64_t per_query;
  __device__ __forceinline__ int64_t count(int) const { return per_query; }
  __device__ __forceinline__ uint64_t get(int q, int64_t i) const { return keys[q * per_query + i]; }
};

struct KeySourceVar {  // GEMM candidates: (b, n_parts, cap) raw {score bits, local row}, first counts[] valid
  const uint2* cand;
  const int32_t* counts;
  int n_parts;
  int cap;
  int per_part;            // entries considered per part (the kernel leaves <= k valid ones)
  const int64_t* exclude;  // optional (b) local row to drop per query
  __device__ __forceinline__ int64_t count(int) const { return static_cast<int64_t>(n_parts) * per_part; }
  __device__ __forceinline__ uint64_t get(int q, int64_t i) const {
    // n_parts * per_part < 2^31 (at most a few thousand lists of <= 1025 entries): 32-bit division
    const int part = static_cast<int>(static_cast<uint32_t>(i) / static_cast<uint32_t>(per_part));
    const int j = static_cast<int>(static_cast<uint32_t>(i) - static_cast<uint32_t>(part) * static_cast<uint32_t>(per_part));
#ifdef MMR_DIAG
    if (part >= n_parts || j >= cap) __trap();  // bounds-checked build
#endif
    // the entry is loaded whether or not it is valid (slots past the count are allocated, just stale): the
    // count and the entry come back in ONE round trip instead of two dependent ones
    const int c = __ldg(counts + static_cast<int64_t>(q) * n_parts + part);
    const uint2 e = __ldcg(cand + (static_cast<int64_t>(q) * n_parts + part) * cap + j);
    if (j >= (c < per_part ? c : per_part)) return 0ull;
    if (exclude != nullptr && static_cast<int64_t>(e.y) == exclude[q]) return 0ull;
    return make_key(__uint_as_float(e.x), e.y);
  }
};

struct KeySourceLists {  // n_lists x (b, k_in) score/row pairs with GLOBAL rows; list l starts at l * stride
  const float* scores;
  const int64_t* rows;
  int n_lists, b, k_in;
  int64_t s_stride, r_stride;  // elements between consecutive lists
  __device__ __forceinline__ int64_t count(int) const { return static_cast<int64_t>(n_lists) * k_in; }
  __device__ __forceinline__ uint64_t get(int q, int64_t i) const {
    const int l = static_cast<int>(i / k_in);
    const int j = static_cast<int>(i - static_cast<int64_t>(l) * k_in);
    const int64_t at = static_cast<int64_t>(q) * k_in + j;
    const int64_t r = rows[l * r_stride + at];
    return r < 0 ? 0ull : make_key(scores[l * s_stride + at], static_cast<uint32_t>(r));
  }
};

template <typename Source>
__global__ void __launch_bounds__(1024, 1)
select_topk_kernel(Source src, int k_out, int sz, int keep, int64_t row_offset, float* __restrict__ out_scores,
                   int64_t* __restrict__ out_rows, int32_t* __restrict__ out_src) {
  extern __shared__ __align__(16) uint64_t s[];
  const int q = blockIdx.x;
  const int64_t total = src.count(q);
  int64_t pos = 0;
  bool first = true;
  do {
    const int base = first ? 0 : keep;
    const int64_t room = sz - base;
    const int64_t chunk = (total - pos) < room ? (total - pos) : room;
    for (int i = threadIdx.x; i < sz - base; i += blockDim.x) {
      s[base + i] = (i < chunk) ? src.get(q, pos + i) : 0ull;
    }
    __syncthreads();
    block_bitonic_sort_desc(s, sz);
    pos += chunk;
    first = false;
  } while (pos < total);

  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const uint64_t key = (i < sz) ? s[i] : 0ull;
    const int64_t o = static_cast<int64_t>(q) * k_out + i;
    if (key == 0ull) {
      out_scores[o] = -INFINITY;
      out_rows[o] = -1;
      if (out_src != nullptr) out_src[o] = -1;
    } else {
      out_scores[o] = key_score(key);
      out_rows[o] = row_offset + static_cast<int64_t>(key_row(key));
      if (out_src != nullptr) {
        int found = -1;
        for (int64_t j = 0; j < total; ++j) {
          if (src.get(q, j) == key) {
            found = static_cast<int>(j);
            break;
          }
        }
        out_src[o] = found;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path (k <= 128, <= 4096 candidates per query): hierarchical warp selection, no block-wide
// sort.  Each of 8 warps holds 512 candidates in registers (16 keys per lane), finds the threshold
// of its local top-k with an MSB-first radix descent (warp reductions only) and keeps <= k
// survivors; warp 0 repeats the selection over the 8 x 128 survivors and bitonic-sorts the final
// <= 128 keys.  ~10x cheaper than sorting 4096 keys with a CTA-wide bitonic network.
// ---------------------------------------------------------------------------------------------
// count(key >= T) == k threshold, MSB-first radix descent on the score word first and on the row
// word only to break ties on the k-th score (see gemm_topk.cu: warp_rank_threshold)
template <int E>
__device__ __forceinline__ uint64_t warp_topk_threshold(const uint64_t (&key)[E], int k) {
  uint32_t prefix = 0;
  bool exact = false;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (static_cast<uint32_t>(key[e] >> 32) >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) {
      prefix = cand;
      if (c == k) {
        exact = true;
        break;
      }
    }
  }
  if (exact) return static_cast<uint64_t>(prefix) << 32;
  int above = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) above += (static_cast<uint32_t>(key[e] >> 32) > prefix) ? 1 : 0;
  above = __reduce_add_sync(0xffffffffu, above);
  const int need = k - above;
  uint32_t low = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = low | (1u << bit);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e)
      c += (static_cast<uint32_t>(key[e] >> 32) == prefix && static_cast<uint32_t>(key[e]) >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= need) {
      low = cand;
      if (c == need) break;
    }
  }
  return (static_cast<uint64_t>(prefix) << 32) | low;  // 0 when fewer than k valid keys: everything non-zero survives
}

// survivors (key >= thr) of a warp's register-resident keys -> dst[0..cap), with the index each key
// came from (idx_of(e) for register slot e of this lane) carried along; the tail is zero padded
template <int E, typename IdxFn>
__device__ __forceinline__ void warp_write_survivors(const uint64_t (&key)[E], uint64_t thr, uint64_t* dst,
                                                     int32_t* dst_idx, int cap, int lane, IdxFn idx_of) {
  int base = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool keep = key[e] >= thr && key[e] != 0ull;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    const int pos = base + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < cap) {
      dst[pos] = key[e];
      dst_idx[pos] = idx_of(e);
    }
    base += __popc(m);
  }
  for (int i = base + lane; i < cap; i += 32) {
    dst[i] = 0ull;
    dst_idx[i] = -1;
  }
}

// warp bitonic sort (descending) of n keys with a 32-bit payload
__device__ __forceinline__ void warp_bitonic_sort_desc_kv(uint64_t* s, int32_t* v, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += kWarp) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = s[i], b = s[ixj];
          if ((a < b) == ((i & k) == 0)) {
            s[i] = b;
            s[ixj] = a;
            const int32_t t = v[i];
            v[i] = v[ixj];
            v[ixj] = t;
          }
        }
      }
      __syncwarp();
    }
  }
}

// grid = (queries, chunks): chunk c selects among candidates [c * 4096, (c + 1) * 4096).  With
// `stage_keys` != nullptr the (<= 128, zero padded) surviving keys go to stage_keys[q][c][128] for a
// second pass instead of being decoded into scores / rows.
//
// Pruning bound: the GEMM kernel publishes, per (list, query), a lower bound of the list's r-th best
// score with r * n_lists >= k (gemm_topk.cu); when every list has published, the union holds >= k
// candidates >= g = min over lists, so everything below g is dropped before any selection work.
// When at most kRankCap candidates survive (typically 150-300 of 1800), the CTA ranks them by
// counting (each thread counts the keys larger than its own: keys are unique, so the count IS the
// output position) -- no radix descent and no sort.
struct NoBound {
  __device__ __forceinline__ uint32_t get(int, int) const { return 0u; }
};
struct TauBound {  // tau_pub[list][b_pad] of ordered-u32 scores, 0 = not published
  const uint32_t* tau_pub;
  int n_lists, b_pad;
  __device__ __forceinline__ uint32_t get(int q, int lane) const {
    if (tau_pub == nullptr) return 0u;
    uint32_t g = 0xFFFFFFFFu;
    for (int p = lane; p < n_lists; p += 32) {
      const uint32_t v = __ldcg(tau_pub + static_cast<int64_t>(p) * b_pad + q);
      g = v < g ? v : g;
    }
    return __reduce_min_sync(0xffffffffu, g);
  }
};
constexpr int kRankCap = 512;

// Where the selected (score, global row) pairs go: dense (b, k_out) arrays, or -- the fused select + scatter
// of the sharded path -- straight into the exchange region of the rank that owns the query (16 x 4/8-byte
// NVLink stores per query instead of a second kernel that re-reads the dense arrays).
// The kernel stages a query's result (<= 128 slots) in shared memory and the sink flushes it with `nt` threads.
struct DenseSink {
  float* scores;
  int64_t* rows;
  int32_t* src;
  int k_out;
  __device__ __forceinline__ int limit() const { return k_out; }
  __device__ __forceinline__ void flush(int q, const float* o_sc, const int64_t* o_row, const int32_t* o_src, int t,
                                        int nt) const {
    const int64_t base = static_cast<int64_t>(q) * k_out;
    for (int i = t; i < k_out; i += nt) {
      scores[base + i] = o_sc[i];
      rows[base + i] = o_row[i];
      if (src != nullptr) src[base + i] = o_src[i];
    }
  }
};
struct RemoteSink {
  PeerSink s;
  __device__ __forceinline__ int limit() const { return s.kp; }
  // 16-byte NVLink stores (kp is a multiple of 4 and the lists are 256-byte aligned in the region): 4- and 8-byte
  // peer stores, one packet per element, made this kernel twice as slow as its dense form
  __device__ __forceinline__ void flush(int q, const float* o_sc, const int64_t* o_row, const int32_t*, int t,
                                        int nt) const {
    const int dest = q / s.per;
    const int64_t slot = (static_cast<int64_t>(s.my_rank) * s.per + (q - dest * s.per)) * s.kp;
#ifdef MMR_DIAG
    if (dest < 0 || dest >= kMaxWorld || s.peers.base[dest] == nullptr) __trap();  // bounds-checked build
#endif
    float4* d_sc = reinterpret_cast<float4*>(s.peers.base[dest] + s.off_scores + slot * 4);
    longlong2* d_row = reinterpret_cast<longlong2*>(s.peers.base[dest] + s.off_rows + slot * 8);
    const float4* s_sc4 = reinterpret_cast<const float4*>(o_sc);
    const longlong2* s_row2 = reinterpret_cast<const longlong2*>(o_row);
    for (int v = t; v < s.kp / 4; v += nt) d_sc[v] = s_sc4[v];
    for (int v = t; v < s.kp / 2; v += nt) d_row[v] = s_row2[v];
    // one system-scope fence per warp that stored something (<= 75 vectors: the first three warps), not one
    // per thread of the CTA: the signal kernel's release follows in stream order
    if (t < ((s.kp / 2 + 31) & ~31)) __threadfence_system();
  }
};

template <typename Source, typename Bound, typename Sink>
#ifndef MMR_SELECT_MIN_CTAS
#define MMR_SELECT_MIN_CTAS 4   // 64 registers: four CTAs per SM (the kernel is a chain of L2 round trips)
#endif
__global__ void __launch_bounds__(256, MMR_SELECT_MIN_CTAS)
select_fast_kernel(Source src, Bound bound, int k_out, int64_t row_offset, Sink sink, uint64_t* __restrict__ stage_keys) {
  __shared__ __align__(16) uint64_t lvl2[8 * 128];
  __shared__ int32_t lvl2_idx[8 * 128];
  __shared__ __align__(16) uint64_t fin[128];
  __shared__ int32_t fin_idx[128];
  __shared__ __align__(16) float o_sc[128];      // the query's result, staged for the sink's (vectorised) flush
  __shared__ __align__(16) int64_t o_row[128];
  __shared__ int32_t o_src[128];
  __shared__ int s_count;
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = src.count(q);
  const int64_t chunk_base = static_cast<int64_t>(blockIdx.y) * 4096;
  const uint64_t gkey = static_cast<uint64_t>(bound.get(q, lane)) << 32;  // keys below are out of the top-k
  const bool empty_slice = chunk_base + static_cast<int64_t>(warp) * 512 >= total;
  uint64_t key[16];
  int mine = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int64_t i = chunk_base + static_cast<int64_t>(warp) * 512 + e * 32 + lane;
    uint64_t kk = (!empty_slice && i < total) ? src.get(q, i) : 0ull;
    kk = kk < gkey ? 0ull : kk;
    key[e] = kk;
    mine += kk != 0ull ? 1 : 0;
  }
  if (stage_keys == nullptr) {
    // ---- rank-by-counting path for small survivor sets ----
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int wsum = __reduce_add_sync(0xffffffffu, mine);
    int wbase = 0;
    if (lane == 0 && wsum > 0) wbase = atomicAdd(&s_count, wsum);
    __syncthreads();
    const int m = s_count;
    if (m <= kRankCap) {
      // compact the survivors into lvl2 (keys) / lvl2_idx (source index); order is irrelevant
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int pos = wbase + incl - mine;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (key[e] != 0ull) {
          lvl2[pos] = key[e];
          lvl2_idx[pos] = warp * 512 + e * 32 + lane;
          ++pos;
        }
      }
      __syncthreads();
      const uint64_t k0 = static_cast<int>(threadIdx.x) < m ? lvl2[threadIdx.x] : 0ull;
      const uint64_t k1 = static_cast<int>(threadIdx.x) + 256 < m ? lvl2[threadIdx.x + 256] : 0ull;
      int r0 = 0, r1 = 0;
      if ((m & 1) != 0 && threadIdx.x == 0) lvl2[m] = 0ull;   // pad to an even count (kRankCap < 8 * 128 slots)
      __syncthreads();
      // warps without a key skip the count; 16-byte shared loads (two keys, broadcast); the second key of a
      // thread exists only when more than 256 candidates survived
      if (warp * 32 < m) {
        const ulonglong2* pairs = reinterpret_cast<const ulonglong2*>(lvl2);
        const int np = (m + 1) >> 1;
        if (m <= 256) {
#pragma unroll 4
          for (int i = 0; i < np; ++i) {
            const ulonglong2 kk = pairs[i];
            r0 += (kk.x > k0 ? 1 : 0) + (kk.y > k0 ? 1 : 0);
          }
        } else {
#pragma unroll 2
          for (int i = 0; i < np; ++i) {
            const ulonglong2 kk = pairs[i];
            r0 += (kk.x > k0 ? 1 : 0) + (kk.y > k0 ? 1 : 0);
            r1 += (kk.x > k1 ? 1 : 0) + (kk.y > k1 ? 1 : 0);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t k64 = h == 0 ? k0 : k1;
        const int r = h == 0 ? r0 : r1;
        if (k64 != 0ull && r < k_out) {
          o_sc[r] = key_score(k64);
          o_row[r] = row_offset + static_cast<int64_t>(key_row(k64));
          o_src[r] = lvl2_idx[threadIdx.x + h * 256];
        }
      }
      // fewer candidates than k_out, and the sink's padding beyond k_out
      for (int i = (m < k_out ? m : k_out) + threadIdx.x; i < sink.limit(); i += blockDim.x) {
        o_sc[i] = -INFINITY;
        o_row[i] = -1;
        o_src[i] = -1;
      }
      __syncthreads();
      sink.flush(q, o_sc, o_row, o_src, threadIdx.x, blockDim.x);
      return;
    }
    __syncthreads();  // lvl2 is reused below
  }
  if (empty_slice) {  // nothing in this warp's slice
    for (int i = lane; i < 128; i += 32) {
      lvl2[warp * 128 + i] = 0ull;
      lvl2_idx[warp * 128 + i] = -1;
    }
  } else {
    const uint64_t thr = warp_topk_threshold<16>(key, k_out);
    warp_write_survivors<16>(key, thr, lvl2 + warp * 128, lvl2_idx + warp * 128, 128, lane,
                             [&](int e) { return warp * 512 + e * 32 + lane; });
  }
  __syncthreads();
  if (warp == 0) {
    uint64_t key[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) key[e] = lvl2[e * 32 + lane];
    const uint64_t thr = warp_topk_threshold<32>(key, k_out);
    warp_write_survivors<32>(key, thr, fin, fin_idx, 128, lane, [&](int e) { return lvl2_idx[e * 32 + lane]; });
    __syncwarp();
    if (stage_keys != nullptr) {  // first pass of a two-pass selection: hand the survivors on unsorted
      uint64_t* dst = stage_keys + (static_cast<int64_t>(q) * gridDim.y + blockIdx.y) * 128;
      for (int i = lane; i < 128; i += 32) dst[i] = fin[i];
      return;
    }
    warp_bitonic_sort_desc_kv(fin, fin_idx, 128, lane);
    for (int i = lane; i < sink.limit(); i += 32) {
      const uint64_t k64 = (i < k_out) ? fin[i] : 0ull;
      o_sc[i] = k64 == 0ull ? -INFINITY : key_score(k64);
      o_row[i] = k64 == 0ull ? -1 : row_offset + static_cast<int64_t>(key_row(k64));
      o_src[i] = k64 == 0ull ? -1 : fin_idx[i];
    }
    __syncwarp();
    sink.flush(q, o_sc, o_row, o_src, lane, 32);
  }
}

// ---------------------------------------------------------------------------------------------
// One WARP per query for the GEMM's candidate lists (k <= 128): the default selection of a batched search step.
//
// select_fast_kernel above spends a CTA on a query: every thread loads 16 of the n_parts x per_part SLOTS (most
// of them past their list's count), and five block-wide barriers separate the phases -- 63 M warp instructions
// and 92 us for 4096 queries x 18 lists, half in the slot loads and half in the O(m^2) rank-by-counting over
// the m ~ 220 survivors of the pruning bound.  Here a warp reads the lists' counts first and then only the valid
// entries (8 lists in flight per round trip), keeps the survivors in 4 KB of shared memory (a radix-descent cut
// to the top-k whenever 512 are held, and once more if more than 128 are left at the end), and ranks the <= 128
// finalists by counting with four keys per lane.  No block-wide barrier; 4096 queries are one wave of warps.
// ---------------------------------------------------------------------------------------------
#ifndef MMR_SELECT_WARP
#define MMR_SELECT_WARP 1
#endif
#ifndef MMR_SELECT_WARP_CTAS
#define MMR_SELECT_WARP_CTAS 5   // register cap 102: 72 (dense sink) / 84 (peer sink) used, no spills; measured for 4096 queries x 18 lists: 5 -> 34 us, 6 -> 42, 8 (64 registers) -> 36
#endif
constexpr int kWsCap = 512;   // keys a warp holds (16 per lane in the cut)
constexpr int kWsWarps = 4;   // queries per CTA

// the top-k of buf[0..cnt), cnt <= 512, moved to the front of buf; returns how many there are (k, or all valid ones)
__device__ __forceinline__ int warp_keep_topk(uint64_t* buf, int cnt, int k, int lane) {
  uint64_t key[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int i = e * 32 + lane;
    key[e] = i < cnt ? buf[i] : 0ull;
  }
  const uint64_t thr = warp_topk_threshold<16>(key, k);
  __syncwarp();
  int base = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const bool keep = key[e] >= thr && key[e] != 0ull;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(m & ((1u << lane) - 1u))] = key[e];
    base += __popc(m);
  }
  __syncwarp();
  return base;
}

template <typename Bound, typename Sink>
__global__ void __launch_bounds__(kWsWarps * 32, MMR_SELECT_WARP_CTAS)
select_warp_kernel(KeySourceVar src, Bound bound, int b, int k_out, int64_t row_offset, Sink sink) {
  __shared__ __align__(16) uint64_t s_key[kWsWarps][kWsCap];
  __shared__ __align__(16) float s_sc[kWsWarps][128];
  __shared__ __align__(16) int64_t s_row[kWsWarps][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * kWsWarps + warp;
  if (q >= b) return;  // warps are independent: no block-wide barrier below
  uint64_t* buf = s_key[warp];
  const uint64_t gkey = static_cast<uint64_t>(bound.get(q, lane)) << 32;  // keys below are out of the top-k
  const int64_t excl = src.exclude != nullptr ? src.exclude[q] : -1;
#ifdef MMR_DIAG
  if (src.per_part > src.cap) __trap();  // bounds-checked build
#endif
  int cnt = 0;  // keys held in buf (warp-uniform)
  for (int p0 = 0; p0 < src.n_parts; p0 += 32) {
    const int np = src.n_parts - p0 < 32 ? src.n_parts - p0 : 32;
    int c = 0;  // lane l: valid entries of list p0 + l
    if (lane < np) {
      c = __ldg(src.counts + static_cast<int64_t>(q) * src.n_parts + p0 + lane);
      c = c < 0 ? 0 : (c < src.per_part ? c : src.per_part);
    }
    const int max_c = __reduce_max_sync(0xffffffffu, c);
    const uint2* base = src.cand + (static_cast<int64_t>(q) * src.n_parts + p0) * src.cap;
    for (int j0 = 0; j0 < max_c; j0 += 32) {      // one pass for lists of <= 32 entries (the usual case)
      for (int pp = 0; pp < np; pp += 8) {        // 8 lists' loads in flight
        uint2 e[8];
        bool ok[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int cp = __shfl_sync(0xffffffffu, c, (pp + u) & 31);
          ok[u] = pp + u < np && j0 + lane < cp;
          e[u] = ok[u] ? __ldcg(base + static_cast<int64_t>(pp + u) * src.cap + j0 + lane) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint64_t key = ok[u] ? make_key(__uint_as_float(e[u].x), e[u].y) : 0ull;
          if (static_cast<int64_t>(e[u].y) == excl) key = 0ull;
          key = key < gkey ? 0ull : key;
          const uint32_t m = __ballot_sync(0xffffffffu, key != 0ull);
          if (m != 0u) {  // warp-uniform
            if (cnt + 32 > kWsCap) {
              __syncwarp();
              cnt = warp_keep_topk(buf, cnt, k_out, lane);
            }
            if (key != 0ull) buf[cnt + __popc(m & ((1u << lane) - 1u))] = key;
            cnt += __popc(m);
          }
        }
      }
    }
  }
  __syncwarp();
#ifdef MMR_DIAG
  if (cnt > kWsCap) __trap();
#endif
  if (cnt > 128) cnt = warp_keep_topk(buf, cnt, k_out, lane);  // k_out <= 128 finalists
  // rank by counting: keys are unique, so the number of larger keys IS the output position
  if ((cnt & 1) != 0 && lane == 0) buf[cnt] = 0ull;  // pad to an even count (cnt <= 128 < kWsCap)
  __syncwarp();
  uint64_t k4[4];
  int r4[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = e * 32 + lane;
    k4[e] = i < cnt ? buf[i] : 0ull;
    r4[e] = 0;
  }
  {
    const ulonglong2* pairs = reinterpret_cast<const ulonglong2*>(buf);
    const int npairs = (cnt + 1) >> 1;
#pragma unroll 2
    for (int i = 0; i < npairs; ++i) {
      const ulonglong2 kk = pairs[i];
#pragma unroll
      for (int e = 0; e < 4; ++e) r4[e] += (kk.x > k4[e] ? 1 : 0) + (kk.y > k4[e] ? 1 : 0);
    }
  }
  float* o_sc = s_sc[warp];
  int64_t* o_row = s_row[warp];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (k4[e] != 0ull && r4[e] < k_out) {
      o_sc[r4[e]] = key_score(k4[e]);
      o_row[r4[e]] = row_offset + static_cast<int64_t>(key_row(k4[e]));
    }
  }
  // fewer candidates than k_out, and the sink's padding beyond k_out
  for (int i = (cnt < k_out ? cnt : k_out) + lane; i < sink.limit(); i += 32) {
    o_sc[i] = -INFINITY;
    o_row[i] = -1;
  }
  __syncwarp();
  sink.flush(q, o_sc, o_row, nullptr, lane, 32);
}

// dense (b, k) lists -> the owner ranks' regions (shapes the fused select + scatter does not cover)
__global__ void scatter_lists_kernel(const float* __restrict__ scores, const int64_t* __restrict__ rows, int b, int k,
                                     RemoteSink sink) {
  __shared__ __align__(16) float o_sc[MMR_MAX_K];
  __shared__ __align__(16) int64_t o_row[MMR_MAX_K];
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < sink.limit(); i += blockDim.x) {
    const bool ok = i < k && rows[static_cast<int64_t>(q) * k + i] >= 0;
    o_sc[i] = ok ? scores[static_cast<int64_t>(q) * k + i] : -INFINITY;
    o_row[i] = ok ? rows[static_cast<int64_t>(q) * k + i] : -1;
  }
  __syncthreads();
  sink.flush(q, o_sc, o_row, nullptr, threadIdx.x, blockDim.x);
}

// payload carried through a merge: out[q][i] = payload[list * stride + q * k_in + j] for the source
// position src[q][i] = list * k_in + j reported by the merge (0 where src < 0)
__global__ void gather_payload_kernel(const float* __restrict__ payload, int64_t list_stride,
                                      const int32_t* __restrict__ src, int64_t total, int k_in, int k_out,
                                      float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int s = src[i];
  const int64_t q = i / k_out;
  out[i] = s < 0 ? 0.f : payload[static_cast<int64_t>(s / k_in) * list_stride + q * k_in + (s % k_in)];
}

// what retrieve(..., reranker=...) returns (Retrieval/retrieval.py:257-269): candidate ids in the
// reranked order + the combined score
__global__ void apply_order_kernel(const int64_t* __restrict__ rows, const int32_t* __restrict__ order,
                                   const double* __restrict__ scores4, int64_t total, int k, int keep,
                                   int64_t* __restrict__ out_rows, double* __restrict__ out_final) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int o = order[i];
  const int64_t q = i / keep;
  out_rows[i] = o < 0 ? -1 : rows[q * k + o];
  out_final[i] = scores4[i * 4];
}

template <typename Source, typename Bound = NoBound>
int launch_select(Source src, int b, int64_t per_query, int k_out, int64_t row_offset, float* out_scores,
                  int64_t* out_rows, int32_t* out_src, cudaStream_t stream, Bound bound = Bound(),
                  const PeerSink* peer = nullptr) {
  if (b == 0 || k_out == 0) return MMR_OK;
  if (peer != nullptr) {
    if (k_out <= 128 && per_query <= 4096) {  // fused: the selection's output stores ARE the exchange
      select_fast_kernel<Source, Bound, RemoteSink><<<b, 256, 0, stream>>>(src, bound, k_out, row_offset,
                                                                          RemoteSink{*peer}, nullptr);
      MMR_LAUNCHED();
      return MMR_OK;
    }
    MMR_REQUIRE(out_scores != nullptr && out_rows != nullptr, "select: staging buffers needed for the scatter");
    MMR_TRY(launch_select(src, b, per_query, k_out, row_offset, out_scores, out_rows, nullptr, stream, bound, nullptr));
    return launch_scatter_lists(out_scores, out_rows, b, k_out, *peer, stream);
  }
  if (k_out <= 128 && per_query <= 4096) {
    select_fast_kernel<Source, Bound, DenseSink><<<b, 256, 0, stream>>>(
        src, bound, k_out, row_offset, DenseSink{out_scores, out_rows, out_src, k_out}, nullptr);
    MMR_LAUNCHED();
    return MMR_OK;
  }
  if (k_out <= 128 && out_src == nullptr && per_query <= 32 * 4096) {
    // two passes: chunks of 4096 candidates -> <= 128 survivors each -> one final selection
    const int chunks = static_cast<int>((per_query + 4095) / 4096);
    uint64_t* stage = nullptr;
    MMR_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&stage), static_cast<size_t>(b) * chunks * 128 * sizeof(uint64_t),
                                 stream));
    select_fast_kernel<Source, Bound, DenseSink><<<dim3(b, chunks), 256, 0, stream>>>(
        src, bound, k_out, row_offset, DenseSink{nullptr, nullptr, nullptr, k_out}, stage);
    count_launch();
    KeySourceFlat flat{stage, static_cast<int64_t>(chunks) * 128};
    select_fast_kernel<KeySourceFlat, NoBound, DenseSink><<<b, 256, 0, stream>>>(
        flat, NoBound(), k_out, row_offset, DenseSink{out_scores, out_rows, nullptr, k_out}, nullptr);
    count_launch();
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(stage, stream);
    if (e != cudaSuccess) return fail(MMR_ECUDA, std::string("select: ") + cudaGetErrorString(e));
    return MMR_OK;
  }
  const int keep = next_pow2(k_out);
  int sz;
  if (per_query <= kMaxSortKeys) {
    sz = next_pow2(static_cast<int>(per_query < 2 ? 2 : per_query));
    if (sz < keep) sz = keep;
  } else {
    sz = kMaxSortKeys;
    if (keep * 2 > sz) return fail(MMR_EUNSUP, "select: k too large for the merge kernel");
  }
  if (sz < 2) sz = 2;
  int threads = sz / 2;
  threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
  const size_t smem = static_cast<size_t>(sz) * sizeof(uint64_t);
  auto kern = select_topk_kernel<Source>;
  MMR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<b, threads, smem, stream>>>(src, k_out, sz, keep, row_offset, out_scores, out_rows, out_src);
  MMR_LAUNCHED();
  return MMR_OK;
}

}  // namespace

int launch_select_keys(const uint64_t* keys, int b, int64_t keys_per_query, int k_out, int64_t row_offset,
                       float* out_scores, int64_t* out_rows, cudaStream_t stream, const PeerSink* sink) {
  KeySourceFlat src{keys, keys_per_query};
  return launch_select(src, b, keys_per_query, k_out, row_offset, out_scores, out_rows, nullptr, stream, NoBound(),
                       sink);
}

int launch_scatter_lists(const float* scores, const int64_t* rows, int b, int k, const PeerSink& sink,
                         cudaStream_t stream) {
  if (b == 0) return MMR_OK;
  scatter_lists_kernel<<<b, 128, 0, stream>>>(scores, rows, b, k, RemoteSink{sink});
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_select_var(const uint64_t* cand, const int32_t* counts, int b, int n_parts, int cap, int per_part,
                      int k_out, int64_t row_offset, const int64_t* exclude_local, const uint32_t* tau_pub, int b_pad,
                      float* out_scores, int64_t* out_rows, cudaStream_t stream, const PeerSink* sink) {
  KeySourceVar src{reinterpret_cast<const uint2*>(cand), counts, n_parts, cap, per_part, exclude_local};
  TauBound bound{tau_pub, n_parts, b_pad};
#if MMR_SELECT_WARP
  if (k_out <= 128) {  // a warp per query (any number of lists)
    if (b == 0 || k_out == 0) return MMR_OK;
    const unsigned grid = static_cast<unsigned>((b + kWsWarps - 1) / kWsWarps);
    if (sink != nullptr) {  // fused: the selection's output stores ARE the exchange
      select_warp_kernel<TauBound, RemoteSink><<<grid, kWsWarps * 32, 0, stream>>>(src, bound, b, k_out, row_offset,
                                                                                   RemoteSink{*sink});
    } else {
      select_warp_kernel<TauBound, DenseSink><<<grid, kWsWarps * 32, 0, stream>>>(
          src, bound, b, k_out, row_offset, DenseSink{out_scores, out_rows, nullptr, k_out});
    }
    MMR_LAUNCHED();
    return MMR_OK;
  }
#endif
  return launch_select(src, b, static_cast<int64_t>(n_parts) * per_part, k_out, row_offset, out_scores, out_rows,
                       nullptr, stream, bound, sink);
}

int launch_gather_payload(const float* payload, int64_t list_stride, const int32_t* src, int b, int k_in, int k_out,
                          float* out, cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(b) * k_out;
  if (total == 0) return MMR_OK;
  gather_payload_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(payload, list_stride, src, total,
                                                                                        k_in, k_out, out);
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_apply_order(const int64_t* rows, const int32_t* order, const double* scores4, int b, int k, int keep,
                       int64_t* out_rows, double* out_final, cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(b) * keep;
  if (total == 0) return MMR_OK;
  apply_order_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(rows, order, scores4, total, k,
                                                                                     keep, out_rows, out_final);
  MMR_LAUNCHED();
  return MMR_OK;
}

int launch_merge_lists(const float* scores, const int64_t* rows, int n_lists, int b, int k_in,
                       int64_t scores_list_stride, int64_t rows_list_stride, int k_out, float* out_scores,
                       int64_t* out_rows, int32_t* out_src, cudaStream_t stream) {
  KeySourceLists src{scores, rows, n_lists, b, k_in, scores_list_stride, rows_list_stride};
  return launch_select(src, b, static_cast<int64_t>(n_lists) * k_in, k_out, 0, out_scores, out_rows, out_src,
                       stream);
}

}  // namespace mmr
