// K1 + K3: tcgen05 / TMEM bf16 GEMM with a fused inverse-norm scale + top-K epilogue.
//
// Replaces cosine_similarity(Q, G) + np.argsort(row)[::-1][:k] for a BATCH of queries (reference
// Evaluate/retrieval_overlap.py:85,90; Retrieval/retrieval.py:128,134): the (Q, N) score matrix is
// never written to HBM.
//
//   scores[m, n] = sum_d Q[m, d] * G[n, d]      Q: (b, d_pad) bf16, G: (n, d_pad) bf16, both K-major
//
// Work unit = a CTA PAIR (2-CTA cluster, cta_group::2; a single CTA when the batch has one query tile):
// 2 x 128 queries x a contiguous range of gallery tiles of 256 rows.  Per CTA, 12 warps:
//   warps 0-7 epilogue, two warpgroups; thread <-> query row (TMEM lane).  BOTH warpgroups drain every
//             accumulator tile, warpgroup g taking columns [128 g, 128 g + 128): with two TMEM stages the MMA of
//             tile t + 2 can only start when tile t has been drained, so the drain has one tile time and is halved
//             by splitting the columns.  tcgen05.ld 32 columns -> t = acc * inv_norm(g) -> chunk max; only if some
//             lane's max beats its threshold does the warp run the append path: one REDUX ORs the lanes' masks of
//             hit 4-column groups, each hit group is re-read from TMEM, rescaled and appended with predicated
//             stores (a 32-bit cursor) to the (query, part, group) candidate list in global memory (L2-resident,
//             rarely written); when a list fills, the warp selects the exact top-k with a bitwise radix descent
//             over 64-bit keys held in registers, compacts the list and tightens the threshold.  The stage goes
//             back to the MMA issuer with a RELAXED mbarrier arrive (a release would fence on the warp's
//             outstanding list stores on the critical path of the tile).
//   warp 8    TMA producer: cp.async.bulk.tensor, 128B-swizzled gallery chunks into a stage ring (mbarrier
//             complete_tx).  In a pair each CTA stages only ITS half of the gallery tile ([128 x 64] chunks)
//             and all bytes are credited to the leader CTA's barriers.  For d_pad <= 512 the CTA's
//             [128 x d_pad] query tile is loaded once and stays resident in shared memory; otherwise query
//             chunks stream with the gallery.  Leaders pace themselves on the slowest sharer of their gallery
//             part (a 12- or 64-tile window) so that the part is served from L2.
//   warp 9    MMA issuer (leader CTA only in a pair): one lane issues tcgen05.mma.kind::f16 (M=128 per CTA,
//             256 per pair, N=256, K=16), fp32 accumulators in TMEM (2 stages x 256 columns);
//             tcgen05.commit (multicast to both CTAs of a pair) frees smem stages and publishes finished
//             accumulators.  Warps 10-11 idle (they donate registers).
//   Registers are rebalanced with setmaxnreg (40 for warps 8-11, 232 for the epilogue).
// Thresholds: per list, the score of its own k-th best (strict: gallery rows reach a list in
// increasing order, so on an exact tie the earlier row is already there => "score desc, row asc");
// across lists, every list publishes the score of its r-th best, r = ceil(k / n_lists): if every
// list holds >= r candidates >= g = min over lists, the union holds >= k, so nothing below g can be
// in the global top-k and all CTAs of the query tile prune with g.  select.cu merges the lists.
// The default (kProbe) instantiation: a probe pass over the first tile seeds g before anything is appended, the
// leaders pace themselves on the slowest sharer of their gallery part, and the final per-list pass filters by
// g before any exact selection (see the comments at those sites and DESIGN.md section 4); kProbe = false is
// the round-1 "long launch" form, kept selectable (mmr_index_tune) and parity-tested.
//
// Roofline: tensor-core bound, 2 * b * n * d_pad FLOP per launch.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "internal.h"

namespace mmr {
namespace {

constexpr int kBlockM = 128;          // queries per CTA tile (TMEM lanes)
constexpr int kBlockN = 256;          // gallery rows per accumulator stage (UMMA N)
constexpr int kBlockK = 64;           // bf16 elements per smem chunk row = 128 B (swizzle span)
constexpr int kUmmaK = 16;
constexpr int kMaxStages = 8;
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kBlockN;  // 512
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kBBytes = kBlockN * kBlockK * 2;   // 32 KiB
constexpr int kMaxSmemOptin = 232448;             // 227 KiB per CTA on sm_100
constexpr int kEpiGroups = 2;                     // epilogue warpgroups (one per accumulator stage)
constexpr int kEpiThreads = 128;                  // threads per epilogue warpgroup
// Warp roles.  The issue arbiter of an SM sub-partition is said to prefer the HIGHEST eligible warp id, so the
// two single-thread control warps (TMA producer, MMA issuer) sit ABOVE the eight epilogue warps they share their
// sub-partitions with: a late TMA or MMA issue is a tensor-pipe bubble, a late epilogue instruction is not.
// Measured neutral (3.549 vs 3.541 ms at 1.25M x 4096; MMR_CTRL_HIGH=0 restores the round-1 order -- control
// warps 0-3, epilogue 4-11 -- for A/B builds).
#ifndef MMR_CTRL_HIGH
#define MMR_CTRL_HIGH 1
#endif
constexpr int kFirstEpiWarp = MMR_CTRL_HIGH ? 0 : 4;
constexpr int kTmaWarp = MMR_CTRL_HIGH ? 8 : 0;
constexpr int kMmaWarp = kTmaWarp + 1;
constexpr int kNumThreads = 128 + kEpiGroups * kEpiThreads;  // 384
constexpr int kRegsLow = 40, kRegsHigh = 232;     // 128*40 + 256*232 = 64512 <= 65536
constexpr int kTrack = 8;                         // per-thread running top-8 used to publish pruning bounds
// Epilogue split.  1: BOTH epilogue warpgroups drain EVERY accumulator tile, warpgroup g taking columns
// [g * 128, g * 128 + 128) -- the tile's epilogue latency is halved.  With two TMEM stages the MMA of tile
// t + 2 can only start when tile t has been drained, i.e. the drain has ONE MMA tile time, not two, whatever
// the number of warpgroups that alternate on whole tiles (0: the round-1 scheme, kept for A/B builds:
// warpgroup g drains tiles g, g + 2, ... on its own: 3.2-3.4 us per tile against 2.8 us for TMA + MMA alone).
#ifndef MMR_EPI_SPLIT
#define MMR_EPI_SPLIT 1
#endif
constexpr bool kSplit = MMR_EPI_SPLIT != 0;
// Append path.  1: the chunk's scaled scores are NOT kept in registers; the warp ORs its lanes' masks of hit
// column groups (one REDUX) and, for every hit group, re-reads those 4 accumulator columns from TMEM (the stage is
// still ours), rescales and appends -- straight-line code in a loop over the hit groups (typically one), and the
// tracker update is predicated instead of branched.  0: the round-1 form (32 scaled scores live, eight
// vote + branch pairs per chunk) for A/B builds.  The append path runs for ~40 % of the chunks of a short launch.
#ifndef MMR_SLOW_RELOAD
#define MMR_SLOW_RELOAD 1
#endif
constexpr bool kReload = MMR_SLOW_RELOAD != 0;
constexpr int kEpiCols = kSplit ? kBlockN / kEpiGroups : kBlockN;   // accumulator columns a warpgroup drains per tile
constexpr int kEpiStep = kSplit ? 1 : kEpiGroups;                    // tile stride of a warpgroup

struct __align__(16) SmemAux {
  float ginv[kAccStages][kBlockN];  // first: read as float4
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t q_full;   // resident-Q variant: all query chunks have landed
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint32_t tmem_base;
  uint32_t pad;
};
// Shared-memory plan.  kResident (d_pad <= 512): the 128-query tile stays in shared memory for
// the whole CTA (num_kc x 16 KiB) and only gallery chunks stream through the stage ring, which
// removes a third of the L2->SM and TMA->smem traffic.  Otherwise both operands stream per chunk.
// kPair (cta_group::2): two CTAs on the SMs of one TPC work on 256 queries x the same gallery tile.
// Each CTA stages only HALF of the gallery tile (128 rows) and the pair's MMA reads both halves, so
// the L2 -> SM and shared-memory traffic per FLOP of the gallery operand halves.
struct SmemPlan {
  bool resident;
  int num_stages;
  int stage_bytes;
  int q_bytes;
  size_t total;
};
inline SmemPlan plan_smem(int d_pad, bool pair) {
  SmemPlan p;
  const int num_kc = d_pad / kBlockK;
  const int b_bytes = pair ? kBBytes / 2 : kBBytes;
  p.resident = static_cast<size_t>(num_kc) * kABytes + 2 * static_cast<size_t>(b_bytes) + sizeof(SmemAux) <=
               static_cast<size_t>(kMaxSmemOptin);
  p.q_bytes = p.resident ? num_kc * kABytes : 0;
  p.stage_bytes = p.resident ? b_bytes : kABytes + b_bytes;
  int st = static_cast<int>((kMaxSmemOptin - p.q_bytes - static_cast<int>(sizeof(SmemAux))) / p.stage_bytes);
  p.num_stages = st > kMaxStages ? kMaxStages : st;
  p.total = static_cast<size_t>(p.q_bytes) + static_cast<size_t>(p.num_stages) * p.stage_bytes + sizeof(SmemAux);
  return p;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// try_wait with a suspend-time hint: the waiting thread sleeps in hardware until the phase
// completes (or the hint expires) instead of re-issuing the poll every few cycles -- the producer
// and MMA warps share their SM sub-partitions with epilogue warps and must not steal issue slots.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA-pair (cta_group::2) helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Arrivals that hand a TMEM stage back to the MMA issuer order nothing but tcgen05 accesses (tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync precede them): relaxed, so that the arriving thread does not sit in a
// memory fence waiting for its warp's outstanding candidate-list stores (L2 write round trips) on the critical
// path of the stage release.
#ifndef MMR_RELAXED_ARRIVE
#define MMR_RELAXED_ARRIVE 1
#endif
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
#if MMR_RELAXED_ARRIVE
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  mbar_arrive_cluster(cluster_addr);
#endif
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
#if MMR_RELAXED_ARRIVE
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
#else
  mbar_arrive_n(bar, 1);
#endif
}
// TMA load into THIS CTA's shared memory whose completion bytes are credited to a barrier that may
// live in the peer CTA (`bar_cluster` is a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* smem, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major, 128-byte swizzle, rows of 64 bf16: 8-row atoms 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(0) << 16;                      // leading byte offset (ignored for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                      // layout: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// pair variants: one instruction drives the tensor cores of both SMs (M = 256: 128 rows per CTA, each
// CTA supplies its own A tile and half of the B tile); commit arrives on the barrier at the same
// shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// candidate entry in global memory: {fp32 score bits, local gallery row}
__device__ __forceinline__ uint64_t entry_key(uint2 e) { return make_key(__uint_as_float(e.x), e.y); }

// ---------------------------------------------------------------------------------------------
// Warp-cooperative exact top-k of one query's candidate buffer (cnt <= E*32 entries).
// The buffer front ends up holding the k best; returns the score of the k-th best (the new strict
// local threshold) and, through *rth_score, the score of the r-th best (r <= k), which the CTA
// publishes so that all CTAs scanning other gallery parts for the same query can prune with it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other < v ? other : v;
  }
  return v;
}

// largest-prefix threshold T with count(key >= T) == rank (keys are distinct), found MSB first
template <int E>
__device__ __forceinline__ uint64_t warp_rank_threshold(const uint64_t (&key)[E], int rank) {
  uint64_t prefix = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const uint64_t cand = prefix | (1ull << bit);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (key[e] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= rank) {
      prefix = cand;
      if (c == rank) break;
    }
  }
  return prefix;
}

template <int E>
__device__ __forceinline__ float warp_compact(uint2* buf, int cnt, int k, int r, int lane, float* rth_score) {
  uint64_t key[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    key[e] = (i < cnt) ? entry_key(buf[i]) : 0ull;
  }
  __syncwarp();
  const uint64_t tk = warp_rank_threshold<E>(key, k);
  // keep key >= tk (exactly k of them when cnt >= k), compacting in place
  uint64_t kmin = ~0ull;
  int base = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool keep = key[e] >= tk && key[e] != 0ull;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      buf[pos] = make_uint2(__float_as_uint(key_score(key[e])), key_row(key[e]));
      kmin = key[e] < kmin ? key[e] : kmin;
    }
    base += __popc(m);
  }
  kmin = warp_min_u64(kmin);
  if (rth_score != nullptr) {
    const uint64_t tr = (r >= k) ? tk : warp_rank_threshold<E>(key, r);
    uint64_t rmin = ~0ull;
#pragma unroll
    for (int e = 0; e < E; ++e) rmin = (key[e] >= tr && key[e] < rmin) ? key[e] : rmin;
    *rth_score = key_score(warp_min_u64(rmin));
  }
  __syncwarp();
  return key_score(kmin);
}

// generic version for large k: keys are re-read from the (L2-resident) buffer every pass
__device__ __noinline__ uint64_t warp_rank_threshold_generic(const uint2* buf, int cnt, int rank, int lane) {
  uint64_t prefix = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const uint64_t cand = prefix | (1ull << bit);
    int c = 0;
    for (int i = lane; i < cnt; i += 32) c += (entry_key(buf[i]) >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= rank) {
      prefix = cand;
      if (c == rank) break;
    }
  }
  return prefix;
}

__device__ __noinline__ float warp_compact_generic(uint2* buf, int cnt, int k, int r, int lane, float* rth_score) {
  const uint64_t tk = warp_rank_threshold_generic(buf, cnt, k, lane);
  if (rth_score != nullptr) {
    const uint64_t tr = (r >= k) ? tk : warp_rank_threshold_generic(buf, cnt, r, lane);
    uint64_t rmin = ~0ull;
    for (int i = lane; i < cnt; i += 32) {
      const uint64_t key = entry_key(buf[i]);
      rmin = (key >= tr && key < rmin) ? key : rmin;
    }
    *rth_score = key_score(warp_min_u64(rmin));
  }
  // stable in-place compaction: a kept element never moves to a higher index, and rounds proceed
  // in increasing index order, so reads of later rounds are never clobbered
  uint64_t kmin = ~0ull;
  int base = 0;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int i = i0 + lane;
    const uint2 e = (i < cnt) ? buf[i] : make_uint2(0u, 0u);
    const uint64_t key = (i < cnt) ? entry_key(e) : 0ull;
    const bool keep = key >= tk && key != 0ull;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) {
      buf[base + __popc(m & ((1u << lane) - 1u))] = e;
      kmin = key < kmin ? key : kmin;
    }
    base += __popc(m);
    __syncwarp();
  }
  return key_score(warp_min_u64(kmin));
}

__device__ __forceinline__ float warp_compact_dispatch(uint2* buf, int cnt, int k, int r, int cap, int lane,
                                                       float* rth_score) {
  if (cap == 512) return warp_compact<16>(buf, cnt, k, r, lane, rth_score);
  if (cap == 1024) return warp_compact<32>(buf, cnt, k, r, lane, rth_score);
  return warp_compact_generic(buf, cnt, k, r, lane, rth_score);
}

// Slow path of the epilogue for one score t = acc * inv_norm(g), branch-free: if t beats the
// (conservative) pre-threshold, append {t * inv_norm(q), col} at the cursor and advance it.
// Everything after the compare is predicated -- no BSSY/BRA per score (a divergent branch per
// score made the first version of this kernel epilogue-bound at 5x the MMA time).
// The cursor is a 32-bit entry count: the only loop-carried dependency between consecutive scores is one
// predicated 32-bit add (a predicated 64-bit pointer increment compiles to IADD3 + IMAD.X + two SELs per score,
// a ~12-cycle chain link; the append path runs for about every second chunk of a short launch).
__device__ __forceinline__ void score_step(uint2* buf, uint32_t& cnt, float t, float tau_pre, float qinv, uint32_t col) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .f32 s;\n\t"
      ".reg .u64 a;\n\t"
      "setp.gt.f32 p, %2, %3;\n\t"
      "mad.wide.u32 a, %0, 8, %1;\n\t"
      "@p mul.f32 s, %2, %4;\n\t"
      "@p st.global.v2.b32 [a], {s, %5};\n\t"
      "@p add.u32 %0, %0, 1;\n\t"
      "}"
      : "+r"(cnt)
      : "l"(buf), "f"(t), "f"(tau_pre), "f"(qinv), "r"(col)
      : "memory");
}

// Threshold on t = acc * inv_norm(g) that never rejects a score whose final value
// fl(t * qinv) exceeds tau: tau / qinv rounded down, then one more part in 2^22 below.  False
// positives (a few ulps) are harmless -- compaction orders by the exact final key.
__device__ __forceinline__ float pre_threshold(float tau, float qinv) {
  const float r = __fdiv_rd(tau, qinv);  // qinv == 0: -inf stays -inf, 0/0 = NaN => nothing passes (all ties lose)
  return r > 0.f ? __fmul_rd(r, 1.0f - 0x1p-22f) : __fmul_rd(r, 1.0f + 0x1p-22f);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// kProbe: the short-launch variant (probe pass on the first tile, pacing of the sharers, bound-filtered
// final pass).  It is a compile-time switch because the long-launch code must stay exactly as it was: the
// same source with the extra paths merely disabled at run time measured 4-7 % slower at 10M x 4096.
template <bool kResident, bool kPair, bool kProbe>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
                 const float* __restrict__ inv_norm, const float* __restrict__ q_inv, int64_t n, int b, int d_pad,
                 int k, int cap, int m_tiles, int m_group, int n_parts, int tiles_per_part, int tiles_total,
                 int num_stages, int pace_w,
                 int pub_rank, int refresh_tiles, int early_tiles, int debug_flags, uint2* __restrict__ cand,
                 int32_t* __restrict__ counts, uint32_t* __restrict__ tau_pub, unsigned long long* __restrict__ trace) {
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment; the declaration requests it and the
  // kernel traps loudly if the runtime did not honour it (no slack bytes are budgeted).
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
#ifndef MMR_DIAG
  debug_flags = 0;  // the diagnostic switches exist only in -DMMR_DIAG builds
#endif
#ifdef MMR_GEMM_TRACE  // slot 0: kernel entry of this CTA (low 48 bits) and its SM id (high 16 bits)
  if (trace != nullptr && threadIdx.x == 0) {
    unsigned long long now;
    uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace[static_cast<size_t>(blockIdx.x) * 64] = (now & 0xFFFFFFFFFFFFull) | (static_cast<unsigned long long>(smid) << 48);
  }
#endif
  const int num_kc = d_pad / kBlockK;
  constexpr int kBStage = kPair ? kBBytes / 2 : kBBytes;  // gallery bytes this CTA stages per chunk
  constexpr int kStageBytes = kResident ? kBStage : kABytes + kBStage;
  uint8_t* const smem_q = smem_raw;                                             // resident query chunks
  uint8_t* const smem = smem_raw + (kResident ? num_kc * kABytes : 0);          // stage ring
  SmemAux* aux = reinterpret_cast<SmemAux*>(smem + static_cast<size_t>(num_stages) * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA order: waves of `m_group` query tiles x all gallery parts, query tile fastest.  All parts of a
  // query tile are co-resident (they exchange pruning thresholds through tau_pub) and the CTAs that
  // stream the SAME gallery part sit on neighbouring SMs and run in step, so the part is fetched
  // from HBM once and served to the others from L2.
  // Pair mode: the scheduling unit is the 2-CTA cluster (two consecutive query tiles, same part);
  // m_group / m_tiles then count PAIRS of query tiles.  A pair whose second tile lies beyond the batch
  // still runs both CTAs (the query TMA zero-fills, no query row is valid).
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const int unit = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int m_units = kPair ? (m_tiles + 1) / 2 : m_tiles;
  const int per_group = n_parts * m_group;
  const int rem = unit % per_group;
  const int m_unit = (unit / per_group) * m_group + rem % m_group;
  const int part = rem / m_group;
  if (m_unit >= m_units) return;  // padding units of the last group (whole CTA / pair exits before any barrier)
  const int m_tile = kPair ? m_unit * 2 + static_cast<int>(cta_rank) : m_unit;
  const bool leader = cta_rank == 0u;
  const int tile_begin = part * tiles_per_part;
  const int tile_end = min(tile_begin + tiles_per_part, tiles_total);
  const int num_tiles = tile_end - tile_begin;

  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_g);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&aux->full[s], 1);
      mbar_init(&aux->empty[s], 1);
    }
    mbar_init(&aux->q_full, 1);
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&aux->tmem_full[s], 1);
      // one arrival per epilogue warpgroup (that drains the stage) of each CTA
      mbar_init(&aux->tmem_empty[s], (kPair ? 2 : 1) * (kSplit ? kEpiGroups : 1));
    }
    fence_barrier_init();
  }
  if (kPair) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  if (warp == kMmaWarp) {  // TMEM allocation (whole warp; in pair mode the same warp of both CTAs), address in smem
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&aux->tmem_base)),
                   "r"(static_cast<uint32_t>(kTmemCols))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&aux->tmem_base)),
                   "r"(static_cast<uint32_t>(kTmemCols))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = aux->tmem_base;

  if (warp >= kTmaWarp && warp < kTmaWarp + 4) {  // the control warpgroup (two of its warps only donate registers)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLow));
    if (warp == kTmaWarp) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        // pair mode: both CTAs load into their own shared memory, but every byte is credited to the
        // LEADER's barriers (the leader issues the MMAs); the leader alone posts the expected bytes
        if (kResident) {  // the query tile is loaded once and reused for every gallery tile
          if (leader) mbar_expect_tx(&aux->q_full, static_cast<uint32_t>((kPair ? 2 : 1) * num_kc * kABytes));
          const uint32_t qbar = kPair ? mapa_rank(smem_u32(&aux->q_full), 0u) : 0u;
          for (int kc = 0; kc < num_kc; ++kc) {
            if (kPair)
              tma_load_2d_pair(&tmap_q, qbar, smem_q + static_cast<size_t>(kc) * kABytes, kc * kBlockK,
                               m_tile * kBlockM);
            else
              tma_load_2d(&tmap_q, &aux->q_full, smem_q + static_cast<size_t>(kc) * kABytes, kc * kBlockK,
                          m_tile * kBlockM);
          }
        }
        // Pacing.  The units that stream the same gallery part share it through L2 only while they
        // stay within a few tens of tiles of each other (126 MB of L2 / n_parts); once they drift apart
        // every unit pulls its own copy from HBM and the ring (2 us of work) no longer hides the latency
        // (measured: DRAM reads x2-3, -6 % at 10M rows).  Every 8 tiles the leader publishes its tile
        // index and holds its loads while the slowest sharer of its part is more than pace_w tiles
        // behind.  Advisory only: a sharer that makes no progress for ~100 us (not resident yet) switches
        // pacing off for this unit; finished units publish "done".
        uint32_t* const progress = (tau_pub != nullptr && pace_w > 0 && leader)
                                       ? tau_pub + static_cast<size_t>(m_tiles) * kBlockM * (n_parts * kEpiGroups)
                                       : nullptr;
        const int n_share = min(m_group, m_units - (unit / per_group) * m_group);
        volatile uint32_t* const shared_prog = progress != nullptr ? progress + (unit - rem) + part * m_group : nullptr;
        bool pace_live = progress != nullptr && n_share > 1;
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < num_tiles; ++t) {
          if (pace_live && (t & 7) == 0) {
            *reinterpret_cast<volatile uint32_t*>(progress + unit) = static_cast<uint32_t>(t) + 1u;
            if (t >= pace_w) {
              for (int spins = 0;; ++spins) {
                uint32_t slowest = 0xFFFFFFFFu;
                for (int j = 0; j < n_share; ++j) {
                  const uint32_t v = shared_prog[j];
                  slowest = v < slowest ? v : slowest;
                }
                if (slowest + static_cast<uint32_t>(pace_w) >= static_cast<uint32_t>(t) + 1u) break;
                if (spins > 256) {
                  pace_live = false;
                  break;
                }
                __nanosleep(256);
              }
            }
          }
          // pair mode: this CTA stages gallery rows [n0, n0 + 128) of the pair's 256-row tile
          const int n0 = (tile_begin + t) * kBlockN + (kPair ? static_cast<int>(cta_rank) * (kBlockN / 2) : 0);
          for (int kc = 0; kc < num_kc; ++kc) {
            mbar_wait(&aux->empty[stage], phase ^ 1u);
            uint8_t* sa = smem + static_cast<size_t>(stage) * kStageBytes;
            uint8_t* sb = kResident ? sa : sa + kABytes;
            if (kPair) {
              if (leader) mbar_expect_tx(&aux->full[stage], 2 * kStageBytes);
              const uint32_t fbar = mapa_rank(smem_u32(&aux->full[stage]), 0u);
              if (!kResident) tma_load_2d_pair(&tmap_q, fbar, sa, kc * kBlockK, m_tile * kBlockM);
              tma_load_2d_pair(&tmap_g, fbar, sb, kc * kBlockK, n0);
              if (++stage == num_stages) {
                stage = 0;
                phase ^= 1u;
              }
              continue;
            }
            mbar_expect_tx(&aux->full[stage], kStageBytes);
            if (!kResident) tma_load_2d(&tmap_q, &aux->full[stage], sa, kc * kBlockK, m_tile * kBlockM);
            tma_load_2d(&tmap_g, &aux->full[stage], sb, kc * kBlockK, n0);
            if (++stage == num_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        if (progress != nullptr) *reinterpret_cast<volatile uint32_t*>(progress + unit) = 0xFFFFFFFFu;  // done
      }
    } else if (warp == kMmaWarp) {
      // ===================== MMA issuer =====================
      if (lane == 0 && leader) {  // pair mode: only the leader CTA issues (for both SMs)
        constexpr uint32_t idesc = make_idesc(kPair ? 2 * kBlockM : kBlockM, kBlockN);
        int stage = 0;
        uint32_t phase = 0;
        if (kResident) mbar_wait(&aux->q_full, 0u);
        for (int t = 0; t < num_tiles; ++t) {
          const int acc = t & 1;
          const uint32_t acc_phase = (t >> 1) & 1u;
          mbar_wait(&aux->tmem_empty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBlockN);
          for (int kc = 0; kc < num_kc; ++kc) {
            mbar_wait(&aux->full[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + static_cast<size_t>(stage) * kStageBytes);
            const uint32_t sa = kResident ? smem_u32(smem_q + static_cast<size_t>(kc) * kABytes) : st;
            const uint32_t sb = kResident ? st : st + kABytes;
            const uint64_t adesc = make_umma_desc(sa);
            const uint64_t bdesc = make_umma_desc(sb);
#pragma unroll
            for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle span: +2 in the (>>4) address field
              if (kPair)
                umma_f16_pair(tmem_d, adesc + static_cast<uint64_t>(kk * 2), bdesc + static_cast<uint64_t>(kk * 2),
                              idesc, (kc | kk) != 0 ? 1u : 0u);
              else
                umma_f16(tmem_d, adesc + static_cast<uint64_t>(kk * 2), bdesc + static_cast<uint64_t>(kk * 2), idesc,
                         (kc | kk) != 0 ? 1u : 0u);
            }
            // frees the smem stage (in both CTAs of a pair) when these MMAs retire
            if (kPair) umma_commit_pair(&aux->empty[stage]); else umma_commit(&aux->empty[stage]);
            if (++stage == num_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          // accumulator complete (each CTA of a pair holds its own 128 query rows of it)
          if (kPair) umma_commit_pair(&aux->tmem_full[acc]); else umma_commit(&aux->tmem_full[acc]);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsHigh));
    // ===================== epilogue: scale + threshold + top-k =====================
    const int group = (warp - kFirstEpiWarp) >> 2;  // split: the column half this warpgroup drains; else its TMEM stage
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;    // query row within the tile == TMEM lane
    const int q = m_tile * kBlockM + row_in_tile;
    const int b_pad = m_tiles * kBlockM;
    const int n_lists = n_parts * kEpiGroups;
    const int list = part * kEpiGroups + group;
    const bool q_ok = q < b;
    const float qinv = q_ok ? q_inv[q] : 0.f;
    uint2* const buf = cand + (static_cast<int64_t>(q_ok ? q : 0) * n_lists + list) * cap;
    uint32_t cnt = 0;                                // append cursor: entries in this thread's list
#ifdef MMR_GEMM_TRACE
    uint32_t dbg_chunks = 0, dbg_slow = 0, dbg_groups = 0;
#endif
    float tau_local = q_ok ? -INFINITY : INFINITY;   // pre-threshold from this list's own k-th best (strict)
    float tau_pre = tau_local;                       // effective conservative threshold on acc * inv_norm(g)
    // running top-8 of the best-of-chunk final scores this thread has appended (a subset of its list,
    // so its r-th entry is a valid, always-current lower bound on the list's r-th best): lets the list
    // publish its pruning bound without waiting for the next compaction
    // (slots [0, kTrack - r) are pinned at +inf so that the r-th best always sits in the LAST slot:
    // a compile-time register index -- a runtime index would push the array to local memory)
    float top[kTrack];
#pragma unroll
    for (int i = 0; i < kTrack; ++i) top[i] = (i < kTrack - pub_rank) ? INFINITY : -INFINITY;
    float published = -INFINITY;
    const bool track = tau_pub != nullptr && pub_rank <= kTrack && q_ok;
    const bool probe = kProbe && tau_pub != nullptr && pub_rank <= kTrack && (debug_flags & 2) == 0;  // warp-uniform
    const int epi_tid = threadIdx.x - (kFirstEpiWarp * 32 + group * kEpiThreads);  // 0..127
    const int64_t range_end = min(static_cast<int64_t>(tile_end) * kBlockN, n);
    const uint32_t full_cnt = static_cast<uint32_t>(cap - 32);  // a chunk appends at most 32 entries
    const uint32_t nan_bits = 0x7FC00000u;           // marks gallery rows outside the range: never a hit
    const int bar_id = 1 + group;                    // named barrier of this warpgroup
    float* const ginv = aux->ginv[group];
    const uint32_t ginv_s = smem_u32(ginv);
    // TMEM address of this warp's lanes; the column offset of the stage / half is added per tile
    const uint32_t taddr_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int first_tile = kSplit ? 0 : group;       // first tile this warpgroup drains
    const int col0 = kSplit ? group * kEpiCols : 0;  // first accumulator column this warpgroup drains

    // inverse norms of this group's columns of its first tile; later tiles are prefetched one (group) tile ahead
    float gnext0 = __uint_as_float(nan_bits), gnext1 = gnext0;
    if (first_tile < num_tiles) {
      const int64_t r0 = static_cast<int64_t>(tile_begin + first_tile) * kBlockN + col0 + epi_tid, r1 = r0 + kEpiThreads;
      if (r0 < range_end) gnext0 = __ldg(inv_norm + r0);
      if (!kSplit && r1 < range_end) gnext1 = __ldg(inv_norm + r1);
    }

    for (int t = first_tile; t < num_tiles; t += kEpiStep) {
      const int acc = kSplit ? (t & 1) : group;      // accumulator stage of tile t
      const uint32_t acc_phase = (t >> 1) & 1u;
      const uint32_t taddr = taddr_lane + static_cast<uint32_t>(acc * kBlockN + col0);
      const int64_t n0 = static_cast<int64_t>(tile_begin + t) * kBlockN + col0;   // first gallery row of the columns drained
      ginv[epi_tid] = gnext0;
      if (!kSplit) ginv[epi_tid + kEpiThreads] = gnext1;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // ginv of this tile is complete
      gnext0 = gnext1 = __uint_as_float(nan_bits);
      if (t + kEpiStep < num_tiles) {
        const int64_t r0 = n0 + kEpiStep * kBlockN + epi_tid, r1 = r0 + kEpiThreads;
        if (r0 < range_end) gnext0 = __ldg(inv_norm + r0);
        if (!kSplit && r1 < range_end) gnext1 = __ldg(inv_norm + r1);
      }
      // refresh the cross-CTA pruning bound (0 = "not published yet" => no pruning); purely
      // advisory and monotone, so stale reads are safe
      // (on every tile for the first `early_tiles` tiles of the group: right after start-up the bound
      // moves fastest and an early bound keeps the lists from filling with the first few hundred rows)
      // (split: warpgroup g refreshes on the tiles with t % 2 == g, the cadence of the unsplit scheme)
      if (tau_pub != nullptr && t >= kEpiGroups && q_ok && (!kSplit || (t & 1) == group) && (debug_flags & 16) == 0 &&
          ((t >> 1) < early_tiles || ((t >> 1) % refresh_tiles) == 0)) {
        uint32_t g = 0xFFFFFFFFu;
        for (int p = 0; p < n_lists; ++p) {
          const uint32_t v = __ldcg(tau_pub + static_cast<int64_t>(p) * b_pad + q);
          g = v < g ? v : g;
        }
        if (g != 0u) tau_pre = fmaxf(tau_local, pre_threshold(ordered_to_f32(g), qinv));
      }
      mbar_wait(&aux->tmem_full[acc], acc_phase);
      tc_fence_after();
#ifdef MMR_GEMM_TRACE  // diagnostic build only (scripts/trace_gemm.py): when did accumulator tile t become ready
      if (trace != nullptr && epi_tid == 0 && (t < 40 || (group == 0 && (t & 31) == 0 && (t >> 5) < 23))) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        // slots 1..40: tiles 0..39 (both warpgroups); slots 41..: every 32nd tile from 32 on
        trace[static_cast<size_t>(blockIdx.x) * 64 + (t < 40 ? 1 + t : 40 + (t >> 5))] = now;
      }
#endif
      // ---- probe pass (first tile of this warpgroup only) -------------------------------------
      // A list that starts with no threshold appends every score of its first two tiles (512 stores of
      // 8 bytes scattered over 32 lists per instruction: ~60 us per tile, L2-write-bound) and then all 32
      // lists of the warp are compacted (~175 us).  Instead the accumulator tile -- it stays in TMEM until
      // this warpgroup releases the stage -- is read twice: the first pass only takes the 8 chunk maxima
      // into the running top-8 and publishes the list's r-th best; after a short bounded wait for the
      // other lists of the query (all parts run this step at the same time) the cross-list bound g prunes
      // the REAL pass over the same tile and everything after it.  The maxima are not stored by the probe:
      // the tracker is reset, the real pass re-appends them (they are >= this list's bound >= g), so the
      // published claim "r entries >= bound in this list" holds from the end of the real pass on.
      if (kProbe && t == first_tile && probe && (debug_flags & 1) == 0) {
#pragma unroll 1
        for (int c = 0; c < kEpiCols / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + static_cast<uint32_t>(c * 32), v);
          tmem_ld_wait();
          // maxima of the two 16-column halves of the chunk (split: a half tile has only 4 chunks of 32,
          // and the published rank r = ceil(k / n_lists) needs up to kTrack = 8 distinct entries)
          float mh[2] = {-INFINITY, -INFINITY};
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float g0, g1, g2, g3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(g3)
                         : "r"(ginv_s + static_cast<uint32_t>((c * 32 + j4 * 4) * 4)));
            mh[j4 >> 2] = fmaxf(mh[j4 >> 2],
                                fmaxf(fmaxf(__uint_as_float(v[j4 * 4 + 0]) * g0, __uint_as_float(v[j4 * 4 + 1]) * g1),
                                      fmaxf(__uint_as_float(v[j4 * 4 + 2]) * g2, __uint_as_float(v[j4 * 4 + 3]) * g3)));
          }
#pragma unroll
          for (int h = 0; h < (kSplit ? 2 : 1); ++h) {
            float x = (kSplit ? mh[h] : fmaxf(mh[0], mh[1])) * qinv;
#pragma unroll
            for (int i = 0; i < kTrack; ++i) {
              const float hi = fmaxf(top[i], x);
              x = fminf(top[i], x);
              top[i] = hi;
            }
          }
        }
        if (track) {
          const float rth = top[kTrack - 1];
          if (rth > published) {
            published = rth;
            __stcg(tau_pub + static_cast<int64_t>(list) * b_pad + q, f32_to_ordered(rth));
          }
        }
#pragma unroll
        for (int i = 0; i < kTrack; ++i) top[i] = (i < kTrack - pub_rank) ? INFINITY : -INFINITY;
        // bounded wait for the other lists of this query (lists of CTAs that are not resident yet never
        // answer: after ~140 us the tile is processed without a bound, exactly as before; the host disables the
        // probe when some list of the launch gets no tile at all)
        for (int spin = 0; spin < 96; ++spin) {
          uint32_t g = 0xFFFFFFFFu;
          if (q_ok) {
            for (int p = 0; p < n_lists; ++p) {
              const uint32_t v = __ldcg(tau_pub + static_cast<int64_t>(p) * b_pad + q);
              g = v < g ? v : g;
            }
            if (g != 0u) tau_pre = fmaxf(tau_pre, pre_threshold(ordered_to_f32(g), qinv));
          }
          if (__all_sync(0xffffffffu, !q_ok || g != 0u)) break;
          __nanosleep(500);
        }
      }
#pragma unroll 1
      for (int c = 0; c < ((debug_flags & 1) ? 0 : kEpiCols / 32); ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
#ifdef MMR_DIAG
        if (debug_flags & 4) {  // TMEM loads only: no scaling, no filter (results are garbage)
          if ((v[0] ^ v[31]) == 0x7FFFFFFFu) cnt += 1;
          continue;
        }
#endif
        // fast path: t_j = acc_j * inv_norm(g_j) and the chunk maximum (NaN = out-of-range column,
        // ignored by fmaxf); almost every chunk ends here once the threshold has tightened
        float tv[32];
        float gm[8];  // maxima of the 8 groups of 4 columns
        float m = -INFINITY;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float g0, g1, g2, g3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(g3)
                       : "r"(ginv_s + static_cast<uint32_t>((c * 32 + j4 * 4) * 4)));
          tv[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) * g0;
          tv[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) * g1;
          tv[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) * g2;
          tv[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) * g3;
          gm[j4] = fmaxf(fmaxf(tv[j4 * 4 + 0], tv[j4 * 4 + 1]), fmaxf(tv[j4 * 4 + 2], tv[j4 * 4 + 3]));
          m = fmaxf(m, gm[j4]);
        }
#ifdef MMR_DIAG
        if (debug_flags & 8) {  // scaling + maxima only: never take the append path (results are garbage)
          if (m == 123.456f) cnt += 1;
          continue;
        }
#endif
#ifdef MMR_GEMM_TRACE
        ++dbg_chunks;
#endif
        if (__any_sync(0xffffffffu, m > tau_pre)) {  // warp-uniform
          const uint32_t colbase = static_cast<uint32_t>(n0) + static_cast<uint32_t>(c * 32);
#ifdef MMR_GEMM_TRACE
          ++dbg_slow;
#endif
          if (kReload) {
            uint32_t hm = 0u;  // this lane's groups of 4 columns with a hit
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) hm |= (gm[j4] > tau_pre) ? (1u << j4) : 0u;
            uint32_t um = __reduce_or_sync(0xffffffffu, hm);
            if (debug_flags & 32) um = 0u;
            while (um != 0u) {  // typically one group
#ifdef MMR_GEMM_TRACE
              ++dbg_groups;
#endif
              const uint32_t j4 = static_cast<uint32_t>(__ffs(um) - 1);
              um &= um - 1u;
              uint32_t a4[4];
              tmem_ld4(taddr + static_cast<uint32_t>(c * 32) + j4 * 4u, a4);
              float g0, g1, g2, g3;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(g3)
                           : "r"(ginv_s + (static_cast<uint32_t>(c * 32) + j4 * 4u) * 4u));
              tmem_ld_wait();
              const uint32_t col = colbase + j4 * 4u;
              score_step(buf, cnt, __uint_as_float(a4[0]) * g0, tau_pre, qinv, col + 0u);
              score_step(buf, cnt, __uint_as_float(a4[1]) * g1, tau_pre, qinv, col + 1u);
              score_step(buf, cnt, __uint_as_float(a4[2]) * g2, tau_pre, qinv, col + 2u);
              score_step(buf, cnt, __uint_as_float(a4[3]) * g3, tau_pre, qinv, col + 3u);
            }
            // m was appended above: fold it into the running top-8 (a no-op unless it beats the smallest slot;
            // predicated with -inf instead of branched) and publish when the r-th best moved
            if (track && (debug_flags & 64) == 0) {
              float x = (m > tau_pre) ? m * qinv : -INFINITY;
#pragma unroll
              for (int i = 0; i < kTrack; ++i) {
                const float hi = fmaxf(top[i], x);
                x = fminf(top[i], x);
                top[i] = hi;
              }
              const float rth = top[kTrack - 1];
              if (rth > published) {
                published = rth;
                __stcg(tau_pub + static_cast<int64_t>(list) * b_pad + q, f32_to_ordered(rth));
              }
            }
          } else {
          // second-level filter: only the groups of 4 columns in which some lane has a hit run the
          // predicated append (typically 1-3 of 8)
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            if (__any_sync(0xffffffffu, gm[j4] > tau_pre) && (debug_flags & 32) == 0) {
              score_step(buf, cnt, tv[j4 * 4 + 0], tau_pre, qinv, colbase + j4 * 4 + 0);
              score_step(buf, cnt, tv[j4 * 4 + 1], tau_pre, qinv, colbase + j4 * 4 + 1);
              score_step(buf, cnt, tv[j4 * 4 + 2], tau_pre, qinv, colbase + j4 * 4 + 2);
              score_step(buf, cnt, tv[j4 * 4 + 3], tau_pre, qinv, colbase + j4 * 4 + 3);
            }
          }
          // m was appended above: fold it into the running top-8 and publish.  The insertion only changes the
          // array when x beats its last (smallest) slot -- the common case is a hit that does not
          if (track && m > tau_pre && m * qinv > top[kTrack - 1] && (debug_flags & 64) == 0) {
            float x = m * qinv;
#pragma unroll
            for (int i = 0; i < kTrack; ++i) {
              const float hi = fmaxf(top[i], x);
              x = fminf(top[i], x);
              top[i] = hi;
            }
            const float rth = top[kTrack - 1];
            if (rth > published) {
              published = rth;
              __stcg(tau_pub + static_cast<int64_t>(list) * b_pad + q, f32_to_ordered(rth));
            }
          }
          }
          // make room for the next 32 columns: compact every list of this warp that is nearly full
#ifdef MMR_DIAG
          if (cnt > static_cast<uint32_t>(cap)) __trap();   // bounds-checked build: a list never outgrows its buffer
#endif
          uint32_t need = (debug_flags & 128) ? 0u : __ballot_sync(0xffffffffu, cnt > full_cnt);
          while (need != 0u) {
            const int src = __ffs(need) - 1;
            need &= need - 1u;
            uint2* sbuf = reinterpret_cast<uint2*>(__shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(buf), src));
            const int scnt = static_cast<int>(__shfl_sync(0xffffffffu, cnt, src));
            __syncwarp();
            float rth = 0.f;
            const float new_tau = warp_compact_dispatch(sbuf, scnt, k, pub_rank, cap, lane,
                                                        tau_pub != nullptr ? &rth : nullptr);
            if (lane == src) {
              cnt = static_cast<uint32_t>(k);
              tau_local = pre_threshold(new_tau, qinv);
              tau_pre = fmaxf(tau_pre, tau_local);
              if (tau_pub != nullptr && rth > published) {
                published = rth;
                __stcg(tau_pub + static_cast<int64_t>(list) * b_pad + q, f32_to_ordered(rth));
              }
            }
          }
        }
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // all 4 warps are done with TMEM stage + ginv
      if (epi_tid == 0) {  // the MMA issuer (the leader CTA's in pair mode) may overwrite this stage
        if (kPair)
          mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&aux->tmem_empty[acc]), 0u));
        else
          mbar_arrive_relaxed(&aux->tmem_empty[acc]);
      }
    }
#ifdef MMR_GEMM_TRACE
    if (trace != nullptr && epi_tid == 0 && group == 0) {  // slot 63: this CTA's last tile has been drained
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      trace[static_cast<size_t>(blockIdx.x) * 64 + 63] = now;
    }
#endif
    // final pass per list so that select.cu merges lists of <= k entries: first drop everything below the
    // cross-list bound (one sweep, no selection; typically leaves 10-30 entries), and only if more than k
    // remain select the exact top-k
    {
      int fcnt = static_cast<int>(cnt);
      if (kProbe && tau_pub != nullptr && q_ok && fcnt > k) {
        uint32_t gfin = 0xFFFFFFFFu;  // ordered score; 0 = no bound
        for (int p = 0; p < n_lists; ++p) {
          const uint32_t v = __ldcg(tau_pub + static_cast<int64_t>(p) * b_pad + q);
          gfin = v < gfin ? v : gfin;
        }
        if (gfin != 0u) {
          // Lane-private stable in-place filter of THIS thread's list (kept entries only move to lower
          // indices).  Eight loads are in flight before anything is stored: the warp-cooperative version
          // (one list at a time, a dependent L2 round trip per 32 entries) took 40-80 us per CTA for lists of
          // 100-150 entries -- 2 x 80 us of a 4.2 ms launch at the N = 8 shard size.
          // (16-byte loads through L1: two entries per request, a 128-byte line serves 16 -- 8-byte .cg loads
          // fetched a 32-byte L2 sector per entry and made this sweep L2-bandwidth-bound)
          int w = 0;
          const uint4* buf2 = reinterpret_cast<const uint4*>(buf);   // lists are 16-byte aligned (cap is even)
          for (int i0 = 0; i0 < fcnt; i0 += 16) {
            uint4 e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] = (i0 + 2 * j < fcnt) ? buf2[(i0 >> 1) + j] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (i0 + 2 * j < fcnt && f32_to_ordered(__uint_as_float(e[j].x)) >= gfin) buf[w++] = make_uint2(e[j].x, e[j].y);
              if (i0 + 2 * j + 1 < fcnt && f32_to_ordered(__uint_as_float(e[j].z)) >= gfin) buf[w++] = make_uint2(e[j].z, e[j].w);
            }
          }
          fcnt = w;
        }
      }
      __syncwarp();
      uint32_t need = __ballot_sync(0xffffffffu, q_ok && fcnt > k);   // rare: still more than k after the filter
      while (need != 0u) {
        const int src = __ffs(need) - 1;
        need &= need - 1u;
        uint2* sbuf = reinterpret_cast<uint2*>(__shfl_sync(0xffffffffu, reinterpret_cast<uint64_t>(buf), src));
        const int scnt = __shfl_sync(0xffffffffu, fcnt, src);
        __syncwarp();
        (void)warp_compact_dispatch(sbuf, scnt, k, k, cap, lane, nullptr);
        if (lane == src) fcnt = k;
      }
      if (q_ok) counts[static_cast<int64_t>(q) * n_lists + list] = fcnt;
    }
#ifdef MMR_GEMM_TRACE
    if (trace != nullptr && epi_tid == 0 && group == 0) {  // slot 62: final per-list pass done
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      trace[static_cast<size_t>(blockIdx.x) * 64 + 62] = now;
      // slots 58-61: chunks drained by this warp, chunks that took the append path, hit groups, entries lane 0 appended
      trace[static_cast<size_t>(blockIdx.x) * 64 + 58] = dbg_chunks;
      trace[static_cast<size_t>(blockIdx.x) * 64 + 59] = dbg_slow;
      trace[static_cast<size_t>(blockIdx.x) * 64 + 60] = dbg_groups;
      trace[static_cast<size_t>(blockIdx.x) * 64 + 61] = cnt;
    }
#endif
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();  // pair: neither CTA leaves while the other may still touch it
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(static_cast<uint32_t>(kTmemCols))
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(static_cast<uint32_t>(kTmemCols))
                   : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeFn* out) {
  static EncodeFn cached = nullptr;
  if (cached == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MMR_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return fail(MMR_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cached = reinterpret_cast<EncodeFn>(fn);
  }
  *out = cached;
  return MMR_OK;
}

// (rows, d_pad) bf16 row-major -> 2-D map, box = [box_rows x 64 elements], 128-byte swizzle, zero OOB fill
int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int d_pad, int box_rows) {
  EncodeFn enc;
  MMR_TRY(get_encode_fn(&enc));
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(d_pad) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MMR_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return MMR_OK;
}

#ifdef MMR_DIAG
// Diagnostic build only (-DMMR_DIAG, scripts/*): environment overrides for A/B experiments.  The shipping
// library reads no environment variable on the search path.
int diag_env(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v != nullptr ? std::atoi(v) : dflt;
}
#endif

}  // namespace

int plan_gemm(int64_t n, int d_pad, int b, int k, int num_sms, const GemmTune& tune, GemmPlan* plan) {
  // k + 1 candidates are kept when a row is excluded per query (dropped again by the select step)
  if (k < 1 || k > MMR_MAX_K + 1) return fail(MMR_EUNSUP, "gemm: k must be in [1, 1024]");
  if (n < 1) return fail(MMR_EINVAL, "gemm: empty gallery");
  if (d_pad % kBlockK != 0) return fail(MMR_EINVAL, "gemm: d_pad must be a multiple of 64");
  const int m_tiles = (b + kBlockM - 1) / kBlockM;
  const int64_t tiles_total = (n + kBlockN - 1) / kBlockN;
  // CTA pairs (cta_group::2) whenever there are at least two query tiles
  bool pair_ok = tune.pair != MMR_GEMM_PAIR_OFF;
#ifdef MMR_DIAG
  static const bool pair_env = diag_env("MMR_B200_GEMM_PAIR", 1) != 0;
  pair_ok = pair_ok && pair_env;
#endif
  const bool pair = pair_ok && m_tiles >= 2;
  const int ctas_per_unit = pair ? 2 : 1;
  const int m_units = pair ? (m_tiles + 1) / 2 : m_tiles;   // scheduling units along the batch
  const int slots = num_sms / ctas_per_unit;                // units resident at once
  // pick the number of gallery parts so that m_units * parts fills whole waves of SMs
  int best_parts = 1;
  double best_eff = -1.0;
  for (int w = 1; w <= 8; ++w) {
    int64_t parts = static_cast<int64_t>(w) * slots / m_units;
    if (parts < 1) parts = 1;
    if (parts > tiles_total) parts = tiles_total;
    const int64_t ctas = parts * m_units;
    const int64_t waves = (ctas + slots - 1) / slots;
    const double eff = static_cast<double>(ctas) / static_cast<double>(waves * slots);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best_parts = static_cast<int>(parts);
    }
    if (eff >= 0.97 || parts == tiles_total) break;
  }
  if (tune.parts > 0) best_parts = static_cast<int>(tune.parts < tiles_total ? tune.parts : tiles_total);
  const int tiles_per_part = static_cast<int>((tiles_total + best_parts - 1) / best_parts);
  const int n_parts = static_cast<int>((tiles_total + tiles_per_part - 1) / tiles_per_part);
  int cap = 512;
  if (k > 128) cap = 1024;
  if (k > 256) cap = 4096;
  plan->m_tiles = m_tiles;
  plan->n_parts = n_parts;
  plan->n_lists = n_parts * kEpiGroups;
  plan->pair = pair ? 1 : 0;
  {  // query tiles (pairs of them in pair mode) per wave: as many as fit next to all parts, at least 1
    int g = slots / n_parts;
    if (g < 1) g = 1;
    if (g > m_units) g = m_units;
    plan->m_group = g;
  }
  plan->tiles_per_part = tiles_per_part;
  plan->cap = cap;
  // Which instantiation.  AUTO = the kProbe kernel (probe pass on the first tile, pacing of the sharers of a
  // gallery part, bound-filtered final pass) whenever every list of the launch gets at least one full tile.
  // Round 1 kept the other instantiation ("long": lists fill and compact, the sharers stay aligned through their
  // synchronous first compactions) for launches of more than 2048 tiles per part, because the probe variant with
  // its 64-tile pacing window re-read the gallery more (36 vs 29 GB at 10M rows).  With a 12-tile window the
  // kProbe kernel reads 24.5 GB there (the long one 41 GB since its epilogue got faster and its sharers drift)
  // and is 6 % faster in the power-capped bench loop (29.7 vs 32.5 ms; isolated 26.3 vs 27.6 ms, tensor pipe
  // 96.6 % active).  mmr_index_tune(MMR_TUNE_GEMM_VARIANT) pins either instantiation (tests, A/B runs).
  {
    int probe_max = 0x7FFFFFFF;
#ifdef MMR_DIAG
    static const int probe_max_env = diag_env("MMR_B200_GEMM_PROBE_MAX_TILES", 0x7FFFFFFF);
    probe_max = probe_max_env;
#endif
    const int tiles_last = static_cast<int>(tiles_total) - (n_parts - 1) * tiles_per_part;
    // the probe needs every list of the launch to get at least one tile (else its bound is never published)
    const bool probe_ok = tiles_per_part >= kEpiGroups && tiles_last >= kEpiGroups;
    const bool want = tune.variant == MMR_GEMM_VARIANT_SHORT ||
                      (tune.variant == MMR_GEMM_VARIANT_AUTO && tiles_per_part <= probe_max);
    plan->probe = (want && probe_ok) ? 1 : 0;
  }
  plan->cand_bytes = static_cast<size_t>(b) * plan->n_lists * cap * sizeof(uint2);
  plan->count_bytes = static_cast<size_t>(b) * plan->n_lists * sizeof(int32_t);
  // published pruning bounds [n_lists][m_tiles * 128] followed by one pacing counter per scheduling unit
  {
    const int64_t groups = (m_units + plan->m_group - 1) / plan->m_group;
    plan->pub_bytes = (static_cast<size_t>(m_tiles) * kBlockM * plan->n_lists +
                       static_cast<size_t>(groups) * plan->m_group * plan->n_parts) * sizeof(uint32_t);
  }
  return MMR_OK;
}

int launch_gemm_topk(const void* emb_bf16, const float* inv_norm, int64_t n, int d_pad, const void* q_bf16,
                     const float* q_inv, int b, int k, const int64_t* /*exclude_local: applied by select*/,
                     const GemmPlan& plan, uint64_t* cand, int32_t* counts, uint32_t* tau_pub, cudaStream_t stream) {
  CUtensorMap tmap_q, tmap_g;
  MMR_TRY(make_tmap(&tmap_q, q_bf16, b, d_pad, kBlockM));
  const bool pair = plan.pair != 0;
  MMR_TRY(make_tmap(&tmap_g, emb_bf16, n, d_pad, pair ? kBlockN / 2 : kBlockN));
  const SmemPlan sp = plan_smem(d_pad, pair);
  if (sp.num_stages < 2) return fail(MMR_EUNSUP, "gemm: embedding dimension too large for the shared-memory plan");
  const int tiles_total = static_cast<int>((n + kBlockN - 1) / kBlockN);
  const int m_units = pair ? (plan.m_tiles + 1) / 2 : plan.m_tiles;
  const int grid = ((m_units + plan.m_group - 1) / plan.m_group) * plan.m_group * plan.n_parts * (pair ? 2 : 1);
  const bool probe = plan.probe != 0;
  auto kern = probe ? (pair ? (sp.resident ? gemm_topk_kernel<true, true, true> : gemm_topk_kernel<false, true, true>)
                            : (sp.resident ? gemm_topk_kernel<true, false, true> : gemm_topk_kernel<false, false, true>))
                    : (pair ? (sp.resident ? gemm_topk_kernel<true, true, false> : gemm_topk_kernel<false, true, false>)
                            : (sp.resident ? gemm_topk_kernel<true, false, false> : gemm_topk_kernel<false, false, false>));
  MMR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sp.total)));
  // threshold exchange: rank published per part and refresh period (in gallery tiles)
  const int pub_rank = (k + plan.n_lists - 1) / plan.n_lists;
  // every 4th tile of a warpgroup is enough (the bound only tightens; a stale value just prunes less)
  const int refresh = plan.n_lists <= 64 ? 2 : 2 * ((plan.n_lists + 63) / 64);
  // Early per-tile refresh: measured +28 % at 1M x 1024 (8 CTAs share a gallery part) but -15 % at
  // 1M x 2048/4096 (16 sharers): without the synchronous first compaction the sharers drift apart
  // and the part is re-read from HBM ~5x (ncu: dram read 1.18 -> 5.36 GB).
  int early_tiles = plan.m_group * (pair ? 2 : 1) <= 8 ? 8 : 0;
  // Pacing window (tiles a leader may run ahead of the slowest sharer of its gallery part; 0 = off).  The sharers
  // of a part share it through L2 only while they stay within the part's share of the 126 MB: measured (ncu DRAM
  // bytes + bench-loop time, scripts/exp_pace.sh) 12 tiles is best from ~1000 tiles per part up (2.5M rows: 5.7 vs
  // 14.0 GB read; 5M: 14.5 vs 15.3 ms; 10M: 29.7 vs 31.8 ms), while launches of a few hundred tiles per part, which
  // run at full clocks with no power cap, lose 4 % to a tight window (1.25M rows: 3.51 ms at 64, 3.66 at 16).
  // The long-launch instantiation relies on its lockstep start-up and is not paced.
  int pace_w = plan.probe ? (plan.tiles_per_part <= 768 ? 64 : 12) : 0;
  int flags = 0;    // diagnostic builds: bit 0 skips the score processing, bit 1 the probe pass
#ifdef MMR_DIAG
  // MMR_B200_GEMM_DEBUG bit 0: skip the epilogue's score processing (results are garbage) to measure the
  // TMA + MMA pipeline alone; MMR_B200_EARLY_TILES / MMR_B200_GEMM_PACE override the heuristics above
  static const int debug_env = diag_env("MMR_B200_GEMM_DEBUG", 0);
  static const int early_env = diag_env("MMR_B200_EARLY_TILES", -1);
  static const int pace_env = diag_env("MMR_B200_GEMM_PACE", -1);
  flags = debug_env;
  if (early_env >= 0) early_tiles = early_env;
  if (pace_env >= 0) pace_w = pace_env;
#endif
  if (tau_pub != nullptr) MMR_CUDA_TRY(cudaMemsetAsync(tau_pub, 0, plan.pub_bytes, stream));
  // a list whose warpgroup gets no tile (single-tile parts) must still report an empty list
  MMR_CUDA_TRY(cudaMemsetAsync(counts, 0, plan.count_bytes, stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = sp.total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;  // the CTA pair is a 2-CTA cluster (one TPC)
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  uint2* cand2 = reinterpret_cast<uint2*>(cand);
  unsigned long long* trace = nullptr;
#ifdef MMR_GEMM_TRACE
  // MMR_B200_GEMM_TRACE=<file>: per-CTA tile timestamps, dumped after the launch (synchronises!)
  static const char* trace_path = std::getenv("MMR_B200_GEMM_TRACE");
  if (trace_path != nullptr) {
    MMR_CUDA_TRY(cudaMalloc(&trace, static_cast<size_t>(grid) * 64 * sizeof(unsigned long long)));
    MMR_CUDA_TRY(cudaMemsetAsync(trace, 0, static_cast<size_t>(grid) * 64 * sizeof(unsigned long long), stream));
  }
#endif
  MMR_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmap_q, tmap_g, inv_norm, q_inv, n, b, d_pad, k, plan.cap, plan.m_tiles,
                                  plan.m_group, plan.n_parts, plan.tiles_per_part, tiles_total, sp.num_stages, pace_w, pub_rank,
                                  refresh, early_tiles, flags, cand2, counts, tau_pub, trace));
  MMR_LAUNCHED();
#ifdef MMR_GEMM_TRACE
  if (trace != nullptr) {
    MMR_CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<unsigned long long> h(static_cast<size_t>(grid) * 64);
    MMR_CUDA_TRY(cudaMemcpy(h.data(), trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    if (FILE* f = std::fopen(trace_path, "wb")) {
      std::fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
      std::fclose(f);
    }
  }
#endif
  return MMR_OK;
}

}  // namespace mmr
