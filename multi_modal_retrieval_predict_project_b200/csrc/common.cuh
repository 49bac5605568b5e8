// Shared device/host helpers for the retrieval kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/mmr_b200.h"

namespace mmr {

// ---------------------------------------------------------------------------------------------
// error plumbing (no exception crosses the C ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
void count_launch();  // every kernel launch site calls this (mmr_launch_count)

#define MMR_CUDA_TRY(expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      return ::mmr::fail(_e == cudaErrorMemoryAllocation ? MMR_ENOMEM : MMR_ECUDA,              \
                         std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                             ":" + std::to_string(__LINE__) + ")");                             \
    }                                                                                           \
  } while (0)

// after a <<<...>>> launch: count it and surface launch errors
#define MMR_LAUNCHED()          \
  do {                          \
    ::mmr::count_launch();      \
    MMR_CUDA_TRY(cudaGetLastError()); \
  } while (0)

#define MMR_TRY(expr)             \
  do {                            \
    int _s = (expr);              \
    if (_s != MMR_OK) return _s;  \
  } while (0)

#define MMR_REQUIRE(cond, msg)                                 \
  do {                                                         \
    if (!(cond)) return ::mmr::fail(MMR_EINVAL, (msg));        \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Ordering keys.  Every selection in the library orders candidates by
//     (final fp32 score descending, row id ascending)
// encoded as ONE unsigned 64-bit key (larger key = better):
//     key = (monotone_u32(score) << 32) | (0xFFFFFFFF - local_row)
// so top-K sets are unique and independent of the order kernels visit rows in.
// key 0 is the padding sentinel (smaller than any real key).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t local_row) {
  return (static_cast<uint64_t>(f32_to_ordered(score)) << 32) |
         static_cast<uint64_t>(0xFFFFFFFFu - local_row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
  return ordered_to_f32(static_cast<uint32_t>(key >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) {
  return 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
}

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__host__ __device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Block-wide bitonic sort of `n` (power of two) keys in shared memory, DESCENDING.
// All threads of the block must call it.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t a = s[i], b = s[ixj];
          bool desc_block = ((i & k) == 0);
          if ((a < b) == desc_block) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// Warp-wide bitonic sort of `n` (power of two, >= 2) keys in shared memory, DESCENDING.
// Called by all 32 lanes of one warp on a warp-private region.
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t* s, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += kWarp) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t a = s[i], b = s[ixj];
          bool desc_block = ((i & k) == 0);
          if ((a < b) == desc_block) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace mmr
