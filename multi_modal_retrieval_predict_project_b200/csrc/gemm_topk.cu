// placeholder until the tcgen05 kernel lands (next commit)
#include "internal.h"
namespace mmr {
int plan_gemm(int64_t, int, int, int, int, GemmPlan*) { return fail(MMR_EUNSUP, "gemm path not built"); }
int launch_gemm_topk(const void*, const float*, int64_t, int, const void*, const float*, int, int, const int64_t*,
                     const GemmPlan&, uint64_t*, int32_t*, cudaStream_t) { return fail(MMR_EUNSUP, "gemm path not built"); }
}  // namespace mmr
