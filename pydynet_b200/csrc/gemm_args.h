// gemm_args.h — argument block shared by the pdn_gemm dispatcher (gemm.cu) and the tcgen05 path (gemm_tc.cu).
#pragma once
#include <stdint.h>
namespace pdn {
struct GemmArgs {
  const void* A; const void* B; void* C; const void* bias;
  int64_t M, N, K;
  int64_t a_rs, a_cs, b_rs, b_cs, ldc;
  int64_t nb[3], a_bs[3], b_bs[3], c_bs[3];
  int accumulate;
};
bool gemm_tc_eligible(const GemmArgs& g);
// a_version / b_version: the caller's write counters of the operand buffers (>= 0: operand planes may be cached), -1 = transient
int  gemm_tc_launch(const GemmArgs& g, long long a_version = -1, long long b_version = -1);
void plane_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* bytes, uint64_t* entries);
}  // namespace pdn
