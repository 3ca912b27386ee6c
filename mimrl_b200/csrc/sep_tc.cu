// tcgen05 path of the separable-critic sweeps (placeholder until the kernels land).
#include "common.cuh"

namespace mimrl {
bool sep_tc_supported(int, int, int) { return false; }
size_t sep_tc_workspace_bytes(int, int, int) { return 0; }
int sep_row_stats_tc(const float *, const float *, int, int, int, int, int, float *, float *, float *, void *, size_t,
                     cudaStream_t) {
  set_error("tcgen05 path not built");
  return 9;
}
int sep_weighted_sum_tc(const float *, const float *, int, int, int, int, int, int, const float *, int, const float *,
                        const float *, float *, void *, size_t, cudaStream_t) {
  set_error("tcgen05 path not built");
  return 9;
}
}  // namespace mimrl
