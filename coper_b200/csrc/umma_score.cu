// tcgen05/TMEM engine for the 1-N scorer (placeholder until the UMMA kernels land).
#include "common.cuh"
namespace coper {
int umma_score1n_fwd(const float*, const float*, const float*, int, int64_t, int, float*, int64_t, void*, size_t, int,
                     cudaStream_t) {
  return COPER_ERR_UNSUPPORTED;
}
size_t umma_score1n_workspace_bytes(int, int64_t, int, int) { return 256; }
}  // namespace coper
