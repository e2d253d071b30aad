// Instantiations of sweep_big_kernel (own translation unit: parallel build).
#include "sweep_big.cuh"

namespace tppmx {
// nth = threads per (chain, arm): 128 or 256 (one covariate row per thread: P <= nth)
cudaError_t big_launch(const DevBatch &d, const FastArgs &fa, int nth, size_t smem, uint64_t seed, int iter0, int n_sweeps,
                       cudaStream_t stream) {
  if (nth == 128) return big_launch_impl<128, 1>(d, fa, smem, seed, iter0, n_sweeps, stream);
  if (nth == 256) return big_launch_impl<256, 1>(d, fa, smem, seed, iter0, n_sweeps, stream);
  return cudaErrorInvalidConfiguration;
}
}  // namespace tppmx
