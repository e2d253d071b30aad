// Instantiations of sweep_big_kernel (own translation unit: parallel build).
#include "sweep_big.cuh"

namespace tppmx {
// nth = threads per (chain, arm): 32 (covariates P <= 128 in up to four rows per thread), 128 or 256 (P <= nth)
cudaError_t big_launch(const DevBatch &d, const FastArgs &fa, int nth, size_t smem, uint64_t seed, int iter0, int n_sweeps,
                       cudaStream_t stream) {
  if (nth == 32) {
    if (d.P <= 64) return big_launch_impl<32, 2>(d, fa, smem, seed, iter0, n_sweeps, stream);
    return big_launch_impl<32, 4>(d, fa, smem, seed, iter0, n_sweeps, stream);
  }
  if (nth == 128) return big_launch_impl<128, 1>(d, fa, smem, seed, iter0, n_sweeps, stream);
  if (nth == 256) return big_launch_impl<256, 1>(d, fa, smem, seed, iter0, n_sweeps, stream);
  return cudaErrorInvalidConfiguration;
}
}  // namespace tppmx
