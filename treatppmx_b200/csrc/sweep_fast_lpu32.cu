// Instantiations of sweep_fast_kernel for 32 lanes per (chain, arm) (own translation unit: parallel build).
#include "sweep_fast.cuh"

namespace tppmx {
cudaError_t fast_launch_lpu32(const DevBatch &d, const FastArgs &fa, size_t smem, uint64_t seed, int iter0, int n_sweeps,
                              cudaStream_t stream) {
  return fast_launch_impl<32>(d, fa, smem, seed, iter0, n_sweeps, stream);
}
}  // namespace tppmx
