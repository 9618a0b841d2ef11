// tcgen05 GEMM kernels with 64-wide tiles (all epilogue modes); see gemm_kernel.cuh
#include "gemm_kernel.cuh"

namespace eavqa {
void gemm_dispatch_bn64(int mode, int kind, const GemmArgs& a, cudaStream_t s) { gk::dispatch_bn<64, false>(mode, kind, a, s); }
}  // namespace eavqa
