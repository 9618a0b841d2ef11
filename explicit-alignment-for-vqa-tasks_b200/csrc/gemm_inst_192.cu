// tcgen05 GEMM kernels with 192-wide tiles (all epilogue modes, 1-CTA and CTA-pair); see gemm_kernel.cuh
#include "gemm_kernel.cuh"

namespace eavqa {
void gemm_dispatch_bn192(int mode, int kind, const GemmArgs& a, cudaStream_t s) { gk::dispatch_bn<192, true>(mode, kind, a, s); }
}  // namespace eavqa
