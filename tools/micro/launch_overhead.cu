// Where do the ~8 us of a trivial GEMM launch go?  Back-to-back launch cost of progressively richer kernels.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k_empty() {}
__global__ void k_smem(int* p) { extern __shared__ uint8_t s[]; if (threadIdx.x == 999) p[0] = s[0]; }
__global__ void __launch_bounds__(320, 1) k_tmem(int* p) {
    extern __shared__ uint8_t s[];
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
    if (threadIdx.x == 999) p[0] = s[0];
}
template <typename F> float timeit(F f, int n = 200) {
    for (int i = 0; i < 20; ++i) f();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < n; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / n;
}
int main() {
    int* p; cudaMalloc(&p, 4);
    const int big = 225 * 1024;
    cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    for (int grid : {1, 148}) {
        printf("grid %d\n", grid);
        printf("  empty 320 thr, 0 smem        : %.2f us/launch\n", timeit([&] { k_empty<<<grid, 320>>>(); }));
        printf("  320 thr, 225 KB dyn smem     : %.2f us/launch\n", timeit([&] { k_smem<<<grid, 320, big>>>(p); }));
        printf("  + tcgen05 alloc/dealloc 512  : %.2f us/launch\n", timeit([&] { k_tmem<<<grid, 320, big>>>(p); }));
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = big;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        printf("  same via LaunchKernelEx + cluster(1) attr: %.2f us/launch\n", timeit([&] { cudaLaunchKernelEx(&cfg, k_tmem, p); }));
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        printf("  same via LaunchKernelEx + PDL attr       : %.2f us/launch\n", timeit([&] { cudaLaunchKernelEx(&cfg, k_tmem, p); }));
        // alternate small-smem and big-smem kernels (carve-out reconfiguration?)
        printf("  alternating empty / 225 KB kernel pair   : %.2f us/pair\n", timeit([&] { k_empty<<<grid, 320>>>(); k_smem<<<grid, 320, big>>>(p); }));
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
