// Shared helpers for the sm_100a kernels of the CLIP-prefix LM step.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

namespace eavqa {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing: kernels never throw across the C ABI; api.cu converts to status + last_error
// ---------------------------------------------------------------------------------------------
struct Error : public std::runtime_error {
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define EAVQA_CHECK(cond, msg)                                                                         \
    do {                                                                                               \
        if (!(cond)) throw ::eavqa::Error(std::string(msg) + " [" #cond "] at " __FILE__ ":" + std::to_string(__LINE__)); \
    } while (0)

#define CUDA_CHECK(expr)                                                                               \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            throw ::eavqa::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " in " #expr " at " __FILE__ ":" + std::to_string(__LINE__)); \
    } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

int num_sms();   // cached device SM count (148 on B200)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the library is launched with the stream-serialisation attribute
// and begins with pdl_trigger() (lets the NEXT kernel's grid be scheduled while this one runs) and pdl_wait()
// (blocks until the PREVIOUS kernel has completed and its writes are visible).  Launch latency, block scheduling
// and -- in the GEMM -- barrier/TMEM/descriptor set-up overlap the predecessor's tail.  (Round-1 measurement: a
// trivial GEMM costs 8 us per back-to-back launch; the step has ~440 launches.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// ---------------------------------------------------------------------------------------------
// device math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// hardware tanh (MUFU.TANH, max rel err 2^-11): every use feeds a bf16 value (2^-9) or a product rounded to bf16,
// and it keeps the GEMM epilogues under the MMA time of a tile (ex2 + rcp + fixups tripled their ALU cost)
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// HF NewGELUActivation (transformers/activations.py:59-66): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))).
// Written in FMA form (3 FMUL + 2 FFMA + 1 MUFU): the GEMM epilogue that applies it is instruction-issue bound
// (round-1 ncu: 11 FP instructions per element in the naive form, 910 warp-instructions per 32-column chunk).
__device__ __forceinline__ float gelu_new(float x) {
    const float c = 0.7978845608028654f, c3 = 0.7978845608028654f * 0.044715f;
    const float x2 = x * x;
    const float u = x * __fmaf_rn(x2, c3, c);
    const float hx = 0.5f * x;
    return __fmaf_rn(hx, fast_tanh(u), hx);
}
// d/dx gelu_new = 0.5 (1 + t) + 0.5 x (1 - t^2) c (1 + 3 * 0.044715 x^2),  t = tanh(u)
__device__ __forceinline__ float gelu_new_grad(float x) {
    const float c = 0.7978845608028654f, c3 = 0.7978845608028654f * 0.044715f;
    const float x2 = x * x;
    const float t = fast_tanh(x * __fmaf_rn(x2, c3, c));
    const float du = __fmaf_rn(x2, 3.0f * c3, c);              // du/dx
    const float sech2 = __fmaf_rn(-t, t, 1.0f);
    const float hx = 0.5f * x;
    return __fmaf_rn(hx * sech2, du, __fmaf_rn(0.5f, t, 0.5f));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

}  // namespace eavqa
