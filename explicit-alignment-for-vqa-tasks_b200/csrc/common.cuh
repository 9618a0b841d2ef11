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

// HF NewGELUActivation (transformers/activations.py:59-66)
__device__ __forceinline__ float gelu_new(float x) {
    const float c = 0.7978845608028654f;
    float t = fast_tanh(c * (x + 0.044715f * x * x * x));
    return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float gelu_new_grad(float x) {
    const float c = 0.7978845608028654f;
    float x2 = x * x;
    float t = fast_tanh(c * (x + 0.044715f * x * x2));
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x2);
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
