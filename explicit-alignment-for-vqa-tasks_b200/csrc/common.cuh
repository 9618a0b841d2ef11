// Shared helpers for the sm_100a kernels of the CLIP-prefix LM step.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// ---------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2): one instruction per TWO floats.  The GEMM epilogues are bound by
// the instruction stream of their 8 warps (round-1 ncu: ~620 warp-instructions per 32-column chunk, 6.4 clk each), so
// the element-wise math runs on register pairs.  pk / upk are register renames, not instructions.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 pk1(float x) { return pk(x, x); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 tanh2(f32x2 x) {
    float a, b;
    upk(x, a, b);
    return pk(fast_tanh(a), fast_tanh(b));
}
// gelu_new on a pair: 3 FMUL2 + 2 FFMA2 + 2 MUFU
__device__ __forceinline__ f32x2 gelu_new2(f32x2 x) {
    const f32x2 c = pk1(0.7978845608028654f), c3 = pk1(0.7978845608028654f * 0.044715f), half = pk1(0.5f);
    const f32x2 x2 = mul2(x, x);
    const f32x2 t = tanh2(mul2(x, fma2(x2, c3, c)));
    const f32x2 hx = mul2(x, half);
    return fma2(hx, t, hx);
}
// d/dx gelu_new on a pair: 4 FMUL2 + 5 FFMA2 + 2 MUFU (signs folded into constants: q = t^2 - 1, ndu = -du/dx)
__device__ __forceinline__ f32x2 gelu_new_grad2(f32x2 x) {
    const float cc = 0.7978845608028654f, cc3 = 0.7978845608028654f * 0.044715f;
    const f32x2 c = pk1(cc), c3 = pk1(cc3), n3c3 = pk1(-3.0f * cc3), nc = pk1(-cc), half = pk1(0.5f), m1 = pk1(-1.0f);
    const f32x2 x2 = mul2(x, x);
    const f32x2 t = tanh2(mul2(x, fma2(x2, c3, c)));
    const f32x2 ndu = fma2(x2, n3c3, nc);
    const f32x2 q = fma2(t, t, m1);
    const f32x2 hx = mul2(x, half);
    return fma2(mul2(hx, q), ndu, fma2(half, t, half));
}
// two bf16 packed in a 32-bit word -> a float pair (bf16 -> fp32 is a 16-bit shift)
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t u) { return pk(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_bf16x2(f32x2 v) {
    float lo, hi;
    upk(v, lo, hi);
    return pack_bf16x2(lo, hi);
}
// fp16 pair (10 mantissa bits): the LM head's STORED logits.  They are read once more, by d logits = softmax - onehot, and a
// trained GPT-2's logits sit at |z| ~ 30-150 where a bf16 ulp is 0.25-1.0 -- a 25 % error on every probability (round-2 test
// at that scale: mapper-gradient cosine 0.9989); fp16 is 8x finer at the same 2 bytes and its range (65504) is ample.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(f32x2 v) {
    float lo, hi;
    upk(v, lo, hi);
    return pack_f16x2(lo, hi);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

}  // namespace eavqa
