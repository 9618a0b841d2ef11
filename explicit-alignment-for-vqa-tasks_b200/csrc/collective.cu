// Gradient exchange fused with the optimiser over NVLink 5 / NVSwitch peer memory (the ONE exchange step of the
// data-parallel mapper training step).
//
// Reference: Lightning DDP averages the mapper gradients with an NCCL all-reduce, then torch.optim.AdamW updates every
// parameter on every rank (main.py:133-138, clipcap_exector.py:79-81).  Here ONE kernel per step does reduce-scatter + AdamW +
// all-gather: rank r owns the contiguous shard [r * S, (r + 1) * S) of the flat parameter buffer, reads the SUM of all ranks'
// gradients for that shard straight out of the switch (`multimem.ld_reduce` on the NVLS multicast address of the gradient
// buffers; without multicast: one peer load per rank, summed in rank order), applies AdamW with its local shard of the
// moments, and broadcasts the updated parameters to every rank's buffer (`multimem.st`; without multicast: one peer store per
// rank).  Link traffic per GPU is what an all-reduce moves (the gradient once out, the parameters once in), but the
// optimiser's 28 B / parameter of HBM traffic shrinks to 1 / world of it and the all-reduce's second HBM pass disappears.
//
// Cross-GPU ordering: two flag barriers inside the kernel (start: every rank's gradients are final; end: every rank's stores
// have landed) over a small symmetric flag buffer -- release stores to the peers' flags, acquire spins on the own ones.
// Buffers and flags are symmetric-memory allocations made by the host side (torch.distributed._symmetric_memory); this file
// only sees raw pointers.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace eavqa {
namespace {

constexpr int kMaxRanks = 16;
struct PeerPtrs {
    void* p[kMaxRanks];
};
// layout of a rank's flag buffer (uint32): [0, 16) start-barrier tokens by source rank, [16, 32) end-barrier tokens,
// [32] the local arrival counter of the kernel's CTAs, [33] set to 1 when a spin timed out
constexpr int kFlagStart = 0, kFlagEnd = 16, kFlagCounter = 32, kFlagTimeout = 33;

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= token (tokens only grow); gives up after 10 s and records it, so a lost peer cannot hang the GPU
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t token, uint32_t* timeout_flag) {
    const unsigned long long t0 = global_ns();
    while (static_cast<int32_t>(ld_acquire_sys(flag) - token) < 0) {
        __nanosleep(64);
        if (global_ns() - t0 > 10000000000ull) {
            *timeout_flag = 1u;
            break;
        }
    }
}

__device__ __forceinline__ void adamw_update(float4& pp, const float4& gg, float4& mm, float4& vv, float lr, float beta1, float beta2,
                                             float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {       // the same arithmetic, in the same order, as adamw_kernel (elementwise.cu)
        const float gr = ga[k] * gscale;
        pa[k] *= 1.0f - lr * wd;
        ma[k] = beta1 * ma[k] + (1.0f - beta1) * gr;
        va[k] = beta2 * va[k] + (1.0f - beta2) * gr * gr;
        const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
        pa[k] -= (lr / bc1) * (ma[k] / denom);
    }
}

// MC: gradients / parameters through the NVLS multicast mapping; otherwise through per-rank peer pointers.
template <bool MC, int U>
__global__ void __launch_bounds__(512, 1) sharded_adamw_kernel(const __grid_constant__ PeerPtrs grads, const __grid_constant__ PeerPtrs params,
                                                            const float4* __restrict__ mc_grads, float4* __restrict__ mc_params,
                                                            float4* __restrict__ m, float4* __restrict__ v,
                                                            const __grid_constant__ PeerPtrs flags, uint32_t token,
                                                            int64_t begin4, int64_t end4, int rank, int world,
                                                            float lr, float beta1, float beta2, float eps, float wd, float bc1,
                                                            float bc2_sqrt, float gscale) {
    pdl_trigger();
    pdl_wait();                          // this rank's backward is complete: its gradient buffer is final
    uint32_t* my_flags = static_cast<uint32_t*>(flags.p[rank]);
    if (my_flags != nullptr) {
        // ---- start barrier: every rank's gradients are final before anyone reads them
        if (blockIdx.x == 0 && threadIdx.x < world)
            st_release_sys(static_cast<uint32_t*>(flags.p[threadIdx.x]) + kFlagStart + rank, token);
        if (threadIdx.x < world) wait_flag(my_flags + kFlagStart + threadIdx.x, token, my_flags + kFlagTimeout);
        __syncthreads();
    }
    float4* my_params = static_cast<float4*>(params.p[rank]);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    // U gradient loads in flight per thread: NVLink round trips are several microseconds
    for (int64_t i0 = begin4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < end4; i0 += U * stride) {
        float4 gg[U], pp[U], mm[U], vv[U];
        // every load of the iteration is issued before the first result is needed: the local parameter / moment loads travel
        // with the gradient loads instead of starting when those arrive (round-2 SASS: LDGMC x4, wait, then LDG per element)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < end4) {
                pp[u] = my_params[i];
                mm[u] = m[i];
                vv[u] = v[i];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            gg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < end4) {
                if (MC) {
                    gg[u] = multimem_ld_reduce_add(mc_grads + i);
                } else {
                    gg[u] = static_cast<const float4*>(grads.p[0])[i];
                    for (int r = 1; r < world; ++r) {      // summed in rank order
                        const float4 o = static_cast<const float4*>(grads.p[r])[i];
                        gg[u].x += o.x; gg[u].y += o.y; gg[u].z += o.z; gg[u].w += o.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < end4) {
                adamw_update(pp[u], gg[u], mm[u], vv[u], lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, gscale);
                m[i] = mm[u];
                v[i] = vv[u];
                if (MC) {
                    multimem_st(mc_params + i, pp[u]);
                } else {
                    for (int r = 0; r < world; ++r) static_cast<float4*>(params.p[r])[i] = pp[u];
                }
            }
        }
    }
    if (my_flags != nullptr) {
        // ---- end barrier: the kernel does not complete before every rank's parameter stores have landed here and every rank
        //      has finished reading this rank's gradients (the next step overwrites them)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned arrived = atomicAdd(my_flags + kFlagCounter, 1u);
            if (arrived == gridDim.x - 1) {                     // last CTA of this rank
                my_flags[kFlagCounter] = 0u;
                __threadfence_system();
                for (int r = 0; r < world; ++r) st_release_sys(static_cast<uint32_t*>(flags.p[r]) + kFlagEnd + rank, token);
                for (int r = 0; r < world; ++r) wait_flag(my_flags + kFlagEnd + r, token, my_flags + kFlagTimeout);
            }
        }
    }
}

}  // namespace

void sharded_adamw_range(int64_t n, int rank, int world, int64_t* begin, int64_t* end) {
    const int64_t n4 = n / 4;
    const int64_t per = ceil_div64(n4, world);
    *begin = std::min<int64_t>(n4, per * rank) * 4;
    *end = std::min<int64_t>(n4, per * (rank + 1)) * 4;
}

void sharded_adamw_step(void* const* grad_ptrs, void* const* param_ptrs, const void* mc_grads, void* mc_params, void* const* flag_ptrs,
                        uint32_t token, int rank, int world, float* m, float* v, int64_t offset, int64_t n, int max_ctas, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t s) {
    EAVQA_CHECK(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "sharded_adamw_step: rank / world");
    EAVQA_CHECK(n % 4 == 0 && offset % 4 == 0 && offset >= 0 && step >= 1 && max_ctas >= 0,
                "sharded_adamw_step: offset and n must be multiples of 4 and step >= 1");
    EAVQA_CHECK(grad_ptrs != nullptr && param_ptrs != nullptr && m != nullptr && v != nullptr, "sharded_adamw_step: null argument");
    EAVQA_CHECK((mc_grads == nullptr) == (mc_params == nullptr), "sharded_adamw_step: both multicast addresses or neither");
    PeerPtrs g = {}, p = {}, f = {};
    for (int r = 0; r < world; ++r) {
        EAVQA_CHECK(grad_ptrs[r] != nullptr && param_ptrs[r] != nullptr, "sharded_adamw_step: null peer pointer");
        g.p[r] = grad_ptrs[r];
        p.p[r] = param_ptrs[r];
        if (flag_ptrs != nullptr) {
            EAVQA_CHECK(flag_ptrs[r] != nullptr, "sharded_adamw_step: null flag pointer");
            f.p[r] = flag_ptrs[r];
        }
    }
    // the range [offset, offset + n) of the flat buffers is what this call exchanges (the whole buffer, or one gradient bucket
    // while the rest of the backward still runs); rank r owns its r-th shard
    int64_t begin = 0, end = 0;
    sharded_adamw_range(n, rank, world, &begin, &end);
    begin += offset;
    end += offset;
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
    // every CTA must be resident while it spins in the start barrier: one CTA per SM at most.
    // EAVQA_SHARD_CFG="threads,U" (tuning knob): threads per CTA (<= 512), gradient loads in flight per thread (2 / 4 / 8)
    int threads = 512, unroll = 4;
    if (const char* cfg = getenv("EAVQA_SHARD_CFG")) sscanf(cfg, "%d,%d", &threads, &unroll);
    EAVQA_CHECK(threads >= 64 && threads <= 512 && threads % 32 == 0 && (unroll == 2 || unroll == 4 || unroll == 8), "EAVQA_SHARD_CFG");
    const int64_t work4 = std::max<int64_t>((end - begin) / 4, 1);
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(work4, threads), max_ctas > 0 ? std::min(max_ctas, num_sms()) : num_sms()));
    const bool mc = mc_grads != nullptr;
#define EAVQA_SHARD_LAUNCH(MC_, U_)                                                                                            \
    launch_kernel(sharded_adamw_kernel<MC_, U_>, dim3(grid), dim3(threads), 0, s, g, p, static_cast<const float4*>(mc_grads),   \
                  static_cast<float4*>(mc_params), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), f, token,       \
                  begin / 4, end / 4, rank, world, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale)
    if (mc) {
        if (unroll == 2) EAVQA_SHARD_LAUNCH(true, 2); else if (unroll == 4) EAVQA_SHARD_LAUNCH(true, 4); else EAVQA_SHARD_LAUNCH(true, 8);
    } else {
        if (unroll == 2) EAVQA_SHARD_LAUNCH(false, 2); else if (unroll == 4) EAVQA_SHARD_LAUNCH(false, 4); else EAVQA_SHARD_LAUNCH(false, 8);
    }
#undef EAVQA_SHARD_LAUNCH
    KERNEL_CHECK();
    count_launch();
}

}  // namespace eavqa
