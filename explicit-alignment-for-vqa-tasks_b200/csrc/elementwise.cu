// Fused, vectorised, coalesced memory-bound kernels of the CLIP-prefix LM step:
// weight/activation packing (fp32->bf16, transpose, column sums), LayerNorm forward/backward,
// embedding gather + prefix concat/splice + position add, LM-head cross-entropy bookkeeping,
// greedy-decode argmax/EOS bookkeeping.  All are HBM/L2-bound: 16-byte accesses, warp-shuffle
// reductions, no shared-memory round trips except the transpose tile and cross-warp reductions.
#include <algorithm>
#include <vector>
#include <atomic>

#include "kernels.cuh"

namespace eavqa {

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n); }
int64_t kernel_launch_count() { return g_launches.load(); }

namespace {

// ------------------------------------------------------------------------------------------ packing
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256) convert_transpose_kernel(const T* __restrict__ src, int ld_src, int R, int C,
                                                                bf16* __restrict__ dst, int ld_dst,
                                                                bf16* __restrict__ dst_t, int ld_t,
                                                                float* __restrict__ colsum) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[32][33];
    __shared__ float cs[8][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        float v = 0.f;
        if (r < R && c < C) {
            v = to_f32<T>(src[static_cast<size_t>(r) * ld_src + c]);
            if (dst != nullptr) dst[static_cast<size_t>(r) * ld_dst + c] = __float2bfloat16(v);
        }
        tile[ty + 8 * i][tx] = v;
        part += v;
    }
    if (colsum != nullptr) cs[ty][tx] = part;
    __syncthreads();
    if (colsum != nullptr && ty == 0 && c0 + tx < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += cs[i][tx];
        atomicAdd(colsum + c0 + tx, s);
    }
    if (dst_t != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + ty + 8 * i, r = r0 + tx;
            if (c < C && r < R) dst_t[static_cast<size_t>(c) * ld_t + r] = __float2bfloat16(tile[tx][ty + 8 * i]);
        }
    }
}

// All weight matrices of the trainable mapper in ONE launch: fp32 [rows, cols] -> bf16 copy and/or bf16 transpose
// (round-1 profile: 51 separate convert launches of ~7 us each = 0.36 ms / step for 0.33 GB of traffic).
// 64 x 32 tiles: each thread converts float2 -> bf16x2 (8-byte loads, 4-byte stores on the natural copy).
__global__ void __launch_bounds__(256) pack_batch_kernel(const __grid_constant__ PackJobs jobs) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[32][65];
    int j = 0;
    while (j + 1 < jobs.n && static_cast<int>(blockIdx.x) >= jobs.job[j + 1].tile_begin) ++j;
    const PackJob& jb = jobs.job[j];
    const int t = blockIdx.x - jb.tile_begin;
    const int tiles_x = (jb.cols + 63) >> 6;
    const int r0 = (t / tiles_x) * 32, c0 = (t % tiles_x) * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const int R = jb.rows, C = jb.cols;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
        float2 v = make_float2(0.f, 0.f);
        if (r < R && c < C) {                                     // cols are even (checked on the host)
            v = *reinterpret_cast<const float2*>(jb.src + static_cast<size_t>(r) * C + c);
            if (jb.dst != nullptr)
                *reinterpret_cast<__nv_bfloat162*>(jb.dst + static_cast<size_t>(r) * C + c) = __floats2bfloat162_rn(v.x, v.y);
        }
        tile[ty + 8 * i][2 * tx] = v.x;
        tile[ty + 8 * i][2 * tx + 1] = v.y;
    }
    if (jb.dst_t == nullptr) return;
    __syncthreads();
    // transposed copy [C, R]: thread writes rows (r0 + 2 * (tx & 15), +1) of column c0 + ...; 16 lanes x bf16x2 = 64-byte runs
    const int half = tx >> 4, rr = 2 * (tx & 15);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty * 8 + i * 2 + half;
        const int r = r0 + rr;
        if (c < C && r < R) {
            const float a = tile[rr][c - c0], b = tile[rr + 1][c - c0];
            if (r + 1 < R) *reinterpret_cast<__nv_bfloat162*>(jb.dst_t + static_cast<size_t>(c) * R + r) = __floats2bfloat162_rn(a, b);
            else jb.dst_t[static_cast<size_t>(c) * R + r] = __float2bfloat16(a);
        }
    }
}

__global__ void broadcast_rows_kernel(const float4* __restrict__ src, int n4, float4* __restrict__ dst,
                                      int64_t batch_stride4, int B) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = static_cast<int64_t>(B) * n4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / n4), c = static_cast<int>(i % n4);
        dst[b * batch_stride4 + c] = __ldg(src + c);
    }
}

__global__ void sum_over_batch_kernel(const float* __restrict__ src, int64_t batch_stride, int B, int n,
                                      float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += src[b * batch_stride + c];
    out[c] = s;
}

// ------------------------------------------------------------------------------------------ LayerNorm
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, int ld_x,
                                                            const int* __restrict__ row_index,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y,
                                                            int ld_y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int M, int d, float eps) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= M) return;
    const int xr = row_index ? row_index[m] : m;
    const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(xr) * ld_x);
    const int n4 = d >> 2;
    float4 v[MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            v[i] = xp[c];
            sum += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    }
    const float mean = warp_sum(sum) / d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
            sq += a * a + b * b + e * e + f * f;
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / d + eps);
    if (lane == 0) {
        if (mean_out) mean_out[m] = mean;
        if (rstd_out) rstd_out[m] = rstd;
    }
    uint2* yp = reinterpret_cast<uint2*>(y + static_cast<size_t>(m) * ld_y);
    const float4* gp = reinterpret_cast<const float4*>(gamma);
    const float4* bp = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            const float4 g = __ldg(gp + c), b = __ldg(bp + c);
            uint2 o;
            o.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
            o.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            yp[c] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------ single-token decode glue
// Decode steps run their M = batch-row GEMMs split along K over all SMs, accumulating fp32 partials into scratch rows
// (profiles/: unsplit, the projections used 16 of 148 SMs and took 16-20 us each).  This kernel is what sits between
// those GEMMs: it applies bias / residual / LayerNorm to the accumulated rows and zeroes the scratch rows of the NEXT
// split-K GEMM.  One 128-thread CTA per row.
__device__ __forceinline__ float block_sum_128(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                       // red may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

// x[row] += acc[row] + bias (when acc != null); u[row] = LN(x[row]) * gamma + beta (bf16); zero[row, 0..zero_n) = 0
__global__ void __launch_bounds__(128) decode_residual_ln_kernel(float* __restrict__ x, const float* __restrict__ acc,
                                                                 const float* __restrict__ bias, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, bf16* __restrict__ u, int d,
                                                                 float eps, float* __restrict__ zero, int zero_n) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[4];
    const int row = blockIdx.x, n4 = d >> 2;
    float4* xp = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * d);
    // every global load of the row (x, accumulator, bias, gamma, beta) is issued before the first dependent store: issue is
    // in order, so a store between two loads serialises their L2 round trips (measured in the persistent decode kernel:
    // 8 dependent trips cost 4 us; this kernel is pure latency)
    float4 v[4], a[4], bb[4], g[4], be[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + 128 * i;
        if (c < n4) {
            v[i] = xp[c];
            if (acc != nullptr) {
                a[i] = reinterpret_cast<const float4*>(acc + static_cast<size_t>(row) * d)[c];
                bb[i] = __ldg(reinterpret_cast<const float4*>(bias) + c);
            }
            g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + c);
            be[i] = __ldg(reinterpret_cast<const float4*>(beta) + c);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + 128 * i;
        if (c < n4) {
            if (acc != nullptr) {
                v[i].x += a[i].x + bb[i].x; v[i].y += a[i].y + bb[i].y; v[i].z += a[i].z + bb[i].z; v[i].w += a[i].w + bb[i].w;
                xp[c] = v[i];
            }
            sum += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    }
    const float mean = block_sum_128(sum, red) / d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + 128 * i;
        if (c < n4) {
            const float e0 = v[i].x - mean, e1 = v[i].y - mean, e2 = v[i].z - mean, e3 = v[i].w - mean;
            sq += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
        }
    }
    const float rstd = rsqrtf(block_sum_128(sq, red) / d + eps);
    uint2* up = reinterpret_cast<uint2*>(u + static_cast<size_t>(row) * d);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + 128 * i;
        if (c < n4) {
            uint2 o;
            o.x = pack_bf16x2((v[i].x - mean) * rstd * g[i].x + be[i].x, (v[i].y - mean) * rstd * g[i].y + be[i].y);
            o.y = pack_bf16x2((v[i].z - mean) * rstd * g[i].z + be[i].z, (v[i].w - mean) * rstd * g[i].w + be[i].w);
            up[c] = o;
        }
    }
    if (zero != nullptr) {
        float4* zp = reinterpret_cast<float4*>(zero + static_cast<size_t>(row) * zero_n);
        for (int c = threadIdx.x; c < (zero_n >> 2); c += 128) zp[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int MAXV, bool PARAM_GRADS>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ dy, int ld_dy,
                                                            const float* __restrict__ x, int ld_x,
                                                            const int* __restrict__ row_index,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, float* __restrict__ dx,
                                                            int ld_dx, int accumulate, bf16* __restrict__ dx_bf16,
                                                            int ld_dxb, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int M, int d) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float red[];   // PARAM_GRADS: [2][d] block-level dgamma / dbeta
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int n4 = d >> 2;
    if (PARAM_GRADS) {
        for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) red[i] = 0.f;
        __syncthreads();
    }
    float4 ag[PARAM_GRADS ? MAXV : 1], ab[PARAM_GRADS ? MAXV : 1];
    if (PARAM_GRADS) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float4* gp = reinterpret_cast<const float4*>(gamma);
    for (int m = blockIdx.x * wpb + warp; m < M; m += gridDim.x * wpb) {
        const int xr = row_index ? row_index[m] : m;
        const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(xr) * ld_x);
        const uint2* dyp = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(m) * ld_dy);
        const float mean = mean_in[m], rstd = rstd_in[m];
        float4 xh[MAXV], g[MAXV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = lane + 32 * i;
            if (c < n4) {
                const float4 xv = xp[c];
                const uint2 u = dyp[c];
                const float2 d0 = unpack_bf16x2(u.x), d1 = unpack_bf16x2(u.y);
                const float4 gm = __ldg(gp + c);
                xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
                if (PARAM_GRADS) {
                    ag[i].x += d0.x * xh[i].x; ag[i].y += d0.y * xh[i].y; ag[i].z += d1.x * xh[i].z; ag[i].w += d1.y * xh[i].w;
                    ab[i].x += d0.x; ab[i].y += d0.y; ab[i].z += d1.x; ab[i].w += d1.y;
                }
                g[i] = make_float4(d0.x * gm.x, d0.y * gm.y, d1.x * gm.z, d1.y * gm.w);
                s1 += g[i].x + g[i].y + g[i].z + g[i].w;
                s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
            }
        }
        s1 = warp_sum(s1) / d;
        s2 = warp_sum(s2) / d;
        float4* dxp = reinterpret_cast<float4*>(dx + static_cast<size_t>(xr) * ld_dx);
        uint2* dbp = dx_bf16 ? reinterpret_cast<uint2*>(dx_bf16 + static_cast<size_t>(xr) * ld_dxb) : nullptr;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = lane + 32 * i;
            if (c < n4) {
                float4 r;
                r.x = rstd * (g[i].x - s1 - xh[i].x * s2);
                r.y = rstd * (g[i].y - s1 - xh[i].y * s2);
                r.z = rstd * (g[i].z - s1 - xh[i].z * s2);
                r.w = rstd * (g[i].w - s1 - xh[i].w * s2);
                if (accumulate) {
                    const float4 o = dxp[c];
                    r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
                }
                dxp[c] = r;
                if (dbp) {
                    uint2 o;
                    o.x = pack_bf16x2(r.x, r.y);
                    o.y = pack_bf16x2(r.z, r.w);
                    dbp[c] = o;
                }
            }
        }
    }
    if (PARAM_GRADS) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = lane + 32 * i;
            if (c < n4) {
                atomicAdd(red + 4 * c + 0, ag[i].x); atomicAdd(red + 4 * c + 1, ag[i].y);
                atomicAdd(red + 4 * c + 2, ag[i].z); atomicAdd(red + 4 * c + 3, ag[i].w);
                atomicAdd(red + d + 4 * c + 0, ab[i].x); atomicAdd(red + d + 4 * c + 1, ab[i].y);
                atomicAdd(red + d + 4 * c + 2, ab[i].z); atomicAdd(red + d + 4 * c + 3, ab[i].w);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < d; i += blockDim.x) {
            atomicAdd(dgamma + i, red[i]);
            atomicAdd(dbeta + i, red[d + i]);
        }
    }
}

// Frozen-LM / dx-only LayerNorm backward, tuned for occupancy (round-1 ncu: the fused variant used 103 registers ->
// 2 CTAs/SM, 28 % of DRAM peak).  x and dy stay in registers (x fp32, dy packed bf16); xhat and gamma*dy are
// recomputed in the second sweep instead of being kept.
template <int MAXV>
__global__ void __launch_bounds__(256, (MAXV <= 6) ? 4 : (MAXV <= 10) ? 3 : 2) layernorm_bwd_lean_kernel(const bf16* __restrict__ dy, int ld_dy,
                                                                    const float* __restrict__ x, int ld_x,
                                                                    const int* __restrict__ row_index,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ mean_in,
                                                                    const float* __restrict__ rstd_in, float* __restrict__ dx,
                                                                    int ld_dx, int accumulate, bf16* __restrict__ dx_bf16,
                                                                    int ld_dxb, int M, int d) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= M) return;
    const int n4 = d >> 2;
    const int xr = row_index ? row_index[m] : m;
    const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(xr) * ld_x);
    const uint2* dyp = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(m) * ld_dy);
    const float4* gp = reinterpret_cast<const float4*>(gamma);
    const float mean = mean_in[m], rstd = rstd_in[m];
    float4 xv[MAXV];
    uint2 dv[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            xv[i] = xp[c];
            dv[i] = dyp[c];
        }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            const float4 gm = __ldg(gp + c);
            const float2 d0 = unpack_bf16x2(dv[i].x), d1 = unpack_bf16x2(dv[i].y);
            const float g0 = d0.x * gm.x, g1 = d0.y * gm.y, g2 = d1.x * gm.z, g3 = d1.y * gm.w;
            s1 += g0 + g1 + g2 + g3;
            s2 += g0 * (xv[i].x - mean) + g1 * (xv[i].y - mean) + g2 * (xv[i].z - mean) + g3 * (xv[i].w - mean);
        }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) * rstd / d;
    float4* dxp = reinterpret_cast<float4*>(dx + static_cast<size_t>(xr) * ld_dx);
    uint2* dbp = dx_bf16 ? reinterpret_cast<uint2*>(dx_bf16 + static_cast<size_t>(xr) * ld_dxb) : nullptr;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < n4) {
            const float4 gm = __ldg(gp + c);
            const float2 d0 = unpack_bf16x2(dv[i].x), d1 = unpack_bf16x2(dv[i].y);
            float4 r;
            r.x = rstd * (d0.x * gm.x - s1 - (xv[i].x - mean) * rstd * s2);
            r.y = rstd * (d0.y * gm.y - s1 - (xv[i].y - mean) * rstd * s2);
            r.z = rstd * (d1.x * gm.z - s1 - (xv[i].z - mean) * rstd * s2);
            r.w = rstd * (d1.y * gm.w - s1 - (xv[i].w - mean) * rstd * s2);
            if (accumulate) {
                const float4 o = dxp[c];
                r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
            }
            dxp[c] = r;
            if (dbp) {
                uint2 o;
                o.x = pack_bf16x2(r.x, r.y);
                o.y = pack_bf16x2(r.z, r.w);
                dbp[c] = o;
            }
        }
    }
}

// column reduction over rows: block = 32 column-octets (256 columns, 16-byte dy loads) x 8 row lanes, 64 rows per block
__global__ void __launch_bounds__(256) ln_param_grad_kernel(const bf16* __restrict__ dy, int ld_dy, const float* __restrict__ x,
                                                            int ld_x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int M, int d) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sg[8][256], sb[8][256];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 256 + 8 * tx;
    const int r0 = blockIdx.y * 64;
    float g[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { g[k] = 0.f; b[k] = 0.f; }
    if (c < d) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + ty + 8 * i;
            if (r < M) {
                const uint4 u = *reinterpret_cast<const uint4*>(dy + static_cast<size_t>(r) * ld_dy + c);
                const float4 x0 = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_x + c);
                const float4 x1 = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_x + c + 4);
                const float mu = mean[r], rs = rstd[r];
                float dv[8];
                float2 t;
                t = unpack_bf16x2(u.x); dv[0] = t.x; dv[1] = t.y;
                t = unpack_bf16x2(u.y); dv[2] = t.x; dv[3] = t.y;
                t = unpack_bf16x2(u.z); dv[4] = t.x; dv[5] = t.y;
                t = unpack_bf16x2(u.w); dv[6] = t.x; dv[7] = t.y;
                const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    g[k] += dv[k] * (xv[k] - mu) * rs;
                    b[k] += dv[k];
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sg[ty][8 * tx + k] = g[k];
        sb[ty][8 * tx + k] = b[k];
    }
    __syncthreads();
    const int t = ty * 32 + tx;      // 256 threads: one column each
    const int col = blockIdx.x * 256 + t;
    if (col < d) {
        float gs = 0.f, bs = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { gs += sg[i][t]; bs += sb[i][t]; }
        atomicAdd(dgamma + col, gs);
        atomicAdd(dbeta + col, bs);
    }
}

// ------------------------------------------------------------------------------------------ embedding / splice
__global__ void prepend_plan_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ mask, int B, int Tt,
                                    int P, int* __restrict__ plan, int* __restrict__ valid) {
    pdl_trigger();
    pdl_wait();
    const int T = P + Tt;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    const int b = i / T, t = i % T;
    if (t < P) {
        plan[i] = -(t + 1);
        valid[i] = 1;
    } else {
        plan[i] = static_cast<int>(tokens[b * Tt + t - P]);
        valid[i] = mask ? (mask[b * Tt + t - P] != 0) : 1;
    }
}

// one warp per prompt row: ballot prefix-sum of sentinel flags -> destination index of every text
// token and every prefix row (vct0.py:494-533: dest = (j - c) + P * c).
__global__ void splice_plan_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ mask, int B, int Tt,
                                   int P, int n_img, int64_t sent_lo, int64_t sent_hi, int* __restrict__ plan,
                                   int* __restrict__ valid, int* __restrict__ err_flag) {
    pdl_trigger();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int b = warp;
    const int T_out = Tt + (P - 1) * n_img;
    int base = 0;
    for (int j0 = 0; j0 < Tt; j0 += 32) {
        const int j = j0 + lane;
        int64_t tok = 0;
        bool sent = false;
        if (j < Tt) {
            tok = tokens[b * Tt + j];
            sent = tok >= sent_lo && tok <= sent_hi;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, sent);
        const int c = base + __popc(bal & ((1u << lane) - 1u));
        if (j < Tt) {
            const int dst = (j - c) + P * c;
            if (sent) {
                if (c < n_img)
                    for (int r = 0; r < P; ++r)
                        if (dst + r < T_out) {
                            plan[b * T_out + dst + r] = -(c * P + r + 1);
                            valid[b * T_out + dst + r] = 1;
                        }
            } else if (dst < T_out) {
                plan[b * T_out + dst] = static_cast<int>(tok);
                valid[b * T_out + dst] = mask ? (mask[b * Tt + j] != 0) : 1;
            }
        }
        base += __popc(bal);
    }
    if (lane == 0 && base != n_img) atomicExch(err_flag, 1);
}

__global__ void __launch_bounds__(256) embed_rows_kernel(const int* __restrict__ plan, int rows, int T, int d,
                                                         const float* __restrict__ wte, int vocab,
                                                         const float* __restrict__ prefix, int64_t prefix_batch_stride,
                                                         int prefix_row_stride, const float* __restrict__ wpe,
                                                         float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= rows) return;
    const int b = row / T, t = row % T;
    const int p = plan[row];
    const float4* src;
    if (p >= 0) {
        const int tok = p < vocab ? p : vocab - 1;
        src = reinterpret_cast<const float4*>(wte + static_cast<size_t>(tok) * d);
    } else {
        src = reinterpret_cast<const float4*>(prefix + b * prefix_batch_stride + static_cast<size_t>(-p - 1) * prefix_row_stride);
    }
    const float4* pe = wpe ? reinterpret_cast<const float4*>(wpe + static_cast<size_t>(t) * d) : nullptr;
    float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(row) * d);
    for (int c = lane; c < (d >> 2); c += 32) {
        float4 v = __ldg(src + c);
        if (pe) {
            const float4 w = __ldg(pe + c);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        o[c] = v;
    }
}

// ------------------------------------------------------------------------------------------ cross-entropy
__global__ void ce_plan_kernel(const int64_t* __restrict__ labels, int B, int Tt, int T, int P, int vocab,
                               int* __restrict__ row_index, int* __restrict__ label, int* __restrict__ n_valid) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (r < B * Tt) {
        const int b = r / Tt, j = r % Tt;
        const int64_t lab = labels[r];
        ok = lab >= 0 && lab < vocab;        // -100 = ignore_index
        row_index[r] = b * T + P - 1 + j;
        label[r] = ok ? static_cast<int>(lab) : -1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_valid, __popc(bal));
}

__global__ void __launch_bounds__(256) ce_finalize_kernel(const float2* __restrict__ partial, int tiles,
                                                          const float* __restrict__ target,
                                                          const int* __restrict__ label, float* __restrict__ lse,
                                                          float* __restrict__ loss_sum, int M) {
    pdl_trigger();
    pdl_wait();
    __shared__ float block_loss[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    float row_loss = 0.f;
    if (r < M) {
        float m = -INFINITY, s = 0.f;
        for (int i = lane; i < tiles; i += 32) {
            const float2 p = partial[static_cast<size_t>(r) * tiles + i];
            if (p.x > -INFINITY) {
                const float nm = fmaxf(m, p.x);
                s = s * __expf(m - nm) + p.y * __expf(p.x - nm);
                m = nm;
            }
        }
        const float gm = warp_max(m);
        s = (m > -INFINITY) ? s * __expf(m - gm) : 0.f;
        s = warp_sum(s);
        const float l = gm + logf(s);
        if (lane == 0) {
            lse[r] = l;
            if (label[r] >= 0) row_loss = l - target[r];
        }
    }
    if (lane == 0) block_loss[warp] = row_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += block_loss[i];
        if (t != 0.f) atomicAdd(loss_sum, t);
    }
}

__global__ void ce_loss_kernel(const float* loss_sum, const int* n_valid, float* loss_out) {
    pdl_trigger();
    pdl_wait();
    *loss_out = *loss_sum / static_cast<float>(*n_valid);    // 0/0 = NaN, like the mean over no targets
}

__global__ void __launch_bounds__(256) ce_dlogits_kernel(bf16* __restrict__ z, int ld, int vocab, int n_cols,
                                                         const float* __restrict__ lse, const int* __restrict__ label,
                                                         const int* __restrict__ n_valid) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x;
    const int lab = label[r];
    uint4* zp = reinterpret_cast<uint4*>(z + static_cast<size_t>(r) * ld);
    const int n8 = n_cols >> 3;
    if (lab < 0) {
        for (int c = threadIdx.x; c < n8; c += blockDim.x) zp[c] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const float l = lse[r];
    const float w = 1.0f / static_cast<float>(*n_valid);
    for (int c = threadIdx.x; c < n8; c += blockDim.x) {
        const uint4 u = zp[c];
        float f[8];
        float2 t;
        t = unpack_f16x2(u.x); f[0] = t.x; f[1] = t.y;        // the head stores its logits as fp16; d logits go back as bf16,
        t = unpack_f16x2(u.y); f[2] = t.x; f[3] = t.y;        // the dgrad GEMM's operand type, in place
        t = unpack_f16x2(u.z); f[4] = t.x; f[5] = t.y;
        t = unpack_f16x2(u.w); f[6] = t.x; f[7] = t.y;
        const int v0 = c * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int v = v0 + k;
            float g = 0.f;
            if (v < vocab) {
                g = __expf(f[k] - l);
                if (v == lab) g -= 1.0f;
                g *= w;
            }
            f[k] = g;
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
        zp[c] = o;
    }
}

// ------------------------------------------------------------------------------------------ greedy decode
__global__ void __launch_bounds__(256) greedy_step_kernel(const float* __restrict__ logits, int ld, int vocab, int step,
                                                          int max_new, int has_eos, int64_t pad_id, int64_t eos_id,
                                                          int* __restrict__ unfinished,
                                                          int64_t* __restrict__ tokens_out,
                                                          int* __restrict__ n_unfinished,
                                                          float* __restrict__ top_logit,
                                                          float* __restrict__ token_logprob,
                                                          const float* __restrict__ wte,
                                                          const float* __restrict__ wpe_row, int d,
                                                          float* __restrict__ x_next, int* __restrict__ valid_next,
                                                          int valid_stride, float* __restrict__ part_val,
                                                          int* __restrict__ part_idx, unsigned* __restrict__ arrivals) {
    // grid (B, split): every row's vocabulary is scanned by `split` CTAs (one CTA per row left 20 of 148 SMs idle and took
    // 94 us for 26 MB); each posts its (max, index) and the LAST one to arrive merges them and does the bookkeeping.
    pdl_trigger();
    pdl_wait();
    __shared__ float s_val[8];
    __shared__ int s_idx[8];
    __shared__ int s_next;
    __shared__ float s_best;
    __shared__ int s_last;
    const int b = blockIdx.x;
    const int split = gridDim.y, sl = blockIdx.y;
    const float* lp = logits + static_cast<size_t>(b) * ld;
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    // rows are 16-byte aligned (ld % 4 == 0): float4 loads, 4 of them in flight per thread; ties keep the lowest index
    const float4* lp4 = reinterpret_cast<const float4*>(lp);
    const int n4_all = vocab >> 2;
    const int per = (n4_all + split - 1) / split;
    const int n4_lo = sl * per, n4 = min(n4_all, n4_lo + per);
    for (int v0 = n4_lo + threadIdx.x; v0 < n4; v0 += 4 * blockDim.x) {
        float4 x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = v0 + k * blockDim.x;
            x[k] = v < n4 ? lp4[v] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = (v0 + k * blockDim.x) * 4;
            if (x[k].x > best) { best = x[k].x; best_i = v; }
            if (x[k].y > best) { best = x[k].y; best_i = v + 1; }
            if (x[k].z > best) { best = x[k].z; best_i = v + 2; }
            if (x[k].w > best) { best = x[k].w; best_i = v + 3; }
        }
    }
    if (sl == split - 1)
        for (int v = (n4_all << 2) + threadIdx.x; v < vocab; v += blockDim.x) {
            const float x = lp[v];
            if (x > best || (x == best && v < best_i)) { best = x; best_i = v; }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) {
            best = ov;
            best_i = oi;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_val[warp] = best;
        s_idx[warp] = best_i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i)
            if (s_val[i] > best || (s_val[i] == best && s_idx[i] < best_i)) {
                best = s_val[i];
                best_i = s_idx[i];
            }
        if (split > 1) {
            part_val[b * split + sl] = best;
            part_idx[b * split + sl] = best_i;
            __threadfence();
            const unsigned t = atomicAdd(arrivals + b, 1u);
            s_last = (t == static_cast<unsigned>(split - 1));
            if (s_last) {
                __threadfence();
                arrivals[b] = 0;                                  // ready for the next step
                for (int i = 0; i < split; ++i) {
                    const float pv = __ldcg(part_val + b * split + i);
                    const int pi = __ldcg(part_idx + b * split + i);
                    if (i == 0 || pv > best || (pv == best && pi < best_i)) { best = pv; best_i = pi; }
                }
            }
        } else {
            s_last = 1;
        }
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) {
        if (best_i == 0x7fffffff) best_i = 0;
        int unf = unfinished[b];
        int64_t out = best_i;
        if (has_eos) {
            out = unf ? static_cast<int64_t>(best_i) : pad_id;      // clipcap.py:431-434
            unf = unf && (out != eos_id);                           // clipcap.py:458-461
        }
        tokens_out[static_cast<size_t>(b) * max_new + step] = out;
        unfinished[b] = unf;
        if (unf) atomicAdd(n_unfinished + step, 1);
        if (top_logit) top_logit[static_cast<size_t>(b) * max_new + step] = best;
        if (valid_next) valid_next[static_cast<size_t>(b) * valid_stride] = 1;   // appended position attends as 1
        s_next = best_i;                                            // the RAW argmax is fed back (clipcap.py:423)
        s_best = best;
    }
    __syncthreads();
    const float4* e = reinterpret_cast<const float4*>(wte + static_cast<size_t>(s_next) * d);
    const float4* pe = reinterpret_cast<const float4*>(wpe_row);
    float4* o = reinterpret_cast<float4*>(x_next + static_cast<size_t>(b) * d);
    for (int c = threadIdx.x; c < (d >> 2); c += blockDim.x) {
        float4 v = __ldg(e + c);
        const float4 w = __ldg(pe + c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        o[c] = v;
    }
    if (token_logprob != nullptr) {
        // log softmax(logits)[argmax] = -log sum_v exp(z_v - z_max): second pass over the row (L2-resident), only for
        // ensemble scoring (few_shot_vqa_executor.py:316, torch.log(softmax(scores)))
        const float mx = s_best;
        float sum = 0.f;
        for (int v = threadIdx.x; v < vocab; v += blockDim.x) sum += __expf(lp[v] - mx);
        sum = warp_sum(sum);
        __syncthreads();                       // s_val is reused
        if (lane == 0) s_val[warp] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; ++i) t += s_val[i];
            token_logprob[static_cast<size_t>(b) * max_new + step] = -logf(t);
        }
    }
}

// ------------------------------------------------------------------------------------------ executor-side steps
// ClipCapExecutor.training_step label construction (clipcap_exector.py:134-150), one thread per caption:
//   labels = input_ids with pads -> -100; everything up to and including each <BOS> -> -100; tokens after the first
//   <BOS> are kept; the FIRST pad position gets the pad (= eos) id back as its target and ends the scan.
__global__ void caption_labels_kernel(const int64_t* __restrict__ tokens, int B, int T, int64_t pad_id, int64_t bos_id,
                                      int64_t* __restrict__ labels) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t* t = tokens + static_cast<size_t>(b) * T;
    int64_t* l = labels + static_cast<size_t>(b) * T;
    bool answer = false, ended = false;
    for (int j = 0; j < T; ++j) {
        const int64_t tok = t[j];
        int64_t out;
        if (ended) {
            out = (tok == pad_id) ? -100 : tok;            // positions after the break keep the first assignment
        } else if (tok == pad_id) {                        // first pad (already -100 in the clone): target = pad id, stop
            out = pad_id;
            ended = true;
        } else if (tok == bos_id) {
            answer = true;
            out = -100;
        } else {
            out = answer ? tok : -100;
        }
        l[j] = out;
    }
}

// FewShotVQAExecutor.generate_from_ensembles scoring (few_shot_vqa_executor.py:316-331): per row b and ensemble member e,
// score = sum over generated tokens not in `skip` of their log-probability; best[b] = first argmax over e; the winning
// member's tokens are copied out.  One warp per row.
__global__ void ensemble_select_kernel(const float* __restrict__ logprob, const int64_t* __restrict__ tokens, int E, int B, int S,
                                       const int64_t* __restrict__ skip, int n_skip, float* __restrict__ scores,
                                       int* __restrict__ best, int64_t* __restrict__ best_tokens) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float best_v = -INFINITY;
    int best_e = 0;
    for (int e = 0; e < E; ++e) {
        const size_t base = (static_cast<size_t>(e) * B + b) * S;
        float acc = 0.f;
        for (int k = lane; k < S; k += 32) {
            const int64_t tok = tokens[base + k];
            bool skipped = false;
            for (int i = 0; i < n_skip; ++i) skipped |= (tok == skip[i]);
            if (!skipped) acc += logprob[base + k];
        }
        acc = warp_sum(acc);
        if (lane == 0 && scores != nullptr) scores[static_cast<size_t>(b) * E + e] = acc;
        if (acc > best_v) {                                // strict: np.argmax keeps the first maximum
            best_v = acc;
            best_e = e;
        }
    }
    if (lane == 0) best[b] = best_e;
    const size_t src = (static_cast<size_t>(best_e) * B + b) * S;
    for (int k = lane; k < S; k += 32) best_tokens[static_cast<size_t>(b) * S + k] = tokens[src + k];
}

// x *= *scale unless *scale == 1 (the upstream gradient of loss.backward() is a device scalar that is almost always 1:
// skipping the pass saves a read + write of the 167 MB gradient buffer per step without a host sync)
__global__ void __launch_bounds__(256) scale_by_device_scalar_kernel(float4* __restrict__ x, int64_t n4, const float* __restrict__ scale) {
    pdl_trigger();
    pdl_wait();
    const float sc = __ldg(scale);
    if (sc == 1.0f) return;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 v = x[i];
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        x[i] = v;
    }
}

__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                    float4* __restrict__ v, int64_t n4, float lr, float beta1, float beta2,
                                                    float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        float* pa = reinterpret_cast<float*>(&pp);
        float* ga = reinterpret_cast<float*>(&gg);
        float* ma = reinterpret_cast<float*>(&mm);
        float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = ga[k] * gscale;
            pa[k] *= 1.0f - lr * wd;
            ma[k] = beta1 * ma[k] + (1.0f - beta1) * gr;
            va[k] = beta2 * va[k] + (1.0f - beta2) * gr * gr;
            const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
            pa[k] -= (lr / bc1) * (ma[k] / denom);
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace

// ============================================================================================ launchers
void adamw_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                float eps, float weight_decay, int step, float grad_scale, cudaStream_t s) {
    EAVQA_CHECK(n % 4 == 0 && step >= 1, "adamw_step: n must be a multiple of 4 and step >= 1");
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(n / 4, 256), static_cast<int64_t>(num_sms()) * 8));
    launch_kernel(adamw_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads),
                                      reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n / 4, lr, beta1, beta2, eps,
                                      weight_decay, bc1, sqrtf(bc2), grad_scale);
    KERNEL_CHECK();
    count_launch();
}

void convert_transpose_f32(const float* src, int ld_src, int R, int C, bf16* dst, int ld_dst, bf16* dst_t, int ld_t,
                           float* colsum, cudaStream_t s) {
    dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
    launch_kernel(convert_transpose_kernel<float>, dim3(grid), dim3(block), 0, s, src, ld_src, R, C, dst, ld_dst, dst_t, ld_t, colsum);
    KERNEL_CHECK();
    count_launch();
}
void pack_batch(const std::vector<PackJob>& all, cudaStream_t s) {
    size_t i = 0;
    while (i < all.size()) {
        PackJobs jobs;
        jobs.n = 0;
        int tiles = 0;
        while (i < all.size() && jobs.n < PackJobs::kMax) {
            PackJob jb = all[i++];
            EAVQA_CHECK(jb.rows > 0 && jb.cols > 0 && jb.cols % 2 == 0 && jb.rows % 2 == 0, "pack_batch: even matrix sizes");
            EAVQA_CHECK((reinterpret_cast<uintptr_t>(jb.src) & 7) == 0, "pack_batch: source must be 8-byte aligned");
            jb.tile_begin = tiles;
            tiles += ceil_div(jb.rows, 32) * ceil_div(jb.cols, 64);
            jobs.job[jobs.n++] = jb;
        }
        launch_kernel(pack_batch_kernel, dim3(tiles), dim3(256), 0, s, jobs);
        KERNEL_CHECK();
        count_launch();
    }
}
void convert_transpose_bf16(const bf16* src, int ld_src, int R, int C, bf16* dst_t, int ld_t, float* colsum,
                            cudaStream_t s) {
    dim3 grid(ceil_div(C, 32), ceil_div(R, 32)), block(32, 8);
    launch_kernel(convert_transpose_kernel<bf16>, dim3(grid), dim3(block), 0, s, src, ld_src, R, C, nullptr, 0, dst_t, ld_t, colsum);
    KERNEL_CHECK();
    count_launch();
}
void broadcast_rows_f32(const float* src, int rows, int d, float* dst, int64_t batch_stride, int B, cudaStream_t s) {
    EAVQA_CHECK(d % 4 == 0 && batch_stride % 4 == 0, "broadcast_rows alignment");
    const int n4 = rows * d / 4;
    const int64_t total = static_cast<int64_t>(B) * n4;
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(total, 256), 148 * 8));
    launch_kernel(broadcast_rows_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const float4*>(src), n4, reinterpret_cast<float4*>(dst),
                                               batch_stride / 4, B);
    KERNEL_CHECK();
    count_launch();
}
void sum_over_batch_f32(const float* src, int64_t batch_stride, int B, int n, float* out, cudaStream_t s) {
    launch_kernel(sum_over_batch_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, s, src, batch_stride, B, n, out);
    KERNEL_CHECK();
    count_launch();
}
void fill_zero(void* p, size_t bytes, cudaStream_t s) { CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, s)); }

namespace {
constexpr int kPrefetchChunk = 8192;
__global__ void __launch_bounds__(128) prefetch_l2_kernel(const uint8_t* __restrict__ base, size_t bytes) {
    pdl_trigger();
    pdl_wait();
    const size_t off = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * kPrefetchChunk;
    if (off < bytes) {
        const size_t left = bytes - off;
        const uint32_t sz = static_cast<uint32_t>(left < static_cast<size_t>(kPrefetchChunk) ? left : static_cast<size_t>(kPrefetchChunk)) & ~15u;
        if (sz > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(sz) : "memory");
    }
}
}  // namespace

void prefetch_l2(const void* p, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return;
    const int threads = static_cast<int>(ceil_div64(static_cast<int64_t>(bytes), kPrefetchChunk));
    launch_kernel(prefetch_l2_kernel, dim3(ceil_div(threads, 128)), dim3(128), 0, s, static_cast<const uint8_t*>(p), bytes);
    KERNEL_CHECK();
    count_launch();
}

void layernorm_fwd(const float* x, int ld_x, const int* row_index, const float* gamma, const float* beta, bf16* y,
                   int ld_y, float* mean, float* rstd, int M, int d, float eps, cudaStream_t s) {
    EAVQA_CHECK(d % 4 == 0 && d <= 2048 && ld_x % 4 == 0 && ld_y % 4 == 0, "layernorm width must be a multiple of 4 and <= 2048");
    const int need = ceil_div(d / 4, 32);
    const int grid = ceil_div(M, 8);
#define EAVQA_LN_FWD(V) launch_kernel(layernorm_fwd_kernel<V>, dim3(grid), dim3(256), 0, s, x, ld_x, row_index, gamma, beta, y, ld_y, mean, rstd, M, d, eps)
    if (need <= 2) EAVQA_LN_FWD(2);
    else if (need <= 4) EAVQA_LN_FWD(4);
    else if (need <= 6) EAVQA_LN_FWD(6);
    else if (need <= 8) EAVQA_LN_FWD(8);
    else if (need <= 10) EAVQA_LN_FWD(10);
    else if (need <= 13) EAVQA_LN_FWD(13);
    else EAVQA_LN_FWD(16);
#undef EAVQA_LN_FWD
    KERNEL_CHECK();
    count_launch();
}

void decode_residual_ln(float* x, const float* acc, const float* bias, const float* gamma, const float* beta, bf16* u, int rows, int d,
                        float eps, float* zero, int zero_n, cudaStream_t s) {
    EAVQA_CHECK(d % 4 == 0 && d <= 2048 && zero_n % 4 == 0, "decode_residual_ln: width must be a multiple of 4 and <= 2048");
    EAVQA_CHECK((acc == nullptr) == (bias == nullptr), "decode_residual_ln: acc and bias come together");
    launch_kernel(decode_residual_ln_kernel, dim3(rows), dim3(128), 0, s, x, acc, bias, gamma, beta, u, d, eps, zero, zero_n);
    KERNEL_CHECK();
    count_launch();
}
template <int MAXV>
static void launch_ln_bwd(const bf16* dy, int ld_dy, const float* x, int ld_x, const int* row_index, const float* gamma,
                          const float* mean, const float* rstd, float* dx, int ld_dx, int accumulate, bf16* dx_bf16,
                          int ld_dxb, float* dgamma, float* dbeta, int M, int d, cudaStream_t s) {
    if (dgamma != nullptr && row_index != nullptr) {
        // gathered rows + parameter gradients: not on the step's path; keep the fused (register-heavy) variant
        const int grid = std::min(ceil_div(M, 8), 2 * num_sms());
        launch_kernel(layernorm_bwd_kernel<MAXV, true>, dim3(grid), dim3(256), 2 * d * sizeof(float), s, 
            dy, ld_dy, x, ld_x, row_index, gamma, mean, rstd, dx, ld_dx, accumulate, dx_bf16, ld_dxb, dgamma, dbeta, M, d);
        KERNEL_CHECK();
        count_launch();
        return;
    }
    // the 157-register fused variant ran at one CTA per SM (round-1 profile: 47 us for 63 MB); the column reduction
    // re-reads dy and x (L2-resident) in a second, occupancy-friendly kernel instead
    if (dgamma != nullptr) layernorm_param_grads(dy, ld_dy, x, ld_x, mean, rstd, dgamma, dbeta, M, d, s);
    const int grid = ceil_div(M, 8);
    launch_kernel(layernorm_bwd_lean_kernel<MAXV>, dim3(grid), dim3(256), 0, s, dy, ld_dy, x, ld_x, row_index, gamma, mean, rstd, dx, ld_dx,
                                                         accumulate, dx_bf16, ld_dxb, M, d);
    KERNEL_CHECK();
    count_launch();
}
void layernorm_bwd(const bf16* dy, int ld_dy, const float* x, int ld_x, const int* row_index, const float* gamma,
                   const float* mean, const float* rstd, float* dx, int ld_dx, int accumulate, bf16* dx_bf16,
                   int ld_dxb, float* dgamma, float* dbeta, int M, int d, float eps, cudaStream_t s) {
    (void)eps;
    EAVQA_CHECK(d % 4 == 0 && d <= 2048, "layernorm width must be a multiple of 4 and <= 2048");
    EAVQA_CHECK(mean != nullptr && rstd != nullptr, "layernorm_bwd needs the saved statistics");
    EAVQA_CHECK((dgamma == nullptr) == (dbeta == nullptr), "dgamma and dbeta go together");
    const int need = ceil_div(d / 4, 32);
#define EAVQA_LN_BWD(V) launch_ln_bwd<V>(dy, ld_dy, x, ld_x, row_index, gamma, mean, rstd, dx, ld_dx, accumulate, dx_bf16, ld_dxb, dgamma, dbeta, M, d, s)
    if (need <= 2) EAVQA_LN_BWD(2);
    else if (need <= 4) EAVQA_LN_BWD(4);
    else if (need <= 6) EAVQA_LN_BWD(6);          // d = 768
    else if (need <= 8) EAVQA_LN_BWD(8);          // d = 1024
    else if (need <= 10) EAVQA_LN_BWD(10);        // d = 1280
    else if (need <= 13) EAVQA_LN_BWD(13);        // d = 1600
    else EAVQA_LN_BWD(16);
#undef EAVQA_LN_BWD
}

void layernorm_param_grads(const bf16* dy, int ld_dy, const float* x, int ld_x, const float* mean, const float* rstd,
                           float* dgamma, float* dbeta, int M, int d, cudaStream_t s) {
    EAVQA_CHECK(d % 8 == 0 && ld_dy % 8 == 0 && ld_x % 4 == 0, "layernorm_param_grads alignment");
    dim3 grid(ceil_div(d, 256), ceil_div(M, 64)), block(32, 8);
    launch_kernel(ln_param_grad_kernel, dim3(grid), dim3(block), 0, s, dy, ld_dy, x, ld_x, mean, rstd, dgamma, dbeta, M, d);
    KERNEL_CHECK();
    count_launch();
}

void prepend_plan(const int64_t* tokens, const int64_t* mask, int B, int Tt, int P, int* plan, int* valid, cudaStream_t s) {
    const int n = B * (P + Tt);
    launch_kernel(prepend_plan_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, s, tokens, mask, B, Tt, P, plan, valid);
    KERNEL_CHECK();
    count_launch();
}
void splice_plan(const int64_t* tokens, const int64_t* mask, int B, int Tt, int P, int n_img, int64_t sent_lo,
                 int64_t sent_hi, int* plan, int* valid, int* err_flag, cudaStream_t s) {
    const int T_out = Tt + (P - 1) * n_img;
    fill_zero(plan, sizeof(int) * static_cast<size_t>(B) * T_out, s);
    fill_zero(valid, sizeof(int) * static_cast<size_t>(B) * T_out, s);
    launch_kernel(splice_plan_kernel, dim3(ceil_div(B * 32, 128)), dim3(128), 0, s, tokens, mask, B, Tt, P, n_img, sent_lo, sent_hi, plan, valid, err_flag);
    KERNEL_CHECK();
    count_launch();
}
void embed_rows(const int* plan, int B, int T, int d, const float* wte, int vocab, const float* prefix,
                int64_t prefix_batch_stride, int prefix_row_stride, const float* wpe, float* out, cudaStream_t s) {
    EAVQA_CHECK(d % 4 == 0 && prefix_batch_stride % 4 == 0 && prefix_row_stride % 4 == 0, "embed alignment");
    const int rows = B * T;
    launch_kernel(embed_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, s, plan, rows, T, d, wte, vocab, prefix, prefix_batch_stride,
                                                        prefix_row_stride, wpe, out);
    KERNEL_CHECK();
    count_launch();
}

void ce_plan(const int64_t* labels, int B, int Tt, int T, int P, int vocab, int* row_index, int* label, int* n_valid,
             cudaStream_t s) {
    fill_zero(n_valid, sizeof(int), s);
    launch_kernel(ce_plan_kernel, dim3(ceil_div(B * Tt, 256)), dim3(256), 0, s, labels, B, Tt, T, P, vocab, row_index, label, n_valid);
    KERNEL_CHECK();
    count_launch();
}
void ce_finalize(const float2* partial, int tiles, const float* target, const int* label, float* lse, float* loss_sum,
                 int M, cudaStream_t s) {
    fill_zero(loss_sum, sizeof(float), s);
    launch_kernel(ce_finalize_kernel, dim3(ceil_div(M, 8)), dim3(256), 0, s, partial, tiles, target, label, lse, loss_sum, M);
    KERNEL_CHECK();
    count_launch();
}
void ce_loss(const float* loss_sum, const int* n_valid, float* loss_out, cudaStream_t s) {
    launch_kernel(ce_loss_kernel, dim3(1), dim3(1), 0, s, loss_sum, n_valid, loss_out);
    KERNEL_CHECK();
    count_launch();
}
void ce_dlogits(bf16* z, int ld, int M, int vocab, int n_cols, const float* lse, const int* label, const int* n_valid,
                cudaStream_t s) {
    EAVQA_CHECK(ld % 8 == 0 && n_cols % 8 == 0 && n_cols <= ld, "ce_dlogits alignment");
    launch_kernel(ce_dlogits_kernel, dim3(M), dim3(256), 0, s, z, ld, vocab, n_cols, lse, label, n_valid);
    KERNEL_CHECK();
    count_launch();
}

void greedy_step(const float* logits, int ld, int B, int vocab, int step, int max_new, int has_eos, int64_t pad_id,
                 int64_t eos_id, int* unfinished, int64_t* tokens_out, int* n_unfinished, float* top_logit, float* token_logprob,
                 const float* wte, const float* wpe_row, int d, float* x_next, int* valid_next, int valid_stride,
                 float* part_val, int* part_idx, unsigned* arrivals, cudaStream_t s) {
    // enough CTAs to cover the SMs ~4 times; scratch (part_*, arrivals) is sized for kGreedySplitMax slices per row
    int split = 1;
    if (part_val != nullptr && part_idx != nullptr && arrivals != nullptr)
        split = std::max(1, std::min(kGreedySplitMax, (4 * num_sms()) / std::max(B, 1)));
    launch_kernel(greedy_step_kernel, dim3(B, split), dim3(256), 0, s, logits, ld, vocab, step, max_new, has_eos, pad_id, eos_id, unfinished,
                  tokens_out, n_unfinished, top_logit, token_logprob, wte, wpe_row, d, x_next, valid_next, valid_stride, part_val,
                  part_idx, arrivals);
    KERNEL_CHECK();
    count_launch();
}

void caption_labels(const int64_t* tokens, int B, int T, int64_t pad_id, int64_t bos_id, int64_t* labels, cudaStream_t s) {
    launch_kernel(caption_labels_kernel, dim3(ceil_div(B, 128)), dim3(128), 0, s, tokens, B, T, pad_id, bos_id, labels);
    KERNEL_CHECK();
    count_launch();
}

void ensemble_select(const float* logprob, const int64_t* tokens, int E, int B, int S, const int64_t* skip, int n_skip,
                     float* scores, int* best, int64_t* best_tokens, cudaStream_t s) {
    launch_kernel(ensemble_select_kernel, dim3(ceil_div(B, 4)), dim3(128), 0, s, logprob, tokens, E, B, S, skip, n_skip, scores, best,
                  best_tokens);
    KERNEL_CHECK();
    count_launch();
}

void scale_by_device_scalar(float* x, int64_t n, const float* scale, cudaStream_t s) {
    EAVQA_CHECK(n % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "scale_by_device_scalar: 16-byte aligned buffer, n % 4 == 0");
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(n / 4, 256), static_cast<int64_t>(num_sms()) * 8));
    launch_kernel(scale_by_device_scalar_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<float4*>(x), n / 4, scale);
    KERNEL_CHECK();
    count_launch();
}

}  // namespace eavqa
