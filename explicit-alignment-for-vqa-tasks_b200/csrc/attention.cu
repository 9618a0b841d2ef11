// Attention kernels of the CLIP-prefix LM step.
//
//  * lm_attention_{fwd,bwd}: GPT-2 causal attention with a key-padding mask (head_dim 64), flash
//    style, 64x64 blocks, bf16 mma.sync.m16n8k16 with fp32 softmax statistics.  Sequences here are
//    short (T = 50 in training, ~130 in few-shot prefill), so attention is <1% of the step's FLOPs
//    (SURVEY.md 8d); it is latency/HBM-bound and written for that: one pass over q/k/v, nothing
//    materialised in HBM but O and the log-sum-exp.
//  * lm_attention_decode_acc: one query per (batch, head) against the head-major KV cache (HBM-bound streaming).
//  * mapper_attention_{fwd,bwd}: the mapper's 8-head, unmasked self-attention over S = 20 rows
//    (clipcap.py:81-104); fp32 CUDA-core math in shared memory.
#include <algorithm>
#include <atomic>

#include "kernels.cuh"

namespace eavqa {
namespace {

constexpr int HD = 64;        // GPT-2 head_dim
constexpr int LDS = 72;       // smem row stride in bf16 (64 + 8 pad: conflict-free ldmatrix)
constexpr int BLK = 64;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (16 rows x 16 k) of a row-major tile X[row][k]
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* tile, int row0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4(a, smem_addr(tile + (row0 + (mi & 1) * 8 + r) * LDS + k0 + (mi >> 1) * 8));
}
// A fragment of X^T where X is stored [k][m]: rows m0.., contraction k0..
__device__ __forceinline__ void load_a_trans(uint32_t (&a)[4], const bf16* tile, int m0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4_trans(a, smem_addr(tile + (k0 + (mi >> 1) * 8 + r) * LDS + m0 + (mi & 1) * 8));
}
// B fragments for two adjacent n-tiles (n0, n0+8) from X[n][k] (contraction contiguous): b[0..1] tile 0, b[2..3] tile 1
__device__ __forceinline__ void load_b(uint32_t (&b)[4], const bf16* tile, int n0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4(b, smem_addr(tile + (n0 + (mi >> 1) * 8 + r) * LDS + k0 + (mi & 1) * 8));
}
// B fragments for two adjacent n-tiles from X[k][n] (contraction = stored rows)
__device__ __forceinline__ void load_b_trans(uint32_t (&b)[4], const bf16* tile, int n0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4_trans(b, smem_addr(tile + (k0 + (mi & 1) * 8 + r) * LDS + n0 + (mi >> 1) * 8));
}

// copy a 64 x 64 bf16 tile (rows row0.. of a [*, ld] matrix, 64 columns from col0) into smem; rows >= nrows -> 0
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int64_t ld, int row0, int nrows, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + 128 * i;
        const int r = idx >> 3, c = (idx & 7) * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row0 + r < nrows) v = *reinterpret_cast<const uint4*>(src + static_cast<int64_t>(row0 + r) * ld + c);
        *reinterpret_cast<uint4*>(dst + r * LDS + c) = v;
    }
}

// same, through cp.async (no register staging: every 16-byte request of the tile is in flight at once)
__device__ __forceinline__ void load_tile_async(bf16* dst, const bf16* src, int64_t ld, int row0, int nrows, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + 128 * i;
        const int r = idx >> 3, c = (idx & 7) * 8;
        bf16* d = dst + r * LDS + c;
        if (row0 + r < nrows) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(d)),
                         "l"(src + static_cast<int64_t>(row0 + r) * ld + c) : "memory");
        } else {
            *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// S(16 x 64 per warp) = A(16 x 64, fragments af) * X[n][k]^T
__device__ __forceinline__ void warp_gemm_nt(float (&acc)[8][4], const uint32_t (&af)[4][4], const bf16* xtile, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            load_b(b, xtile, np * 16, ks * 16, lane);
            mma_bf16(acc[2 * np], af[ks], b[0], b[1]);
            mma_bf16(acc[2 * np + 1], af[ks], b[2], b[3]);
        }
}

// ------------------------------------------------------------------------------------------ forward
// kv_cache != null (generation prefill): the CTA of the LAST query block walks every key block anyway and copies the K / V
// tiles it has in shared memory into the head-major KV cache (K block [B, H, Tmax, 64], then V) -- no separate fill kernel.
template <bool WRITE_KV>
__global__ void __launch_bounds__(128, 4) lm_attention_fwd_kernel(const bf16* __restrict__ qkv, const int* __restrict__ valid,
                                                               bf16* __restrict__ o, float* __restrict__ lse, int T,
                                                               int H, bf16* __restrict__ kv_cache, int Tmax) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) bf16 Qs[BLK * LDS];
    __shared__ __align__(16) bf16 Ks[BLK * LDS];
    __shared__ __align__(16) bf16 Vs[BLK * LDS];
    __shared__ int kvalid[BLK];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * HD;
    const int64_t ld = 3 * d;
    const bf16* base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
    const int q0 = qb * BLK;
    const float scale = 0.125f;     // head_dim ** -0.5  (HF modeling_gpt2.py:96-98)

    // Q and the first K/V block are requested together (for T <= 64 that is everything the CTA ever loads)
    load_tile_async(Qs, base, ld, q0, T, tid);
    load_tile_async(Ks, base + d, ld, 0, T, tid);
    load_tile_async(Vs, base + 2 * d, ld, 0, T, tid);
    if (tid < BLK) kvalid[tid] = (tid < T) ? valid[b * T + tid] : 0;
    cp_async_wait_all();
    __syncthreads();
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) load_a(qf[ks], Qs, warp * 16, ks * 16, lane);

    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
    float oacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;

    for (int kb = 0; kb <= qb; ++kb) {
        const int k0 = kb * BLK;
        if (kb > 0) {
            __syncthreads();
            load_tile_async(Ks, base + d, ld, k0, T, tid);
            load_tile_async(Vs, base + 2 * d, ld, k0, T, tid);
            if (tid < BLK) kvalid[tid] = (k0 + tid < T) ? valid[b * T + k0 + tid] : 0;
            cp_async_wait_all();
            __syncthreads();
        }

        if (WRITE_KV && qb == static_cast<int>(gridDim.x) - 1) {
            const int nb = gridDim.z;
            bf16* kc = kv_cache + (static_cast<int64_t>(b) * H + h) * Tmax * HD;
            bf16* vc = kc + static_cast<int64_t>(nb) * H * Tmax * HD;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = tid + 128 * i;
                const int r = idx >> 3, c = (idx & 7) * 8;
                if (k0 + r < T) {
                    *reinterpret_cast<uint4*>(kc + static_cast<int64_t>(k0 + r) * HD + c) = *reinterpret_cast<const uint4*>(Ks + r * LDS + c);
                    *reinterpret_cast<uint4*>(vc + static_cast<int64_t>(k0 + r) * HD + c) = *reinterpret_cast<const uint4*>(Vs + r * LDS + c);
                }
            }
        }
        float sacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[i][e] = 0.f;
        warp_gemm_nt(sacc, qf, Ks, lane);

        float rmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nt * 8 + 2 * t4 + (e & 1);
                const int row = warp * 16 + g + ((e >> 1) << 3);
                const bool ok = kvalid[col] && (k0 + col <= q0 + row);
                const float s = ok ? sacc[nt][e] * scale : -INFINITY;
                sacc[nt][e] = s;
                rmax[e >> 1] = fmaxf(rmax[e >> 1], s);
            }
        float alpha[2], muse[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 1));
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 2));
            const float mn = fmaxf(m_i[r], rmax[r]);
            muse[r] = (mn == -INFINITY) ? 0.f : mn;
            alpha[r] = __expf(m_i[r] - muse[r]);      // m_i = -inf -> 0
            m_i[r] = mn;
            l_i[r] *= alpha[r];
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = __expf(sacc[nt][0] - muse[0]), p1 = __expf(sacc[nt][1] - muse[0]);
            const float p2 = __expf(sacc[nt][2] - muse[1]), p3 = __expf(sacc[nt][3] - muse[1]);
            l_i[0] += p0 + p1;
            l_i[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            oacc[i][0] *= alpha[0]; oacc[i][1] *= alpha[0];
            oacc[i][2] *= alpha[1]; oacc[i][3] *= alpha[1];
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                load_b_trans(bfr, Vs, np * 16, ks * 16, lane);
                mma_bf16(oacc[2 * np], pf[ks], bfr[0], bfr[1]);
                mma_bf16(oacc[2 * np + 1], pf[ks], bfr[2], bfr[3]);
            }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qpos = q0 + warp * 16 + g + r * 8;
        if (qpos >= T) continue;
        const float inv = l_i[r] > 0.f ? 1.0f / l_i[r] : 0.f;
        bf16* op = o + (static_cast<int64_t>(b) * T + qpos) * d + h * HD;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(op + nt * 8 + 2 * t4) = pack_bf16x2(oacc[nt][2 * r] * inv, oacc[nt][2 * r + 1] * inv);
        if (lse != nullptr && t4 == 0)
            lse[(static_cast<int64_t>(b) * H + h) * T + qpos] = (l_i[r] > 0.f) ? m_i[r] + logf(l_i[r]) : INFINITY;
    }
}

// Generation prefill with up to three key blocks (T <= 192: the few-shot prompts, T0 ~ 125-145): ONE CTA per (sample, head)
// keeps the head's K and V resident in shared memory and walks its query blocks, so K / V are read once (the general kernel
// above runs one CTA per query block and re-reads the key blocks below the diagonal: 6 block pairs instead of 3 at T = 143;
// round-2 launch list: 88 us per layer, 19 % of the prefill), the next query tile streams in while the current one is
// reduced, the KV cache is filled from the resident tiles and O leaves as 16-byte row-contiguous stores.
constexpr int kPrefillMaxBlocks = 3;
__global__ void __launch_bounds__(128, 3) lm_attention_prefill_kernel(const bf16* __restrict__ qkv, const int* __restrict__ valid,
                                                                   bf16* __restrict__ o, int T, int H, bf16* __restrict__ kv_cache,
                                                                   int Tmax) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t pre_smem[];
    const int nblk = (T + BLK - 1) / BLK;
    bf16* Ks = reinterpret_cast<bf16*>(pre_smem);                       // [nblk * 64][LDS]
    bf16* Vs = Ks + nblk * BLK * LDS;
    bf16* Qs0 = Vs + nblk * BLK * LDS;                                  // two query tiles
    int* kvalid = reinterpret_cast<int*>(Qs0 + 2 * BLK * LDS);          // [nblk * 64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int h = blockIdx.x, b = blockIdx.y;
    const int d = H * HD;
    const int64_t ld = 3 * d;
    const bf16* base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
    const float scale = 0.125f;

    for (int kb = 0; kb < nblk; ++kb) {
        load_tile_async(Ks + kb * BLK * LDS, base + d, ld, kb * BLK, T, tid);
        load_tile_async(Vs + kb * BLK * LDS, base + 2 * d, ld, kb * BLK, T, tid);
    }
    for (int t = tid; t < nblk * BLK; t += 128) {
        if (t < T) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(kvalid + t)), "l"(valid + b * T + t) : "memory");
        else kvalid[t] = 0;
    }
    load_tile_async(Qs0, base, ld, 0, T, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int qb = 0; qb < nblk; ++qb) {
        bf16* Qs = Qs0 + (qb & 1) * BLK * LDS;
        const int q0 = qb * BLK;
        cp_async_wait_all();
        __syncthreads();                          // tile qb (and, first time, K / V / validity) visible; the other q tile is free
        if (qb + 1 < nblk) {
            load_tile_async(Qs0 + ((qb + 1) & 1) * BLK * LDS, base, ld, q0 + BLK, T, tid);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (qb == 0 && kv_cache != nullptr) {     // fill the head-major KV cache from the resident tiles
            const int nb = gridDim.y;
            bf16* kc = kv_cache + (static_cast<int64_t>(b) * H + h) * Tmax * HD;
            bf16* vc = kc + static_cast<int64_t>(nb) * H * Tmax * HD;
            for (int idx = tid; idx < T * 8; idx += 128) {
                const int r = idx >> 3, c = (idx & 7) * 8;
                *reinterpret_cast<uint4*>(kc + static_cast<int64_t>(r) * HD + c) = *reinterpret_cast<const uint4*>(Ks + r * LDS + c);
                *reinterpret_cast<uint4*>(vc + static_cast<int64_t>(r) * HD + c) = *reinterpret_cast<const uint4*>(Vs + r * LDS + c);
            }
        }
        uint32_t qf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) load_a(qf[ks], Qs, warp * 16, ks * 16, lane);
        float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
        float oacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
        for (int kb = 0; kb <= qb; ++kb) {
            const int k0 = kb * BLK;
            const bf16* Kt = Ks + kb * BLK * LDS;
            const bf16* Vt = Vs + kb * BLK * LDS;
            float sacc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) sacc[i][e] = 0.f;
            warp_gemm_nt(sacc, qf, Kt, lane);
            float rmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = nt * 8 + 2 * t4 + (e & 1);
                    const int row = warp * 16 + g + ((e >> 1) << 3);
                    const bool ok = kvalid[k0 + col] && (k0 + col <= q0 + row);
                    const float sv = ok ? sacc[nt][e] * scale : -INFINITY;
                    sacc[nt][e] = sv;
                    rmax[e >> 1] = fmaxf(rmax[e >> 1], sv);
                }
            float alpha[2], muse[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 1));
                rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 2));
                const float mn = fmaxf(m_i[r], rmax[r]);
                muse[r] = (mn == -INFINITY) ? 0.f : mn;
                alpha[r] = __expf(m_i[r] - muse[r]);
                m_i[r] = mn;
                l_i[r] *= alpha[r];
            }
            uint32_t pf[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float p0 = __expf(sacc[nt][0] - muse[0]), p1 = __expf(sacc[nt][1] - muse[0]);
                const float p2 = __expf(sacc[nt][2] - muse[1]), p3 = __expf(sacc[nt][3] - muse[1]);
                l_i[0] += p0 + p1;
                l_i[1] += p2 + p3;
                pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
                pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                oacc[i][0] *= alpha[0]; oacc[i][1] *= alpha[0];
                oacc[i][2] *= alpha[1]; oacc[i][3] *= alpha[1];
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t bfr[4];
                    load_b_trans(bfr, Vt, np * 16, ks * 16, lane);
                    mma_bf16(oacc[2 * np], pf[ks], bfr[0], bfr[1]);
                    mma_bf16(oacc[2 * np + 1], pf[ks], bfr[2], bfr[3]);
                }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
            l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
        }
        // O through this warp's own 16 rows of the query tile, then 16-byte row-contiguous stores
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = warp * 16 + g + r * 8;
            const float inv = l_i[r] > 0.f ? 1.0f / l_i[r] : 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<uint32_t*>(Qs + row * LDS + nt * 8 + 2 * t4) = pack_bf16x2(oacc[nt][2 * r] * inv, oacc[nt][2 * r + 1] * inv);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = warp * 16 + (lane >> 3) + 4 * i, c = (lane & 7) * 8;
            if (q0 + row < T)
                *reinterpret_cast<uint4*>(o + (static_cast<int64_t>(b) * T + q0 + row) * d + h * HD + c) = *reinterpret_cast<const uint4*>(Qs + row * LDS + c);
        }
    }
}

// Sequences of at most one block (T <= 64: the training step, T = 50): PERSISTENT CTAs walk the (sample, head) items with
// the next item's q / k / v tiles and validity words already streaming into a second buffer set while the current item is
// reduced, and O leaves through shared memory as 16-byte coalesced stores.  Round-2 reason: the one-CTA-per-item kernel above
// loads, then computes, then stores -- every CTA of a wave in lock-step, so DRAM idles during the math (2.7 TB/s of 6.5).
constexpr int kFwdTile = BLK * LDS;                                 // bf16 elements of one staged tile
constexpr int kFwdSetBytes = 3 * kFwdTile * 2 + BLK * 4;            // q, k, v tiles + validity words
__global__ void __launch_bounds__(128, 4) lm_attention_fwd_single_kernel(const bf16* __restrict__ qkv, const int* __restrict__ valid,
                                                                      bf16* __restrict__ o, float* __restrict__ lse, int T,
                                                                      int H, int n_items) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t fwd_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int d = H * HD;
    const int64_t ld = 3 * d;
    const float scale = 0.125f;     // head_dim ** -0.5  (HF modeling_gpt2.py:96-98)

    auto issue = [&](int item, int set) {
        const int b = item / H, h = item - b * H;
        bf16* Qs = reinterpret_cast<bf16*>(fwd_smem + set * kFwdSetBytes);
        int* kv = reinterpret_cast<int*>(Qs + 3 * kFwdTile);
        const bf16* base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
        load_tile_async(Qs, base, ld, 0, T, tid);
        load_tile_async(Qs + kFwdTile, base + d, ld, 0, T, tid);
        load_tile_async(Qs + 2 * kFwdTile, base + 2 * d, ld, 0, T, tid);
        if (tid < BLK) {
            if (tid < T) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(kv + tid)), "l"(valid + b * T + tid) : "memory");
            else kv[tid] = 0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int item = blockIdx.x;
    if (item < n_items) issue(item, 0);
    for (int it = 0; item < n_items; item += gridDim.x, ++it) {
        const int set = it & 1;
        cp_async_wait_all();
        __syncthreads();                          // this item's tiles are visible; every warp is done with the other set
        if (item + static_cast<int>(gridDim.x) < n_items) issue(item + gridDim.x, set ^ 1);
        bf16* Qs = reinterpret_cast<bf16*>(fwd_smem + set * kFwdSetBytes);
        const bf16* Ks = Qs + kFwdTile;
        const bf16* Vs = Qs + 2 * kFwdTile;
        const int* kvalid = reinterpret_cast<const int*>(Qs + 3 * kFwdTile);
        const int b = item / H, h = item - b * H;

        uint32_t qf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) load_a(qf[ks], Qs, warp * 16, ks * 16, lane);
        float sacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[i][e] = 0.f;
        warp_gemm_nt(sacc, qf, Ks, lane);
        float rmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nt * 8 + 2 * t4 + (e & 1);
                const int row = warp * 16 + g + ((e >> 1) << 3);
                const bool ok = kvalid[col] && (col <= row);
                const float sv = ok ? sacc[nt][e] * scale : -INFINITY;
                sacc[nt][e] = sv;
                rmax[e >> 1] = fmaxf(rmax[e >> 1], sv);
            }
        float muse[2], l_i[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 1));
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 2));
            muse[r] = (rmax[r] == -INFINITY) ? 0.f : rmax[r];
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = __expf(sacc[nt][0] - muse[0]), p1 = __expf(sacc[nt][1] - muse[0]);
            const float p2 = __expf(sacc[nt][2] - muse[1]), p3 = __expf(sacc[nt][3] - muse[1]);
            l_i[0] += p0 + p1;
            l_i[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        float oacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                load_b_trans(bfr, Vs, np * 16, ks * 16, lane);
                mma_bf16(oacc[2 * np], pf[ks], bfr[0], bfr[1]);
                mma_bf16(oacc[2 * np + 1], pf[ks], bfr[2], bfr[3]);
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
            l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
        }
        // O through this warp's own 16 rows of the q tile (only this warp ever reads them), then 16-byte row-contiguous stores
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = warp * 16 + g + r * 8;
            const float inv = l_i[r] > 0.f ? 1.0f / l_i[r] : 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<uint32_t*>(Qs + row * LDS + nt * 8 + 2 * t4) = pack_bf16x2(oacc[nt][2 * r] * inv, oacc[nt][2 * r + 1] * inv);
            if (lse != nullptr && t4 == 0 && row < T)
                lse[(static_cast<int64_t>(b) * H + h) * T + row] = (l_i[r] > 0.f) ? rmax[r] + logf(l_i[r]) : INFINITY;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = warp * 16 + (lane >> 3) + 4 * i, c = (lane & 7) * 8;
            if (row < T)
                *reinterpret_cast<uint4*>(o + (static_cast<int64_t>(b) * T + row) * d + h * HD + c) = *reinterpret_cast<const uint4*>(Qs + row * LDS + c);
        }
    }
}

// ------------------------------------------------------------------------------------------ backward
// One CTA per (batch, head); outer loop over key blocks j, inner over query blocks i >= j.
__global__ void __launch_bounds__(128) lm_attention_bwd_kernel(const bf16* __restrict__ qkv, const int* __restrict__ valid,
                                                               const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                               const float* __restrict__ lse, bf16* __restrict__ dqkv,
                                                               float* __restrict__ dq_scratch, int T, int H) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t smem_bwd[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_bwd);
    bf16* Ks = Qs + BLK * LDS;
    bf16* Vs = Ks + BLK * LDS;
    bf16* dOs = Vs + BLK * LDS;
    bf16* Ps = dOs + BLK * LDS;
    bf16* dSs = Ps + BLK * LDS;
    float* Ds = reinterpret_cast<float*>(dSs + BLK * LDS);
    float* Ls = Ds + BLK;
    int* kvalid = reinterpret_cast<int*>(Ls + BLK);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int h = blockIdx.x, b = blockIdx.y;
    const int d = H * HD;
    const int64_t ld = 3 * d;
    const bf16* base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
    const bf16* obase = o + static_cast<int64_t>(b) * T * d + h * HD;
    const bf16* dobase = d_o + static_cast<int64_t>(b) * T * d + h * HD;
    bf16* dbase = dqkv + static_cast<int64_t>(b) * T * ld + h * HD;
    float* dqs = dq_scratch ? dq_scratch + static_cast<int64_t>(b) * T * d + h * HD : nullptr;
    const float scale = 0.125f;
    const int nblk = (T + BLK - 1) / BLK;

    for (int j = 0; j < nblk; ++j) {
        const int k0 = j * BLK;
        __syncthreads();
        load_tile(Ks, base + d, ld, k0, T, tid);
        load_tile(Vs, base + 2 * d, ld, k0, T, tid);
        if (tid < BLK) kvalid[tid] = (k0 + tid < T) ? valid[b * T + k0 + tid] : 0;
        float dkacc[8][4], dvacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dkacc[i][e] = 0.f; dvacc[i][e] = 0.f; }

        for (int i = j; i < nblk; ++i) {
            const int q0 = i * BLK;
            __syncthreads();       // previous iteration's readers of Qs / dOs / Ps / dSs are done
            load_tile(Qs, base, ld, q0, T, tid);
            load_tile(dOs, dobase, d, q0, T, tid);
            {   // D = rowsum(dO * O), lse: two threads per query row
                const int r = tid >> 1, half = tid & 1;
                float acc = 0.f;
                if (q0 + r < T) {
                    const uint4* op = reinterpret_cast<const uint4*>(obase + static_cast<int64_t>(q0 + r) * d + half * 32);
                    const uint4* dp = reinterpret_cast<const uint4*>(dobase + static_cast<int64_t>(q0 + r) * d + half * 32);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 a = op[c], bb = dp[c];
                        float2 x, y;
                        x = unpack_bf16x2(a.x); y = unpack_bf16x2(bb.x); acc += x.x * y.x + x.y * y.y;
                        x = unpack_bf16x2(a.y); y = unpack_bf16x2(bb.y); acc += x.x * y.x + x.y * y.y;
                        x = unpack_bf16x2(a.z); y = unpack_bf16x2(bb.z); acc += x.x * y.x + x.y * y.y;
                        x = unpack_bf16x2(a.w); y = unpack_bf16x2(bb.w); acc += x.x * y.x + x.y * y.y;
                    }
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                if (half == 0) {
                    Ds[r] = acc;
                    Ls[r] = (q0 + r < T) ? lse[(static_cast<int64_t>(b) * H + h) * T + q0 + r] : 0.f;
                }
            }
            __syncthreads();

            uint32_t af[4][4];
            float sacc[8][4], dpacc[8][4];
#pragma unroll
            for (int x = 0; x < 8; ++x)
#pragma unroll
                for (int e = 0; e < 4; ++e) { sacc[x][e] = 0.f; dpacc[x][e] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) load_a(af[ks], Qs, warp * 16, ks * 16, lane);
            warp_gemm_nt(sacc, af, Ks, lane);                      // S = Q K^T
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) load_a(af[ks], dOs, warp * 16, ks * 16, lane);
            warp_gemm_nt(dpacc, af, Vs, lane);                     // dP = dO V^T

#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float p[4], ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = nt * 8 + 2 * t4 + (e & 1);
                    const int row = warp * 16 + g + ((e >> 1) << 3);
                    const bool ok = kvalid[col] && (k0 + col <= q0 + row) && (q0 + row < T);
                    p[e] = ok ? __expf(sacc[nt][e] * scale - Ls[row]) : 0.f;
                    ds[e] = p[e] * (dpacc[nt][e] - Ds[row]);
                }
                const int r0 = warp * 16 + g, c0 = nt * 8 + 2 * t4;
                *reinterpret_cast<uint32_t*>(Ps + r0 * LDS + c0) = pack_bf16x2(p[0], p[1]);
                *reinterpret_cast<uint32_t*>(Ps + (r0 + 8) * LDS + c0) = pack_bf16x2(p[2], p[3]);
                *reinterpret_cast<uint32_t*>(dSs + r0 * LDS + c0) = pack_bf16x2(ds[0], ds[1]);
                *reinterpret_cast<uint32_t*>(dSs + (r0 + 8) * LDS + c0) = pack_bf16x2(ds[2], ds[3]);
            }
            __syncthreads();

            // dV(keys 16w.. x hd) += P^T dO ;  dK += dS^T Q   (contraction over the 64 queries)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t pa[4], sa[4];
                load_a_trans(pa, Ps, warp * 16, ks * 16, lane);
                load_a_trans(sa, dSs, warp * 16, ks * 16, lane);
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t bfr[4];
                    load_b_trans(bfr, dOs, np * 16, ks * 16, lane);
                    mma_bf16(dvacc[2 * np], pa, bfr[0], bfr[1]);
                    mma_bf16(dvacc[2 * np + 1], pa, bfr[2], bfr[3]);
                    load_b_trans(bfr, Qs, np * 16, ks * 16, lane);
                    mma_bf16(dkacc[2 * np], sa, bfr[0], bfr[1]);
                    mma_bf16(dkacc[2 * np + 1], sa, bfr[2], bfr[3]);
                }
            }
            // dQ(queries 16w.. x hd) = dS K   (contraction over the 64 keys)
            float dqacc[8][4];
#pragma unroll
            for (int x = 0; x < 8; ++x)
#pragma unroll
                for (int e = 0; e < 4; ++e) dqacc[x][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t sa[4];
                load_a(sa, dSs, warp * 16, ks * 16, lane);
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t bfr[4];
                    load_b_trans(bfr, Ks, np * 16, ks * 16, lane);
                    mma_bf16(dqacc[2 * np], sa, bfr[0], bfr[1]);
                    mma_bf16(dqacc[2 * np + 1], sa, bfr[2], bfr[3]);
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int qpos = q0 + warp * 16 + g + r * 8;
                if (qpos >= T) continue;
                if (nblk == 1) {
                    bf16* p = dbase + static_cast<int64_t>(qpos) * ld;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
                        *reinterpret_cast<uint32_t*>(p + nt * 8 + 2 * t4) =
                            pack_bf16x2(dqacc[nt][2 * r] * scale, dqacc[nt][2 * r + 1] * scale);
                } else {
                    float* p = dqs + static_cast<int64_t>(qpos) * d;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        float2* q2 = reinterpret_cast<float2*>(p + nt * 8 + 2 * t4);
                        float2 v = make_float2(dqacc[nt][2 * r] * scale, dqacc[nt][2 * r + 1] * scale);
                        if (j > 0) {
                            const float2 old = *q2;
                            v.x += old.x;
                            v.y += old.y;
                        }
                        *q2 = v;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int kpos = k0 + warp * 16 + g + r * 8;
            if (kpos >= T) continue;
            bf16* pk = dbase + static_cast<int64_t>(kpos) * ld + d;
            bf16* pv = dbase + static_cast<int64_t>(kpos) * ld + 2 * d;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                *reinterpret_cast<uint32_t*>(pk + nt * 8 + 2 * t4) = pack_bf16x2(dkacc[nt][2 * r] * scale, dkacc[nt][2 * r + 1] * scale);
                *reinterpret_cast<uint32_t*>(pv + nt * 8 + 2 * t4) = pack_bf16x2(dvacc[nt][2 * r], dvacc[nt][2 * r + 1]);
            }
        }
    }
    if (nblk > 1) {
        __syncthreads();
        for (int idx = tid; idx < T * (HD / 2); idx += blockDim.x) {
            const int r = idx / (HD / 2), c = (idx % (HD / 2)) * 2;
            const float2 v = *reinterpret_cast<const float2*>(dqs + static_cast<int64_t>(r) * d + c);
            *reinterpret_cast<uint32_t*>(dbase + static_cast<int64_t>(r) * ld + c) = pack_bf16x2(v.x, v.y);
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, T <= 64
// The training shapes (T = 50) fit one 64 x 64 block: no loops, no dQ scratch, and the three gradient GEMMs run
// one after another so that only one 32-register accumulator is live at a time (round-1 ncu: the general kernel
// needs 252 registers -> 2 CTAs / SM, 12 % warp occupancy, 84 us; this one is capped at 128 -> 4 CTAs / SM).
__global__ void __launch_bounds__(128, 4) lm_attention_bwd_single_kernel(const bf16* __restrict__ qkv, const int* __restrict__ valid,
                                                                         const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                                         const float* __restrict__ lse, bf16* __restrict__ dqkv,
                                                                         int T, int H) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t smem_bwd[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_bwd);
    bf16* Ks = Qs + BLK * LDS;
    bf16* Vs = Ks + BLK * LDS;
    bf16* dOs = Vs + BLK * LDS;
    bf16* Ps = dOs + BLK * LDS;
    bf16* dSs = Ps + BLK * LDS;
    float* Ds = reinterpret_cast<float*>(dSs + BLK * LDS);
    float* Ls = Ds + BLK;
    int* kvalid = reinterpret_cast<int*>(Ls + BLK);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int h = blockIdx.x, b = blockIdx.y;
    const int d = H * HD;
    const int64_t ld = 3 * d;
    const bf16* base = qkv + static_cast<int64_t>(b) * T * ld + h * HD;
    const bf16* obase = o + static_cast<int64_t>(b) * T * d + h * HD;
    const bf16* dobase = d_o + static_cast<int64_t>(b) * T * d + h * HD;
    bf16* dbase = dqkv + static_cast<int64_t>(b) * T * ld + h * HD;
    const float scale = 0.125f;

    load_tile_async(Qs, base, ld, 0, T, tid);
    load_tile_async(Ks, base + d, ld, 0, T, tid);
    load_tile_async(Vs, base + 2 * d, ld, 0, T, tid);
    load_tile_async(dOs, dobase, d, 0, T, tid);
    load_tile_async(Ps, obase, d, 0, T, tid);          // O, only needed for D; Ps is overwritten with P afterwards
    if (tid < BLK) kvalid[tid] = (tid < T) ? valid[b * T + tid] : 0;
    cp_async_wait_all();
    __syncthreads();
    {   // D = rowsum(dO * O), lse: two threads per query row
        const int r = tid >> 1, half = tid & 1;
        float acc = 0.f;
        const uint4* op = reinterpret_cast<const uint4*>(Ps + r * LDS + half * 32);
        const uint4* dp = reinterpret_cast<const uint4*>(dOs + r * LDS + half * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 a = op[c], bb = dp[c];
            float2 x, y;
            x = unpack_bf16x2(a.x); y = unpack_bf16x2(bb.x); acc += x.x * y.x + x.y * y.y;
            x = unpack_bf16x2(a.y); y = unpack_bf16x2(bb.y); acc += x.x * y.x + x.y * y.y;
            x = unpack_bf16x2(a.z); y = unpack_bf16x2(bb.z); acc += x.x * y.x + x.y * y.y;
            x = unpack_bf16x2(a.w); y = unpack_bf16x2(bb.w); acc += x.x * y.x + x.y * y.y;
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (half == 0) {
            Ds[r] = acc;
            Ls[r] = (r < T) ? lse[(static_cast<int64_t>(b) * H + h) * T + r] : 0.f;
        }
    }
    __syncthreads();
    {
        uint32_t af[4][4];
        float sacc[8][4], dpacc[8][4];
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int e = 0; e < 4; ++e) { sacc[x][e] = 0.f; dpacc[x][e] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) load_a(af[ks], Qs, warp * 16, ks * 16, lane);
        warp_gemm_nt(sacc, af, Ks, lane);                      // S = Q K^T
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) load_a(af[ks], dOs, warp * 16, ks * 16, lane);
        warp_gemm_nt(dpacc, af, Vs, lane);                     // dP = dO V^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float p[4], ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nt * 8 + 2 * t4 + (e & 1);
                const int row = warp * 16 + g + ((e >> 1) << 3);
                const bool ok = kvalid[col] && (col <= row) && (row < T);
                p[e] = ok ? __expf(sacc[nt][e] * scale - Ls[row]) : 0.f;
                ds[e] = p[e] * (dpacc[nt][e] - Ds[row]);
            }
            const int r0 = warp * 16 + g, c0 = nt * 8 + 2 * t4;
            *reinterpret_cast<uint32_t*>(Ps + r0 * LDS + c0) = pack_bf16x2(p[0], p[1]);
            *reinterpret_cast<uint32_t*>(Ps + (r0 + 8) * LDS + c0) = pack_bf16x2(p[2], p[3]);
            *reinterpret_cast<uint32_t*>(dSs + r0 * LDS + c0) = pack_bf16x2(ds[0], ds[1]);
            *reinterpret_cast<uint32_t*>(dSs + (r0 + 8) * LDS + c0) = pack_bf16x2(ds[2], ds[3]);
        }
    }
    __syncthreads();
    // one gradient at a time: acc(16 rows x 64) = A^T-or-A (from Ps / dSs) * B (from dOs / Qs / Ks)
    // Gradients leave through the warp's own 16 rows of the V tile (nobody reads V after the dP product above) as 16-byte,
    // row-contiguous stores: from the accumulator layout every store instruction would touch 8 rows with 16 bytes each.
    auto store_rows = [&](const float (&acc)[8][4], int col_off, float mul) {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = warp * 16 + g + r * 8;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<uint32_t*>(Vs + row * LDS + nt * 8 + 2 * t4) = pack_bf16x2(acc[nt][2 * r] * mul, acc[nt][2 * r + 1] * mul);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = warp * 16 + (lane >> 3) + 4 * i, c = (lane & 7) * 8;
            if (row < T)
                *reinterpret_cast<uint4*>(dbase + static_cast<int64_t>(row) * ld + col_off + c) = *reinterpret_cast<const uint4*>(Vs + row * LDS + c);
        }
    };
    {   // dV = P^T dO
        float acc[8][4];
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[x][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4];
            load_a_trans(a, Ps, warp * 16, ks * 16, lane);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                load_b_trans(bfr, dOs, np * 16, ks * 16, lane);
                mma_bf16(acc[2 * np], a, bfr[0], bfr[1]);
                mma_bf16(acc[2 * np + 1], a, bfr[2], bfr[3]);
            }
        }
        store_rows(acc, 2 * d, 1.0f);
    }
    {   // dK = dS^T Q
        float acc[8][4];
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[x][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4];
            load_a_trans(a, dSs, warp * 16, ks * 16, lane);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                load_b_trans(bfr, Qs, np * 16, ks * 16, lane);
                mma_bf16(acc[2 * np], a, bfr[0], bfr[1]);
                mma_bf16(acc[2 * np + 1], a, bfr[2], bfr[3]);
            }
        }
        store_rows(acc, d, scale);
    }
    {   // dQ = dS K
        float acc[8][4];
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[x][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4];
            load_a(a, dSs, warp * 16, ks * 16, lane);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bfr[4];
                load_b_trans(bfr, Ks, np * 16, ks * 16, lane);
                mma_bf16(acc[2 * np], a, bfr[0], bfr[1]);
                mma_bf16(acc[2 * np + 1], a, bfr[2], bfr[3]);
            }
        }
        store_rows(acc, 0, scale);
    }
}

// ------------------------------------------------------------------------------------------ KV cache / decode
// KV cache of one layer: K block [B, H, Tmax, 64] followed by the V block of the same shape (bf16).  Head-major: the
// keys / values one (sample, head) attends over are contiguous (Tmax * 128 B), so a decode step streams them with
// fully coalesced loads.  Prefill fills it from inside lm_attention_fwd_kernel; each decode step appends one position.

// Split-K decode path: q | k | v arrive as fp32 GEMM accumulators + bias.  One 128-thread CTA per (sample, head); its four
// warps each take a quarter of the keys (flash-decoding style: per-warp max / sum / partial P.V, merged through shared
// memory), which quadruples the loads in flight over the one-warp-per-head kernel above (24.6 us per layer for 78 MB).
__global__ void __launch_bounds__(128) lm_attention_decode_acc_kernel(const float* __restrict__ qkv_acc, const float* __restrict__ qkv_bias,
                                                                      bf16* __restrict__ cache, const int* __restrict__ valid,
                                                                      int valid_stride, bf16* __restrict__ o, float* __restrict__ zero,
                                                                      int H, int pos, int Tmax) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sc[];     // [Tmax] scores / probabilities
    __shared__ __align__(16) float sq[HD];
    __shared__ float part[4][HD];
    __shared__ float part_m[4], part_l[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int d = H * HD;
    const float* arow = qkv_acc + static_cast<int64_t>(b) * 3 * d;
    const int B = gridDim.x / H;
    bf16* kbase = cache + (static_cast<int64_t>(b) * H + h) * Tmax * HD;                  // this head's keys [Tmax, 64]
    bf16* vhead = kbase + static_cast<int64_t>(B) * H * Tmax * HD;                        // ... and values
    if (tid < HD) {
        const int c = h * HD + tid;
        sq[tid] = (arow[c] + __ldg(qkv_bias + c)) * 0.125f;                              // head_dim ** -0.5 folded into q
        kbase[static_cast<int64_t>(pos) * HD + tid] = __float2bfloat16(arow[d + c] + __ldg(qkv_bias + d + c));
        if (zero != nullptr) zero[static_cast<int64_t>(b) * d + c] = 0.f;
    } else {
        const int c = h * HD + tid - HD;
        vhead[static_cast<int64_t>(pos) * HD + tid - HD] = __float2bfloat16(arow[2 * d + c] + __ldg(qkv_bias + 2 * d + c));
    }
    __syncthreads();
    // q stays in shared memory (broadcast LDS.128 in the dot products): 64 fewer registers per thread -> twice the CTAs
    // per SM, and this kernel is bound by (load latency x waves of CTAs), not by instruction issue
    const int n = pos + 1;
    const int chunk = (n + 3) >> 2;
    const int t0 = warp * chunk, t1 = min(n, t0 + chunk);
    float mx = -INFINITY;
    for (int t = t0 + lane; t < t1; t += 32) {
        float acc = -INFINITY;
        if (valid[static_cast<int64_t>(b) * valid_stride + t]) {
            const uint4* kp = reinterpret_cast<const uint4*>(kbase + static_cast<int64_t>(t) * HD);
            uint4 u[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) u[c] = kp[c];
            acc = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 qa = *reinterpret_cast<const float4*>(sq + c * 8), qb = *reinterpret_cast<const float4*>(sq + c * 8 + 4);
                float2 f;
                f = unpack_bf16x2(u[c].x); acc += qa.x * f.x + qa.y * f.y;
                f = unpack_bf16x2(u[c].y); acc += qa.z * f.x + qa.w * f.y;
                f = unpack_bf16x2(u[c].z); acc += qb.x * f.x + qb.y * f.y;
                f = unpack_bf16x2(u[c].w); acc += qb.z * f.x + qb.w * f.y;
            }
        }
        sc[t] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    const float muse = (mx == -INFINITY) ? 0.f : mx;
    float sum = 0.f;
    for (int t = t0 + lane; t < t1; t += 32) {
        const float p = __expf(sc[t] - muse);
        sc[t] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const int ks = lane >> 3, cg = lane & 7;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const bf16* vbase = vhead + cg * 8;
    // 8 value rows per lane in flight (the warp covers 32 keys per iteration with 512-byte coalesced requests)
    for (int t = t0 + ks; t < t1; t += 32) {
        uint4 u[8];
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int tt = t + 4 * i;
            const bool ok = tt < t1;
            u[i] = ok ? *reinterpret_cast<const uint4*>(vbase + static_cast<int64_t>(tt) * HD) : make_uint4(0u, 0u, 0u, 0u);
            p[i] = ok ? sc[tt] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float2 f;
            f = unpack_bf16x2(u[i].x); acc[0] += p[i] * f.x; acc[1] += p[i] * f.y;
            f = unpack_bf16x2(u[i].y); acc[2] += p[i] * f.x; acc[3] += p[i] * f.y;
            f = unpack_bf16x2(u[i].z); acc[4] += p[i] * f.x; acc[5] += p[i] * f.y;
            f = unpack_bf16x2(u[i].w); acc[6] += p[i] * f.x; acc[7] += p[i] * f.y;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (ks == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) part[warp][cg * 8 + i] = acc[i];
    }
    if (lane == 0) {
        part_m[warp] = mx;
        part_l[warp] = sum;
    }
    __syncthreads();
    if (tid < HD) {
        const float m = fmaxf(fmaxf(part_m[0], part_m[1]), fmaxf(part_m[2], part_m[3]));
        float l = 0.f, v = 0.f;
        if (m > -INFINITY) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float sw = (part_m[w] == -INFINITY) ? 0.f : __expf(part_m[w] - m);
                l += part_l[w] * sw;
                v += part[w][tid] * sw;
            }
        }
        o[static_cast<int64_t>(b) * d + h * HD + tid] = __float2bfloat16(l > 0.f ? v / l : 0.f);
    }
}

// Same step with the (sample, head)'s whole K / V history staged in shared memory by cp.async: every byte the CTA needs is
// requested at once (ncu on the direct-load kernel above: 28 us per layer for 74 MB = 2.6 TB/s, all stalls on the K / V
// loads with ~20 warps per SM in flight).  Used while the history fits (<= kDecodeStageKeys keys).
constexpr int kDecodeStageKeys = 224;      // 2 x 224 x 128 B = 56 KB per CTA -> 3-4 CTAs per SM
__global__ void __launch_bounds__(128) lm_attention_decode_staged_kernel(const float* __restrict__ qkv_acc, const float* __restrict__ qkv_bias,
                                                                         bf16* __restrict__ cache, const int* __restrict__ valid,
                                                                         int valid_stride, bf16* __restrict__ o, float* __restrict__ zero,
                                                                         int H, int pos, int Tmax) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t smem_dec[];
    const int n = pos + 1;
    bf16* Ks = reinterpret_cast<bf16*>(smem_dec);            // [n][64]
    bf16* Vs = Ks + static_cast<size_t>(n) * HD;             // [n][64]
    float* sc = reinterpret_cast<float*>(Vs + static_cast<size_t>(n) * HD);      // [n]
    int* vm = reinterpret_cast<int*>(sc + n);                // [n] key validity (a global load inside the score loop costs an
                                                             // L1-missing round trip per 32 keys on the critical path)
    __shared__ __align__(16) float sq[HD];
    __shared__ float part[4][HD];
    __shared__ float part_m[4], part_l[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int d = H * HD;
    const float* arow = qkv_acc + static_cast<int64_t>(b) * 3 * d;
    const int B = gridDim.x / H;
    bf16* kbase = cache + (static_cast<int64_t>(b) * H + h) * Tmax * HD;
    bf16* vhead = kbase + static_cast<int64_t>(B) * H * Tmax * HD;
    // history: 2 * pos rows of 128 B = 8 x 16-B requests each.  Two cp.async groups: the keys (+ validity words) first, so that
    // the scores and the softmax run while the values are still streaming in (round-2 ncu: DRAM busy only 36 % of the kernel,
    // every CTA of a wave loaded, then computed, in lock-step)
    for (int idx = tid; idx < pos * 8; idx += 128) {
        const int t = idx >> 3, c = (idx & 7) * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(Ks + t * HD + c)), "l"(kbase + static_cast<int64_t>(t) * HD + c) : "memory");
    }
    for (int t = tid; t < n; t += 128)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(vm + t)), "l"(valid + static_cast<int64_t>(b) * valid_stride + t) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int idx = tid; idx < pos * 8; idx += 128) {
        const int t = idx >> 3, c = (idx & 7) * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(Vs + t * HD + c)), "l"(vhead + static_cast<int64_t>(t) * HD + c) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // this step's q / k / v from the fp32 accumulators: k, v go to the cache AND to their shared-memory rows
    if (tid < HD) {
        const int c = h * HD + tid;
        sq[tid] = (arow[c] + __ldg(qkv_bias + c)) * 0.125f;
        const bf16 kv = __float2bfloat16(arow[d + c] + __ldg(qkv_bias + d + c));
        kbase[static_cast<int64_t>(pos) * HD + tid] = kv;
        Ks[pos * HD + tid] = kv;
        if (zero != nullptr) zero[static_cast<int64_t>(b) * d + c] = 0.f;
    } else {
        const int c = h * HD + tid - HD;
        const bf16 vv = __float2bfloat16(arow[2 * d + c] + __ldg(qkv_bias + 2 * d + c));
        vhead[static_cast<int64_t>(pos) * HD + tid - HD] = vv;
        Vs[pos * HD + tid - HD] = vv;
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");        // keys + validity have landed
    __syncthreads();
    // lane = (key slot ks, 16-byte column group cg): 8 lanes read one 128-byte row (conflict-free), 4 keys per warp pass
    const int ks = lane >> 3, cg = lane & 7;
    float q8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q8[i] = sq[cg * 8 + i];
    const int chunk = (n + 3) >> 2;
    const int t0 = warp * chunk, t1 = min(n, t0 + chunk);
    for (int tb = t0; tb < t1; tb += 4) {                   // warp-uniform trip count: the shuffles below need every lane
        const int t = tb + ks;
        const bool ok = t < t1;
        float acc = 0.f;
        if (ok) {
            const uint4 u = *reinterpret_cast<const uint4*>(Ks + t * HD + cg * 8);
            float2 f;
            f = unpack_bf16x2(u.x); acc = q8[0] * f.x + q8[1] * f.y;
            f = unpack_bf16x2(u.y); acc += q8[2] * f.x + q8[3] * f.y;
            f = unpack_bf16x2(u.z); acc += q8[4] * f.x + q8[5] * f.y;
            f = unpack_bf16x2(u.w); acc += q8[6] * f.x + q8[7] * f.y;
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (ok && cg == 0) sc[t] = vm[t] ? acc : -INFINITY;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int t = t0 + lane; t < t1; t += 32) mx = fmaxf(mx, sc[t]);
    mx = warp_max(mx);
    const float muse = (mx == -INFINITY) ? 0.f : mx;
    float sum = 0.f;
    for (int t = t0 + lane; t < t1; t += 32) {
        const float p = __expf(sc[t] - muse);
        sc[t] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    cp_async_wait_all();                                        // the values
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int t = t0 + ks; t < t1; t += 4) {
        const float p = sc[t];
        const uint4 u = *reinterpret_cast<const uint4*>(Vs + t * HD + cg * 8);
        float2 f;
        f = unpack_bf16x2(u.x); acc[0] += p * f.x; acc[1] += p * f.y;
        f = unpack_bf16x2(u.y); acc[2] += p * f.x; acc[3] += p * f.y;
        f = unpack_bf16x2(u.z); acc[4] += p * f.x; acc[5] += p * f.y;
        f = unpack_bf16x2(u.w); acc[6] += p * f.x; acc[7] += p * f.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (ks == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) part[warp][cg * 8 + i] = acc[i];
    }
    if (lane == 0) {
        part_m[warp] = mx;
        part_l[warp] = sum;
    }
    __syncthreads();
    if (tid < HD) {
        const float m = fmaxf(fmaxf(part_m[0], part_m[1]), fmaxf(part_m[2], part_m[3]));
        float l = 0.f, v = 0.f;
        if (m > -INFINITY) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float sw = (part_m[w] == -INFINITY) ? 0.f : __expf(part_m[w] - m);
                l += part_l[w] * sw;
                v += part[w][tid] * sw;
            }
        }
        o[static_cast<int64_t>(b) * d + h * HD + tid] = __float2bfloat16(l > 0.f ? v / l : 0.f);
    }
}

// ------------------------------------------------------------------------------------------ mapper attention, tensor-core path
// S <= 32, head_dim in {16, 32, 64, 96, 128}: one 64-thread CTA per (sample, head), warp w owns query rows 16w..16w+15
// (and, in the backward, key rows 16w..).  bf16 mma.sync with fp32 softmax; replaces the shared-memory-bound scalar
// kernels below (round-1 profile: 67 us forward / 131 us backward per layer for 0.3 GFLOP of work).
template <int LD> __device__ __forceinline__ void ld_a(uint32_t (&a)[4], const bf16* tile, int row0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4(a, smem_addr(tile + (row0 + (mi & 1) * 8 + r) * LD + k0 + (mi >> 1) * 8));
}
template <int LD> __device__ __forceinline__ void ld_a_trans(uint32_t (&a)[4], const bf16* tile, int m0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4_trans(a, smem_addr(tile + (k0 + (mi >> 1) * 8 + r) * LD + m0 + (mi & 1) * 8));
}
template <int LD> __device__ __forceinline__ void ld_b(uint32_t (&b)[4], const bf16* tile, int n0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4(b, smem_addr(tile + (n0 + (mi >> 1) * 8 + r) * LD + k0 + (mi & 1) * 8));
}
template <int LD> __device__ __forceinline__ void ld_b_trans(uint32_t (&b)[4], const bf16* tile, int n0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldmatrix_x4_trans(b, smem_addr(tile + (k0 + (mi & 1) * 8 + r) * LD + n0 + (mi >> 1) * 8));
}

// 32 x HDIM tile through cp.async: every 16-byte request of the CTA's tiles is in flight at once (the first version
// loaded through registers, 18 dependent load -> store round trips per thread: 23 us per layer for 31 MB).
// Callers finish with cp_async_wait_all() + __syncthreads().
template <int HDIM>
__device__ __forceinline__ void load_rows32(bf16* dst, const bf16* src, int64_t ld, int S, int tid) {
    constexpr int LD = HDIM + 8, CPR = HDIM / 8;
#pragma unroll
    for (int idx = tid; idx < 32 * CPR; idx += 64) {
        const int r = idx / CPR, c = (idx % CPR) * 8;
        bf16* d = dst + r * LD + c;
        if (r < S) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(d)), "l"(src + static_cast<int64_t>(r) * ld + c)
                         : "memory");
        } else {
            *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// scores of this warp's 16 queries against 32 keys: acc[4][4]
template <int HDIM>
__device__ __forceinline__ void scores32(float (&acc)[4][4], const bf16* As, const bf16* Bs, int warp, int lane) {
    constexpr int LD = HDIM + 8;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HDIM / 16; ++ks) {
        uint32_t a[4];
        ld_a<LD>(a, As, warp * 16, ks * 16, lane);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            uint32_t b[4];
            ld_b<LD>(b, Bs, np * 16, ks * 16, lane);
            mma_bf16(acc[2 * np], a, b[0], b[1]);
            mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

template <int HDIM>
__global__ void __launch_bounds__(64) mapper_attention_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, int S, int H) {
    pdl_trigger();
    pdl_wait();
    constexpr int LD = HDIM + 8;
    __shared__ __align__(16) bf16 Qs[32 * LD];
    __shared__ __align__(16) bf16 Ks[32 * LD];
    __shared__ __align__(16) bf16 Vs[32 * LD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int h = blockIdx.x, b = blockIdx.y;
    const int d = H * HDIM;
    const int64_t ld = 3 * d;
    const bf16* base = qkv + static_cast<int64_t>(b) * S * ld + h * HDIM;
    load_rows32<HDIM>(Qs, base, ld, S, tid);
    load_rows32<HDIM>(Ks, base + d, ld, S, tid);
    load_rows32<HDIM>(Vs, base + 2 * d, ld, S, tid);
    cp_async_wait_all();
    __syncthreads();
    float sacc[4][4];
    scores32<HDIM>(sacc, Qs, Ks, warp, lane);
    const float scale = rsqrtf(static_cast<float>(HDIM));      // clipcap.py:75
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = nt * 8 + 2 * t4 + (e & 1);
            const float v = col < S ? sacc[nt][e] * scale : -INFINITY;
            sacc[nt][e] = v;
            mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float p = __expf(sacc[nt][e] - mx[e >> 1]);
            sacc[nt][e] = p;
            sum[e >> 1] += p;
        }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
        sum[r] = 1.0f / sum[r];
    }
    uint32_t pf[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(sacc[nt][0] * sum[0], sacc[nt][1] * sum[0]);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(sacc[nt][2] * sum[1], sacc[nt][3] * sum[1]);
    }
    float oacc[HDIM / 8][4];
#pragma unroll
    for (int i = 0; i < HDIM / 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int np = 0; np < HDIM / 16; ++np) {
            uint32_t bfr[4];
            ld_b_trans<LD>(bfr, Vs, np * 16, ks * 16, lane);
            mma_bf16(oacc[2 * np], pf[ks], bfr[0], bfr[1]);
            mma_bf16(oacc[2 * np + 1], pf[ks], bfr[2], bfr[3]);
        }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = warp * 16 + g + r * 8;
        if (q >= S) continue;
        bf16* op = o + (static_cast<int64_t>(b) * S + q) * d + h * HDIM;
#pragma unroll
        for (int nt = 0; nt < HDIM / 8; ++nt)
            *reinterpret_cast<uint32_t*>(op + nt * 8 + 2 * t4) = pack_bf16x2(oacc[nt][2 * r], oacc[nt][2 * r + 1]);
    }
}

template <int HDIM>
__global__ void __launch_bounds__(64) mapper_attention_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                                      bf16* __restrict__ dqkv, int S, int H) {
    pdl_trigger();
    pdl_wait();
    constexpr int LD = HDIM + 8, LP = 40;
    extern __shared__ __align__(16) uint8_t smem_map[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_map);
    bf16* Ks = Qs + 32 * LD;
    bf16* Vs = Ks + 32 * LD;
    bf16* Gs = Vs + 32 * LD;
    bf16* Ps = Gs + 32 * LD;
    bf16* dSs = Ps + 32 * LP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int h = blockIdx.x, b = blockIdx.y;
    const int d = H * HDIM;
    const int64_t ld = 3 * d;
    const bf16* base = qkv + static_cast<int64_t>(b) * S * ld + h * HDIM;
    bf16* dbase = dqkv + static_cast<int64_t>(b) * S * ld + h * HDIM;
    load_rows32<HDIM>(Qs, base, ld, S, tid);
    load_rows32<HDIM>(Ks, base + d, ld, S, tid);
    load_rows32<HDIM>(Vs, base + 2 * d, ld, S, tid);
    load_rows32<HDIM>(Gs, d_o + static_cast<int64_t>(b) * S * d + h * HDIM, d, S, tid);
    cp_async_wait_all();
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(HDIM));
    {
        float sacc[4][4], dpacc[4][4];
        scores32<HDIM>(sacc, Qs, Ks, warp, lane);       // S  = Q K^T
        scores32<HDIM>(dpacc, Gs, Vs, warp, lane);      // dP = dO V^T
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nt * 8 + 2 * t4 + (e & 1);
                const float v = col < S ? sacc[nt][e] * scale : -INFINITY;
                sacc[nt][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        float sum[2] = {0.f, 0.f}, dot[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = __expf(sacc[nt][e] - mx[e >> 1]);
                sacc[nt][e] = p;
                sum[e >> 1] += p;
                dot[e >> 1] += p * dpacc[nt][e];
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
            sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
            dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], 1);
            dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], 2);
            sum[r] = 1.0f / sum[r];
            dot[r] *= sum[r];                            // sum_j p_ij dP_ij with normalised p
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float p[4], ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = warp * 16 + g + ((e >> 1) << 3);
                p[e] = row < S ? sacc[nt][e] * sum[e >> 1] : 0.f;      // padded query rows must not reach dK / dV
                ds[e] = p[e] * (dpacc[nt][e] - dot[e >> 1]) * scale;
            }
            const int r0 = warp * 16 + g, c0 = nt * 8 + 2 * t4;
            *reinterpret_cast<uint32_t*>(Ps + r0 * LP + c0) = pack_bf16x2(p[0], p[1]);
            *reinterpret_cast<uint32_t*>(Ps + (r0 + 8) * LP + c0) = pack_bf16x2(p[2], p[3]);
            *reinterpret_cast<uint32_t*>(dSs + r0 * LP + c0) = pack_bf16x2(ds[0], ds[1]);
            *reinterpret_cast<uint32_t*>(dSs + (r0 + 8) * LP + c0) = pack_bf16x2(ds[2], ds[3]);
        }
    }
    __syncthreads();
    auto store_rows = [&](const float (&acc)[HDIM / 8][4], int col_off) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int pos = warp * 16 + g + r * 8;
            if (pos >= S) continue;
            bf16* p = dbase + static_cast<int64_t>(pos) * ld + col_off;
#pragma unroll
            for (int nt = 0; nt < HDIM / 8; ++nt)
                *reinterpret_cast<uint32_t*>(p + nt * 8 + 2 * t4) = pack_bf16x2(acc[nt][2 * r], acc[nt][2 * r + 1]);
        }
    };
    // which: 0 -> dV = P^T dO, 1 -> dK = dS^T Q (A transposed from smem), 2 -> dQ = dS K
#pragma unroll
    for (int which = 0; which < 3; ++which) {
        float acc[HDIM / 8][4];
#pragma unroll
        for (int i = 0; i < HDIM / 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
        const bf16* Asrc = which == 0 ? Ps : dSs;
        const bf16* Bsrc = which == 0 ? Gs : (which == 1 ? Qs : Ks);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t a[4];
            if (which == 2) ld_a<LP>(a, Asrc, warp * 16, ks * 16, lane);
            else ld_a_trans<LP>(a, Asrc, warp * 16, ks * 16, lane);
#pragma unroll
            for (int np = 0; np < HDIM / 16; ++np) {
                uint32_t bfr[4];
                ld_b_trans<LD>(bfr, Bsrc, np * 16, ks * 16, lane);
                mma_bf16(acc[2 * np], a, bfr[0], bfr[1]);
                mma_bf16(acc[2 * np + 1], a, bfr[2], bfr[3]);
            }
        }
        store_rows(acc, which == 0 ? 2 * d : (which == 1 ? d : 0));
    }
}

// ------------------------------------------------------------------------------------------ mapper attention
// smem (fp32): q, k, v [S][hd+1]  (+ do for backward), p [S][S+1] (+ dp)
__global__ void __launch_bounds__(128) mapper_attention_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, int S,
                                                                   int H, int hd) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    const int ldh = hd + 1, ldp = S + 1;
    float* q = sm;
    float* k = q + S * ldh;
    float* v = k + S * ldh;
    float* p = v + S * ldh;
    const int h = blockIdx.x, b = blockIdx.y;
    const int d = H * hd;
    const bf16* base = qkv + static_cast<int64_t>(b) * S * 3 * d + h * hd;
    for (int idx = threadIdx.x; idx < S * hd; idx += blockDim.x) {
        const int i = idx / hd, c = idx % hd;
        const bf16* row = base + static_cast<int64_t>(i) * 3 * d + c;
        q[i * ldh + c] = __bfloat162float(row[0]);
        k[i * ldh + c] = __bfloat162float(row[d]);
        v[i * ldh + c] = __bfloat162float(row[2 * d]);
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(hd));      // clipcap.py:75
    for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
        const int i = idx / S, j = idx % S;
        float acc = 0.f;
        for (int c = 0; c < hd; ++c) acc += q[i * ldh + c] * k[j * ldh + c];
        p[i * ldp + j] = acc * scale;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < S; i += 4) {
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) mx = fmaxf(mx, p[i * ldp + j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float e = __expf(p[i * ldp + j] - mx);
            p[i * ldp + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < S; j += 32) p[i * ldp + j] *= inv;
    }
    __syncthreads();
    bf16* ob = o + static_cast<int64_t>(b) * S * d + h * hd;
    for (int idx = threadIdx.x; idx < S * hd; idx += blockDim.x) {
        const int i = idx / hd, c = idx % hd;
        float acc = 0.f;
        for (int j = 0; j < S; ++j) acc += p[i * ldp + j] * v[j * ldh + c];
        ob[static_cast<int64_t>(i) * d + c] = __float2bfloat16(acc);
    }
}

__global__ void __launch_bounds__(128) mapper_attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                                   bf16* __restrict__ dqkv, int S, int H, int hd) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    const int ldh = hd + 1, ldp = S + 1;
    float* q = sm;
    float* k = q + S * ldh;
    float* v = k + S * ldh;
    float* go = v + S * ldh;
    float* p = go + S * ldh;
    float* ds = p + S * ldp;
    const int h = blockIdx.x, b = blockIdx.y;
    const int d = H * hd;
    const bf16* base = qkv + static_cast<int64_t>(b) * S * 3 * d + h * hd;
    const bf16* gbase = d_o + static_cast<int64_t>(b) * S * d + h * hd;
    for (int idx = threadIdx.x; idx < S * hd; idx += blockDim.x) {
        const int i = idx / hd, c = idx % hd;
        const bf16* row = base + static_cast<int64_t>(i) * 3 * d + c;
        q[i * ldh + c] = __bfloat162float(row[0]);
        k[i * ldh + c] = __bfloat162float(row[d]);
        v[i * ldh + c] = __bfloat162float(row[2 * d]);
        go[i * ldh + c] = __bfloat162float(gbase[static_cast<int64_t>(i) * d + c]);
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(hd));
    for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
        const int i = idx / S, j = idx % S;
        float acc = 0.f, acc2 = 0.f;
        for (int c = 0; c < hd; ++c) {
            acc += q[i * ldh + c] * k[j * ldh + c];
            acc2 += go[i * ldh + c] * v[j * ldh + c];
        }
        p[i * ldp + j] = acc * scale;
        ds[i * ldp + j] = acc2;            // dP
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < S; i += 4) {
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) mx = fmaxf(mx, p[i * ldp + j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float e = __expf(p[i * ldp + j] - mx);
            p[i * ldp + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float dot = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float pv = p[i * ldp + j] * inv;
            p[i * ldp + j] = pv;
            dot += pv * ds[i * ldp + j];
        }
        dot = warp_sum(dot);
        for (int j = lane; j < S; j += 32) ds[i * ldp + j] = p[i * ldp + j] * (ds[i * ldp + j] - dot) * scale;
    }
    __syncthreads();
    bf16* dbase = dqkv + static_cast<int64_t>(b) * S * 3 * d + h * hd;
    for (int idx = threadIdx.x; idx < S * hd; idx += blockDim.x) {
        const int i = idx / hd, c = idx % hd;
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < S; ++j) {
            dq += ds[i * ldp + j] * k[j * ldh + c];
            dk += ds[j * ldp + i] * q[j * ldh + c];
            dv += p[j * ldp + i] * go[j * ldh + c];
        }
        bf16* row = dbase + static_cast<int64_t>(i) * 3 * d + c;
        row[0] = __float2bfloat16(dq);
        row[d] = __float2bfloat16(dk);
        row[2 * d] = __float2bfloat16(dv);
    }
}

}  // namespace

// ============================================================================================ launchers
void lm_attention_fwd(const bf16* qkv, const int* valid, bf16* o, float* lse, int B, int T, int H, cudaStream_t s, bf16* kv_cache,
                      int Tmax) {
    if (T <= BLK && kv_cache == nullptr) {
        static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
        if (!configured) {
            CUDA_CHECK(cudaFuncSetAttribute(lm_attention_fwd_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kFwdSetBytes));
            configured = true;
        }
        const int n_items = B * H;
        launch_kernel(lm_attention_fwd_single_kernel, dim3(std::min(n_items, 4 * num_sms())), dim3(128), 2 * kFwdSetBytes, s, qkv, valid, o,
                      lse, T, H, n_items);
        KERNEL_CHECK();
        count_launch();
        return;
    }
    if (lse == nullptr && T > BLK && T <= kPrefillMaxBlocks * BLK) {
        const int nblk = ceil_div(T, BLK);
        const size_t smem = static_cast<size_t>(2 * nblk + 2) * BLK * LDS * sizeof(bf16) + static_cast<size_t>(nblk) * BLK * sizeof(int);
        static std::atomic<bool> configured{false};
        if (!configured) {
            CUDA_CHECK(cudaFuncSetAttribute(lm_attention_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (2 * kPrefillMaxBlocks + 2) * BLK * LDS * static_cast<int>(sizeof(bf16)) + kPrefillMaxBlocks * BLK * 4));
            configured = true;
        }
        launch_kernel(lm_attention_prefill_kernel, dim3(H, B), dim3(128), smem, s, qkv, valid, o, T, H, kv_cache, Tmax);
        KERNEL_CHECK();
        count_launch();
        return;
    }
    dim3 grid(ceil_div(T, BLK), H, B);
    if (kv_cache != nullptr) launch_kernel(lm_attention_fwd_kernel<true>, dim3(grid), dim3(128), 0, s, qkv, valid, o, lse, T, H, kv_cache, Tmax);
    else launch_kernel(lm_attention_fwd_kernel<false>, dim3(grid), dim3(128), 0, s, qkv, valid, o, lse, T, H, kv_cache, Tmax);
    KERNEL_CHECK();
    count_launch();
}

void lm_attention_bwd(const bf16* qkv, const int* valid, const bf16* o, const bf16* d_o, const float* lse, bf16* dqkv,
                      float* dq_scratch, int B, int T, int H, cudaStream_t s) {
    const int smem = 6 * BLK * LDS * sizeof(bf16) + 2 * BLK * sizeof(float) + BLK * sizeof(int);
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(lm_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    EAVQA_CHECK(T <= BLK || dq_scratch != nullptr, "lm_attention_bwd needs dq_scratch for T > 64");
    dim3 grid(H, B);
    if (T <= BLK) {
        static std::atomic<bool> configured1{false};
        if (!configured1) {
            CUDA_CHECK(cudaFuncSetAttribute(lm_attention_bwd_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured1 = true;
        }
        launch_kernel(lm_attention_bwd_single_kernel, dim3(grid), dim3(128), smem, s, qkv, valid, o, d_o, lse, dqkv, T, H);
        KERNEL_CHECK();
        count_launch();
        return;
    }
    launch_kernel(lm_attention_bwd_kernel, dim3(grid), dim3(128), smem, s, qkv, valid, o, d_o, lse, dqkv, dq_scratch, T, H);
    KERNEL_CHECK();
    count_launch();
}

void lm_attention_decode_acc(const float* qkv_acc, const float* qkv_bias, bf16* cache, const int* valid, int valid_stride, bf16* o,
                             float* zero, int B, int H, int pos, int Tmax, cudaStream_t s) {
    EAVQA_CHECK(pos < Tmax, "decode position beyond the KV cache");
    const int n = pos + 1;
    if (n <= kDecodeStageKeys) {
        const size_t smem = static_cast<size_t>(n) * (2 * HD * sizeof(bf16) + sizeof(float) + sizeof(int));
        static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
        if (!configured) {
            CUDA_CHECK(cudaFuncSetAttribute(lm_attention_decode_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kDecodeStageKeys * (2 * HD * static_cast<int>(sizeof(bf16)) + 2 * static_cast<int>(sizeof(float)))));
            configured = true;
        }
        launch_kernel(lm_attention_decode_staged_kernel, dim3(B * H), dim3(128), smem, s, qkv_acc, qkv_bias, cache, valid, valid_stride, o,
                      zero, H, pos, Tmax);
    } else {
        const size_t smem = sizeof(float) * Tmax;
        EAVQA_CHECK(smem <= 40 * 1024, "decode: KV length exceeds the score buffer");
        launch_kernel(lm_attention_decode_acc_kernel, dim3(B * H), dim3(128), smem, s, qkv_acc, qkv_bias, cache, valid, valid_stride, o, zero,
                      H, pos, Tmax);
    }
    KERNEL_CHECK();
    count_launch();
}

template <int HDIM>
static void launch_mapper_bwd_mma(const bf16* qkv, const bf16* d_o, bf16* dqkv, int B, int S, int H, cudaStream_t s) {
    const int smem = (4 * 32 * (HDIM + 8) + 2 * 32 * 40) * sizeof(bf16);
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured && smem > 48 * 1024) {
        CUDA_CHECK(cudaFuncSetAttribute(mapper_attention_bwd_mma_kernel<HDIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    launch_kernel(mapper_attention_bwd_mma_kernel<HDIM>, dim3(dim3(H, B)), dim3(64), smem, s, qkv, d_o, dqkv, S, H);
}

void mapper_attention_fwd(const bf16* qkv, bf16* o, int B, int S, int H, int hd, cudaStream_t s) {
    if (S <= 32 && (hd == 16 || hd == 32 || hd == 64 || hd == 96 || hd == 128)) {
        dim3 grid(H, B);
        switch (hd) {
            case 16: launch_kernel(mapper_attention_fwd_mma_kernel<16>, dim3(grid), dim3(64), 0, s, qkv, o, S, H); break;
            case 32: launch_kernel(mapper_attention_fwd_mma_kernel<32>, dim3(grid), dim3(64), 0, s, qkv, o, S, H); break;
            case 64: launch_kernel(mapper_attention_fwd_mma_kernel<64>, dim3(grid), dim3(64), 0, s, qkv, o, S, H); break;
            case 96: launch_kernel(mapper_attention_fwd_mma_kernel<96>, dim3(grid), dim3(64), 0, s, qkv, o, S, H); break;
            default: launch_kernel(mapper_attention_fwd_mma_kernel<128>, dim3(grid), dim3(64), 0, s, qkv, o, S, H); break;
        }
        KERNEL_CHECK();
        count_launch();
        return;
    }
    const int smem = (3 * S * (hd + 1) + S * (S + 1)) * sizeof(float);
    EAVQA_CHECK(smem <= 200 * 1024, "mapper attention tile does not fit in shared memory");
    static int configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(mapper_attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    dim3 grid(H, B);
    launch_kernel(mapper_attention_fwd_kernel, dim3(grid), dim3(128), smem, s, qkv, o, S, H, hd);
    KERNEL_CHECK();
    count_launch();
}

void mapper_attention_bwd(const bf16* qkv, const bf16* d_o, bf16* dqkv, int B, int S, int H, int hd, cudaStream_t s) {
    if (S <= 32 && (hd == 16 || hd == 32 || hd == 64 || hd == 96 || hd == 128)) {
        switch (hd) {
            case 16: launch_mapper_bwd_mma<16>(qkv, d_o, dqkv, B, S, H, s); break;
            case 32: launch_mapper_bwd_mma<32>(qkv, d_o, dqkv, B, S, H, s); break;
            case 64: launch_mapper_bwd_mma<64>(qkv, d_o, dqkv, B, S, H, s); break;
            case 96: launch_mapper_bwd_mma<96>(qkv, d_o, dqkv, B, S, H, s); break;
            default: launch_mapper_bwd_mma<128>(qkv, d_o, dqkv, B, S, H, s); break;
        }
        KERNEL_CHECK();
        count_launch();
        return;
    }
    const int smem = (4 * S * (hd + 1) + 2 * S * (S + 1)) * sizeof(float);
    EAVQA_CHECK(smem <= 200 * 1024, "mapper attention tile does not fit in shared memory");
    static int configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(mapper_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    dim3 grid(H, B);
    launch_kernel(mapper_attention_bwd_kernel, dim3(grid), dim3(128), smem, s, qkv, d_o, dqkv, S, H, hd);
    KERNEL_CHECK();
    count_launch();
}

}  // namespace eavqa
