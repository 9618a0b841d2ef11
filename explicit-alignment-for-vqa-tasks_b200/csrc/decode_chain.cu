// Persistent single-token decode step: ONE cooperative kernel walks a list of phases -- the split-K projection GEMMs
// (tcgen05 / TMEM, operands by TMA), the KV-cache attention and the residual + LayerNorm glue of every LM layer, then the
// tied head -- with a grid-wide barrier between phases.  It replaces the ~170 dependent launches of one decode step
// (clipcap.py:416-423 re-runs the whole sequence instead; same function, KV-cached here), whose cost was pure kernel
// latency: round-1 profile, 1750 launches x ~6 us per 128-answer batch against 0.38 ms / step of HBM traffic.
//
// One CTA per SM (cooperative launch: all co-resident, so the barrier cannot deadlock), 320 threads:
//   warp 0 / lane 0 : TMA producer of the GEMM phases        warp 1 / lane 0 : MMA issuer (warp 1 owns TMEM)
//   warps 2..9      : GEMM epilogue (TMEM -> registers -> swizzled staging -> TMA store / reduce-add);
//                     in attention phases two groups of four warps, one (sample, head) at a time each;
//   all 10 warps    : one row per warp in the residual + LayerNorm phases.
// Data crossing a phase boundary lives in global memory (L2): generic-proxy writers fence towards the async proxy before
// the barrier, the TMA producer fences after it; bulk stores are waited for (complete, not just read) before the barrier.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <vector>

#include "gemm_kernel.cuh"
#include "kernels.cuh"

namespace eavqa {

namespace dc {

constexpr int BM = 128, BN = 64, BK = 64;
constexpr int STAGES = 6;
constexpr int STAGE_A = BM * BK * 2, STAGE_B = BN * BK * 2;
constexpr int THREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BUF = 4096;                                   // 32 rows x 128 B (fp32) or 32 x 64 B (bf16)
constexpr int PIPE_BYTES = STAGES * (STAGE_A + STAGE_B);        // 147456
constexpr int WORK_BYTES = PIPE_BYTES + EPI_WARPS * EPI_BUF;    // 180224: also the attention phases' staging area
constexpr int BAR_BYTES = 512;
constexpr int SMEM = WORK_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 128;                                  // two 64-column fp32 accumulators
constexpr int ATT_GROUP_BYTES = WORK_BYTES / 2;                 // per four-warp group
constexpr int ATT_FIXED = (64 + 4 * 64 + 8) * 4;                // q, per-warp partial outputs, per-warp (max, sum)
constexpr int HD = 64;
constexpr int ATT_BAR_OFS = 160;                                // four mbarriers (two groups x two buffers) inside the barrier block
constexpr int GLUE_RED_OFS = 256;                               // 32 floats of reduction scratch for the glue phases

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// Grid-wide barrier on a monotonically increasing counter (zeroed by the host once per generate call).  `target` is the
// counter value after every CTA has arrived at this barrier.
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned target) {
    fence_proxy_async_all();           // this thread's generic writes -> ordered before later async-proxy (TMA) reads
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();               // cumulative: covers the CTA's writes ordered before it by the barrier above
        atomicAdd(counter, 1u);
        while (static_cast<int>(ld_acquire_gpu(counter) - target) < 0) {
        }
        __threadfence();               // gpu-scope fence: also drops this SM's stale L1 lines
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------- GEMM phase
struct PipeState {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
};
struct AccState {
    int acc = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        acc ^= 1;
        if (acc == 0) phase ^= 1;
    }
};

__device__ __forceinline__ void item_range(const ChainPhase& p, int item, int& n_idx, int& kb0, int& kb1) {
    const int num_kb = (p.K + BK - 1) / BK;
    n_idx = item / p.split;
    const int sp = item - n_idx * p.split;
    kb0 = static_cast<int>(static_cast<int64_t>(sp) * num_kb / p.split);
    kb1 = static_cast<int>(static_cast<int64_t>(sp + 1) * num_kb / p.split);
}

__device__ __forceinline__ void gemm_producer(const ChainPhase& p, PipeState& st, uint32_t smem_a, uint32_t smem_b, uint32_t full_bar,
                                              uint32_t empty_bar) {
    fence_proxy_async_all();           // operands written by generic stores of the previous phase (other SMs) -> TMA reads
    const int n_items = (p.N / BN) * p.split;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int n_idx, kb0, kb1;
        item_range(p, item, n_idx, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(empty_bar + 8 * st.stage, st.phase ^ 1);
            ptx::mbar_arrive_expect_tx(full_bar + 8 * st.stage, STAGE_A + STAGE_B);
            ptx::tma_load_2d(smem_a + st.stage * STAGE_A, &p.map_a, full_bar + 8 * st.stage, kb * BK, 0);
            ptx::tma_load_2d(smem_b + st.stage * STAGE_B, &p.map_b, full_bar + 8 * st.stage, kb * BK, n_idx * BN);
            st.advance();
        }
    }
}

__device__ __forceinline__ void gemm_mma(const ChainPhase& p, PipeState& st, AccState& as, uint32_t smem_a, uint32_t smem_b,
                                         uint32_t full_bar, uint32_t empty_bar, uint32_t tfull_bar, uint32_t tempty_bar,
                                         uint32_t tmem_base) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM, BN);
    const int n_items = (p.N / BN) * p.split;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int n_idx, kb0, kb1;
        item_range(p, item, n_idx, kb0, kb1);
        ptx::mbar_wait(tempty_bar + 8 * as.acc, as.phase ^ 1);
        ptx::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as.acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(full_bar + 8 * st.stage, st.phase);
            ptx::tcgen05_fence_after();
            const uint64_t da = ptx::make_kmajor_sw128_desc(smem_a + st.stage * STAGE_A);
            const uint64_t db = ptx::make_kmajor_sw128_desc(smem_b + st.stage * STAGE_B);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) ptx::umma_bf16(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            ptx::umma_commit(empty_bar + 8 * st.stage);
            st.advance();
        }
        ptx::umma_commit(tfull_bar + 8 * as.acc);
        as.advance();
    }
}

// warps 2..9: warp w reads TMEM lanes [32 (w & 3), +32) = 32 output rows and columns [32 half, +32) of the 64-wide tile
__device__ __forceinline__ void gemm_epilogue(const ChainPhase& p, AccState& as, uint32_t tfull_bar, uint32_t tempty_bar,
                                              uint32_t tmem_base, uint32_t buf, int warp, int lane) {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int n_items = (p.N / BN) * p.split;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int n_idx = item / p.split;
        const int n0 = n_idx * BN + half * 32, row0 = quarter * 32;
        ptx::mbar_wait(tfull_bar + 8 * as.acc, as.phase);
        ptx::tcgen05_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as.acc * BN + half * 32, r);
        ptx::tmem_ld_wait();
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar + 8 * as.acc);
        as.advance();
        if (lane == 0) gk::tma_store_wait_read<0>();        // the staging buffer's previous store has been read out
        __syncwarp();
        if (p.mode == CHAIN_GELU_BF16) {
            f32x2 v[16];
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b = __ldg(b4 + q);
                v[2 * q] = gelu_new2(add2(pk(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), pk(b.x, b.y)));
                v[2 * q + 1] = gelu_new2(add2(pk(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), pk(b.z, b.w)));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                gk::st_shared_v4(buf + gk::swz64(lane, q), pack_bf16x2(v[4 * q]), pack_bf16x2(v[4 * q + 1]), pack_bf16x2(v[4 * q + 2]),
                                 pack_bf16x2(v[4 * q + 3]));
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) gk::st_shared_v4(buf + gk::swz128(lane, q), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (p.mode == CHAIN_REDUCE_F32) gk::tma_reduce_add_2d(&p.map_out, buf, n0, row0);
            else gk::tma_store_2d(&p.map_out, buf, n0, row0);
            gk::tma_store_commit();
        }
    }
    // every bulk store of this phase must have COMPLETED (not only been read from shared memory) before the grid barrier
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------- attention phase
// One (sample, head) per four-warp group at a time: this step's q / k / v come from the fp32 split-K accumulator + bias (k, v
// are appended to the cache), the head's K / V history is staged in shared memory by cp.async -- double-buffered when two
// histories fit, so the next item's bytes are in flight while this one is reduced -- and the four warps split the keys
// (flash-decoding style merge through shared memory).
struct AttGroup {
    uint8_t* base;       // this group's staging area
    int gid, gtid;       // group id in the CTA (0 / 1), thread id in the group (0..127)
    uint32_t bar[2];     // mbarriers of the two staging buffers (bulk-copy completion)
    uint32_t parity[2];  // their current phase parities (persist across the attention phases of the kernel)
};

__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int KV_BOX = 16;                // cache rows per TMA box (2 KB); staged histories are padded to a multiple of it

// staging layout of one item (1024-byte aligned: the rows are written by TMA with SWIZZLE_128B, i.e. the 16-byte chunk c of
// row t sits at chunk c ^ (t & 7)):  K rows [nr][64] bf16 | V rows [nr][64] bf16 | scores [n] fp32 | key validity [n] int32
__device__ __forceinline__ int att_rows(int n) { return (n + KV_BOX - 1) / KV_BOX * KV_BOX; }
__device__ __forceinline__ int att_hist_bytes(int n) { return (2 * att_rows(n) * HD * 2 + n * 8 + 1023) & ~1023; }

// The head's K / V history [0, pos) by TMA boxes of 16 rows through the per-layer tensor map of the cache (one thread, ~2 x 9
// instructions per item; the first version issued 144 LDGSTS warp-instructions per item and kept the LSU busy for ~2800 cycles),
// swizzled so that a LANE can later read a whole key row without bank conflicts.  The validity words travel by cp.async.
__device__ __forceinline__ void att_issue_loads(const ChainPhase& p, int item, int pos, uint8_t* buf, int gtid, uint32_t bar) {
    const int b = item / p.H;
    const int n = pos + 1, nr = att_rows(n);
    if (gtid == 0 && pos > 0) {
        // the buffer was last READ through the generic proxy (the item two back): order those reads before the async writes
        ptx::fence_proxy_async_smem();
        const int nbox = (pos + KV_BOX - 1) / KV_BOX;
        const int row_k = item * p.Tmax, row_v = row_k + p.B * p.H * p.Tmax;
        const uint32_t ks = ptx::smem_u32(buf), vs = ks + static_cast<uint32_t>(nr) * HD * 2;
        ptx::mbar_arrive_expect_tx(bar, static_cast<uint32_t>(2 * nbox * KV_BOX * HD * 2));
        for (int i = 0; i < nbox; ++i) {
            ptx::tma_load_2d(ks + i * KV_BOX * HD * 2, &p.map_a, bar, 0, row_k + i * KV_BOX);
            ptx::tma_load_2d(vs + i * KV_BOX * HD * 2, &p.map_a, bar, 0, row_v + i * KV_BOX);
        }
    }
    int* vm = reinterpret_cast<int*>(buf + 2 * nr * HD * 2) + n;
    const int* vrow = p.valid + static_cast<int64_t>(b) * p.valid_stride;
    for (int t = gtid; t < n; t += 128) cp_async_4(ptx::smem_u32(vm + t), vrow + t);
    cp_async_commit();
}

// this thread's share of an item's new q | k | v row, RAW (accumulator and bias separately: the add happens an item later, so
// that the in-order issue never waits for these loads): threads 0..63 hold (q, k) of one head dimension, 64..127 hold v
struct QkvRaw { float a0, b0, a1, b1; };
__device__ __forceinline__ QkvRaw att_load_qkv(const ChainPhase& p, int item, int gtid) {
    const int b = item / p.H, h = item - b * p.H;
    const int d = p.H * HD;
    const float* __restrict__ arow = p.qkv_acc + static_cast<int64_t>(b) * 3 * d;
    const float* __restrict__ bias = p.bias;
    QkvRaw r;
    if (gtid < HD) {
        const int c = h * HD + gtid;
        r.a0 = __ldcg(arow + c); r.b0 = __ldg(bias + c);
        r.a1 = __ldcg(arow + d + c); r.b1 = __ldg(bias + d + c);
    } else {
        const int c = h * HD + gtid - HD;
        r.a0 = __ldcg(arow + 2 * d + c); r.b0 = __ldg(bias + 2 * d + c);
        r.a1 = 0.f; r.b1 = 0.f;
    }
    return r;
}

__device__ __forceinline__ void attention_phase(const ChainPhase& p, int pos, AttGroup& g) {
    const int n = pos + 1, nr = att_rows(n);
    const int hist_bytes = att_hist_bytes(n);
    const bool dbl = 2 * hist_bytes + ATT_FIXED <= ATT_GROUP_BYTES;
    uint8_t* bufs[2] = {g.base, g.base + hist_bytes};
    float* fixed = reinterpret_cast<float*>(g.base + (dbl ? 2 : 1) * hist_bytes);
    float* sq = fixed;                    // [64]
    float* part = fixed + 64;             // [4][64]
    float* part_m = part + 256;           // [4]
    float* part_l = part_m + 4;           // [4]
    const int warp = g.gtid >> 5, lane = g.gtid & 31;
    const int n_items = p.B * p.H;
    const int first = blockIdx.x * 2 + g.gid, stride = gridDim.x * 2;
    const int d = p.H * HD;
    const int bar_id = 1 + g.gid;

    QkvRaw cur = {0.f, 0.f, 0.f, 0.f};    // this item's q / k / v element(s), fetched one item ahead
    if (first < n_items) {
        att_issue_loads(p, first, pos, bufs[0], g.gtid, g.bar[0]);
        cur = att_load_qkv(p, first, g.gtid);
    }
    int k = 0;
    for (int item = first; item < n_items; item += stride, ++k) {
        const int bi = dbl ? (k & 1) : 0;
        uint8_t* buf = bufs[bi];
        const int next = item + stride;
        const bool prefetch = dbl && next < n_items;
        if (prefetch) att_issue_loads(p, next, pos, bufs[bi ^ 1], g.gtid, g.bar[bi ^ 1]);
        QkvRaw nxt = {0.f, 0.f, 0.f, 0.f};
        if (next < n_items) nxt = att_load_qkv(p, next, g.gtid);              // latency hidden behind this item's reduction
        const float a0 = cur.a0 + cur.b0, a1 = cur.a1 + cur.b1;
        const int b = item / p.H, h = item - b * p.H;
        uint8_t* Ks = buf;
        uint8_t* Vs = buf + nr * HD * 2;
        float* sc = reinterpret_cast<float*>(Vs + nr * HD * 2);
        const int* vm = reinterpret_cast<const int*>(sc + n);
        bf16* kbase = p.cache + static_cast<int64_t>(item) * p.Tmax * HD;
        bf16* vhead = kbase + static_cast<int64_t>(p.B) * p.H * p.Tmax * HD;
        if (prefetch) cp_async_wait_group<1>();        // (the validity words travel by cp.async)
        else cp_async_wait_group<0>();
        if (pos > 0) {
            ptx::mbar_wait(g.bar[bi], g.parity[bi]);   // history boxes have landed (the last one may cover row `pos` with stale
            g.parity[bi] ^= 1;                         // bytes: the new row is written only now)
        }
        // this step's q / k / v: k, v go to the cache and to row `pos` of the staged history (swizzled like the TMA rows)
        {
            const int e = g.gtid & 63, sw = ((e >> 3) ^ (pos & 7)) << 4;
            if (g.gtid < HD) {
                sq[e] = a0 * 0.125f;                                               // head_dim ** -0.5 folded into q
                const bf16 kv = __float2bfloat16(a1);
                kbase[static_cast<int64_t>(pos) * HD + e] = kv;
                *reinterpret_cast<bf16*>(Ks + pos * 128 + sw + (e & 7) * 2) = kv;
                if (p.zero != nullptr) p.zero[static_cast<int64_t>(b) * d + h * HD + e] = 0.f;
            } else {
                const bf16 vv = __float2bfloat16(a0);
                vhead[static_cast<int64_t>(pos) * HD + e] = vv;
                *reinterpret_cast<bf16*>(Vs + pos * 128 + sw + (e & 7) * 2) = vv;
            }
        }
        named_bar_sync(bar_id, 128);
        // scores: one KEY per lane (no shuffles): warp w takes the 32-key blocks w, w + 4, ...
        float q[64];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 f = *reinterpret_cast<const float4*>(sq + 4 * c);
            q[4 * c] = f.x; q[4 * c + 1] = f.y; q[4 * c + 2] = f.z; q[4 * c + 3] = f.w;
        }
        float mx = -INFINITY;
        for (int t = warp * 32 + lane; t < n; t += 128) {
            const uint8_t* row = Ks + t * 128;
            const int x = t & 7;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 u = *reinterpret_cast<const uint4*>(row + ((c ^ x) << 4));
                float2 f;
                f = unpack_bf16x2(u.x); acc0 += q[8 * c] * f.x; acc1 += q[8 * c + 1] * f.y;
                f = unpack_bf16x2(u.y); acc0 += q[8 * c + 2] * f.x; acc1 += q[8 * c + 3] * f.y;
                f = unpack_bf16x2(u.z); acc0 += q[8 * c + 4] * f.x; acc1 += q[8 * c + 5] * f.y;
                f = unpack_bf16x2(u.w); acc0 += q[8 * c + 6] * f.x; acc1 += q[8 * c + 7] * f.y;
            }
            const float sv = vm[t] ? acc0 + acc1 : -INFINITY;
            sc[t] = sv;
            mx = fmaxf(mx, sv);
        }
        mx = warp_max(mx);
        const float muse = (mx == -INFINITY) ? 0.f : mx;
        float sum = 0.f;
        for (int t = warp * 32 + lane; t < n; t += 128) {
            const float pr = __expf(sc[t] - muse);
            sc[t] = pr;
            sum += pr;
        }
        sum = warp_sum(sum);
        __syncwarp();
        // P.V: lane owns output dimensions 2 lane, 2 lane + 1; one broadcast probability + one conflict-free row read per key
        float o0 = 0.f, o1 = 0.f;
        const int vchunk = lane >> 2, vofs = (lane & 3) * 4;
        for (int t0 = warp * 32; t0 < n; t0 += 128) {
            const int cnt = min(32, n - t0);
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const int t = t0 + j;
                const float pr = sc[t];
                const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Vs + t * 128 + ((vchunk ^ (t & 7)) << 4) + vofs));
                o0 += pr * f.x;
                o1 += pr * f.y;
            }
        }
        part[warp * 64 + 2 * lane] = o0;
        part[warp * 64 + 2 * lane + 1] = o1;
        if (lane == 0) {
            part_m[warp] = mx;
            part_l[warp] = sum;
        }
        named_bar_sync(bar_id, 128);
        if (g.gtid < HD) {
            const float m = fmaxf(fmaxf(part_m[0], part_m[1]), fmaxf(part_m[2], part_m[3]));
            float l = 0.f, v = 0.f;
            if (m > -INFINITY) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float sw = (part_m[w] == -INFINITY) ? 0.f : __expf(part_m[w] - m);
                    l += part_l[w] * sw;
                    v += part[w * 64 + g.gtid] * sw;
                }
            }
            p.o[static_cast<int64_t>(b) * d + h * HD + g.gtid] = __float2bfloat16(l > 0.f ? v / l : 0.f);
        }
        named_bar_sync(bar_id, 128);           // sq / part / this buffer are rewritten by the next item
        if (!dbl && next < n_items) att_issue_loads(p, next, pos, bufs[0], g.gtid, g.bar[0]);
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------- glue phase
// x[row] += acc[row] + bias (when acc != null); u[row] = LN(x[row]) * gamma + beta (bf16); zero[row, 0..zero_n) = 0.  One row per warp.
// One row per CTA at a time, spread over all 320 threads (<= 2 float4 each): every global load of the row -- x, accumulator,
// bias, gamma, beta -- is in flight at once and the two reductions go through shared memory, so the phase costs about one L2
// round trip.  (First version: one row per warp, 8 float4 per lane with the dependent store of each vector between the loads
// of the next -- in-order issue serialised the round trips: 9 us per phase.)
__device__ __forceinline__ void glue_phase(const ChainPhase& p, int warp, int lane, float* red) {
    const int d = p.d, n4 = d >> 2;
    const int tid = threadIdx.x;
    for (int row = blockIdx.x; row < p.B; row += gridDim.x) {
        float4* __restrict__ xp = reinterpret_cast<float4*>(p.x + static_cast<size_t>(row) * d);
        const float4* __restrict__ ap = p.acc != nullptr ? reinterpret_cast<const float4*>(p.acc + static_cast<size_t>(row) * d) : nullptr;
        const float4* __restrict__ bp = reinterpret_cast<const float4*>(p.bias);
        const float4* __restrict__ gp = reinterpret_cast<const float4*>(p.gamma);
        const float4* __restrict__ tp = reinterpret_cast<const float4*>(p.beta);
        uint2* __restrict__ up = reinterpret_cast<uint2*>(p.u + static_cast<size_t>(row) * d);
        float4 v[2], a[2], b[2], gm[2], bt[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = tid + THREADS * j;
            if (c < n4) {
                v[j] = __ldcg(xp + c);
                if (ap != nullptr) {
                    a[j] = __ldcg(ap + c);
                    b[j] = __ldg(bp + c);
                }
                gm[j] = __ldg(gp + c);
                bt[j] = __ldg(tp + c);
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = tid + THREADS * j;
            if (c < n4) {
                if (ap != nullptr) {
                    v[j].x += a[j].x + b[j].x; v[j].y += a[j].y + b[j].y; v[j].z += a[j].z + b[j].z; v[j].w += a[j].w + b[j].w;
                    __stcg(xp + c, v[j]);
                }
                sum += v[j].x + v[j].y + v[j].z + v[j].w;
            }
        }
        sum = warp_sum(sum);
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) tot += red[w];
        const float mean = tot / d;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = tid + THREADS * j;
            if (c < n4) {
                const float e0 = v[j].x - mean, e1 = v[j].y - mean, e2 = v[j].z - mean, e3 = v[j].w - mean;
                sq += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
            }
        }
        sq = warp_sum(sq);
        if (lane == 0) red[16 + warp] = sq;
        __syncthreads();
        float tot2 = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) tot2 += red[16 + w];
        const float rstd = rsqrtf(tot2 / d + 1e-5f);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = tid + THREADS * j;
            if (c < n4) {
                uint2 o;
                o.x = pack_bf16x2((v[j].x - mean) * rstd * gm[j].x + bt[j].x, (v[j].y - mean) * rstd * gm[j].y + bt[j].y);
                o.y = pack_bf16x2((v[j].z - mean) * rstd * gm[j].z + bt[j].z, (v[j].w - mean) * rstd * gm[j].w + bt[j].w);
                __stcg(up + c, o);
            }
        }
        __syncthreads();                      // `red` is rewritten by the next row
    }
    if (p.zero != nullptr) {                  // [rows, zero_n] is dense: every thread of the grid clears a few float4
        float4* zp = reinterpret_cast<float4*>(p.zero);
        const int total = p.B * (p.zero_n >> 2);
        for (int i = blockIdx.x * THREADS + threadIdx.x; i < total; i += gridDim.x * THREADS) __stcg(zp + i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

// ---------------------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(THREADS, 1)
decode_chain_kernel(const ChainPhase* __restrict__ phases, int n_phases, int pos, unsigned* bar_counter, unsigned epoch0,
                    unsigned long long* trace) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const uint32_t smem_a = base, smem_b = base + STAGES * STAGE_A;
    const uint32_t smem_epi = base + PIPE_BYTES;
    const uint32_t bars = base + WORK_BYTES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
    const uint32_t tmem_slot = tempty_bar + 16;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(full_bar + 8 * i, 1);
            ptx::mbar_init(empty_bar + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar + 8 * i, 1);
            ptx::mbar_init(tempty_bar + 8 * i, EPI_WARPS);
        }
        for (int i = 0; i < 4; ++i) ptx::mbar_init(bars + ATT_BAR_OFS + 8 * i, 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async_smem();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                               // the previous kernel of the stream (greedy pick / prefill) has completed

    PipeState pipe;                           // producer's and MMA issuer's own copies advance in lock-step
    AccState accs;
    AttGroup grp;
    grp.gid = warp >= 6 ? 1 : 0;
    grp.gtid = (threadIdx.x - 64) & 127;
    grp.base = smem + grp.gid * ATT_GROUP_BYTES;
    grp.bar[0] = bars + ATT_BAR_OFS + 16 * grp.gid;
    grp.bar[1] = grp.bar[0] + 8;
    grp.parity[0] = grp.parity[1] = 0;
    const unsigned G = gridDim.x;

    if (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) trace[0] = globaltimer_ns();
    for (int ph = 0; ph < n_phases; ++ph) {
        const ChainPhase& p = phases[ph];
        if (p.type == CHAIN_GEMM) {
            if (warp == 0) {
                if (lane == 0) gemm_producer(p, pipe, smem_a, smem_b, full_bar, empty_bar);
            } else if (warp == 1) {
                if (lane == 0) gemm_mma(p, pipe, accs, smem_a, smem_b, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base);
            } else {
                gemm_epilogue(p, accs, tfull_bar, tempty_bar, tmem_base, smem_epi + (warp - 2) * EPI_BUF, warp, lane);
            }
        } else if (p.type == CHAIN_ATTN) {
            if (warp >= 2) attention_phase(p, pos, grp);
        } else {
            glue_phase(p, warp, lane, reinterpret_cast<float*>(smem + WORK_BYTES + GLUE_RED_OFS));
        }
        if (threadIdx.x < 3 && ph + 1 < n_phases && phases[ph + 1].type == CHAIN_GEMM) {
            // the TMA unit fetches a descriptor from global memory on first use: start that before the barrier
            const ChainPhase& nx = phases[ph + 1];
            ptx::prefetch_tensormap(threadIdx.x == 0 ? &nx.map_a : threadIdx.x == 1 ? &nx.map_b : &nx.map_out);
        }
        if (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) trace[1 + 2 * ph] = globaltimer_ns();      // CTA 0's own work done
        grid_sync(bar_counter, (epoch0 + static_cast<unsigned>(ph) + 1u) * G);
        if (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) trace[2 + 2 * ph] = globaltimer_ns();      // everyone's work done
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace dc

// ---------------------------------------------------------------------------------------------------------- host side
static_assert(sizeof(ChainPhase) % 128 == 0, "tensor maps inside an array of phases must stay 64-byte aligned");

bool decode_chain_supported(int max_keys) {
    // the attention phases stage one head's whole K / V history per four-warp group
    const int nr = (max_keys + dc::KV_BOX - 1) / dc::KV_BOX * dc::KV_BOX;
    return ((2 * nr * dc::HD * 2 + max_keys * 8 + 1023) & ~1023) + dc::ATT_FIXED <= dc::ATT_GROUP_BYTES;
}

void chain_gemm_phase(ChainPhase& p, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, int split, int mode, void* out,
                      int ldo, const float* bias) {
    EAVQA_CHECK(M >= 1 && M <= dc::BM, "decode chain: at most 128 rows per step");
    EAVQA_CHECK(N % dc::BN == 0 && K % 8 == 0, "decode chain GEMM: N must be a multiple of 64");
    EAVQA_CHECK(mode == CHAIN_REDUCE_F32 || mode == CHAIN_STORE_F32 || mode == CHAIN_GELU_BF16, "decode chain GEMM: unknown mode");
    EAVQA_CHECK(mode != CHAIN_GELU_BF16 || (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15) == 0), "decode chain GEMM: gelu needs a bias");
    const int num_kb = ceil_div(K, dc::BK);
    p = ChainPhase();
    p.type = CHAIN_GEMM;
    p.M = M; p.N = N; p.K = K; p.mode = mode;
    p.split = std::max(1, std::min(mode == CHAIN_REDUCE_F32 ? split : 1, num_kb));
    p.bias = bias;
    p.map_a = gemm_make_map(A, M, K, lda, dc::BM, MAP_OPERAND);
    p.map_b = gemm_make_map(W, N, K, ldw, dc::BN, MAP_OPERAND);
    p.map_out = gemm_make_map(out, M, N, ldo, 32, mode == CHAIN_GELU_BF16 ? MAP_EPI_BF16 : MAP_EPI_F32);
}

void chain_attn_phase(ChainPhase& p, const float* qkv_acc, const float* qkv_bias, bf16* cache, const int* valid, int valid_stride, bf16* o,
                      float* zero, int B, int H, int Tmax) {
    p = ChainPhase();
    p.type = CHAIN_ATTN;
    p.qkv_acc = qkv_acc; p.bias = qkv_bias; p.cache = cache; p.valid = valid; p.valid_stride = valid_stride; p.o = o; p.zero = zero;
    p.B = B; p.H = H; p.Tmax = Tmax;
    // the layer's cache as a [2 * B * H * Tmax, 64] bf16 matrix (K block, then V block), boxes of 16 rows, SWIZZLE_128B
    p.map_a = gemm_make_map(cache, 2 * B * H * Tmax, dc::HD, dc::HD, dc::KV_BOX, MAP_OPERAND);
}

void chain_glue_phase(ChainPhase& p, float* x, const float* acc, const float* bias, const float* gamma, const float* beta, bf16* u, int rows,
                      int d, float* zero, int zero_n) {
    EAVQA_CHECK(d % 4 == 0 && d <= 2048 && zero_n % 4 == 0, "decode chain glue: width must be a multiple of 4 and <= 2048");
    EAVQA_CHECK((acc == nullptr) == (bias == nullptr), "decode chain glue: acc and bias come together");
    p = ChainPhase();
    p.type = CHAIN_GLUE;
    p.x = x; p.acc = acc; p.bias = bias; p.gamma = gamma; p.beta = beta; p.u = u; p.B = rows; p.d = d; p.zero = zero; p.zero_n = zero_n;
}

void launch_decode_chain(const ChainPhase* dev_phases, int n_phases, int pos, unsigned* bar_counter, unsigned epoch0, cudaStream_t s,
                         unsigned long long* trace) {
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(dc::decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::SMEM));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms());
    cfg.blockDim = dim3(dc::THREADS);
    cfg.dynamicSmemBytes = dc::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;      // every CTA co-resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, dc::decode_chain_kernel, dev_phases, n_phases, pos, bar_counter, epoch0, trace));
    KERNEL_CHECK();
    count_launch();
}

}  // namespace eavqa
