// Device side of the persistent warp-specialised bf16 GEMM for sm_100a (host side: gemm_tcgen05.cu):
//   TMA -> shared (SWIZZLE_128B) -> tcgen05.mma (UMMA 128 x BN x 16, fp32 accumulators in TMEM, double-buffered)
//   -> tcgen05.ld -> fused epilogue in registers -> swizzled shared staging -> TMA store.
//
// Roles (320 threads, one CTA per SM):
//   warp 0 / lane 0 : TMA producer   (ring of STAGES smem slots, full/empty mbarriers)
//   warp 1 / lane 0 : MMA issuer     (also owns TMEM alloc/dealloc, whole warp)
//   warps 2..9      : epilogue       (warp w owns TMEM lanes [32*(w%4), +32) = 32 output rows, and one half of the
//                                     tile's columns; one thread = one row, 32 columns per chunk)
// The accumulator of tile i+1 is produced into the other TMEM buffer while the epilogue drains tile i.
// Every global access of the epilogue is a TMA transfer: outputs are staged row-per-thread into 128-B / 64-B
// swizzled shared tiles and stored with cp.async.bulk.tensor (fully coalesced, tails clipped by the hardware);
// the residual / activation-derivative operand is fetched the same way, one chunk ahead.
//
// The epilogue is specialised at COMPILE time (EpiMode).  Round-1 profiles (profiles/README.md): the K = 768 GEMMs of
// the LM are bound by the instruction stream of the 8 epilogue warps, not by MMA or stores; with run-time mode
// branches ptxas copied every accumulator chunk (32 MOV) to merge the branch results and the scalar math cost
// ~620 warp-instructions per 32-column chunk.  Compile-time modes + packed f32x2 math (FADD2/FMUL2/FFMA2) +
// ping-pong tcgen05.ld register blocks bring the gelu + two-output epilogue to ~190.
#pragma once
#include <atomic>
#include <string>

#include "gemm.cuh"
#include "ptx.cuh"

namespace eavqa {

// ---- host services implemented in gemm_tcgen05.cu
enum MapKind { MAP_OPERAND = 0, MAP_EPI_BF16 = 1, MAP_EPI_F32 = 2 };
CUtensorMap gemm_make_map(const void* ptr, int rows, int cols, int ld, int box_rows, int kind);
void gemm_prof_before(cudaStream_t stream, const GemmArgs& a, int bn_tag, void** token);
void gemm_prof_after(cudaStream_t stream, void* token);

// Epilogue modes (what happens to the fp32 accumulator tile before it is stored)
enum EpiMode {
    EM_BF16 = 0,            // bf16 out
    EM_BF16_BIAS,           // bf16 out, + bias
    EM_BF16_BIAS_GELU,      // bf16 out = gelu_new(acc + bias); optional second output = acc + bias (bf16)
    EM_BF16_BIAS_RELU,
    EM_BF16_BIAS_TANH,
    EM_BF16_DGELU,          // bf16 out = acc * gelu_new'(aux)
    EM_BF16_DRELU,          // bf16 out = acc * (aux > 0)
    EM_BF16_DTANH,          // bf16 out = acc * (1 - aux^2)
    EM_F32,                 // fp32 out (also split-K partials, added with TMA reduce)
    EM_F32_BIAS,            // fp32 out, + bias
    EM_F32_BIAS_RES,        // fp32 out = acc + bias + residual
    EM_CE,                  // fp16 (!) logits out + online-softmax statistics and the label's logit (LM head)
    EM_COUNT
};

namespace gk {

constexpr int BM = 128;
constexpr int BK = 64;        // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int GROUP_M = 8;    // tile rasterisation: 8 M-blocks share each B tile while it is hot in L2
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int CHUNK = 32;     // accumulator columns per epilogue step
constexpr int EPI_BUF = 4096; // one staging tile: 32 rows x 128 B
constexpr int CLC_STAGES = 6; // in-flight "next work item" responses (producer <= MMA + 1 tile <= epilogue + 3 tiles)
// offsets inside the barrier block (1024-B aligned): pipeline barriers first (<= 208 B), then the CLC ring
constexpr int CLC_RESP_OFS = 208;                          // CLC_STAGES x 16 B responses
constexpr int CLC_FULL_OFS = CLC_RESP_OFS + 16 * CLC_STAGES;
constexpr int CLC_EMPTY_OFS = CLC_FULL_OFS + 8 * CLC_STAGES;

template <int MODE>
struct Epi {
    static constexpr bool ce = MODE == EM_CE;
    static constexpr bool out_f32 = MODE == EM_F32 || MODE == EM_F32_BIAS || MODE == EM_F32_BIAS_RES;
    static constexpr bool bias = MODE == EM_BF16_BIAS || MODE == EM_BF16_BIAS_GELU || MODE == EM_BF16_BIAS_RELU ||
                                 MODE == EM_BF16_BIAS_TANH || MODE == EM_F32_BIAS || MODE == EM_F32_BIAS_RES;
    static constexpr int act = MODE == EM_BF16_BIAS_GELU ? ACT_GELU_NEW : MODE == EM_BF16_BIAS_RELU ? ACT_RELU
                               : MODE == EM_BF16_BIAS_TANH ? ACT_TANH : ACT_NONE;
    static constexpr int dact = MODE == EM_BF16_DGELU ? DACT_GELU_NEW : MODE == EM_BF16_DRELU ? DACT_RELU
                                : MODE == EM_BF16_DTANH ? DACT_TANH : DACT_NONE;
    static constexpr bool res = MODE == EM_F32_BIAS_RES;
    static constexpr bool aux = dact != DACT_NONE;
    static constexpr bool has_in = res || aux;      // an operand tile is TMA-fetched into the staging buffer
};

template <int BN>
struct Cfg {
    static constexpr int STAGE_A = BM * BK * 2;
    static constexpr int STAGE_B = BN * BK * 2;
    static constexpr int STAGES = (BN == 256) ? 3 : (BN == 192) ? 4 : (BN == 128) ? 5 : 6;
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int HALF = BN / 2;              // columns per epilogue warp
    static constexpr int NCHUNK = HALF / CHUNK;
    static constexpr int EPI_SMEM = EPI_WARPS * 2 * EPI_BUF;     // per warp: bufA (out) + bufB (in / out2)
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM = STAGES * (STAGE_A + STAGE_B) + EPI_SMEM + BAR_BYTES + 1024;
    static_assert(16 * STAGES + 32 + 8 * EPI_WARPS + 4 <= CLC_RESP_OFS && CLC_EMPTY_OFS + 8 * CLC_STAGES <= BAR_BYTES, "barrier block layout");
    static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "UMMA N / epilogue split");
    static_assert(STAGE_B % 1024 == 0, "B stage must keep 1024-B alignment");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// byte offset of 16-byte unit j of row r inside a TMA-swizzled staging tile
__device__ __forceinline__ uint32_t swz128(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }        // SWIZZLE_128B, 128-B rows
__device__ __forceinline__ uint32_t swz64(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }    // SWIZZLE_64B, 64-B rows

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& m_idx, int& n_idx) {
    const int per_group = GROUP_M * num_n;
    const int group = tile / per_group;
    const int first_m = group * GROUP_M;
    const int gsize = min(num_m - first_m, GROUP_M);
    const int in_group = tile - group * per_group;
    m_idx = first_m + in_group % gsize;
    n_idx = in_group / gsize;
}

struct TmaMaps {
    CUtensorMap a, b, out, out2, in;
};

// Dynamic persistent scheduling (cluster launch control, ptx.cuh).  The grid has one CTA (pair) per work item; a CTA starts
// with its own block index and then keeps asking the hardware for the index of a CTA that has not been launched yet.
// The producer thread issues one query per work item it starts (so the answer is there long before anyone needs it);
// every role thread / epilogue warp reads the same sequence of answers from a ring in shared memory.
struct WorkFeed {
    uint32_t resp, full, empty;     // shared addresses of the ring (responses, full / empty barriers)
    int slot = 0;
    uint32_t phase = 0;
    bool pair;                      // CTA pair: answers are multicast to both CTAs, `empty` lives in the leader
    __device__ __forceinline__ WorkFeed(uint32_t bars, bool pair_) : resp(bars + CLC_RESP_OFS), full(bars + CLC_FULL_OFS),
                                                                     empty(bars + CLC_EMPTY_OFS), pair(pair_) {}
    __device__ __forceinline__ void advance() {
        if (++slot == CLC_STAGES) { slot = 0; phase ^= 1; }
    }
    // producer side: request the next work item (leader only talks to the hardware; every CTA arms its own barrier)
    __device__ __forceinline__ void request(bool leader) {
        if (leader) ptx::mbar_wait(empty + 8 * slot, phase ^ 1);          // every reader is done with this slot's last answer
        ptx::mbar_arrive_expect_tx(full + 8 * slot, 16);
        if (leader) {
            if (pair) ptx::clc_try_cancel_multicast(resp + 16 * slot, full + 8 * slot);
            else ptx::clc_try_cancel(resp + 16 * slot, full + 8 * slot);
        }
        advance();
    }
    // consumer side: block index of the next work item's first CTA, or -1 when the grid is exhausted.
    // `arrive` = this thread reports the slot as read (one thread per role / epilogue warp).
    __device__ __forceinline__ int next(bool arrive) {
        ptx::mbar_wait(full + 8 * slot, phase);
        const int id = ptx::clc_decode(resp + 16 * slot);
        ptx::fence_proxy_async_smem();                                     // generic read before the async proxy rewrites the slot
        if (arrive) {
            if (pair) ptx::mbar_arrive_cluster(empty + 8 * slot, 0);
            else ptx::mbar_arrive(empty + 8 * slot);
        }
        advance();
        return id;
    }
};

// ---------------------------------------------------------------------------------------------
// epilogue role (8 warps), shared by the 1-CTA and the 2-CTA (cta_group::2) kernels
// ---------------------------------------------------------------------------------------------
template <int BN, int MODE, class Coords, class Release>
__device__ __forceinline__ void epilogue_role(const TmaMaps& maps, const GemmEpilogue& ep, int M, int N, int num_n, int first_tile,
                                              uint32_t bars, bool pair, Coords coords, Release release_tmem,
                                              uint32_t tmem_base, uint32_t tfull_bar, uint32_t smem_epi, uint32_t in_bar0,
                                              int warp, int lane) {
    using C = Cfg<BN>;
    using E = Epi<MODE>;
    const int ew = warp - 2;                      // 0..7
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = ew >> 2;                     // which half of the tile's columns
    const uint32_t bufA = smem_epi + ew * 2 * EPI_BUF;
    const uint32_t bufB = bufA + EPI_BUF;
    const uint32_t in_bar = in_bar0 + 8 * ew;
    constexpr uint32_t in_bytes = E::res ? 32u * 128u : 32u * 64u;
    uint32_t in_phase = 0;
    uint32_t out_slot = 0;                        // bf16 outputs alternate between two 2-KB halves of the buffers
    pdl_wait();           // first global access of these warps comes next (operand prefetch, bias, labels, stores)

    auto n_valid_chunks = [&](int n_idx) {
        const int col0 = n_idx * BN + half * C::HALF;
        const int rem = N - col0;
        return rem <= 0 ? 0 : min(C::NCHUNK, (rem + CHUNK - 1) / CHUNK);
    };
    auto issue_in = [&](int tile, int c) {        // lane 0 only
        int m_idx, n_idx;
        coords(tile, m_idx, n_idx);
        ptx::mbar_arrive_expect_tx(in_bar, in_bytes);
        ptx::tma_load_2d(bufB, &maps.in, in_bar, n_idx * BN + half * C::HALF + c * CHUNK, m_idx * BM + quarter * 32);
    };
    // the operand of a tile's first chunk is requested when the tile becomes known (below, at the end of the previous tile)
    auto prefetch_in = [&](int tile) {            // lane 0 only
        int m_idx, n_idx;
        coords(tile, m_idx, n_idx);
        if (n_valid_chunks(n_idx) > 0) issue_in(tile, 0);
    };
    if (E::has_in && lane == 0) prefetch_in(first_tile);
    WorkFeed feed(bars, pair);

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile;;) {
        int m_idx, n_idx;
        coords(tile, m_idx, n_idx);
        const int row0 = m_idx * BM + quarter * 32;
        const int row = row0 + lane;
        const bool row_ok = row < M;
        const int col_base = n_idx * BN + half * C::HALF;
        const int nvalid = n_valid_chunks(n_idx);

        ptx::mbar_wait(tfull_bar + 8 * acc, acc_phase);
        ptx::tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * C::HALF;
        if (nvalid == 0) {
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) release_tmem(acc);
        }
        float ce_m = -INFINITY, ce_s = 0.f;
        int label = -1;
        if (E::ce && row_ok && ep.ce_label != nullptr) label = __ldg(ep.ce_label + row);

        // One 32-column chunk of this thread's row.  `r` holds the chunk (tcgen05.ld issued one chunk earlier), `rn`
        // receives the next one while this chunk's math runs: the two register blocks ping-pong, no copies.
        // All element-wise math is done on packed float pairs (FADD2 / FMUL2 / FFMA2).
        auto do_chunk = [&](const int c, uint32_t (&r)[32], uint32_t (&rn)[32]) {
            const int n0 = col_base + c * CHUNK;
            ptx::tmem_ld_wait();
            f32x2 p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = pk(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            if (c + 1 < nvalid) {
                ptx::tmem_ld_32x32(taddr + (c + 1) * CHUNK, rn);          // overlaps this chunk's math / staging
            } else {
                ptx::tcgen05_fence_before();                              // accumulator fully read: release the TMEM buffer
                __syncwarp();
                if (lane == 0) release_tmem(acc);
            }
            if (E::bias) {
                // every lane reads the same 16 bytes: one broadcast transaction per load, served by L1.  (Issuing these
                // before the TMEM wait hides ~35 clk per chunk but costs 32 live registers -> spills at the 168-register
                // cap of a 10-warp CTA; measured not worth it.)
                if (n0 + CHUNK <= N) {
                    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = __ldg(b4 + q);
                        p[2 * q] = add2(p[2 * q], pk(b.x, b.y));
                        p[2 * q + 1] = add2(p[2 * q + 1], pk(b.z, b.w));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float b0 = (n0 + 2 * j < N) ? __ldg(ep.bias + n0 + 2 * j) : 0.f;
                        const float b1 = (n0 + 2 * j + 1 < N) ? __ldg(ep.bias + n0 + 2 * j + 1) : 0.f;
                        p[j] = add2(p[j], pk(b0, b1));
                    }
                }
            }
            if (E::ce) {
                // running max / sum-exp over the valid vocabulary columns; the label's logit in fp32
                float v[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) upk(p[j], v[2 * j], v[2 * j + 1]);
                const float log2e = 1.4426950408889634f;
                if (n0 + CHUNK <= ep.n_valid) {       // warp-uniform: every column of the chunk is a vocabulary entry
                    float m4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
                    for (int j = 4; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], v[j]);
                    const float nm = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), ce_m);
                    const f32x2 sc = pk1(log2e), off = pk1(-nm * log2e);
                    f32x2 a4[4] = {pk1(0.f), pk1(0.f), pk1(0.f), pk1(0.f)};
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float e0, e1;
                        upk(fma2(p[j], sc, off), e0, e1);
                        a4[j & 3] = add2(a4[j & 3], pk(exp2f(e0), exp2f(e1)));
                    }
                    float s0, s1;
                    upk(add2(add2(a4[0], a4[1]), add2(a4[2], a4[3])), s0, s1);
                    ce_s = ce_s * exp2f((ce_m - nm) * log2e) + (s0 + s1);
                    ce_m = nm;
                } else {
                    float cm = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + j < ep.n_valid) cm = fmaxf(cm, v[j]);
                    if (cm > -INFINITY) {
                        const float nm = fmaxf(ce_m, cm);
                        float a = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < ep.n_valid) a += exp2f((v[j] - nm) * log2e);
                        ce_s = ce_s * exp2f((ce_m - nm) * log2e) + a;
                        ce_m = nm;
                    }
                }
                if (row_ok && label >= n0 && label < n0 + 32) {
                    float t = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + j == label) t = v[j];
                    ep.ce_target[row] = t;
                }
            }
            const uint32_t slot_off = E::out_f32 ? 0u : (out_slot & 1u) * 2048u;
            // staging buffers of this slot must have been read out by their previous TMA store
            if (lane == 0) {
                if (E::out_f32) tma_store_wait_read<0>();
                else tma_store_wait_read<1>();
            }
            __syncwarp();
            if (E::act == ACT_GELU_NEW) {
                if (ep.out2 != nullptr) {         // pre-activation, bf16 (bufB is free: this mode has no input operand)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_v4(bufB + slot_off + swz64(lane, q), pack_bf16x2(p[q * 4 + 0]), pack_bf16x2(p[q * 4 + 1]),
                                     pack_bf16x2(p[q * 4 + 2]), pack_bf16x2(p[q * 4 + 3]));
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) p[j] = gelu_new2(p[j]);
            } else if (E::act == ACT_TANH) {
#pragma unroll
                for (int j = 0; j < 16; ++j) p[j] = tanh2(p[j]);
            } else if (E::act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float lo, hi;
                    upk(p[j], lo, hi);
                    p[j] = pk(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
                }
            }
            if (E::has_in) {
                ptx::mbar_wait(in_bar, in_phase);
                in_phase ^= 1;
                if (E::res) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint4 u = ld_shared_v4(bufB + swz128(lane, q));
                        p[2 * q] = add2(p[2 * q], pk(__uint_as_float(u.x), __uint_as_float(u.y)));
                        p[2 * q + 1] = add2(p[2 * q + 1], pk(__uint_as_float(u.z), __uint_as_float(u.w)));
                    }
                } else {
                    uint32_t a[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 u = ld_shared_v4(bufB + swz64(lane, q));
                        a[q * 4 + 0] = u.x; a[q * 4 + 1] = u.y; a[q * 4 + 2] = u.z; a[q * 4 + 3] = u.w;
                    }
                    if (E::dact == DACT_GELU_NEW) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) p[j] = mul2(p[j], gelu_new_grad2(bf16x2_to_f32x2(a[j])));
                    } else if (E::dact == DACT_TANH) {
                        const f32x2 one = pk1(1.0f);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const f32x2 t = bf16x2_to_f32x2(a[j]);
                            p[j] = mul2(p[j], sub2(one, mul2(t, t)));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float lo, hi, alo, ahi;
                            upk(p[j], lo, hi);
                            upk(bf16x2_to_f32x2(a[j]), alo, ahi);
                            p[j] = pk(alo > 0.f ? lo : 0.f, ahi > 0.f ? hi : 0.f);
                        }
                    }
                }
                __syncwarp();                     // every lane has consumed bufB: fetch the next chunk's operand
                if (lane == 0 && c + 1 < nvalid) issue_in(tile, c + 1);
            }
            const bool store_out = !E::ce || ep.out != nullptr;      // loss-only evaluation keeps no logits
            if (ep.debug != 3 && store_out) {
                if (E::out_f32) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float x0, x1, x2, x3;
                        upk(p[2 * q], x0, x1);
                        upk(p[2 * q + 1], x2, x3);
                        st_shared_v4(bufA + swz128(lane, q), __float_as_uint(x0), __float_as_uint(x1), __float_as_uint(x2),
                                     __float_as_uint(x3));
                    }
                } else if (E::ce) {
                    // the stored logits are fp16, not bf16 (common.cuh: pack_f16x2); same 2-byte tiles, same TMA store
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_v4(bufA + slot_off + swz64(lane, q), pack_f16x2(p[q * 4 + 0]), pack_f16x2(p[q * 4 + 1]),
                                     pack_f16x2(p[q * 4 + 2]), pack_f16x2(p[q * 4 + 3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_v4(bufA + slot_off + swz64(lane, q), pack_bf16x2(p[q * 4 + 0]), pack_bf16x2(p[q * 4 + 1]),
                                     pack_bf16x2(p[q * 4 + 2]), pack_bf16x2(p[q * 4 + 3]));
                }
            }
            ptx::fence_proxy_async_smem();        // generic-proxy writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0 && store_out && !(ep.debug == 2 || ep.debug == 3 || (ep.debug == 1 && (out_slot & 1)))) {
                if (MODE == EM_F32 && ep.split_k > 1) tma_reduce_add_2d(&maps.out, bufA + slot_off, n0, row0);
                else tma_store_2d(&maps.out, bufA + slot_off, n0, row0);
                if (E::act == ACT_GELU_NEW && ep.out2 != nullptr) tma_store_2d(&maps.out2, bufB + slot_off, n0, row0);
                tma_store_commit();
            }
            ++out_slot;
        };

        uint32_t ra[32], rb[32];
        if (nvalid > 0) ptx::tmem_ld_32x32(taddr, ra);
#pragma unroll
        for (int c = 0; c < C::NCHUNK; ++c) {
            if (c >= nvalid) break;               // warp-uniform
            if (c & 1) do_chunk(c, rb, ra);
            else do_chunk(c, ra, rb);
        }
        if (E::ce && row_ok && n_idx < num_n)
            ep.ce_partial[static_cast<size_t>(row) * ep.ce_tiles + n_idx * 2 + half] = make_float2(ce_m, ce_s);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        // next work item (every lane reads the same answer; lane 0 reports the slot as read)
        const int id = feed.next(lane == 0);
        if (id < 0) break;
        tile = pair ? (id >> 1) : id;
        if (E::has_in && lane == 0) prefetch_in(tile);
    }
    if (lane == 0) tma_store_wait_read<0>();      // staging smem must outlive the last bulk stores
}

// ---------------------------------------------------------------------------------------------
// 1-CTA kernel.  MN = operands stored [K, MN] (weight-gradient form dW = dY^T X), read as MN-major UMMA operands.
// (Round 1 also measured 2x1 / 2x2 clusters with TMA multicast of the shared operand tile: within +-5 % / 30-45 %
//  slower, i.e. L2 -> SM operand delivery is not the limiter; those variants were removed.)
// ---------------------------------------------------------------------------------------------
template <int BN, int MODE, bool MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ TmaMaps maps, int M, int N, int K, const GemmEpilogue ep) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-B aligned bases (descriptor base_offset = 0)
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);

    const uint32_t smem_a = base;
    const uint32_t smem_b = base + C::STAGES * C::STAGE_A;
    const uint32_t smem_epi = smem_b + C::STAGES * C::STAGE_B;          // 1024-aligned (stage sizes are)
    const uint32_t bars = smem_epi + C::EPI_SMEM;
    const uint32_t full_bar = bars;                       // STAGES x 8 B
    const uint32_t empty_bar = bars + 8 * C::STAGES;      // STAGES x 8 B
    const uint32_t tfull_bar = bars + 16 * C::STAGES;     // 2 x 8 B
    const uint32_t tempty_bar = tfull_bar + 16;           // 2 x 8 B
    const uint32_t in_bar0 = tempty_bar + 16;             // EPI_WARPS x 8 B
    const uint32_t tmem_slot = in_bar0 + 8 * EPI_WARPS;   // 4 B
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();        // the next kernel may be scheduled; it blocks in its own pdl_wait() until this grid completes

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&maps.a);
        ptx::prefetch_tensormap(&maps.b);
        ptx::prefetch_tensormap(&maps.out);
        for (int i = 0; i < C::STAGES; ++i) {
            ptx::mbar_init(full_bar + 8 * i, 1);
            ptx::mbar_init(empty_bar + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar + 8 * i, 1);
            ptx::mbar_init(tempty_bar + 8 * i, EPI_WARPS);     // one arrive per epilogue warp
        }
        for (int i = 0; i < EPI_WARPS; ++i) ptx::mbar_init(in_bar0 + 8 * i, 1);
        for (int i = 0; i < CLC_STAGES; ++i) {
            ptx::mbar_init(bars + CLC_FULL_OFS + 8 * i, 1);
            ptx::mbar_init(bars + CLC_EMPTY_OFS + 8 * i, 2 + EPI_WARPS);       // producer + MMA + epilogue warps
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async_smem();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int split = ep.split_k > 1 ? ep.split_k : 1;      // split-K: `split` consecutive work items share an output tile
    const int num_kb = (K + BK - 1) / BK;
    const int first_tile = blockIdx.x;                      // the grid has one CTA per work item (see WorkFeed)
    auto coords = [&](int tile, int& m_idx, int& n_idx) { tile_coords(tile / split, num_m, num_n, m_idx, n_idx); };
    auto kb_range = [&](int tile, int& kb0, int& kb1) {     // balanced, never empty (split <= num_kb)
        const int sp = tile % split;
        kb0 = static_cast<int>(static_cast<int64_t>(sp) * num_kb / split);
        kb1 = static_cast<int>(static_cast<int64_t>(sp + 1) * num_kb / split);
    };

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            pdl_wait();       // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel
            int stage = 0;
            uint32_t phase = 0;
            WorkFeed req(bars, false), feed(bars, false);
            for (int tile = first_tile; tile >= 0; tile = feed.next(true)) {
                req.request(true);                                        // ask for the work item after this one
                int m_idx, n_idx, kb0, kb1;
                coords(tile, m_idx, n_idx);
                kb_range(tile, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    ptx::mbar_arrive_expect_tx(full_bar + 8 * stage, C::STAGE_A + C::STAGE_B);
                    if (MN) {
                        // operands stored [K, MN]: boxes of 64 (MN) x 64 (K rows), one per 64 columns of the tile
#pragma unroll
                        for (int bx = 0; bx < BM / 64; ++bx)
                            ptx::tma_load_2d(smem_a + stage * C::STAGE_A + bx * 8192, &maps.a, full_bar + 8 * stage,
                                             m_idx * BM + bx * 64, kb * BK);
#pragma unroll
                        for (int bx = 0; bx < BN / 64; ++bx)
                            ptx::tma_load_2d(smem_b + stage * C::STAGE_B + bx * 8192, &maps.b, full_bar + 8 * stage,
                                             n_idx * BN + bx * 64, kb * BK);
                    } else {
                        ptx::tma_load_2d(smem_a + stage * C::STAGE_A, &maps.a, full_bar + 8 * stage, kb * BK, m_idx * BM);
                        ptx::tma_load_2d(smem_b + stage * C::STAGE_B, &maps.b, full_bar + 8 * stage, kb * BK, n_idx * BN);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM, BN, MN, MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            WorkFeed feed(bars, false);
            for (int tile = first_tile; tile >= 0; tile = feed.next(true)) {
                ptx::mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);     // epilogue drained this buffer
                ptx::tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                int kb0, kb1;
                kb_range(tile, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(full_bar + 8 * stage, phase);          // TMA bytes landed
                    ptx::tcgen05_fence_after();
                    const uint64_t da = MN ? ptx::make_mnmajor_sw128_desc(smem_a + stage * C::STAGE_A)
                                           : ptx::make_kmajor_sw128_desc(smem_a + stage * C::STAGE_A);
                    const uint64_t db = MN ? ptx::make_mnmajor_sw128_desc(smem_b + stage * C::STAGE_B)
                                           : ptx::make_kmajor_sw128_desc(smem_b + stage * C::STAGE_B);
                    // one UMMA_K = 16 step: K-major +32 B inside the 128-B swizzle row; MN-major +16 rows = 2048 B
                    constexpr uint32_t kstep = MN ? (2048u >> 4) : (32u >> 4);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        ptx::umma_bf16(tmem_d, da + kstep * k, db + kstep * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(empty_bar + 8 * stage);              // frees the smem slot when the MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(tfull_bar + 8 * acc);                    // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue warps =====================
        epilogue_role<BN, MODE>(maps, ep, M, N, num_n, first_tile, bars, false, coords,
                                [&](int acc) { ptx::mbar_arrive(tempty_bar + 8 * acc); }, tmem_base, tfull_bar, smem_epi, in_bar0,
                                warp, lane);
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// 2-CTA variant: a CTA pair (cluster of 2, adjacent SMs) computes one 256 x BN tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and only HALF of the B tile (BN/2 rows); the pair's tensor cores read both
// halves, so per CTA the shared-memory traffic per 64-deep K block drops from 2 x (16 + BN/8) KB to 2 x (16 + BN/16) KB
// (BN = 256: 96 -> 64 KB per 512 MMA clocks) -- the port that capped the 1-CTA kernel at ~55-60 % of the tensor pipe.
// The leader CTA (rank 0) issues every MMA; both CTAs run producer and epilogue roles on their own rows.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct Cfg2 {
    static constexpr int STAGE_A = BM * BK * 2;
    static constexpr int STAGE_B = (BN / 2) * BK * 2;
    static constexpr int STAGES = (BN == 256) ? 5 : (BN == 192) ? 5 : 6;
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int EPI_SMEM = EPI_WARPS * 2 * EPI_BUF;
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM = STAGES * (STAGE_A + STAGE_B) + EPI_SMEM + BAR_BYTES + 1024;
    static_assert(16 * STAGES + 32 + 8 * EPI_WARPS + 4 <= CLC_RESP_OFS, "barrier block layout");
    static_assert(BN == 128 || BN == 192 || BN == 256, "pair tile width");
    static_assert(STAGE_B % 1024 == 0, "B stage must keep 1024-B alignment");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int BN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ TmaMaps maps, int M, int N, int K, const GemmEpilogue ep) {
    using C = Cfg2<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);
    const uint32_t rank = ptx::cluster_ctarank();      // 0 = leader
    const bool leader = rank == 0;

    const uint32_t smem_a = base;
    const uint32_t smem_b = base + C::STAGES * C::STAGE_A;
    const uint32_t smem_epi = smem_b + C::STAGES * C::STAGE_B;
    const uint32_t bars = smem_epi + C::EPI_SMEM;
    const uint32_t full_bar = bars;                       // leader's is used: bytes of BOTH CTAs land on it
    const uint32_t empty_bar = bars + 8 * C::STAGES;      // per CTA: released by the leader's multicast commit
    const uint32_t tfull_bar = bars + 16 * C::STAGES;     // per CTA: accumulator ready (multicast commit)
    const uint32_t tempty_bar = tfull_bar + 16;           // leader's is used: 2 x EPI_WARPS arrivals
    const uint32_t in_bar0 = tempty_bar + 16;
    const uint32_t tmem_slot = in_bar0 + 8 * EPI_WARPS;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&maps.a);
        ptx::prefetch_tensormap(&maps.b);
        ptx::prefetch_tensormap(&maps.out);
        for (int i = 0; i < C::STAGES; ++i) {
            ptx::mbar_init(full_bar + 8 * i, 1);
            ptx::mbar_init(empty_bar + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar + 8 * i, 1);
            ptx::mbar_init(tempty_bar + 8 * i, 2 * EPI_WARPS);
        }
        for (int i = 0; i < EPI_WARPS; ++i) ptx::mbar_init(in_bar0 + 8 * i, 1);
        for (int i = 0; i < CLC_STAGES; ++i) {
            ptx::mbar_init(bars + CLC_FULL_OFS + 8 * i, 1);
            ptx::mbar_init(bars + CLC_EMPTY_OFS + 8 * i, 3 + 2 * EPI_WARPS);   // leader's: 2 producers + 1 MMA + 2 x 8 epilogue warps
        }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async_smem();
    }
    if (warp == 1) {                                      // the same warp of BOTH CTAs allocates collectively
        ptx::tmem_alloc_2cta(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish_2cta();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_pm = (num_m + 1) / 2;                   // pair tiles along M (256 rows)
    const int num_kb = (K + BK - 1) / BK;
    const int first_tile = blockIdx.x / 2;                // the grid has one CTA pair per tile (see WorkFeed)
    auto coords = [&](int tile, int& m_idx, int& n_idx) {
        int pm, pn;
        tile_coords(tile, num_pm, num_n, pm, pn);
        m_idx = pm * 2 + static_cast<int>(rank);
        n_idx = pn;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer (both CTAs) =====================
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            WorkFeed req(bars, true), feed(bars, true);
            for (int tile = first_tile, id = 0; id >= 0; id = feed.next(true), tile = id >> 1) {
                req.request(leader);                                      // the leader asks; both CTAs receive the answer
                int m_idx, n_idx;
                coords(tile, m_idx, n_idx);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    if (leader) ptx::mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * (C::STAGE_A + C::STAGE_B));
                    ptx::tma_load_2d_2cta(smem_a + stage * C::STAGE_A, &maps.a, full_bar + 8 * stage, kb * BK, m_idx * BM);
                    ptx::tma_load_2d_2cta(smem_b + stage * C::STAGE_B, &maps.b, full_bar + 8 * stage, kb * BK,
                                          n_idx * BN + static_cast<int>(rank) * (BN / 2));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ===================== MMA issuer (leader CTA only) =====================
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            WorkFeed feed(bars, true);
            for (int id = 0; id >= 0; id = feed.next(true)) {             // the MMA loop does not need the tile coordinates
                ptx::mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);     // both CTAs' epilogues drained this buffer
                ptx::tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(full_bar + 8 * stage, phase);          // both CTAs' bytes landed
                    ptx::tcgen05_fence_after();
                    const uint64_t da = ptx::make_kmajor_sw128_desc(smem_a + stage * C::STAGE_A);
                    const uint64_t db = ptx::make_kmajor_sw128_desc(smem_b + stage * C::STAGE_B);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        ptx::umma_bf16_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit_2cta(empty_bar + 8 * stage, 0x3);   // frees the slot in both CTAs
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_2cta(tfull_bar + 8 * acc, 0x3);         // accumulator ready in both CTAs
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs, own 128 rows) =====================
        epilogue_role<BN, MODE>(maps, ep, M, N, num_n, first_tile, bars, true, coords,
                                [&](int acc) { ptx::mbar_arrive_cluster(tempty_bar + 8 * acc, 0); }, tmem_base, tfull_bar, smem_epi,
                                in_bar0, warp, lane);
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// host launchers (templates; instantiated per tile width in gemm_inst_*.cu)
// ---------------------------------------------------------------------------------------------
inline void fill_epi_maps(const GemmArgs& a, bool out_f32, TmaMaps& maps) {
    const GemmEpilogue& e = a.ep;
    maps.out = e.out ? gemm_make_map(e.out, a.M, a.N, e.ldo, 32, out_f32 ? MAP_EPI_F32 : MAP_EPI_BF16) : maps.a;
    maps.out2 = e.out2 ? gemm_make_map(e.out2, a.M, a.N, e.ldo2, 32, MAP_EPI_BF16) : maps.a;
    if (e.residual) maps.in = gemm_make_map(e.residual, a.M, a.N, e.ld_res, 32, MAP_EPI_F32);
    else if (e.dact != DACT_NONE) maps.in = gemm_make_map(e.aux, a.M, a.N, e.ld_aux, 32, MAP_EPI_BF16);
    else maps.in = maps.a;
}

template <class Kernel>
inline void launch_with_attrs(Kernel kernel, int grid, int smem, int cluster, cudaStream_t stream, const TmaMaps& maps,
                              const GemmArgs& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, maps, a.M, a.N, a.K, a.ep));
    KERNEL_CHECK();
}

template <int BN, int MODE, bool MN>
void launch_1cta(const GemmArgs& a, cudaStream_t stream) {
    using C = Cfg<BN>;
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, MODE, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    TmaMaps maps;
    if (MN) {
        // At [K, M] / Bt [K, N] row-major: rows = contraction index, 64-column x 64-row boxes
        maps.a = gemm_make_map(a.A, a.K, a.M, a.lda, 64, MAP_OPERAND);
        maps.b = gemm_make_map(a.B, a.K, a.N, a.ldb, 64, MAP_OPERAND);
    } else {
        maps.a = gemm_make_map(a.A, a.M, a.K, a.lda, BM, MAP_OPERAND);
        maps.b = gemm_make_map(a.B, a.N, a.K, a.ldb, BN, MAP_OPERAND);
    }
    fill_epi_maps(a, Epi<MODE>::out_f32, maps);
    const int split = a.ep.split_k > 1 ? a.ep.split_k : 1;
    const int grid = ceil_div(a.M, BM) * ceil_div(a.N, BN) * split;      // one CTA per work item; running CTAs cancel + adopt the rest
    void* token = nullptr;
    gemm_prof_before(stream, a, BN, &token);
    launch_with_attrs(gemm_bf16_tn_kernel<BN, MODE, MN>, grid, C::SMEM, 1, stream, maps, a);
    gemm_prof_after(stream, token);
}

template <int BN, int MODE>
void launch_2cta(const GemmArgs& a, cudaStream_t stream) {
    using C = Cfg2<BN>;
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    TmaMaps maps;
    maps.a = gemm_make_map(a.A, a.M, a.K, a.lda, BM, MAP_OPERAND);
    maps.b = gemm_make_map(a.B, a.N, a.K, a.ldb, BN / 2, MAP_OPERAND);      // each CTA of the pair stages half of the B tile
    fill_epi_maps(a, Epi<MODE>::out_f32, maps);
    const int grid = 2 * ceil_div(ceil_div(a.M, BM), 2) * ceil_div(a.N, BN);      // one CTA pair per 256 x BN tile
    void* token = nullptr;
    gemm_prof_before(stream, a, BN + 1000, &token);
    launch_with_attrs(gemm_bf16_tn_2cta_kernel<BN, MODE>, grid, C::SMEM, 2, stream, maps, a);
    gemm_prof_after(stream, token);
}

// one switch per tile width; PAIR = also instantiate the cta_group::2 kernels for this width
template <int BN, bool PAIR>
void dispatch_bn(int mode, int kind, const GemmArgs& a, cudaStream_t s) {     // kind: 0 = 1-CTA, 1 = 1-CTA MN-major, 2 = CTA pair
    if (kind == 1) {
        EAVQA_CHECK(mode == EM_F32, "MN-major (wgrad form) GEMM: plain fp32 epilogue only");
        launch_1cta<BN, EM_F32, true>(a, s);
        return;
    }
#define EAVQA_MODE_CASE(MODE_)                                         \
    case MODE_:                                                        \
        if constexpr (PAIR) {                                          \
            if (kind == 2) { launch_2cta<BN, MODE_>(a, s); break; }    \
        }                                                              \
        launch_1cta<BN, MODE_, false>(a, s);                           \
        break;
    switch (mode) {
        EAVQA_MODE_CASE(EM_BF16)
        EAVQA_MODE_CASE(EM_BF16_BIAS)
        EAVQA_MODE_CASE(EM_BF16_BIAS_GELU)
        EAVQA_MODE_CASE(EM_BF16_BIAS_RELU)
        EAVQA_MODE_CASE(EM_BF16_BIAS_TANH)
        EAVQA_MODE_CASE(EM_BF16_DGELU)
        EAVQA_MODE_CASE(EM_BF16_DRELU)
        EAVQA_MODE_CASE(EM_BF16_DTANH)
        EAVQA_MODE_CASE(EM_F32)
        EAVQA_MODE_CASE(EM_F32_BIAS)
        EAVQA_MODE_CASE(EM_F32_BIAS_RES)
        EAVQA_MODE_CASE(EM_CE)
        default: throw Error("GEMM: unknown epilogue mode");
    }
#undef EAVQA_MODE_CASE
}

}  // namespace gk

// defined in gemm_inst_{64,128,192,256}.cu
void gemm_dispatch_bn64(int mode, int kind, const GemmArgs& a, cudaStream_t s);
void gemm_dispatch_bn128(int mode, int kind, const GemmArgs& a, cudaStream_t s);
void gemm_dispatch_bn192(int mode, int kind, const GemmArgs& a, cudaStream_t s);
void gemm_dispatch_bn256(int mode, int kind, const GemmArgs& a, cudaStream_t s);

}  // namespace eavqa
