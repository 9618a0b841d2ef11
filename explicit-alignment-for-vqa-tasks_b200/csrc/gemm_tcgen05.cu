// Persistent warp-specialised bf16 GEMM for sm_100a:
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
// the residual / activation-derivative operand is fetched the same way, one chunk ahead.  (Round-1 profile: the
// first version stored straight from registers, one row per thread -> 32 LSU wavefronts per store instruction,
// and K=768 GEMMs ran at 340-470 TFLOP/s; see profiles/.)
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "gemm.cuh"
#include "ptx.cuh"

namespace eavqa {

static std::atomic<int64_t> g_gemm_launches{0};
int64_t gemm_launch_count() { return g_gemm_launches.load(); }

struct ProfRec {
    cudaEvent_t start, stop;
    int M, N, K, bn;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

void gemm_profile_begin() {
    g_prof.clear();
    g_prof_on = true;
}
void gemm_profile_end(double* total_ms, double* total_flops, int64_t* launches, std::string* report) {
    g_prof_on = false;
    CUDA_CHECK(cudaDeviceSynchronize());
    std::map<std::tuple<int, int, int, int>, std::pair<double, int>> by_shape;
    double ms_sum = 0, fl_sum = 0;
    for (auto& r : g_prof) {
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, r.start, r.stop));
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
        ms_sum += ms;
        fl_sum += 2.0 * r.M * r.N * r.K;
        auto& e = by_shape[std::make_tuple(r.M, r.N, r.K, r.bn)];
        e.first += ms;
        e.second += 1;
    }
    if (total_ms) *total_ms = ms_sum;
    if (total_flops) *total_flops = fl_sum;
    if (launches) *launches = static_cast<int64_t>(g_prof.size());
    if (report) {
        report->clear();
        for (auto& kv : by_shape) {
            const int M = std::get<0>(kv.first), N = std::get<1>(kv.first), K = std::get<2>(kv.first), bn = std::get<3>(kv.first);
            const double ms = kv.second.first / kv.second.second;
            char line[256];
            snprintf(line, sizeof(line), "M=%d N=%d K=%d bn=%d launches=%d avg_ms=%.4f tflops=%.1f\n", M, N, K, bn,
                     kv.second.second, ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
            *report += line;
        }
    }
    g_prof.clear();
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

namespace {

constexpr int BM = 128;
constexpr int BK = 64;        // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int GROUP_M = 8;    // tile rasterisation: 8 M-blocks share each B tile while it is hot in L2
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int CHUNK = 32;     // accumulator columns per epilogue step
constexpr int EPI_BUF = 4096; // one staging tile: 32 rows x 128 B

template <int BN>
struct Cfg {
    static constexpr int STAGE_A = BM * BK * 2;
    static constexpr int STAGE_B = BN * BK * 2;
    static constexpr int STAGES = (BN == 256) ? 3 : (BN == 192) ? 4 : (BN == 128) ? 5 : 6;
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int HALF = BN / 2;              // columns per epilogue warp
    static constexpr int NCHUNK = HALF / CHUNK;
    static constexpr int EPI_SMEM = EPI_WARPS * 2 * EPI_BUF;     // per warp: bufA (out) + bufB (in / out2)
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = STAGES * (STAGE_A + STAGE_B) + EPI_SMEM + BAR_BYTES + 1024;
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

// ---------------------------------------------------------------------------------------------
// epilogue role (8 warps), shared by the 1-CTA and the 2-CTA (cta_group::2) kernels
// ---------------------------------------------------------------------------------------------
template <int BN, bool CE, class Coords, class Release>
__device__ __forceinline__ void epilogue_role(const TmaMaps& maps, const GemmEpilogue& ep, int M, int N, int num_n, int num_tiles,
                                              int tile_begin, int tile_step, Coords coords, Release release_tmem,
                                              uint32_t tmem_base, uint32_t tfull_bar, uint32_t smem_epi, uint32_t in_bar0,
                                              int warp, int lane) {
    using C = Cfg<BN>;
    const int ew = warp - 2;                      // 0..7
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = ew >> 2;                     // which half of the tile's columns
    const uint32_t bufA = smem_epi + ew * 2 * EPI_BUF;
    const uint32_t bufB = bufA + EPI_BUF;
    const uint32_t in_bar = in_bar0 + 8 * ew;
    const bool has_res = ep.residual != nullptr;
    const bool has_aux = ep.dact != DACT_NONE;
    const bool has_in = has_res || has_aux;
    const bool out_f32 = ep.out_fp32 != 0;
    const uint32_t in_bytes = has_res ? 32u * 128u : 32u * 64u;
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
    auto next_tile_with_work = [&](int tile) {
        int t = tile;
        while (t < num_tiles) {
            int m_idx, n_idx;
            coords(t, m_idx, n_idx);
            if (n_valid_chunks(n_idx) > 0) break;
            t += tile_step;
        }
        return t;
    };
    if (has_in && lane == 0) {
        const int t0 = next_tile_with_work(tile_begin);
        if (t0 < num_tiles) issue_in(t0, 0);
    }

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_begin; tile < num_tiles; tile += tile_step) {
        int m_idx, n_idx;
        coords(tile, m_idx, n_idx);
        const int row0 = m_idx * BM + quarter * 32;
        const int row = row0 + lane;
        const bool row_ok = row < M;
        const int col_base = n_idx * BN + half * C::HALF;
        const int nvalid = n_valid_chunks(n_idx);

        // this warp's slice of the bias vector: lane l keeps column (32k + l) of every chunk, broadcast by shuffle
        float bias_reg[C::NCHUNK];
#pragma unroll
        for (int k = 0; k < C::NCHUNK; ++k) {
            const int n = col_base + k * 32 + lane;
            bias_reg[k] = (ep.bias != nullptr && n < N) ? __ldg(ep.bias + n) : 0.f;
        }

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
        if (CE && row_ok && ep.ce_label != nullptr) label = __ldg(ep.ce_label + row);

        uint32_t r[32];
        if (nvalid > 0) ptx::tmem_ld_32x32(taddr, r);
#pragma unroll
        for (int c = 0; c < C::NCHUNK; ++c) {
            if (c >= nvalid) break;               // warp-uniform
            const int n0 = col_base + c * CHUNK;
            ptx::tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (c + 1 < nvalid) {
                ptx::tmem_ld_32x32(taddr + (c + 1) * CHUNK, r);       // overlaps this chunk's math / staging
            } else {
                ptx::tcgen05_fence_before();                          // accumulator fully read: release the TMEM buffer
                __syncwarp();
                if (lane == 0) release_tmem(acc);
            }
            if (ep.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_reg[c], j);
            }
            if (CE) {
                // running max / sum-exp over the valid vocabulary columns; the label's logit in fp32
                float cm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n0 + j < ep.n_valid) cm = fmaxf(cm, v[j]);
                if (cm > -INFINITY) {
                    const float nm = fmaxf(ce_m, cm);
                    float a = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + j < ep.n_valid) a += __expf(v[j] - nm);
                    ce_s = ce_s * __expf(ce_m - nm) + a;
                    ce_m = nm;
                }
                if (row_ok && label >= n0 && label < n0 + 32) {
                    float t = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + j == label) t = v[j];
                    ep.ce_target[row] = t;
                }
            }
            const uint32_t slot_off = out_f32 ? 0u : (out_slot & 1u) * 2048u;
            // staging buffers of this slot must have been read out by their previous TMA store
            if (lane == 0) {
                if (out_f32) tma_store_wait_read<0>();
                else tma_store_wait_read<1>();
            }
            __syncwarp();
            if (ep.out2 != nullptr) {             // pre-activation, bf16 (bufB is free: no mode has out2 and an input operand)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    st_shared_v4(bufB + slot_off + swz64(lane, q), pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]),
                                 pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]), pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]),
                                 pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
            }
            if (ep.act == ACT_GELU_NEW) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_new(v[j]);
            } else if (ep.act == ACT_TANH) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[j]);
            } else if (ep.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (has_in) {
                ptx::mbar_wait(in_bar, in_phase);
                in_phase ^= 1;
                if (has_res) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint4 u = ld_shared_v4(bufB + swz128(lane, q));
                        v[q * 4 + 0] += __uint_as_float(u.x); v[q * 4 + 1] += __uint_as_float(u.y);
                        v[q * 4 + 2] += __uint_as_float(u.z); v[q * 4 + 3] += __uint_as_float(u.w);
                    }
                } else {
                    float a[32];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 u = ld_shared_v4(bufB + swz64(lane, q));
                        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                        a[q * 8 + 0] = f0.x; a[q * 8 + 1] = f0.y; a[q * 8 + 2] = f1.x; a[q * 8 + 3] = f1.y;
                        a[q * 8 + 4] = f2.x; a[q * 8 + 5] = f2.y; a[q * 8 + 6] = f3.x; a[q * 8 + 7] = f3.y;
                    }
                    if (ep.dact == DACT_GELU_NEW) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= gelu_new_grad(a[j]);
                    } else if (ep.dact == DACT_TANH) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= (1.0f - a[j] * a[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (a[j] > 0.f) ? v[j] : 0.f;
                    }
                }
                __syncwarp();                     // every lane has consumed bufB: fetch the next chunk's operand
                if (lane == 0) {
                    if (c + 1 < nvalid) {
                        issue_in(tile, c + 1);
                    } else {
                        const int tn = next_tile_with_work(tile + tile_step);
                        if (tn < num_tiles) issue_in(tn, 0);
                    }
                }
            }
            if (ep.out != nullptr && ep.debug != 3) {
                if (out_f32) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        st_shared_v4(bufA + swz128(lane, q), __float_as_uint(v[q * 4 + 0]), __float_as_uint(v[q * 4 + 1]),
                                     __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_shared_v4(bufA + slot_off + swz64(lane, q), pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]),
                                     pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]), pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]),
                                     pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
                }
            }
            ptx::fence_proxy_async_smem();        // generic-proxy writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0 && !(ep.debug == 2 || ep.debug == 3 || (ep.debug == 1 && (out_slot & 1)))) {
                if (ep.out != nullptr) {
                    if (ep.split_k > 1) tma_reduce_add_2d(&maps.out, bufA + slot_off, n0, row0);
                    else tma_store_2d(&maps.out, bufA + slot_off, n0, row0);
                }
                if (ep.out2 != nullptr) tma_store_2d(&maps.out2, bufB + slot_off, n0, row0);
                tma_store_commit();
            }
            ++out_slot;
        }
        if (CE && row_ok && n_idx < num_n)
            ep.ce_partial[static_cast<size_t>(row) * ep.ce_tiles + n_idx * 2 + half] = make_float2(ce_m, ce_s);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_read<0>();      // staging smem must outlive the last bulk stores
}

// CM x CN thread-block cluster: the CM CTAs of a cluster column share one B tile and the CN CTAs of a cluster row
// share one A tile; each CTA fetches 1/CM of B (1/CN of A) and TMA-multicasts it to its peers, so every operand
// byte crosses the L2 -> SM fabric once per cluster instead of once per CTA.  (Round-1 measurement: with single
// CTAs the 128x192 tile needs 40 KB per 64-deep K block, and L2 delivers ~46 B/clk/SM -> ~890 clk per K block
// against 384 clk of MMA: every large GEMM plateaued at ~1000 TFLOP/s.)
template <int BN, bool CE, int CM, int CN, bool MN = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ TmaMaps maps, int M, int N, int K, const GemmEpilogue ep) {
    using C = Cfg<BN>;
    constexpr int CLUSTER = CM * CN;
    static_assert(!MN || CLUSTER == 1, "MN-major operands are implemented for independent CTAs only");
    const uint32_t crank = CLUSTER > 1 ? ptx::cluster_ctarank() : 0u;
    const int cm = crank % CM, cn = crank / CM;
    // peers that share my A tile (same cm) / my B tile (same cn), as cluster-rank bit masks
    uint16_t mask_a = 0, mask_b = 0;
#pragma unroll
    for (int j = 0; j < CN; ++j) mask_a |= static_cast<uint16_t>(1u << (cm + CM * j));
#pragma unroll
    for (int i = 0; i < CM; ++i) mask_b |= static_cast<uint16_t>(1u << (i + CM * cn));
    const uint16_t mask_release = mask_a | mask_b;       // everyone whose producer writes into my stages
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
            ptx::mbar_init(empty_bar + 8 * i, CM + CN - 1);    // my MMA warp + every peer that multicasts into this slot
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar + 8 * i, 1);
            ptx::mbar_init(tempty_bar + 8 * i, EPI_WARPS);     // one arrive per epilogue warp
        }
        for (int i = 0; i < EPI_WARPS; ++i) ptx::mbar_init(in_bar0 + 8 * i, 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async_smem();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER > 1) ptx::cluster_sync();                 // peers' barriers are initialised before anyone signals them
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    // the cluster walks "super tiles" of CM x CN tiles; CTA (cm, cn) owns tile (sm * CM + cm, sn * CN + cn).  Tiles
    // past the edge are computed on zero-filled operands and clipped by the TMA stores.
    const int num_sm = (num_m + CM - 1) / CM;
    const int num_sn = (num_n + CN - 1) / CN;
    const int split = ep.split_k > 1 ? ep.split_k : 1;      // split-K: `split` consecutive work items share an output tile
    const int num_tiles = num_sm * num_sn * split;
    const int num_kb = (K + BK - 1) / BK;
    const int tile_begin = blockIdx.x / CLUSTER;
    const int tile_step = gridDim.x / CLUSTER;
    auto coords = [&](int tile, int& m_idx, int& n_idx) {
        int sm, sn;
        tile_coords(tile / split, num_sm, num_sn, sm, sn);
        m_idx = sm * CM + cm;
        n_idx = sn * CN + cn;
    };
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
            for (int tile = tile_begin; tile < num_tiles; tile += tile_step) {
                int m_idx, n_idx, kb0, kb1;
                coords(tile, m_idx, n_idx);
                kb_range(tile, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1);           // the slot is free here AND in every peer
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
                    } else if (CN == 1)
                        ptx::tma_load_2d(smem_a + stage * C::STAGE_A, &maps.a, full_bar + 8 * stage, kb * BK, m_idx * BM);
                    else
                        ptx::tma_load_2d_multicast(smem_a + stage * C::STAGE_A + cn * (C::STAGE_A / CN), &maps.a,
                                                   full_bar + 8 * stage, kb * BK, m_idx * BM + cn * (BM / CN), mask_a);
                    if (MN) {
                    } else if (CM == 1)
                        ptx::tma_load_2d(smem_b + stage * C::STAGE_B, &maps.b, full_bar + 8 * stage, kb * BK, n_idx * BN);
                    else
                        ptx::tma_load_2d_multicast(smem_b + stage * C::STAGE_B + cm * (C::STAGE_B / CM), &maps.b,
                                                   full_bar + 8 * stage, kb * BK, n_idx * BN + cm * (BN / CM), mask_b);
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
            for (int tile = tile_begin; tile < num_tiles; tile += tile_step) {
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
                    // frees the smem slot (here and for the peers that multicast into it) when the MMAs retire
                    if (CLUSTER == 1) ptx::umma_commit(empty_bar + 8 * stage);
                    else ptx::umma_commit_multicast(empty_bar + 8 * stage, mask_release);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(tfull_bar + 8 * acc);                    // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue warps =====================
        epilogue_role<BN, CE>(maps, ep, M, N, num_n, num_tiles, tile_begin, tile_step, coords,
                              [&](int acc) { ptx::mbar_arrive(tempty_bar + 8 * acc); }, tmem_base, tfull_bar, smem_epi, in_bar0,
                              warp, lane);
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (CLUSTER > 1) ptx::cluster_sync();                 // no CTA may exit while a peer can still signal its barriers
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
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = STAGES * (STAGE_A + STAGE_B) + EPI_SMEM + BAR_BYTES + 1024;
    static_assert(BN == 128 || BN == 192 || BN == 256, "pair tile width");
    static_assert(STAGE_B % 1024 == 0, "B stage must keep 1024-B alignment");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int BN, bool CE>
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
    const int num_tiles = num_pm * num_n;
    const int num_kb = (K + BK - 1) / BK;
    const int tile_begin = blockIdx.x / 2;
    const int tile_step = gridDim.x / 2;
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
            for (int tile = tile_begin; tile < num_tiles; tile += tile_step) {
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
            for (int tile = tile_begin; tile < num_tiles; tile += tile_step) {
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
        epilogue_role<BN, CE>(maps, ep, M, N, num_n, num_tiles, tile_begin, tile_step, coords,
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
// host: tensor maps (cached) and launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        EAVQA_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

typedef std::tuple<const void*, int, int, int, int, int> MapKey;   // ptr, rows, cols, ld, box_rows, kind
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

enum MapKind { MAP_OPERAND = 0, MAP_EPI_BF16 = 1, MAP_EPI_F32 = 2 };

// [rows, cols] row-major, row stride ld elements.
//   MAP_OPERAND : bf16, box = box_rows x 64 cols (128 B), SWIZZLE_128B   (UMMA K-major operand tiles)
//   MAP_EPI_BF16: bf16, box = 32 x 32 cols (64 B),  SWIZZLE_64B          (epilogue staging tiles)
//   MAP_EPI_F32 : fp32, box = 32 x 32 cols (128 B), SWIZZLE_128B
// Out-of-bounds elements read as zero and are not written.
CUtensorMap make_map(const void* ptr, int rows, int cols, int ld, int box_rows, int kind) {
    MapKey key(ptr, rows, cols, ld, box_rows, kind);
    std::lock_guard<std::mutex> lock(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) return it->second;
    const bool f32 = kind == MAP_EPI_F32;
    const int esz = f32 ? 4 : 2;
    EAVQA_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand must be 16-byte aligned");
    EAVQA_CHECK((static_cast<int64_t>(ld) * esz) % 16 == 0, "GEMM row strides must be multiples of 16 bytes");
    EAVQA_CHECK(ld >= cols, "GEMM row stride smaller than the row width");
    CUtensorMap m;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kind == MAP_OPERAND ? BK : CHUNK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = (kind == MAP_EPI_BF16) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = encode_fn()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                             const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EAVQA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (code " + std::to_string(static_cast<int>(r)) + ")");
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps[key] = m;
    return m;
}

template <int BN, bool CE, int CM, int CN, bool MN = false>
void launch(const GemmArgs& a, cudaStream_t stream) {
    using C = Cfg<BN>;
    constexpr int CLUSTER = CM * CN;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, CE, CM, CN, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    const GemmEpilogue& e = a.ep;
    TmaMaps maps;
    if (MN) {
        // At [K, M] / Bt [K, N] row-major: rows = contraction index, 64-column x 64-row boxes
        maps.a = make_map(a.A, a.K, a.M, a.lda, 64, MAP_OPERAND);
        maps.b = make_map(a.B, a.K, a.N, a.ldb, 64, MAP_OPERAND);
    } else {
        maps.a = make_map(a.A, a.M, a.K, a.lda, BM / CN, MAP_OPERAND);  // each CTA fetches its 1/CN slice of the A tile
        maps.b = make_map(a.B, a.N, a.K, a.ldb, BN / CM, MAP_OPERAND);  // ... and its 1/CM slice of the B tile
    }
    maps.out = e.out ? make_map(e.out, a.M, a.N, e.ldo, 32, e.out_fp32 ? MAP_EPI_F32 : MAP_EPI_BF16) : maps.a;
    maps.out2 = e.out2 ? make_map(e.out2, a.M, a.N, e.ldo2, 32, MAP_EPI_BF16) : maps.a;
    if (e.residual) maps.in = make_map(e.residual, a.M, a.N, e.ld_res, 32, MAP_EPI_F32);
    else if (e.dact != DACT_NONE) maps.in = make_map(e.aux, a.M, a.N, e.ld_aux, 32, MAP_EPI_BF16);
    else maps.in = maps.a;
    const int split = e.split_k > 1 ? e.split_k : 1;
    const int tiles = ceil_div(ceil_div(a.M, BM), CM) * ceil_div(ceil_div(a.N, BN), CN) * split;      // super tiles x K splits
    const int max_clusters = num_sms() / CLUSTER;
    const int grid = (tiles < max_clusters ? tiles : max_clusters) * CLUSTER;
    ProfRec rec;
    if (g_prof_on) {
        CUDA_CHECK(cudaEventCreate(&rec.start));
        CUDA_CHECK(cudaEventCreate(&rec.stop));
        rec.M = a.M; rec.N = a.N; rec.K = a.K; rec.bn = BN;
        CUDA_CHECK(cudaEventRecord(rec.start, stream));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<BN, CE, CM, CN, MN>, maps, a.M, a.N, a.K, a.ep));
    KERNEL_CHECK();
    if (g_prof_on) {
        CUDA_CHECK(cudaEventRecord(rec.stop, stream));
        g_prof.push_back(rec);
    }
    g_gemm_launches.fetch_add(1);
}

template <int BN, bool CE>
void launch_2cta(const GemmArgs& a, cudaStream_t stream) {
    using C = Cfg2<BN>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_2cta_kernel<BN, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    const GemmEpilogue& e = a.ep;
    TmaMaps maps;
    maps.a = make_map(a.A, a.M, a.K, a.lda, BM, MAP_OPERAND);
    maps.b = make_map(a.B, a.N, a.K, a.ldb, BN / 2, MAP_OPERAND);      // each CTA of the pair stages half of the B tile
    maps.out = e.out ? make_map(e.out, a.M, a.N, e.ldo, 32, e.out_fp32 ? MAP_EPI_F32 : MAP_EPI_BF16) : maps.a;
    maps.out2 = e.out2 ? make_map(e.out2, a.M, a.N, e.ldo2, 32, MAP_EPI_BF16) : maps.a;
    if (e.residual) maps.in = make_map(e.residual, a.M, a.N, e.ld_res, 32, MAP_EPI_F32);
    else if (e.dact != DACT_NONE) maps.in = make_map(e.aux, a.M, a.N, e.ld_aux, 32, MAP_EPI_BF16);
    else maps.in = maps.a;
    const int tiles = ceil_div(ceil_div(a.M, BM), 2) * ceil_div(a.N, BN);
    const int max_clusters = num_sms() / 2;
    const int grid = (tiles < max_clusters ? tiles : max_clusters) * 2;
    ProfRec rec;
    if (g_prof_on) {
        CUDA_CHECK(cudaEventCreate(&rec.start));
        CUDA_CHECK(cudaEventCreate(&rec.stop));
        rec.M = a.M; rec.N = a.N; rec.K = a.K; rec.bn = BN + 1000;
        CUDA_CHECK(cudaEventRecord(rec.start, stream));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_tn_2cta_kernel<BN, CE>, maps, a.M, a.N, a.K, a.ep));
    KERNEL_CHECK();
    if (g_prof_on) {
        CUDA_CHECK(cudaEventRecord(rec.stop, stream));
        g_prof.push_back(rec);
    }
    g_gemm_launches.fetch_add(1);
}

}  // namespace

// Pick the N tile that minimises (waves x per-tile time).  Per-tile time ~ BN plus a fixed per-tile cost
// (pipeline fill, accumulator hand-off); BN = 64 is penalised because the operand reads (A 128 rows + B 64 rows
// per K step) exceed the 128 B/clk shared-memory port.
int gemm_pick_block_n(int M, int N, int K, int forced) {
    (void)K;
    if (forced == 64 || forced == 128 || forced == 192 || forced == 256) return forced;
    EAVQA_CHECK(forced == 0, "block_n must be 0, 64, 128, 192 or 256");
    const int cands[4] = {256, 192, 128, 64};
    const double penalty[4] = {1.0, 1.0, 1.04, 1.5};
    const int sms = num_sms();
    const int num_m = ceil_div(M, BM);
    double best = 1e300;
    int best_bn = 128;
    for (int i = 0; i < 4; ++i) {
        const int bn = cands[i];
        const int num_n = ceil_div(N, bn);
        const int64_t tiles = static_cast<int64_t>(num_m) * num_n;
        const int64_t waves = (tiles + sms - 1) / sms;
        const double cost = static_cast<double>(waves) * (bn * penalty[i] + 24.0);
        if (cost < best - 1e-9) {
            best = cost;
            best_bn = bn;
        }
    }
    return best_bn;
}

// Cluster shape: 1 = independent CTAs, 2 / 4 = 2x1 / 2x2 TMA-multicast clusters (kept for reference: measured no
// gain, profiles/r01_gemm_cluster_sweep.txt), 8 = CTA pair with tcgen05.mma.cta_group::2 (256 x BN tiles).
// The pair halves each CTA's shared-memory traffic for B and wins wherever the main loop dominates: measured on
// B200 (tools/gemm_bench.py) 8192^3 1247 -> 1591 TFLOP/s, LM head 1109 -> 1377, head dgrad 1048 -> 1387,
// c_fc 1072 -> 1157; it loses on short-K, few-tile problems whose time is epilogue / launch latency.
int gemm_pick_cluster(int M, int N, int bn, int forced) {
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return bn == 64 ? 1 : forced;
    EAVQA_CHECK(forced == 0, "cluster must be 0, 1, 2, 4 or 8 (8 = CTA pair, tcgen05 cta_group::2)");
    (void)M; (void)N;
    return 1;
}

// Tile width and CTA mode for a problem.  Model: time ~ waves x (K blocks x clocks per K block + fixed per-tile cost);
// clocks per 64-deep K block = max(MMA, shared-memory port): 1 CTA max(2 BN, 256 + 2 BN), pair max(2 BN, 256 + BN).
void gemm_pick_config(int M, int N, int K, int forced_bn, int forced_cluster, int* bn_out, int* cluster_out) {
    const int sms = num_sms();
    const int num_m = ceil_div(M, BM);
    const int kb = ceil_div(K, BK);
    int cluster = forced_cluster;
    if (cluster == 0) {
        // In isolation the pair also wins on the K = 768 shapes (+5..25 %), but inside the step those are bound by
        // their epilogues (CE statistics, gelu + second output, residual) and the coupled pair loses 5-15 %
        // (profiles/r01_gemm_shapes_v5_pair_everywhere.txt): use it where the main loop dominates.
        cluster = (K >= 4096 && M >= 2048 && N >= 256) ? 8 : 1;
    }
    EAVQA_CHECK(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8, "cluster must be 0, 1, 2, 4 or 8");
    int bn = forced_bn;
    if (bn == 0) {
        if (cluster == 8) {
            const int cands[3] = {256, 192, 128};
            double best = 1e300;
            for (int i = 0; i < 3; ++i) {
                const int b = cands[i];
                const int64_t tiles = static_cast<int64_t>(ceil_div(num_m, 2)) * ceil_div(N, b);
                const int64_t waves = (tiles + sms / 2 - 1) / (sms / 2);
                const double per_kb = std::max(2.0 * b, 256.0 + b);
                const double cost = static_cast<double>(waves) * (kb * per_kb + 700.0);
                if (cost < best - 1e-9) { best = cost; bn = b; }
            }
        } else {
            bn = gemm_pick_block_n(M, N, K, 0);
        }
    }
    EAVQA_CHECK(bn == 64 || bn == 128 || bn == 192 || bn == 256, "block_n must be 0, 64, 128, 192 or 256");
    if (bn == 64) cluster = 1;
    *bn_out = bn;
    *cluster_out = cluster;
}

void gemm_bf16_tn(const GemmArgs& a_in, cudaStream_t stream) {
    GemmArgs a = a_in;
    {
        static int dbg = -1;
        if (dbg < 0) {
            const char* e = getenv("EAVQA_GEMM_DEBUG");
            dbg = e ? atoi(e) : 0;
        }
        a.ep.debug = dbg;
    }
    EAVQA_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "GEMM with an empty dimension");
    EAVQA_CHECK(a.A != nullptr && a.B != nullptr, "GEMM operand is null");
    const GemmEpilogue& e = a.ep;
    EAVQA_CHECK(e.out != nullptr || e.out2 != nullptr || e.ce_partial != nullptr, "GEMM without an output");
    EAVQA_CHECK(!(e.out2 != nullptr && (e.residual != nullptr || e.dact != DACT_NONE)),
                "GEMM: a second output cannot be combined with a residual / derivative operand");
    EAVQA_CHECK(!(e.residual != nullptr && e.dact != DACT_NONE), "GEMM: residual and derivative operand are exclusive");
    if (e.dact != DACT_NONE) EAVQA_CHECK(e.aux != nullptr, "GEMM: derivative epilogue needs aux");
    const bool ce = e.ce_partial != nullptr;
    int bn = 0, cl = 1;
    gemm_pick_config(a.M, a.N, a.K, a.block_n, a.mn_major ? 1 : a.cluster, &bn, &cl);
    if (ce) {
        EAVQA_CHECK(e.ce_tiles == 2 * ceil_div(a.N, bn), "ce_tiles must be 2 * ceil(N / block_n)");
        EAVQA_CHECK(e.ce_target != nullptr && e.n_valid > 0 && e.n_valid <= a.N, "CE epilogue arguments");
    }
#define EAVQA_GEMM_CASE(BN_)                                                                   \
    case BN_:                                                                                  \
        if (cl == 8) ce ? launch_2cta<BN_, true>(a, stream) : launch_2cta<BN_, false>(a, stream);          \
        else if (cl == 4) ce ? launch<BN_, true, 2, 2>(a, stream) : launch<BN_, false, 2, 2>(a, stream);      \
        else if (cl == 2) ce ? launch<BN_, true, 2, 1>(a, stream) : launch<BN_, false, 2, 1>(a, stream); \
        else ce ? launch<BN_, true, 1, 1>(a, stream) : launch<BN_, false, 1, 1>(a, stream);              \
        break;
    if (a.ep.split_k > 1) {
        EAVQA_CHECK(e.out != nullptr && e.out_fp32 && !e.bias && !e.residual && !e.out2 && e.act == ACT_NONE && e.dact == DACT_NONE && !ce,
                    "split-K needs a plain fp32 output (partials are added into it)");
        EAVQA_CHECK(cl == 1 && a.ep.split_k <= ceil_div(a.K, BK), "split-K: independent CTAs, at most one split per K block");
    }
    if (a.mn_major) {
        EAVQA_CHECK(!ce && cl == 1, "MN-major (wgrad form) GEMM: plain epilogues, independent CTAs only");
        switch (bn) {
            case 256: launch<256, false, 1, 1, true>(a, stream); break;
            case 192: launch<192, false, 1, 1, true>(a, stream); break;
            case 128: launch<128, false, 1, 1, true>(a, stream); break;
            default: launch<64, false, 1, 1, true>(a, stream); break;
        }
        return;
    }
    switch (bn) {
        EAVQA_GEMM_CASE(256)
        EAVQA_GEMM_CASE(192)
        EAVQA_GEMM_CASE(128)
        default:
            ce ? launch<64, true, 1, 1>(a, stream) : launch<64, false, 1, 1>(a, stream);
            break;
    }
#undef EAVQA_GEMM_CASE
}

}  // namespace eavqa
