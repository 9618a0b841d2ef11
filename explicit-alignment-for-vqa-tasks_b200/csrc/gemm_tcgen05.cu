// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> shared (SWIZZLE_128B) -> tcgen05.mma
// (UMMA 128 x BN x 16, fp32 accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue.
//
// Roles (192 threads, one CTA per SM):
//   warp 0 / lane 0 : TMA producer   (ring of STAGES smem slots, full/empty mbarriers)
//   warp 1 / lane 0 : MMA issuer     (also owns TMEM alloc/dealloc, whole warp)
//   warps 2..5      : epilogue       (warp w reads TMEM lanes [32*(w%4), +32): one thread = one row)
// The accumulator of tile i+1 is produced into the other TMEM buffer while the epilogue drains
// tile i (tmem_full / tmem_empty mbarriers), so the tensor pipe never waits for the epilogue.
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
constexpr int NUM_THREADS = 192;

template <int BN>
struct Cfg {
    static constexpr int STAGE_A = BM * BK * 2;
    static constexpr int STAGE_B = BN * BK * 2;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192) ? 5 : (BN == 128) ? 6 : 8;
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = STAGES * (STAGE_A + STAGE_B) + BAR_BYTES + 1024;  // +1024: manual alignment slack
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
    static_assert(STAGE_B % 1024 == 0, "B stage must keep 1024-B alignment");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// ---------------------------------------------------------------------------------------------
// epilogue: one thread owns one output row; `v` holds 32 consecutive fp32 accumulator columns
// ---------------------------------------------------------------------------------------------
struct CeState {
    float m, s;
};

template <bool CE>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& ep, float (&v)[32], int row, int n0, int N,
                                               CeState& ce, int label) {
    const bool full = (n0 + 32 <= N);
    if (ep.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (full || n0 + j < N) v[j] += __ldg(ep.bias + n0 + j);
    }
    if (CE) {
        // running max / sum-exp over the valid vocabulary columns of this tile; the label's logit in fp32
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < ep.n_valid) cm = fmaxf(cm, v[j]);
        if (cm > -INFINITY) {
            float nm = fmaxf(ce.m, cm);
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < ep.n_valid) acc += __expf(v[j] - nm);
            ce.s = ce.s * __expf(ce.m - nm) + acc;
            ce.m = nm;
        }
        if (label >= n0 && label < n0 + 32) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j == label) t = v[j];
            ep.ce_target[row] = t;
        }
    }
    if (ep.out2 != nullptr) {
        bf16* p = ep.out2 + static_cast<size_t>(row) * ep.ldo2 + n0;
        if (full) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u;
                u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                reinterpret_cast<uint4*>(p)[q] = u;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) p[j] = __float2bfloat16(v[j]);
        }
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
    if (ep.dact != DACT_NONE) {
        const bf16* ap = ep.aux + static_cast<size_t>(row) * ep.ld_aux + n0;
        float a[32];
        if (full) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u = __ldg(reinterpret_cast<const uint4*>(ap) + q);
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                a[q * 8 + 0] = f0.x; a[q * 8 + 1] = f0.y; a[q * 8 + 2] = f1.x; a[q * 8 + 3] = f1.y;
                a[q * 8 + 4] = f2.x; a[q * 8 + 5] = f2.y; a[q * 8 + 6] = f3.x; a[q * 8 + 7] = f3.y;
            }
        } else {
            for (int j = 0; j < 32; ++j) a[j] = (n0 + j < N) ? __bfloat162float(ap[j]) : 0.f;
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
    if (ep.residual != nullptr) {
        const float* rp = ep.residual + static_cast<size_t>(row) * ep.ld_res + n0;
        if (full) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 r = __ldg(reinterpret_cast<const float4*>(rp) + q);
                v[q * 4 + 0] += r.x; v[q * 4 + 1] += r.y; v[q * 4 + 2] += r.z; v[q * 4 + 3] += r.w;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) v[j] += rp[j];
        }
    }
    if (ep.out != nullptr) {
        if (ep.out_fp32) {
            float* p = static_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldo + n0;
            if (full) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    reinterpret_cast<float4*>(p)[q] = make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            } else {
                for (int j = 0; j < 32; ++j)
                    if (n0 + j < N) p[j] = v[j];
            }
        } else {
            bf16* p = static_cast<bf16*>(ep.out) + static_cast<size_t>(row) * ep.ldo + n0;
            if (full) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                    u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                    u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                    u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                    reinterpret_cast<uint4*>(p)[q] = u;
                }
            } else {
                for (int j = 0; j < 32; ++j)
                    if (n0 + j < N) p[j] = __float2bfloat16(v[j]);
            }
        }
    }
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

template <int BN, bool CE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N,
                    int K, const GemmEpilogue ep) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-B aligned bases (descriptor base_offset = 0)
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);

    const uint32_t smem_a = base;
    const uint32_t smem_b = base + C::STAGES * C::STAGE_A;
    const uint32_t bars = smem_b + C::STAGES * C::STAGE_B;
    const uint32_t full_bar = bars;                       // STAGES x 8 B
    const uint32_t empty_bar = bars + 8 * C::STAGES;      // STAGES x 8 B
    const uint32_t tfull_bar = bars + 16 * C::STAGES;     // 2 x 8 B
    const uint32_t tempty_bar = tfull_bar + 16;           // 2 x 8 B
    const uint32_t tmem_slot = tempty_bar + 16;           // 4 B
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tma_a);
        ptx::prefetch_tensormap(&tma_b);
        for (int i = 0; i < C::STAGES; ++i) {
            ptx::mbar_init(full_bar + 8 * i, 1);
            ptx::mbar_init(empty_bar + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar + 8 * i, 1);
            ptx::mbar_init(tempty_bar + 8 * i, 4);     // one arrive per epilogue warp
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
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_idx, n_idx;
                tile_coords(tile, num_m, num_n, m_idx, n_idx);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    ptx::mbar_arrive_expect_tx(full_bar + 8 * stage, C::STAGE_A + C::STAGE_B);
                    ptx::tma_load_2d(smem_a + stage * C::STAGE_A, &tma_a, full_bar + 8 * stage, kb * BK, m_idx * BM);
                    ptx::tma_load_2d(smem_b + stage * C::STAGE_B, &tma_b, full_bar + 8 * stage, kb * BK, n_idx * BN);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                ptx::mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);     // epilogue drained this buffer
                ptx::tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(full_bar + 8 * stage, phase);          // TMA bytes landed
                    ptx::tcgen05_fence_after();
                    const uint64_t da = ptx::make_kmajor_sw128_desc(smem_a + stage * C::STAGE_A);
                    const uint64_t db = ptx::make_kmajor_sw128_desc(smem_b + stage * C::STAGE_B);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing 16 bf16 (32 B) along K inside the 128-B swizzle row: +2 in the >>4 address field
                        ptx::umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
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
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int row_in_tile = quarter * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_idx, n_idx;
            tile_coords(tile, num_m, num_n, m_idx, n_idx);
            const int row = m_idx * BM + row_in_tile;
            const bool row_ok = row < M;
            ptx::mbar_wait(tfull_bar + 8 * acc, acc_phase);
            ptx::tcgen05_fence_after();
            CeState ce;
            ce.m = -INFINITY;
            ce.s = 0.f;
            int label = -1;
            if (CE && row_ok && ep.ce_label != nullptr) label = __ldg(ep.ce_label + row);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                const int n0 = n_idx * BN + c;
                if (n0 >= N) break;                   // warp-uniform
                uint32_t r[32];
                ptx::tmem_ld_32x32(taddr + c, r);
                ptx::tmem_ld_wait();
                if (row_ok) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    epilogue_chunk<CE>(ep, v, row, n0, N, ce, label);
                }
            }
            if (CE && row_ok) ep.ce_partial[static_cast<size_t>(row) * ep.ce_tiles + n_idx] = make_float2(ce.m, ce.s);
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar + 8 * acc);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
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

typedef std::tuple<const void*, int, int, int, int> MapKey;   // ptr, rows, cols, ld, box_rows
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

// [rows, cols] bf16, row stride ld elements, box = box_rows x 64 columns, SWIZZLE_128B, zero OOB fill
CUtensorMap make_map(const bf16* ptr, int rows, int cols, int ld, int box_rows) {
    MapKey key(ptr, rows, cols, ld, box_rows);
    std::lock_guard<std::mutex> lock(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) return it->second;
    EAVQA_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand must be 16-byte aligned");
    EAVQA_CHECK(ld % 8 == 0, "GEMM operand row stride must be a multiple of 8 elements");
    EAVQA_CHECK(ld >= cols, "GEMM operand row stride smaller than its width");
    CUtensorMap m;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(bf16)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EAVQA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (code " + std::to_string(static_cast<int>(r)) + ")");
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps[key] = m;
    return m;
}

template <int BN, bool CE>
void launch(const GemmArgs& a, cudaStream_t stream) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    CUtensorMap ma = make_map(a.A, a.M, a.K, a.lda, BM);
    CUtensorMap mb = make_map(a.B, a.N, a.K, a.ldb, BN);
    const int tiles = ceil_div(a.M, BM) * ceil_div(a.N, BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    ProfRec rec;
    if (g_prof_on) {
        CUDA_CHECK(cudaEventCreate(&rec.start));
        CUDA_CHECK(cudaEventCreate(&rec.stop));
        rec.M = a.M; rec.N = a.N; rec.K = a.K; rec.bn = BN;
        CUDA_CHECK(cudaEventRecord(rec.start, stream));
    }
    gemm_bf16_tn_kernel<BN, CE><<<grid, NUM_THREADS, C::SMEM, stream>>>(ma, mb, a.M, a.N, a.K, a.ep);
    KERNEL_CHECK();
    if (g_prof_on) {
        CUDA_CHECK(cudaEventRecord(rec.stop, stream));
        g_prof.push_back(rec);
    }
    g_gemm_launches.fetch_add(1);
}

}  // namespace

// Pick the N tile that minimises (waves x per-tile time).  Per-tile time ~ BN, inflated when the
// operand reads (A 128 rows + B BN rows per K step) exceed the 128 B/clk shared-memory port:
// BN=64 is port-bound (x1.5), BN>=128 is MMA-bound.
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
        // a fixed per-tile overhead (pipeline fill / epilogue hand-off) keeps tiny tiles from winning on ties
        const double cost = static_cast<double>(waves) * (bn * penalty[i] + 24.0);
        if (cost < best - 1e-9) {
            best = cost;
            best_bn = bn;
        }
    }
    return best_bn;
}

void gemm_bf16_tn(const GemmArgs& a, cudaStream_t stream) {
    EAVQA_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "GEMM with an empty dimension");
    EAVQA_CHECK(a.A != nullptr && a.B != nullptr, "GEMM operand is null");
    const GemmEpilogue& e = a.ep;
    EAVQA_CHECK(e.out != nullptr || e.out2 != nullptr || e.ce_partial != nullptr, "GEMM without an output");
    if (e.out) {
        EAVQA_CHECK(e.ldo % (e.out_fp32 ? 4 : 8) == 0, "GEMM output stride alignment");
        EAVQA_CHECK((reinterpret_cast<uintptr_t>(e.out) & 15) == 0, "GEMM output must be 16-byte aligned");
    }
    if (e.out2) EAVQA_CHECK(e.ldo2 % 8 == 0 && (reinterpret_cast<uintptr_t>(e.out2) & 15) == 0, "GEMM out2 alignment");
    if (e.residual) EAVQA_CHECK(e.ld_res % 4 == 0 && (reinterpret_cast<uintptr_t>(e.residual) & 15) == 0, "GEMM residual alignment");
    if (e.dact != DACT_NONE) EAVQA_CHECK(e.aux != nullptr && e.ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(e.aux) & 15) == 0, "GEMM aux alignment");
    const bool ce = e.ce_partial != nullptr;
    const int bn = gemm_pick_block_n(a.M, a.N, a.K, a.block_n);
    if (ce) {
        EAVQA_CHECK(e.ce_tiles == ceil_div(a.N, bn), "ce_tiles does not match the N tiling");
        EAVQA_CHECK(e.ce_target != nullptr && e.n_valid > 0 && e.n_valid <= a.N, "CE epilogue arguments");
    }
    switch (bn) {
        case 256: ce ? launch<256, true>(a, stream) : launch<256, false>(a, stream); break;
        case 192: ce ? launch<192, true>(a, stream) : launch<192, false>(a, stream); break;
        case 128: ce ? launch<128, true>(a, stream) : launch<128, false>(a, stream); break;
        default:  ce ? launch<64, true>(a, stream) : launch<64, false>(a, stream); break;
    }
}

}  // namespace eavqa
