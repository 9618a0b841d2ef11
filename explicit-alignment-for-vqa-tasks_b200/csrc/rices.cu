// RICES in-context example retrieval: the step BEFORE the few-shot path.
//
// Reference: src/in_context_example_selection/get_question_knn.py:64-76 -- faiss.normalize_L2 on the train (database)
// and val (query) CLIP text embeddings, faiss.IndexFlatIP on the GPU, search(k = 2048) -- and
// get_image_knn_from_text_knn.py:79-92 -- per test question, the same normalised inner product of its image embedding
// against the images of its 2048 text neighbours, fully sorted.  faiss is a third-party dependency that is not under
// /root/reference; its published semantics are restated in oracle/rices.py: rows scaled by 1 / ||x||_2 (zero rows
// untouched), exact inner products, the k largest per query in descending order.
//
// B200 design: the score matrix is a tensor-core GEMM.  To keep fp32-level accuracy (the neighbours' scores differ by
// 1e-4 .. 1e-3; plain bf16 operands would reorder them) every normalised row is split into bf16 hi + lo parts and the
// product hi.hi' + hi.lo' + lo.hi' is ONE tcgen05 GEMM over a 3x longer contraction: queries are packed [hi | hi | lo],
// database rows [hi | lo | hi]; the dropped lo.lo' term is < 2^-17.  Scores are produced in [query block x database
// chunk] tiles that stay L2-resident (64 MB) for the selection kernel, which keeps a per-query candidate pool and a
// running k-th-score threshold: after the first chunks almost nothing passes the threshold, so selection costs one
// streaming read of the tile.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <vector>

#include "gemm.cuh"
#include "kernels.cuh"

namespace eavqa {
namespace {

// ---------------------------------------------------------------------------------------------
// normalise + split:  out[row] = [hi | hi | lo] (queries) or [hi | lo | hi] (database), bf16, 3 * D wide
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rices_normalize_split_kernel(const float* __restrict__ x, int64_t rows, int D,
                                                                   bf16* __restrict__ out, int database_layout) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
    if (row >= rows) return;
    const float* xr = x + row * D;
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float v = xr[c];
        ss += v * v;
    }
    ss = warp_sum(ss);
    const float sc = ss > 0.f ? 1.0f / sqrtf(ss) : 1.0f;      // faiss fvec_renorm_L2: zero rows are left alone
    bf16* o = out + row * 3 * D;
    for (int c = lane; c < D; c += 32) {
        const float v = xr[c] * sc;
        const bf16 hi = __float2bfloat16(v);
        const bf16 lo = __float2bfloat16(v - __bfloat162float(hi));
        o[c] = hi;
        o[D + c] = database_layout ? lo : hi;
        o[2 * D + c] = database_layout ? hi : lo;
    }
}

// ---------------------------------------------------------------------------------------------
// selection.  One CTA per query row.  Pool = unsorted candidates (score, index) in shared memory, capacity `cap`;
// `thr` = k-th best score seen so far (-inf until k candidates exist).  Ordering everywhere: score descending, then
// index ascending (columns arrive in ascending index order, so a later element that merely ties the threshold loses).
// ---------------------------------------------------------------------------------------------
struct Cand {
    float s;
    int i;
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) { return a.s > b.s || (a.s == b.s && a.i < b.i); }

// bitonic sort of p[0..P) (P a power of two) into "best first" order; all threads of the CTA call it
__device__ void bitonic_sort_best_first(Cand* p, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;              // this pair sorts best-first when `up`
                const Cand a = p[lo], b = p[hi];
                const bool swap = up ? better(b, a) : better(a, b);
                if (swap) {
                    p[lo] = b;
                    p[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// keeps the best min(cnt, k) candidates at p[0..), returns the new count and threshold (all threads get them)
__device__ void compact_pool(Cand* p, int& cnt, int k, float& thr, int P) {
    for (int t = cnt + threadIdx.x; t < P; t += blockDim.x) {
        p[t].s = -INFINITY;
        p[t].i = 0x7fffffff;
    }
    bitonic_sort_best_first(p, P);
    cnt = min(cnt, k);
    thr = cnt >= k ? p[k - 1].s : -INFINITY;
    __syncthreads();
}

__global__ void __launch_bounds__(256) rices_select_kernel(const float* __restrict__ S, int64_t ldS, int ncols, int64_t col0, int k,
                                                          int cap, float* __restrict__ pool_s, int* __restrict__ pool_i,
                                                          int* __restrict__ pool_n, float* __restrict__ pool_thr) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(8) uint8_t smem_sel[];
    Cand* pool = reinterpret_cast<Cand*>(smem_sel);
    __shared__ int s_cnt, s_batch;
    const int row = blockIdx.x;
    const int P = next_pow2(cap);
    int cnt = pool_n[row];
    float thr = pool_thr[row];
    if (threadIdx.x == 0) s_batch = 0;
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
        pool[t].s = pool_s[static_cast<int64_t>(row) * cap + t];
        pool[t].i = pool_i[static_cast<int64_t>(row) * cap + t];
    }
    if (threadIdx.x == 0) s_cnt = cnt;
    __syncthreads();
    const float* srow = S + static_cast<int64_t>(row) * ldS;
    // batches of 4 columns per thread (ldS and the row base are 16-byte aligned).  After the first chunks hardly anything
    // beats the threshold: a batch nobody passes costs one barrier and no shared-memory traffic.
    const int batch = 4 * blockDim.x;
    for (int c0 = 0; c0 < ncols; c0 += batch) {
        const int c = c0 + 4 * threadIdx.x;
        float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (c + 3 < ncols) {
            v = *reinterpret_cast<const float4*>(srow + c);
        } else if (c < ncols) {
            v.x = srow[c];
            if (c + 1 < ncols) v.y = srow[c + 1];
            if (c + 2 < ncols) v.z = srow[c + 2];
        }
        // how many elements of this batch beat the threshold (exact count: compaction only when the pool would overflow;
        // the first version compacted whenever a full batch might not fit, i.e. on almost every batch -> 235 us per chunk)
        int mine = (v.x > thr) + (v.y > thr) + (v.z > thr) + (v.w > thr);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine > 0) atomicAdd(&s_batch, mine);
        __syncthreads();
        const int total = s_batch;                                  // CTA-uniform
        if (total == 0) continue;
        if (cnt + total > cap) {
            compact_pool(pool, cnt, k, thr, P);                     // raises thr: fewer than `total` may pass now, never more
            if (threadIdx.x == 0) s_cnt = cnt;
            __syncthreads();
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (e[j] > thr) {
                const int pos = atomicAdd(&s_cnt, 1);
                pool[pos].s = e[j];
                pool[pos].i = static_cast<int>(col0 + c + j);
            }
        __syncthreads();
        cnt = s_cnt;
        if (threadIdx.x == 0) s_batch = 0;
        __syncthreads();                                            // count read / batch counter reset before the next batch
    }
    // The pool goes back unsorted with up to `cap` entries and the threshold of the LAST compaction: a stale (lower)
    // threshold only admits a few extra candidates, whereas compacting at the end of every chunk cost one 4096-element
    // sort per query per chunk (~2/3 of the selection time).  rices_finalize_kernel does the final sort.
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
        pool_s[static_cast<int64_t>(row) * cap + t] = pool[t].s;
        pool_i[static_cast<int64_t>(row) * cap + t] = pool[t].i;
    }
    if (threadIdx.x == 0) {
        pool_n[row] = cnt;
        pool_thr[row] = thr;
    }
}

// sorted output: out_scores / out_index [rows, k]; rows with fewer than k candidates are padded like faiss (-FLT_MAX, -1)
__global__ void __launch_bounds__(256) rices_finalize_kernel(const float* __restrict__ pool_s, const int* __restrict__ pool_i,
                                                            const int* __restrict__ pool_n, int cap, int k,
                                                            float* __restrict__ out_scores, int64_t* __restrict__ out_index) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(8) uint8_t smem_sel[];
    Cand* pool = reinterpret_cast<Cand*>(smem_sel);
    const int row = blockIdx.x;
    int cnt = pool_n[row];
    float thr = -INFINITY;
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
        pool[t].s = pool_s[static_cast<int64_t>(row) * cap + t];
        pool[t].i = pool_i[static_cast<int64_t>(row) * cap + t];
    }
    __syncthreads();
    compact_pool(pool, cnt, k, thr, next_pow2(cap));
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const bool ok = t < cnt;
        out_scores[static_cast<int64_t>(row) * k + t] = ok ? pool[t].s : -3.402823466e38f;
        out_index[static_cast<int64_t>(row) * k + t] = ok ? static_cast<int64_t>(pool[t].i) : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// second stage (get_image_knn_from_text_knn.py:79-92): one CTA per test question; its candidates are rows of the
// train-image table picked by cand[q, 0..C) (-1 = padding).  sim = <q / |q|, x / |x|> in fp32, all candidates sorted.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rices_rerank_kernel(const float* __restrict__ query, const float* __restrict__ table, int D,
                                                          const int* __restrict__ cand, int C, float* __restrict__ out_sim,
                                                          int* __restrict__ out_pos) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(8) uint8_t smem_sel[];
    Cand* pool = reinterpret_cast<Cand*>(smem_sel);
    __shared__ float s_qinv;
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const float* qr = query + static_cast<int64_t>(q) * D;
    if (warp == 0) {
        float ss = 0.f;
        for (int c = lane; c < D; c += 32) ss += qr[c] * qr[c];
        ss = warp_sum(ss);
        if (lane == 0) s_qinv = ss > 0.f ? 1.0f / sqrtf(ss) : 1.0f;
    }
    __syncthreads();
    const float qinv = s_qinv;
    int n_valid = 0;
    for (int j = warp; j < C; j += nwarp) {
        const int id = cand[static_cast<int64_t>(q) * C + j];
        float sim = -INFINITY;
        if (id >= 0) {
            const float* xr = table + static_cast<int64_t>(id) * D;
            float dot = 0.f, ss = 0.f;
            for (int c = lane; c < D; c += 32) {
                const float xv = xr[c];
                dot += (qr[c] * qinv) * xv;
                ss += xv * xv;
            }
            dot = warp_sum(dot);
            ss = warp_sum(ss);
            sim = dot * (ss > 0.f ? 1.0f / sqrtf(ss) : 1.0f);
        }
        if (lane == 0) {
            pool[j].s = sim;
            pool[j].i = id >= 0 ? j : 0x7fffffff;      // padding sorts last
        }
        n_valid += id >= 0;
    }
    __syncthreads();
    int cnt = C;
    float thr;
    compact_pool(pool, cnt, C, thr, next_pow2(C));
    for (int t = threadIdx.x; t < C; t += blockDim.x) {
        const bool ok = pool[t].i != 0x7fffffff;
        out_sim[static_cast<int64_t>(q) * C + t] = ok ? pool[t].s : -3.402823466e38f;
        out_pos[static_cast<int64_t>(q) * C + t] = ok ? pool[t].i : -1;
    }
    (void)n_valid;
}

// per-query pool state at the start of a query block: empty pool, threshold -inf (on the device: no host staging buffer)
__global__ void rices_pool_reset_kernel(int* __restrict__ pool_n, float* __restrict__ pool_thr, int n) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        pool_n[i] = 0;
        pool_thr[i] = -INFINITY;
    }
}

int select_cap(int k) { return k + std::max(k, 1024); }      // the pool always has room for one 1024-column batch beyond k

// workspace kept between calls (grow-only): the index build of a 444k x 768 database needs 2 GB, and cudaMalloc / cudaFree
// of that size cost more than the search of a thousand queries
struct Workspace {
    void* p = nullptr;
    size_t cap = 0;
    ~Workspace() { if (p) cudaFree(p); }
    uint8_t* reserve(size_t bytes, cudaStream_t s) {
        if (bytes > cap) {
            CUDA_CHECK(cudaStreamSynchronize(s));
            if (p) CUDA_CHECK(cudaFree(p));
            p = nullptr; cap = 0;
            CUDA_CHECK(cudaMalloc(&p, bytes));
            cap = bytes;
        }
        return static_cast<uint8_t*>(p);
    }
};
Workspace g_ws;
size_t align256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }

}  // namespace

// queries [M, D], database [N, D] fp32 on the device; out_scores [M, k] fp32, out_index [M, k] int64
void rices_search(const float* queries, const float* database, int64_t M, int64_t N, int D, int k, float* out_scores,
                  int64_t* out_index, cudaStream_t s) {
    EAVQA_CHECK(queries && database && out_scores && out_index, "rices_search: null argument");
    EAVQA_CHECK(M > 0 && N > 0 && N < (1ll << 31) && D > 0 && D % 8 == 0 && D <= 4096, "rices_search: bad shape (D must be a multiple of 8)");
    EAVQA_CHECK(k >= 1 && k <= 2048, "rices_search: 1 <= k <= 2048 (the limit of faiss' GPU k-selection as well)");
    const int K3 = 3 * D;
    const int cap = select_cap(k);
    const int P = 1 << static_cast<int>(std::ceil(std::log2(static_cast<double>(cap))));
    const size_t sel_smem = sizeof(Cand) * static_cast<size_t>(P);
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(rices_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CUDA_CHECK(cudaFuncSetAttribute(rices_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        configured = true;
    }
    // tiles: Mc queries x Nc database rows of fp32 scores (<= 64 MB: stays in the 126 MB L2 between GEMM and selection)
    const int64_t Mc = std::min<int64_t>(M, 1024);
    const int64_t Nc = std::min<int64_t>(N, 16384);
    const int64_t ldS = (Nc + 3) / 4 * 4;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes); return o; };
    const size_t o_db = take(sizeof(bf16) * static_cast<size_t>(N) * K3), o_q = take(sizeof(bf16) * static_cast<size_t>(Mc) * K3);
    const size_t o_S = take(sizeof(float) * static_cast<size_t>(Mc) * ldS), o_ps = take(sizeof(float) * static_cast<size_t>(Mc) * cap);
    const size_t o_pi = take(sizeof(int) * static_cast<size_t>(Mc) * cap), o_pn = take(sizeof(int) * static_cast<size_t>(Mc));
    const size_t o_pt = take(sizeof(float) * static_cast<size_t>(Mc));
    uint8_t* base = g_ws.reserve(off, s);
    struct { bf16* p; } db{reinterpret_cast<bf16*>(base + o_db)}, q{reinterpret_cast<bf16*>(base + o_q)};
    struct { float* p; } S{reinterpret_cast<float*>(base + o_S)}, pool_s{reinterpret_cast<float*>(base + o_ps)},
        pool_thr{reinterpret_cast<float*>(base + o_pt)};
    struct { int* p; } pool_i{reinterpret_cast<int*>(base + o_pi)}, pool_n{reinterpret_cast<int*>(base + o_pn)};
    launch_kernel(rices_normalize_split_kernel, dim3(static_cast<unsigned>(ceil_div64(N, 8))), dim3(256), 0, s, database, N, D, db.p, 1);
    KERNEL_CHECK();
    count_launch();
    for (int64_t m0 = 0; m0 < M; m0 += Mc) {
        const int64_t mc = std::min(Mc, M - m0);
        launch_kernel(rices_normalize_split_kernel, dim3(static_cast<unsigned>(ceil_div64(mc, 8))), dim3(256), 0, s, queries + m0 * D, mc, D,
                      q.p, 0);
        KERNEL_CHECK();
        count_launch();
        launch_kernel(rices_pool_reset_kernel, dim3(static_cast<unsigned>(ceil_div64(mc, 256))), dim3(256), 0, s, pool_n.p, pool_thr.p,
                      static_cast<int>(mc));
        KERNEL_CHECK();
        count_launch();
        for (int64_t n0 = 0; n0 < N; n0 += Nc) {
            const int64_t nc = std::min(Nc, N - n0);
            GemmArgs a;
            a.A = q.p; a.lda = K3; a.B = db.p + n0 * K3; a.ldb = K3;
            a.M = static_cast<int>(mc); a.N = static_cast<int>(nc); a.K = K3;
            a.ep.out = S.p; a.ep.ldo = static_cast<int>(ldS); a.ep.out_fp32 = 1;
            gemm_bf16_tn(a, s);
            launch_kernel(rices_select_kernel, dim3(static_cast<unsigned>(mc)), dim3(256), sel_smem, s, S.p, ldS, static_cast<int>(nc), n0, k, cap,
                          pool_s.p, pool_i.p, pool_n.p, pool_thr.p);
            KERNEL_CHECK();
            count_launch();
        }
        launch_kernel(rices_finalize_kernel, dim3(static_cast<unsigned>(mc)), dim3(256), sel_smem, s, pool_s.p, pool_i.p, pool_n.p, cap, k,
                      out_scores + m0 * k, out_index + m0 * k);
        KERNEL_CHECK();
        count_launch();
    }
    CUDA_CHECK(cudaStreamSynchronize(s));       // the cached workspace may be reused by the next call on any stream
}

// query [M, D], table [n_table, D] fp32; cand [M, C] int32 rows of `table` (-1 = padding); out_sim [M, C], out_pos [M, C] int32
void rices_rerank(const float* query, const float* table, int64_t M, int D, const int* cand, int C, float* out_sim, int* out_pos,
                  cudaStream_t s) {
    EAVQA_CHECK(query && table && cand && out_sim && out_pos, "rices_rerank: null argument");
    EAVQA_CHECK(M > 0 && D > 0 && C >= 1 && C <= 4096, "rices_rerank: bad shape (1 <= candidates <= 4096)");
    const int P = 1 << static_cast<int>(std::ceil(std::log2(static_cast<double>(C))));
    static std::atomic<bool> configured{false};   // idempotent set-up: a race only repeats it
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(rices_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        configured = true;
    }
    launch_kernel(rices_rerank_kernel, dim3(static_cast<unsigned>(M)), dim3(256), sizeof(Cand) * static_cast<size_t>(P), s, query, table, D, cand, C,
                  out_sim, out_pos);
    KERNEL_CHECK();
    count_launch();
}

}  // namespace eavqa
