// Host interface of the tcgen05/TMEM GEMM (gemm_tcgen05.cu).
//
//   D[M, N] = epilogue( A[M, K] * B[N, K]^T )        A, B bf16, K contiguous ("TN"), fp32 accumulate in TMEM
//
// Every matmul on the CLIP-prefix LM path is expressed in this one form by keeping, for each
// weight, the copy whose contraction dimension is contiguous (frozen LM weights are packed once
// in both orientations; the trainable mapper's are re-packed each step by pack_weight_kernel).
// The epilogue combinations the kernels are compiled for are listed in gemm_kernel.cuh (EpiMode);
// any other combination of the fields below is rejected with an error.
#pragma once
#include <string>

#include "common.cuh"

namespace eavqa {

enum GemmAct { ACT_NONE = 0, ACT_GELU_NEW = 1, ACT_TANH = 2, ACT_RELU = 3 };
// derivative applied to the accumulator (dgrad through an activation):
//   DACT_GELU_NEW : aux = pre-activation          D = acc * gelu_new'(aux)
//   DACT_TANH     : aux = tanh output              D = acc * (1 - aux^2)
//   DACT_RELU     : aux = relu output              D = acc * (aux > 0)
enum GemmDact { DACT_NONE = 0, DACT_GELU_NEW = 1, DACT_TANH = 2, DACT_RELU = 3 };

struct GemmEpilogue {
    void* out = nullptr;             // [M, ldo] bf16 (out_fp32 = 0) or fp32 (out_fp32 = 1)
    bf16* out2 = nullptr;            // optional second output, bf16 [M, ldo2]: the PRE-activation value
    const float* bias = nullptr;     // optional fp32 [N]
    const float* residual = nullptr; // optional fp32 [M, ld_res], added last
    const bf16* aux = nullptr;       // bf16 [M, ld_aux], operand of `dact`
    int ldo = 0, ldo2 = 0, ld_res = 0, ld_aux = 0;
    int out_fp32 = 0;
    int act = ACT_NONE;
    int dact = DACT_NONE;
    // ---- fused softmax-cross-entropy statistics (LM head): per row and per N-tile running
    //      (max, sum exp) over columns < n_valid, plus the fp32 logit of the row's label.
    float2* ce_partial = nullptr;    // [M, ce_tiles], ce_tiles = 2 * ceil(N / block_n) (one pair per half tile)
    float* ce_target = nullptr;      // [M]
    const int* ce_label = nullptr;   // [M], -1 = none
    int ce_tiles = 0;
    int n_valid = 0;
    int split_k = 1;                 // > 1: K is split over `split_k` CTAs per output tile which ADD their fp32 partials
                                     // into `out` (TMA reduce-add; out must be pre-initialised; plain fp32 epilogue only)
    int debug = 0;                   // EAVQA_GEMM_DEBUG (timing experiments only; results are wrong when non-zero):
                                     // 1 = issue every other TMA store, 2 = no TMA stores, 3 = no staging and no stores
};

struct GemmArgs {
    const bf16* A = nullptr;
    const bf16* B = nullptr;
    int lda = 0, ldb = 0;            // row strides in elements (multiples of 8)
    int M = 0, N = 0, K = 0;
    // "wgrad form": D[M, N] = At^T * Bt with At [K, M] and Bt [K, N] row-major (contraction over their ROWS), read as
    // MN-major UMMA operands -- no transposed copies.  A / B then point at At / Bt and lda / ldb are their row strides.
    int mn_major = 0;
    int block_n = 0;                 // 0 = pick by wave-quantisation heuristic; else 64/128/192/256
    int cluster = 0;                 // 0 = heuristic; 1 = single CTAs; 8 = CTA pair (tcgen05 cta_group::2, 256 x BN tiles)
    GemmEpilogue ep;
};

// Number of N tiles the heuristic (or block_n) will use: callers size ce_partial with it.
int gemm_pick_block_n(int M, int N, int K, int forced);
// tile width + CTA mode (1 = single CTAs, 8 = CTA pair / cta_group::2) the launcher will use for this problem
void gemm_pick_config(int M, int N, int K, int forced_bn, int forced_cluster, int* bn_out, int* cluster_out);
void gemm_bf16_tn(const GemmArgs& a, cudaStream_t stream);
// kernels this translation unit launched since process start (bench.py's gpu_launches)
int64_t gemm_launch_count();
// per-launch CUDA-event timing of the GEMM kernel (bench.py's roofline leg): begin() arms it, end() synchronises,
// returns total kernel milliseconds / FLOPs (2MNK) / launches and a per-shape text report
void gemm_profile_begin();
bool gemm_profile_active();      // while armed the engine keeps every kernel on one stream, so launch durations do not overlap
void gemm_profile_end(double* total_ms, double* total_flops, int64_t* launches, std::string* report);

}  // namespace eavqa
