// Host launchers of the fused memory-bound kernels and the attention kernels.
// All launch on the given stream and never synchronise.
#pragma once
#include <vector>

#include "common.cuh"

namespace eavqa {

constexpr int kGreedySplitMax = 16;    // CTAs per row of the greedy-decode vocabulary scan (scratch sizing)

void count_launch(int n = 1);
int64_t kernel_launch_count();

// ---------------------------------------------------------------- packing / layout (elementwise.cu)
// fp32 [R, C] (ld_src) -> bf16 copy [R, C] (ld_dst) and/or bf16 transpose [C, R] (ld_t) and/or fp32 column sums
// (atomicAdd into colsum[C]; caller zeroes).  Either destination may be null.
void convert_transpose_f32(const float* src, int ld_src, int R, int C, bf16* dst, int ld_dst, bf16* dst_t, int ld_t,
                           float* colsum, cudaStream_t s);
void convert_transpose_bf16(const bf16* src, int ld_src, int R, int C, bf16* dst_t, int ld_t, float* colsum,
                            cudaStream_t s);
// many fp32 [rows, cols] matrices (dense, ld = cols; even sizes) -> bf16 copies / transposes in ONE launch
struct PackJob {
    const float* src;
    bf16* dst;        // [rows, cols] or null
    bf16* dst_t;      // [cols, rows] or null
    int rows, cols;
    int tile_begin;   // filled by pack_batch
};
struct PackJobs {
    static constexpr int kMax = 48;
    PackJob job[kMax];
    int n;
};
void pack_batch(const std::vector<PackJob>& jobs, cudaStream_t s);
// dst[r, :] = src[:]  for r < R   (prefix_const rows of the mapper input), fp32
void broadcast_rows_f32(const float* src, int rows, int d, float* dst, int64_t batch_stride, int B, cudaStream_t s);
// out[c] (+)= sum_b src[b*batch_stride + c]   (prefix_const gradient), c < n
void sum_over_batch_f32(const float* src, int64_t batch_stride, int B, int n, float* out, cudaStream_t s);
void fill_zero(void* p, size_t bytes, cudaStream_t s);
// hint: pull [p, p + bytes) into L2 (cp.async.bulk.prefetch.L2, fire-and-forget; p 16-byte aligned)
void prefetch_l2(const void* p, size_t bytes, cudaStream_t s);

// ---------------------------------------------------------------- LayerNorm (elementwise.cu)
// y = LN(x[row]) * gamma + beta, bf16 out; x row = row_index ? row_index[m] : m.  Saves mean / rstd (optional).
void layernorm_fwd(const float* x, int ld_x, const int* row_index, const float* gamma, const float* beta, bf16* y,
                   int ld_y, float* mean, float* rstd, int M, int d, float eps, cudaStream_t s);
// dx[row] = (accumulate ? dx[row] : 0) + LN'(dy) ; optional bf16 copy of the result; optional dgamma/dbeta (atomicAdd).
// mean / rstd are recomputed from x when null.
void layernorm_bwd(const bf16* dy, int ld_dy, const float* x, int ld_x, const int* row_index, const float* gamma,
                   const float* mean, const float* rstd, float* dx, int ld_dx, int accumulate, bf16* dx_bf16,
                   int ld_dxb, float* dgamma, float* dbeta, int M, int d, float eps, cudaStream_t s);

// dgamma[c] += sum_m dy[m,c] * xhat[m,c],  dbeta[c] += sum_m dy[m,c]   (trainable mapper LayerNorms; atomicAdd)
void layernorm_param_grads(const bf16* dy, int ld_dy, const float* x, int ld_x, const float* mean, const float* rstd,
                           float* dgamma, float* dbeta, int M, int d, cudaStream_t s);

// ---------------------------------------------------------------- single-token decode glue (elementwise.cu)
// x[row] += acc[row] + bias (acc, bias may both be null); u[row] = LN(x[row]) (bf16); zero[row, 0..zero_n) = 0 (may be null)
void decode_residual_ln(float* x, const float* acc, const float* bias, const float* gamma, const float* beta, bf16* u, int rows, int d,
                        float eps, float* zero, int zero_n, cudaStream_t s);

// ---------------------------------------------------------------- embedding / splice (elementwise.cu)
// plan[b, t] >= 0 : token id ; < 0 : -(prefix_row + 1) ; valid[b, t] = key-validity (attention mask)
void prepend_plan(const int64_t* tokens, const int64_t* mask, int B, int Tt, int P, int* plan, int* valid, cudaStream_t s);
// vct0.py:494-533 semantics; sentinel ids in [sent_lo, sent_hi]; err_flag set to 1 when a row does not hold n_img sentinels
void splice_plan(const int64_t* tokens, const int64_t* mask, int B, int Tt, int P, int n_img, int64_t sent_lo,
                 int64_t sent_hi, int* plan, int* valid, int* err_flag, cudaStream_t s);
// h0[b, t, :] = (plan >= 0 ? wte[plan] : prefix[b, -plan-1, :]) + (wpe ? wpe[t] : 0)
void embed_rows(const int* plan, int B, int T, int d, const float* wte, int vocab, const float* prefix,
                int64_t prefix_batch_stride, int prefix_row_stride, const float* wpe, float* out, cudaStream_t s);

// ---------------------------------------------------------------- LM head cross-entropy (elementwise.cu)
// rows r = b*Tt + j predict text token j from hidden row b*T + P-1+j (labels are shifted inside HF, loss_utils.py:57-60)
void ce_plan(const int64_t* labels, int B, int Tt, int T, int P, int vocab, int* row_index, int* label, int* n_valid,
             cudaStream_t s);
// merges per-tile (max, sumexp) partials -> lse[r]; accumulates sum of (lse - target) over valid rows into loss_sum
void ce_finalize(const float2* partial, int tiles, const float* target, const int* label, float* lse, float* loss_sum,
                 int M, cudaStream_t s);
// loss = loss_sum / n_valid (NaN when no valid target, as the reference's mean over an empty set)
void ce_loss(const float* loss_sum, const int* n_valid, float* loss_out, cudaStream_t s);
// in place: z[r, v] (fp16 logits as the head's CE epilogue stores them) <- bf16 (softmax(z[r])[v] - [v == label[r]]) / n_valid for
// valid rows, 0 otherwise / for v >= vocab
void ce_dlogits(bf16* z, int ld, int M, int vocab, int n_cols, const float* lse, const int* label, const int* n_valid,
                cudaStream_t s);

// ---------------------------------------------------------------- greedy decode bookkeeping (elementwise.cu)
// per row: nxt = argmax(logits[:vocab]); out = unfinished ? nxt : pad; unfinished &= out != eos;
// tokens_out[b, step] = out; x_next[b] = wte[nxt] + wpe[pos]; n_unfinished[step] = sum(unfinished)
void greedy_step(const float* logits, int ld, int B, int vocab, int step, int max_new, int has_eos, int64_t pad_id,
                 int64_t eos_id, int* unfinished, int64_t* tokens_out, int* n_unfinished, float* top_logit,
                 float* token_logprob /* optional: log softmax of the picked token */, const float* wte, const float* wpe_row,
                 int d, float* x_next, int* valid_next, int valid_stride,
                 float* part_val, int* part_idx, unsigned* arrivals /* scratch [B, kGreedySplitMax] x2 + zeroed [B]; may be null */,
                 cudaStream_t s);

// ---------------------------------------------------------------- executor-side steps (elementwise.cu)
// ClipCapExecutor.training_step label construction (clipcap_exector.py:134-150): labels [B, T] int64
void caption_labels(const int64_t* tokens, int B, int T, int64_t pad_id, int64_t bos_id, int64_t* labels, cudaStream_t s);
// FewShotVQAExecutor.generate_from_ensembles (few_shot_vqa_executor.py:316-331): logprob / tokens [E, B, S];
// scores [B, E] (optional), best [B], best_tokens [B, S]
void ensemble_select(const float* logprob, const int64_t* tokens, int E, int B, int S, const int64_t* skip, int n_skip,
                     float* scores, int* best, int64_t* best_tokens, cudaStream_t s);
// x[0..n) *= *scale (device scalar); no memory traffic when *scale == 1
void scale_by_device_scalar(float* x, int64_t n, const float* scale, cudaStream_t s);

// ---------------------------------------------------------------- RICES retrieval (rices.cu)
// faiss.normalize_L2 + IndexFlatIP.search (get_question_knn.py:64-76): queries [M, D], database [N, D] fp32;
// out_scores [M, k] fp32 descending, out_index [M, k] int64 (ties: lower index first; fewer than k rows: -FLT_MAX / -1)
void rices_search(const float* queries, const float* database, int64_t M, int64_t N, int D, int k, float* out_scores,
                  int64_t* out_index, cudaStream_t s);
// per-question re-ranking of candidate rows of `table` (get_image_knn_from_text_knn.py:79-92): cand [M, C] int32, -1 = padding
void rices_rerank(const float* query, const float* table, int64_t M, int D, const int* cand, int C, float* out_sim, int* out_pos,
                  cudaStream_t s);

// ---------------------------------------------------------------- optimiser (elementwise.cu)
// torch.optim.AdamW semantics on the flat mapper buffer (clipcap_exector.py:79-81): decoupled weight decay,
// bias correction; g = grads * grad_scale
void adamw_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                float eps, float weight_decay, int step, float grad_scale, cudaStream_t s);

// ---------------------------------------------------------------- gradient exchange fused with the optimiser (collective.cu)
// Elements [begin, end) of the flat buffer that `rank` owns (multiples of 4; contiguous shards of ceil(n / 4 / world) float4).
void sharded_adamw_range(int64_t n, int rank, int world, int64_t* begin, int64_t* end);
// ONE kernel: sum of all ranks' gradients for this rank's shard (NVLS multicast load-reduce, or peer loads in rank order),
// AdamW on the shard, updated parameters stored to every rank (multicast store, or peer stores).  grad_ptrs / param_ptrs /
// flag_ptrs: rank r's buffer as mapped into THIS process (symmetric memory); mc_*: multicast addresses or null; flag_ptrs
// null = no barriers inside the kernel (the caller brackets the call with its own cross-GPU barriers).  The call exchanges
// elements [offset, offset + n) of the buffers (sharded over the ranks) on at most max_ctas CTAs (0 = one per SM).
void sharded_adamw_step(void* const* grad_ptrs, void* const* param_ptrs, const void* mc_grads, void* mc_params, void* const* flag_ptrs,
                        uint32_t token, int rank, int world, float* m, float* v, int64_t offset, int64_t n, int max_ctas, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t s);

// ---------------------------------------------------------------- attention (attention.cu)
// GPT-2 causal attention with key-padding mask, head_dim 64, qkv [B*T, 3d] bf16 (q | k | v, heads contiguous).
// o [B*T, d] bf16; lse [B, H, T] fp32 (of scaled scores).
// kv_cache != null (generation prefill): also writes every K / V row into the head-major KV cache of this layer
void lm_attention_fwd(const bf16* qkv, const int* valid, bf16* o, float* lse, int B, int T, int H, cudaStream_t s,
                      bf16* kv_cache = nullptr, int Tmax = 0);
// dqkv [B*T, 3d] bf16; dq_scratch fp32 [B*T, d] (only touched when T > 64)
void lm_attention_bwd(const bf16* qkv, const int* valid, const bf16* o, const bf16* d_o, const float* lse, bf16* dqkv,
                      float* dq_scratch, int B, int T, int H, cudaStream_t s);
// KV cache per layer: K block [B, H, Tmax, 64] then V block [B, H, Tmax, 64] (bf16, head-major)
// one new query per (b, h): q | k | v given as fp32 GEMM accumulators [B, 3d] + bias [3d] (split-K decode path); appends this
// step's k, v at position pos and attends over keys [0, pos] where valid; four warps per (sample, head) share the keys;
// also zeroes zero[b, h*64 .. h*64+64) (the next GEMM's accumulator rows, [B, d])
void lm_attention_decode_acc(const float* qkv_acc, const float* qkv_bias, bf16* cache, const int* valid, int valid_stride, bf16* o,
                             float* zero, int B, int H, int pos, int Tmax, cudaStream_t s);

// ---------------------------------------------------------------- persistent decode step (decode_chain.cu)
// One cooperative kernel per single-token step walks an array of phases (in device memory) with a grid barrier between them.
enum ChainPhaseType { CHAIN_GEMM = 0, CHAIN_ATTN = 1, CHAIN_GLUE = 2 };
enum ChainGemmMode {
    CHAIN_REDUCE_F32 = 0,     // split-K: fp32 partial tiles are ADDED into `out` (TMA reduce; `out` zeroed by an earlier phase)
    CHAIN_STORE_F32 = 1,      // fp32 out (LM head logits)
    CHAIN_GELU_BF16 = 2,      // bf16 out = gelu_new(acc + bias) (c_fc)
};
struct alignas(128) ChainPhase {
    CUtensorMap map_a, map_b, map_out;      // GEMM: A [M, K] (box 128 x 64), W [N, K] (box 64 x 64), out [M, N] (box 32 x 32)
    int type, mode;
    int M, N, K, split;                     // GEMM: out[M, N] (+)= A[M, K] W[N, K]^T, M <= 128 rows
    int B, H, Tmax, valid_stride;           // ATTN: batch rows, heads, KV-cache capacity; GLUE: B = rows
    int d, zero_n;                          // GLUE: row width; floats to zero per row of `zero`
    const float* bias;                      // GEMM gelu bias [N] / ATTN qkv bias [3d] / GLUE projection bias [d]
    const float* qkv_acc;                   // ATTN: fp32 q | k | v accumulator rows [B, 3d]
    bf16* cache;                            // ATTN: this layer's KV cache
    const int* valid;                       // ATTN: key validity [B, valid_stride]
    bf16* o;                                // ATTN: attention output [B, d]
    float* zero;                            // ATTN: [B, d] accumulator zeroed per (b, h); GLUE: [rows, zero_n]
    float* x;                               // GLUE: fp32 residual rows [rows, d]
    const float* acc;                       // GLUE: fp32 projection accumulator [rows, d] or null
    const float *gamma, *beta;              // GLUE: LayerNorm affine
    bf16* u;                                // GLUE: LayerNorm output [rows, d]
};
bool decode_chain_supported(int max_keys);  // the attention phases stage one head's whole history in shared memory
void chain_gemm_phase(ChainPhase& p, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, int split, int mode, void* out,
                      int ldo, const float* bias);
void chain_attn_phase(ChainPhase& p, const float* qkv_acc, const float* qkv_bias, bf16* cache, const int* valid, int valid_stride, bf16* o,
                      float* zero, int B, int H, int Tmax);
void chain_glue_phase(ChainPhase& p, float* x, const float* acc, const float* bias, const float* gamma, const float* beta, bf16* u, int rows,
                      int d, float* zero, int zero_n);
// bar_counter: device counter zeroed once per generate call; epoch0 = grid barriers executed by earlier launches since then
// trace (optional, EAVQA_CHAIN_TRACE=1): [1 + 2 n_phases] globaltimer stamps of CTA 0 -- start, then per phase (own work done, barrier passed)
void launch_decode_chain(const ChainPhase* dev_phases, int n_phases, int pos, unsigned* bar_counter, unsigned epoch0, cudaStream_t s,
                         unsigned long long* trace = nullptr);

// mapper self-attention (no mask), S = clip_length + prefix_length small, any head_dim: qkv [B*S, 3d] bf16
void mapper_attention_fwd(const bf16* qkv, bf16* o, int B, int S, int H, int hd, cudaStream_t s);
void mapper_attention_bwd(const bf16* qkv, const bf16* d_o, bf16* dqkv, int B, int S, int H, int hd, cudaStream_t s);

}  // namespace eavqa
