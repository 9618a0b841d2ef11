/*
 * eavqa_b200 -- C ABI of the B200-native CLIP-prefix language-model step.
 *
 * Drop-in boundary for ONE path of rs-anderson/explicit-alignment-for-vqa-tasks: the model object that
 * `ClipCapExecutor` builds and calls (reference `src/trainers/clipcap_exector.py:52-56,165-171,236-243`),
 * i.e. `ClipCaptionPrefix` of `src/models/clipcap.py:240-471,590-599` plus the in-context prefix splice of
 * `src/models/vct0.py:494-533`.  The reference is pure Python and has no FFI; each entry point below names
 * the reference interface it replaces.  The Python host (`eavqa_b200.model.ClipCaptionPrefixB200`) binds
 * these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `eavqa_last_error()` then returns a
 *     thread-local, NUL-terminated message.  Nothing throws across this boundary.
 *   - all tensor arguments are raw DEVICE pointers owned by the caller (torch); the library borrows them for
 *     the duration of the call and never retains them, except weights passed to `eavqa_load_lm_weight`,
 *     which are copied/packed.  Workspace, packed weights and the KV cache are owned by the handle.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; the caller's current torch stream)
 *     and is asynchronous; only `eavqa_generate` synchronises that stream once, at its end, to return the
 *     number of steps taken.  One handle per process/GPU; calls on one handle are not re-entrant.
 *   - there is no CPU path: every entry point fails if no sm_100 device is current.
 */
#ifndef EAVQA_B200_H
#define EAVQA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

#define EAVQA_ABI_VERSION 1

typedef struct eavqa_handle eavqa_handle;

enum { EAVQA_MAPPER_MLP = 0, EAVQA_MAPPER_TRANSFORMER = 1 };
enum { EAVQA_F32 = 0, EAVQA_BF16 = 1 };

/* Mirrors `ClipCaptionModel.__init__` kwargs (clipcap.py:240-249: prefix_length, clip_length, prefix_size,
 * num_layers, mapping_type) and the GPT2Config fields of the LM named by `model_version` (clipcap.py:252). */
typedef struct eavqa_config {
    int32_t n_layer, n_head, d_model;   /* GPT-2: 12/12/768, 24/16/1024, 36/20/1280, 48/25/1600; head_dim must be 64 */
    int32_t vocab, n_positions;         /* 50257 (+ added special tokens, clipcap_exector.py:56), 1024 */
    int32_t prefix_length, clip_length; /* P, and the transformer mapper's clip_length (clipcap.py:213-237) */
    int32_t clip_dim;                   /* prefix_size: 512 (ViT-B/32) or 768 (ViT-L/14) */
    int32_t mapper_type;                /* EAVQA_MAPPER_* ("mlp" vs anything else, clipcap.py:254-271) */
    int32_t mapper_layers;              /* num_layers of the transformer mapper (8) */
} eavqa_config;

const char* eavqa_last_error(void);
int eavqa_abi_version(void);

/* ClipCaptionModel.__init__ (clipcap.py:240-272) */
int eavqa_create(const eavqa_config* cfg, eavqa_handle** out);
int eavqa_destroy(eavqa_handle* h);

/* GPT2LMHeadModel.from_pretrained (clipcap.py:252) + resize_token_embeddings (clipcap_exector.py:56):
 * one call per HF state-dict tensor ("transformer.wte.weight", "transformer.h.0.attn.c_attn.weight", ...),
 * fp32 or bf16, `numel` elements on the device; the tensor is copied and packed.  `eavqa_finalize_lm`
 * checks that every tensor arrived. */
int eavqa_load_lm_weight(eavqa_handle* h, const char* name, const void* dev_ptr, int32_t dtype, int64_t numel, void* stream);
int eavqa_finalize_lm(eavqa_handle* h, void* stream);

/* Layout of the flat fp32 mapper parameter / gradient buffers: entries in the reference's
 * `clip_project.named_parameters()` order (clipcap.py:256-271). */
int64_t eavqa_mapper_param_count(const eavqa_handle* h);   /* total elements */
int32_t eavqa_mapper_num_tensors(const eavqa_handle* h);
int eavqa_mapper_tensor_info(const eavqa_handle* h, int32_t index, char* name_out, size_t name_cap, int64_t* offset,
                             int64_t* rows, int64_t* cols);

/* ClipCaptionModel.forward (clipcap.py:290-342) + loss.backward() restricted to clip_project
 * (ClipCaptionPrefix, clipcap.py:590-599).
 *   clip   [B, clip_dim] fp32        tokens / mask / labels [B, text_len] int64 (labels: -100 = ignore)
 *   params [param_count] fp32        grads [param_count] fp32 (overwritten; NULL = forward only)
 *   loss_out: device fp32 scalar */
int eavqa_train_step(eavqa_handle* h, int32_t batch, int32_t text_len, const float* clip, const int64_t* tokens,
                     const int64_t* mask, const int64_t* labels, const float* params, float* grads, float* loss_out,
                     void* stream);

/* The `.logits` of the object ClipCaptionModel.forward returns (clipcap.py:337-342: HF CausalLMOutput, logits [B, T, V]
 * with T = prefix_length + text_len), for callers that read them.  Forward only: mapper -> concat -> L blocks -> ln_f ->
 * tied head on EVERY position.  logits_out: device fp32 [B * T, ld] with ld >= vocab rounded up to a multiple of 64
 * (columns >= vocab are padding).  The training step itself never materialises this tensor (2.6 GB at B = 256). */
int eavqa_forward_logits(eavqa_handle* h, int32_t batch, int32_t text_len, const float* clip, const int64_t* tokens,
                         const int64_t* mask, const float* params, float* logits_out, int64_t ld, void* stream);

/* ClipCaptionModel.generate + _generate_from_embeddings (clipcap.py:344-471), with the k-shot prompt assembly of
 * VCT0Model.generate / insert_prefix_into_input (vct0.py:446-464,494-533) when n_images > 0.
 *   n_images = 0 : one prefix is prepended (clip [B, clip_dim]);
 *   n_images >= 1: clip [B, n_images, clip_dim]; each token id in [sentinel_lo, sentinel_hi] is replaced by the
 *                  next image's P prefix rows; every row must hold exactly n_images sentinels.
 *   tokens_out [B, max_new] int64 (device); has_eos = 0 reproduces eos_token_id=None.
 *   top_logit  [B, max_new] fp32 (device) or NULL: the winning logit of every step (diagnostics).
 *   token_logprob [B, max_new] fp32 (device) or NULL: log softmax(logits)[picked token] of every step -- what
 *              FewShotVQAExecutor.generate_from_ensembles reads out of `outputs.scores` (few_shot_vqa_executor.py:316-323).
 *   steps_out  (host int32): number of decode steps executed before every row had finished (clipcap.py:463);
 *              only tokens_out[:, :steps_out] is meaningful. */
int eavqa_generate(eavqa_handle* h, int32_t batch, int32_t text_len, int32_t n_images, const float* clip,
                   const int64_t* tokens, const int64_t* mask, int64_t sentinel_lo, int64_t sentinel_hi,
                   const float* params, int32_t max_new, int32_t has_eos, int64_t pad_id, int64_t eos_id,
                   int64_t* tokens_out, float* top_logit, float* token_logprob, int32_t* steps_out, void* stream);

/* VCT0Model.insert_prefix_into_input (vct0.py:494-533) on caller-supplied embeddings (golden-vector parity).
 *   text_table [vocab, d] fp32 rows looked up by token id; prefix [B, n_images*P, d] fp32
 *   out_emb [B, T_out, d] fp32, out_mask [B, T_out] int32, T_out = text_len + (P-1)*n_images */
int eavqa_splice(int32_t batch, int32_t text_len, int32_t n_images, int32_t prefix_length, int32_t d, int32_t vocab,
                 const int64_t* tokens, const int64_t* mask, int64_t sentinel_lo, int64_t sentinel_hi,
                 const float* text_table, const float* prefix, float* out_emb, int32_t* out_mask, void* stream);

/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t eavqa_launch_count(void);

/* Overlapping the data-parallel gradient exchange (Lightning DDP's bucketed all-reduce, main.py:138) with the mapper
 * backward.  The engine reports contiguous ranges [begin, end) of the flat gradient buffer in the order in which
 * eavqa_train_step finishes them (transformer mapper: pairs of layers, last layers first; MLP mapper: none).  The caller
 * installs one cudaEvent_t per range (it keeps owning them); every later eavqa_train_step records event k, on an internal
 * stream, once all gradients of range k are final, so a communication stream can wait on it and all-reduce that range
 * while the rest of the backward still runs.  Gradients outside the ranges are final when eavqa_train_step's work on
 * `stream` completes.  n = 0 uninstalls. */
int32_t eavqa_grad_bucket_count(const eavqa_handle* h);
int eavqa_grad_bucket_range(const eavqa_handle* h, int32_t index, int64_t* begin, int64_t* end);
int eavqa_set_grad_events(eavqa_handle* h, void* const* events, int32_t n);

/* ---- the steps right around the path inside the reference's executors ---- */
/* ClipCapExecutor.training_step label construction (clipcap_exector.py:134-150), replacing its Python double loop over
 * [B, T] tensor elements: labels = input_ids with pads -> -100, everything up to and including <BOS> -> -100, and the
 * FIRST pad position set back to pad_id (the EOS target).  tokens / labels [B, text_len] int64 on the device. */
int eavqa_build_caption_labels(const int64_t* tokens, int32_t batch, int32_t text_len, int64_t pad_id, int64_t bos_id,
                               int64_t* labels, void* stream);

/* FewShotVQAExecutor.generate_from_ensembles (few_shot_vqa_executor.py:293-332): logprob / tokens [n_ensembles, B, steps]
 * (token_logprob / tokens_out of `eavqa_generate`, one call per ensemble member); a row's sequence score is the sum of
 * the log-probabilities of its tokens not listed in skip_ids (the reference skips ids 0, 1, 2 = T5 pad / eos / unk);
 * best[b] = first argmax over members (np.argmax), best_tokens [B, steps] = the winning member's tokens;
 * scores [B, n_ensembles] fp32 or NULL. */
int eavqa_ensemble_select(const float* logprob, const int64_t* tokens, int32_t n_ensembles, int32_t batch, int32_t steps,
                          const int64_t* skip_ids, int32_t n_skip, float* scores, int32_t* best, int64_t* best_tokens,
                          void* stream);

/* RICES in-context example retrieval, the step before the few-shot path.
 * eavqa_rices_search replaces faiss.normalize_L2 + faiss.IndexFlatIP (GPU) .search(k) of
 * src/in_context_example_selection/get_question_knn.py:64-76: queries [n_queries, dim] and database [n_database, dim] are
 * fp32 device arrays (not modified; normalisation happens on packed copies); out_scores [n_queries, k] fp32 in descending
 * order, out_index [n_queries, k] int64 database rows (equal scores: lower row first; when n_database < k the tail is
 * -FLT_MAX / -1 as in faiss).  dim % 8 == 0, k <= 2048 (faiss' GPU limit too).  Scores are tensor-core inner products of
 * hi/lo-split bf16 operands with fp32 accumulation (error < 2e-5).  Synchronises `stream` before returning.
 * eavqa_rices_rerank replaces the per-question index of get_image_knn_from_text_knn.py:79-92: for question q the
 * candidates are rows candidates[q, 0..n_candidates) of `table` [*, dim] (-1 = padding); out_sim / out_pos
 * [n_queries, n_candidates]: cosine similarities in descending order and the candidate POSITIONS (faiss' I for the
 * per-question index); padding comes last as -FLT_MAX / -1.  n_candidates <= 4096. */
int eavqa_rices_search(const float* queries, const float* database, int64_t n_queries, int64_t n_database, int32_t dim, int32_t k,
                       float* out_scores, int64_t* out_index, void* stream);
int eavqa_rices_rerank(const float* query, const float* table, int64_t n_queries, int32_t dim, const int32_t* candidates,
                       int32_t n_candidates, float* out_sim, int32_t* out_pos, void* stream);

/* grads[0..n) *= *scale (device fp32 scalar: the upstream gradient autograd hands to backward()); skips the pass when
 * *scale == 1 without a host sync.  n % 4 == 0. */
int eavqa_scale_grads(float* grads, int64_t n, const float* scale, void* stream);

/* torch.optim.AdamW (clipcap_exector.py:79-81) on the flat mapper buffer, one fused pass:
 * params/grads/exp_avg/exp_avg_sq [n] fp32 on the device; `step` counts from 1; g = grads * grad_scale
 * (grad_scale folds the 1/world_size of the data-parallel mean into the update). */
int eavqa_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream);

/* Data-parallel exchange step fused with the optimiser (replaces Lightning DDP's NCCL all-reduce of the mapper gradients,
 * main.py:133-138, followed by torch.optim.AdamW on every rank, clipcap_exector.py:79-81): ONE kernel per rank does
 * reduce-scatter + AdamW + all-gather over NVLink / NVSwitch peer memory.  Rank r owns elements
 * [begin, end) = eavqa_sharded_adamw_range(n, r, world) of the flat buffer: it reads the SUM over ranks of that shard of the
 * gradients (mc_grads != NULL: `multimem.ld_reduce` on the NVLS multicast address, reduced inside the switch; else peer
 * loads summed in rank order), applies the AdamW update of eavqa_adamw_step with ITS shard of exp_avg / exp_avg_sq (full-size
 * local buffers, only the shard is touched), and stores the updated parameters into EVERY rank's parameter buffer
 * (`multimem.st`, or one peer store per rank).  grad_ptrs / param_ptrs / flag_ptrs [world]: rank r's buffer as mapped into
 * this process (symmetric-memory allocations of the host side); flag_ptrs: 64 zero-initialised uint32 per rank used by the
 * two barriers inside the kernel (all gradients final / all stores landed), `token` must grow by one per call, starting at 1;
 * flag_ptrs == NULL: no barriers inside (the caller brackets the call with its own cross-GPU barriers).
 * One call exchanges elements [offset, offset + n) of the buffers, sharded over the ranks by eavqa_sharded_adamw_range(n, ...)
 * + offset: the whole flat buffer after the step, or one gradient bucket (eavqa_grad_bucket_range) on a communication stream
 * as soon as its event (eavqa_set_grad_events) fires, on at most max_ctas CTAs (0 = one per SM) so that the backward still
 * running keeps its SMs.  Every rank must make the same sequence of calls, each on a stream ordered after the gradients of
 * its range; a call completes only after all ranks' stores into this rank's parameters have landed.  offset, n % 4 == 0,
 * world <= 16. */
int eavqa_sharded_adamw_range(int64_t n, int32_t rank, int32_t world, int64_t* begin, int64_t* end);
int eavqa_sharded_adamw_step(void* const* grad_ptrs, void* const* param_ptrs, const void* mc_grads, void* mc_params,
                             void* const* flag_ptrs, uint32_t token, int32_t rank, int32_t world, float* exp_avg, float* exp_avg_sq,
                             int64_t offset, int64_t n, int32_t max_ctas, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int32_t step, float grad_scale, void* stream);

/* Per-launch CUDA-event timing of the tcgen05 GEMM kernel (the dominant kernel; bench.py's roofline leg).
 * begin() arms it; end() synchronises the device and returns summed kernel milliseconds, FLOPs (2MNK) and
 * launch count since begin(), plus a per-shape text report (with the algorithmic HBM bytes per launch).  While armed,
 * the mapper's weight-gradient GEMMs stay on the caller's stream (normally they overlap the dgrad chain on a side
 * stream), so that per-launch durations do not overlap.  Not for use inside a timed region. */
int eavqa_profile_begin(void);
int eavqa_profile_end(double* total_ms, double* total_flops, int64_t* launches, char* report, size_t report_cap);

/* ---- single-operator entry points (unit parity tests of each kernel; same kernels the step uses) ---- */
/* D[M,N] = epi(A[M,K] * B[N,K]^T): bf16 operands, fp32 accumulate (tcgen05/TMEM); act/dact: see csrc/gemm.cuh.
 * block_n = tile width (0 = auto, 64/128/192/256) + 1000 * cta_mode (0 = auto, 1 = single CTAs, 8 = CTA pair / cta_group::2)
 *           + 100000 * split_k (> 1: K is split over that many CTAs per tile which ADD fp32 partials into a zeroed `out`).
 * Supported epilogues (csrc/gemm_kernel.cuh, EpiMode): bf16 out [+bias] [+gelu_new (+pre-activation out2) | relu | tanh];
 * bf16 out = acc * f'(aux) (dact, no bias); fp32 out [+bias [+residual]].  Other combinations return an error. */
int eavqa_op_gemm(const void* A, int32_t lda, const void* B, int32_t ldb, int32_t M, int32_t N, int32_t K, void* out,
                  int32_t ldo, int32_t out_fp32, const float* bias, const float* residual, int32_t ld_res, int32_t act,
                  const void* aux, int32_t ld_aux, int32_t dact, void* out2, int32_t ldo2, int32_t block_n, void* stream);
/* weight-gradient form D[M,N] (fp32) = At^T * Bt, At [K,M] and Bt [K,N] bf16 row-major (contraction over rows):
 * read as MN-major UMMA operands, no transposed copies (dW = dY^T X of every trainable nn.Linear) */
int eavqa_op_gemm_wgrad(const void* At, int32_t ldat, const void* Bt, int32_t ldbt, int32_t M, int32_t N, int32_t K, float* out,
                        int32_t ldo, int32_t block_n, void* stream);
/* LM-head GEMM with fused softmax statistics: logits [M, ldo] as 2-byte fp16 values (they only feed d logits; fp16 is 8x finer
 * than bf16 at trained-model logit magnitudes), lse [M] and target logit [M] */
int eavqa_op_lmhead_ce(const void* H, const void* W, int32_t M, int32_t vocab, int32_t n_cols, int32_t K, const int32_t* label,
                       void* logits, int32_t ldo, float* lse, float* target, float* loss_sum, void* stream);
int eavqa_op_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                           int32_t M, int32_t d, void* stream);
int eavqa_op_layernorm_bwd(const void* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                           float* dx, int32_t accumulate, void* dx_bf16, float* dgamma, float* dbeta, int32_t M, int32_t d,
                           void* stream);
int eavqa_op_lm_attention_fwd(const void* qkv, const int32_t* valid, void* o, float* lse, int32_t B, int32_t T, int32_t H,
                              void* stream);
int eavqa_op_lm_attention_bwd(const void* qkv, const int32_t* valid, const void* o, const void* d_o, const float* lse,
                              void* dqkv, float* dq_scratch, int32_t B, int32_t T, int32_t H, void* stream);
int eavqa_op_mapper_attention_fwd(const void* qkv, void* o, int32_t B, int32_t S, int32_t H, int32_t hd, void* stream);
int eavqa_op_mapper_attention_bwd(const void* qkv, const void* d_o, void* dqkv, int32_t B, int32_t S, int32_t H, int32_t hd,
                                  void* stream);
int eavqa_op_convert_transpose(const float* src, int32_t R, int32_t C, void* dst, void* dst_t, int32_t ld_t, float* colsum,
                               void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* EAVQA_B200_H */
