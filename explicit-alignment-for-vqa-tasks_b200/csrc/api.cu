// extern "C" boundary (include/eavqa_b200.h): status codes + thread-local error text; no exceptions escape.
#include <cstring>
#include <string>

#include "../../include/eavqa_b200.h"
#include "engine.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

using namespace eavqa;

struct eavqa_handle {
    Engine* engine;
};

static thread_local std::string g_last_error;

#define API_BEGIN try {
#define API_END                                         \
    return 0;                                           \
    }                                                   \
    catch (const std::exception& e) {                   \
        g_last_error = e.what();                        \
        return 1;                                       \
    }                                                   \
    catch (...) {                                       \
        g_last_error = "unknown error";                 \
        return 2;                                       \
    }

static cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* eavqa_last_error(void) { return g_last_error.c_str(); }
int eavqa_abi_version(void) { return EAVQA_ABI_VERSION; }
int64_t eavqa_launch_count(void) { return gemm_launch_count() + kernel_launch_count(); }

int eavqa_create(const eavqa_config* cfg, eavqa_handle** out) {
    API_BEGIN
    EAVQA_CHECK(cfg != nullptr && out != nullptr, "null argument");
    eavqa_handle* h = new eavqa_handle;
    h->engine = nullptr;
    try {
        h->engine = new Engine(*cfg);
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    API_END
}

int eavqa_destroy(eavqa_handle* h) {
    API_BEGIN
    if (h != nullptr) {
        cudaDeviceSynchronize();
        delete h->engine;
        delete h;
    }
    API_END
}

int eavqa_load_lm_weight(eavqa_handle* h, const char* name, const void* dev_ptr, int32_t dtype, int64_t numel, void* stream) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr && name != nullptr, "null argument");
    h->engine->load_lm_weight(name, dev_ptr, dtype, numel, S(stream));
    API_END
}

int eavqa_finalize_lm(eavqa_handle* h, void* stream) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr, "null handle");
    h->engine->finalize_lm(S(stream));
    API_END
}

int64_t eavqa_mapper_param_count(const eavqa_handle* h) { return h ? h->engine->mapper_param_count() : -1; }
int32_t eavqa_mapper_num_tensors(const eavqa_handle* h) { return h ? static_cast<int32_t>(h->engine->mapper_tensors().size()) : -1; }

int eavqa_mapper_tensor_info(const eavqa_handle* h, int32_t index, char* name_out, size_t name_cap, int64_t* offset,
                             int64_t* rows, int64_t* cols) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr, "null handle");
    const auto& ts = h->engine->mapper_tensors();
    EAVQA_CHECK(index >= 0 && index < static_cast<int32_t>(ts.size()), "tensor index out of range");
    const TensorInfo& t = ts[index];
    if (name_out != nullptr && name_cap > 0) {
        std::strncpy(name_out, t.name.c_str(), name_cap - 1);
        name_out[name_cap - 1] = 0;
    }
    if (offset) *offset = t.offset;
    if (rows) *rows = t.rows;
    if (cols) *cols = t.cols;
    API_END
}

int eavqa_train_step(eavqa_handle* h, int32_t batch, int32_t text_len, const float* clip, const int64_t* tokens,
                     const int64_t* mask, const int64_t* labels, const float* params, float* grads, float* loss_out,
                     void* stream) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr, "null handle");
    h->engine->train_step(batch, text_len, clip, tokens, mask, labels, params, grads, loss_out, S(stream));
    API_END
}

int eavqa_forward_logits(eavqa_handle* h, int32_t batch, int32_t text_len, const float* clip, const int64_t* tokens,
                         const int64_t* mask, const float* params, float* logits_out, int64_t ld, void* stream) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr, "null handle");
    EAVQA_CHECK(logits_out != nullptr, "null logits_out");
    h->engine->train_step(batch, text_len, clip, tokens, mask, nullptr, params, nullptr, nullptr, S(stream), logits_out, ld);
    API_END
}

int eavqa_generate(eavqa_handle* h, int32_t batch, int32_t text_len, int32_t n_images, const float* clip,
                   const int64_t* tokens, const int64_t* mask, int64_t sentinel_lo, int64_t sentinel_hi,
                   const float* params, int32_t max_new, int32_t has_eos, int64_t pad_id, int64_t eos_id,
                   int64_t* tokens_out, float* top_logit, float* token_logprob, int32_t* steps_out, void* stream) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr, "null handle");
    h->engine->generate(batch, text_len, n_images, clip, tokens, mask, sentinel_lo, sentinel_hi, params, max_new, has_eos,
                        pad_id, eos_id, tokens_out, top_logit, token_logprob, steps_out, S(stream));
    API_END
}

int32_t eavqa_grad_bucket_count(const eavqa_handle* h) {
    if (h == nullptr) return 0;
    try {
        return static_cast<int32_t>(h->engine->grad_buckets().size());
    } catch (...) {
        return 0;
    }
}

int eavqa_grad_bucket_range(const eavqa_handle* h, int32_t index, int64_t* begin, int64_t* end) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr && begin != nullptr && end != nullptr, "null argument");
    const auto b = h->engine->grad_buckets();
    EAVQA_CHECK(index >= 0 && index < static_cast<int32_t>(b.size()), "bucket index out of range");
    *begin = b[index].first;
    *end = b[index].second;
    API_END
}

int eavqa_set_grad_events(eavqa_handle* h, void* const* events, int32_t n) {
    API_BEGIN
    EAVQA_CHECK(h != nullptr, "null handle");
    h->engine->set_grad_events(events, n);
    API_END
}

int eavqa_build_caption_labels(const int64_t* tokens, int32_t batch, int32_t text_len, int64_t pad_id, int64_t bos_id,
                               int64_t* labels, void* stream) {
    API_BEGIN
    EAVQA_CHECK(tokens && labels && batch > 0 && text_len > 0, "bad argument");
    caption_labels(tokens, batch, text_len, pad_id, bos_id, labels, S(stream));
    API_END
}

int eavqa_ensemble_select(const float* logprob, const int64_t* tokens, int32_t n_ensembles, int32_t batch, int32_t steps,
                          const int64_t* skip_ids, int32_t n_skip, float* scores, int32_t* best, int64_t* best_tokens,
                          void* stream) {
    API_BEGIN
    EAVQA_CHECK(logprob && tokens && best && best_tokens && n_ensembles > 0 && batch > 0 && steps > 0, "bad argument");
    EAVQA_CHECK(n_skip == 0 || skip_ids != nullptr, "skip_ids is null");
    ensemble_select(logprob, tokens, n_ensembles, batch, steps, skip_ids, n_skip, scores, best, best_tokens, S(stream));
    API_END
}

int eavqa_rices_search(const float* queries, const float* database, int64_t n_queries, int64_t n_database, int32_t dim, int32_t k,
                       float* out_scores, int64_t* out_index, void* stream) {
    API_BEGIN
    rices_search(queries, database, n_queries, n_database, dim, k, out_scores, out_index, S(stream));
    API_END
}

int eavqa_rices_rerank(const float* query, const float* table, int64_t n_queries, int32_t dim, const int32_t* candidates,
                       int32_t n_candidates, float* out_sim, int32_t* out_pos, void* stream) {
    API_BEGIN
    rices_rerank(query, table, n_queries, dim, candidates, n_candidates, out_sim, out_pos, S(stream));
    API_END
}

int eavqa_scale_grads(float* grads, int64_t n, const float* scale, void* stream) {
    API_BEGIN
    EAVQA_CHECK(grads && scale && n > 0, "bad argument");
    scale_by_device_scalar(grads, n, scale, S(stream));
    API_END
}

int eavqa_splice(int32_t batch, int32_t text_len, int32_t n_images, int32_t prefix_length, int32_t d, int32_t vocab,
                 const int64_t* tokens, const int64_t* mask, int64_t sentinel_lo, int64_t sentinel_hi,
                 const float* text_table, const float* prefix, float* out_emb, int32_t* out_mask, void* stream) {
    API_BEGIN
    EAVQA_CHECK(batch > 0 && text_len > 0 && n_images > 0 && prefix_length > 0 && d > 0, "bad splice shape");
    EAVQA_CHECK(tokens && text_table && prefix && out_emb && out_mask, "null argument");
    const int T_out = text_len + (prefix_length - 1) * n_images;
    int *plan = nullptr, *err = nullptr;
    CUDA_CHECK(cudaMalloc(&plan, sizeof(int) * static_cast<size_t>(batch) * T_out));
    CUDA_CHECK(cudaMalloc(&err, sizeof(int)));
    int host_err = 0;
    try {
        CUDA_CHECK(cudaMemsetAsync(err, 0, sizeof(int), S(stream)));
        splice_plan(tokens, mask, batch, text_len, prefix_length, n_images, sentinel_lo, sentinel_hi, plan, out_mask, err, S(stream));
        embed_rows(plan, batch, T_out, d, text_table, vocab, prefix, static_cast<int64_t>(n_images) * prefix_length * d, d,
                   nullptr, out_emb, S(stream));
        CUDA_CHECK(cudaMemcpyAsync(&host_err, err, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
        CUDA_CHECK(cudaStreamSynchronize(S(stream)));
    } catch (...) {
        cudaFree(plan);
        cudaFree(err);
        throw;
    }
    cudaFree(plan);
    cudaFree(err);
    EAVQA_CHECK(host_err == 0, "prompt rows must each hold exactly n_images sentinel tokens (vct0.py:512 would fail its .view)");
    API_END
}

// -------------------------------------------------------------------------------------------- operators
int eavqa_op_gemm(const void* A, int32_t lda, const void* B, int32_t ldb, int32_t M, int32_t N, int32_t K, void* out,
                  int32_t ldo, int32_t out_fp32, const float* bias, const float* residual, int32_t ld_res, int32_t act,
                  const void* aux, int32_t ld_aux, int32_t dact, void* out2, int32_t ldo2, int32_t block_n, void* stream) {
    API_BEGIN
    GemmArgs a;
    a.A = static_cast<const bf16*>(A); a.lda = lda; a.B = static_cast<const bf16*>(B); a.ldb = ldb;
    a.M = M; a.N = N; a.K = K; a.block_n = block_n % 1000; a.cluster = (block_n / 1000) % 100;
    a.ep.split_k = block_n / 100000;      // > 1: fp32 partials are ADDED into `out` (caller zero-initialises)
    a.ep.out = out; a.ep.ldo = ldo; a.ep.out_fp32 = out_fp32; a.ep.bias = bias; a.ep.residual = residual; a.ep.ld_res = ld_res;
    a.ep.act = act; a.ep.aux = static_cast<const bf16*>(aux); a.ep.ld_aux = ld_aux; a.ep.dact = dact;
    a.ep.out2 = static_cast<bf16*>(out2); a.ep.ldo2 = ldo2;
    gemm_bf16_tn(a, S(stream));
    API_END
}

int eavqa_op_gemm_wgrad(const void* At, int32_t ldat, const void* Bt, int32_t ldbt, int32_t M, int32_t N, int32_t K, float* out,
                        int32_t ldo, int32_t block_n, void* stream) {
    API_BEGIN
    GemmArgs a;
    a.A = static_cast<const bf16*>(At); a.lda = ldat; a.B = static_cast<const bf16*>(Bt); a.ldb = ldbt;
    a.M = M; a.N = N; a.K = K; a.block_n = block_n % 1000; a.mn_major = 1;
    a.ep.out = out; a.ep.ldo = ldo; a.ep.out_fp32 = 1;
    a.ep.split_k = block_n / 1000;        // > 1: partials are ADDED into `out` (caller zero-initialises)
    gemm_bf16_tn(a, S(stream));
    API_END
}

int eavqa_op_lmhead_ce(const void* H, const void* W, int32_t M, int32_t vocab, int32_t n_cols, int32_t K, const int32_t* label,
                       void* logits, int32_t ldo, float* lse, float* target, float* loss_sum, void* stream) {
    API_BEGIN
    int bn = 0, cluster = 1;
    gemm_pick_config(M, n_cols, K, 0, 0, &bn, &cluster);
    const int tiles = 2 * ceil_div(n_cols, bn);
    float2* partial = nullptr;
    CUDA_CHECK(cudaMalloc(&partial, sizeof(float2) * static_cast<size_t>(M) * tiles));
    try {
        GemmArgs a;
        a.A = static_cast<const bf16*>(H); a.lda = K; a.B = static_cast<const bf16*>(W); a.ldb = K;
        a.M = M; a.N = n_cols; a.K = K; a.block_n = bn; a.cluster = cluster;
        a.ep.out = logits; a.ep.ldo = ldo; a.ep.out_fp32 = 0;
        a.ep.ce_partial = partial; a.ep.ce_target = target; a.ep.ce_label = label; a.ep.ce_tiles = tiles; a.ep.n_valid = vocab;
        gemm_bf16_tn(a, S(stream));
        ce_finalize(partial, tiles, target, label, lse, loss_sum, M, S(stream));
        CUDA_CHECK(cudaStreamSynchronize(S(stream)));
    } catch (...) {
        cudaFree(partial);
        throw;
    }
    cudaFree(partial);
    API_END
}

int eavqa_op_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                           int32_t M, int32_t d, void* stream) {
    API_BEGIN
    layernorm_fwd(x, d, nullptr, gamma, beta, static_cast<bf16*>(y), d, mean, rstd, M, d, 1e-5f, S(stream));
    API_END
}

int eavqa_op_layernorm_bwd(const void* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                           float* dx, int32_t accumulate, void* dx_bf16, float* dgamma, float* dbeta, int32_t M, int32_t d,
                           void* stream) {
    API_BEGIN
    layernorm_bwd(static_cast<const bf16*>(dy), d, x, d, nullptr, gamma, mean, rstd, dx, d, accumulate,
                  static_cast<bf16*>(dx_bf16), d, dgamma, dbeta, M, d, 1e-5f, S(stream));
    API_END
}

int eavqa_op_lm_attention_fwd(const void* qkv, const int32_t* valid, void* o, float* lse, int32_t B, int32_t T, int32_t H,
                              void* stream) {
    API_BEGIN
    lm_attention_fwd(static_cast<const bf16*>(qkv), valid, static_cast<bf16*>(o), lse, B, T, H, S(stream));
    API_END
}

int eavqa_op_lm_attention_bwd(const void* qkv, const int32_t* valid, const void* o, const void* d_o, const float* lse,
                              void* dqkv, float* dq_scratch, int32_t B, int32_t T, int32_t H, void* stream) {
    API_BEGIN
    lm_attention_bwd(static_cast<const bf16*>(qkv), valid, static_cast<const bf16*>(o), static_cast<const bf16*>(d_o), lse,
                     static_cast<bf16*>(dqkv), dq_scratch, B, T, H, S(stream));
    API_END
}

int eavqa_op_mapper_attention_fwd(const void* qkv, void* o, int32_t B, int32_t S_, int32_t H, int32_t hd, void* stream) {
    API_BEGIN
    mapper_attention_fwd(static_cast<const bf16*>(qkv), static_cast<bf16*>(o), B, S_, H, hd, S(stream));
    API_END
}

int eavqa_op_mapper_attention_bwd(const void* qkv, const void* d_o, void* dqkv, int32_t B, int32_t S_, int32_t H, int32_t hd,
                                  void* stream) {
    API_BEGIN
    mapper_attention_bwd(static_cast<const bf16*>(qkv), static_cast<const bf16*>(d_o), static_cast<bf16*>(dqkv), B, S_, H, hd,
                         S(stream));
    API_END
}

int eavqa_op_convert_transpose(const float* src, int32_t R, int32_t C, void* dst, void* dst_t, int32_t ld_t, float* colsum,
                               void* stream) {
    API_BEGIN
    convert_transpose_f32(src, C, R, C, static_cast<bf16*>(dst), C, static_cast<bf16*>(dst_t), ld_t, colsum, S(stream));
    API_END
}

int eavqa_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream) {
    API_BEGIN
    EAVQA_CHECK(params && grads && exp_avg && exp_avg_sq, "null argument");
    adamw_step(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, S(stream));
    API_END
}

int eavqa_sharded_adamw_range(int64_t n, int32_t rank, int32_t world, int64_t* begin, int64_t* end) {
    API_BEGIN
    EAVQA_CHECK(begin && end && world >= 1 && rank >= 0 && rank < world && n >= 0 && n % 4 == 0, "sharded_adamw_range: bad argument");
    sharded_adamw_range(n, rank, world, begin, end);
    API_END
}

int eavqa_sharded_adamw_step(void* const* grad_ptrs, void* const* param_ptrs, const void* mc_grads, void* mc_params,
                             void* const* flag_ptrs, uint32_t token, int32_t rank, int32_t world, float* exp_avg, float* exp_avg_sq,
                             int64_t offset, int64_t n, int32_t max_ctas, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int32_t step, float grad_scale, void* stream) {
    API_BEGIN
    sharded_adamw_step(grad_ptrs, param_ptrs, mc_grads, mc_params, flag_ptrs, token, rank, world, exp_avg, exp_avg_sq, offset, n,
                       max_ctas, lr, beta1, beta2, eps, weight_decay, step, grad_scale, S(stream));
    API_END
}

int eavqa_profile_begin(void) {
    API_BEGIN
    gemm_profile_begin();
    API_END
}

int eavqa_profile_end(double* total_ms, double* total_flops, int64_t* launches, char* report, size_t report_cap) {
    API_BEGIN
    std::string rep;
    gemm_profile_end(total_ms, total_flops, launches, &rep);
    if (report != nullptr && report_cap > 0) {
        std::strncpy(report, rep.c_str(), report_cap - 1);
        report[report_cap - 1] = 0;
    }
    API_END
}

}  // extern "C"
