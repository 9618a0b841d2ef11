// The step engine: owns packed frozen-LM weights, the workspace arena and the KV cache, and composes the
// kernels into the training step (mapper fwd -> LM fwd -> LM-head CE -> LM dgrad -> mapper bwd) and the
// greedy generation loop (prefill + KV-cached decode).  One Engine per process / GPU.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/eavqa_b200.h"
#include "common.cuh"

namespace eavqa {

struct TensorInfo {
    std::string name;
    int64_t offset, rows, cols;    // cols = 1 for vectors
};

// device bump allocator; reset() at the start of every call, grows (re-allocates) on demand
class Arena {
public:
    ~Arena();
    void reset() { off_ = 0; }
    // counting mode: alloc() only advances the offset (returns null); end_measure() restores the real block
    void begin_measure() { saved_base_ = base_; saved_off_ = off_; base_ = nullptr; off_ = 0; }
    size_t end_measure() { size_t n = off_; base_ = saved_base_; off_ = saved_off_; return n; }
    void reserve(size_t bytes, cudaStream_t s);
    void* alloc(size_t bytes);
    template <typename T> T* get(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
    size_t capacity() const { return cap_; }
    const void* base() const { return base_; }
    size_t used() const { return off_; }
private:
    uint8_t* base_ = nullptr;
    uint8_t* saved_base_ = nullptr;
    size_t cap_ = 0, off_ = 0, saved_off_ = 0;
};

struct LmLayer {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    float *b_qkv, *b_o, *b_fc, *b_pr;
    bf16 *w_qkv, *w_qkv_t;     // native Conv1D [d, 3d] (dgrad) and transposed [3d, d] (forward)
    bf16 *w_o, *w_o_t;         // [d, d]
    bf16 *w_fc, *w_fc_t;       // [d, 4d] / [4d, d]
    bf16 *w_pr, *w_pr_t;       // [4d, d] / [d, 4d]
};

class Engine {
public:
    explicit Engine(const eavqa_config& cfg);
    ~Engine();

    void load_lm_weight(const std::string& name, const void* dev_ptr, int dtype, int64_t numel, cudaStream_t s);
    void finalize_lm(cudaStream_t s);

    const std::vector<TensorInfo>& mapper_tensors() const { return mapper_tensors_; }
    // Gradient buckets for overlapping the data-parallel all-reduce with the mapper backward: contiguous ranges of the
    // flat gradient buffer in the order in which the backward finishes them (transformer mapper: pairs of layers, last
    // layers first; MLP mapper: none).  set_grad_events() installs one caller-owned cudaEvent_t per bucket; train_step
    // records event k once every gradient of bucket k is final.
    std::vector<std::pair<int64_t, int64_t>> grad_buckets() const;
    void set_grad_events(void* const* events, int n);
    int64_t mapper_param_count() const { return mapper_count_; }

    // logits_all != null (eavqa_forward_logits): forward only, the tied head runs on every position into
    // logits_all [B * T, ld_logits] fp32 and no loss is computed (labels / loss_out may then be null)
    void train_step(int B, int Tt, const float* clip, const int64_t* tokens, const int64_t* mask, const int64_t* labels,
                    const float* params, float* grads, float* loss_out, cudaStream_t s, float* logits_all = nullptr,
                    int64_t ld_logits = 0);
    void generate(int B, int Tt, int n_images, const float* clip, const int64_t* tokens, const int64_t* mask,
                  int64_t sent_lo, int64_t sent_hi, const float* params, int max_new, int has_eos, int64_t pad_id,
                  int64_t eos_id, int64_t* tokens_out, float* top_logit, float* token_logprob, int32_t* steps_out, cudaStream_t s);

private:
    struct MapperFwd;   // activations of one mapper forward (arena-backed)
    struct MapperW;     // per-step bf16 copies of the trainable weights

    int64_t pofs(const std::string& name) const;
    void pack_mapper_weights(const float* params, bool need_bwd, MapperW& w, cudaStream_t s);
    // returns pointer to prefix rows fp32: [N, P, d] with the given strides
    void mapper_forward(const float* params, const MapperW& w, const float* clip, int N, bool save, MapperFwd& f,
                        cudaStream_t s);
    void mapper_backward(const float* params, const MapperW& w, const MapperFwd& f, const float* dprefix,
                         int64_t dprefix_batch_stride, int N, float* grads, cudaStream_t s);

    eavqa_config cfg_;
    int d_, L_, H_, V_, Vpad_, P_, S_, D_;
    bool finalized_ = false;
    std::map<std::string, bool> loaded_;
    std::vector<void*> owned_;           // cudaMalloc'd weight storage
    // frozen LM
    float *wte_f32_ = nullptr, *wpe_f32_ = nullptr, *lnf_g_ = nullptr, *lnf_b_ = nullptr;
    bf16 *wte_bf16_ = nullptr, *wte_t_bf16_ = nullptr;     // [Vpad, d] and [d, Vpad]
    std::vector<LmLayer> layers_;
    // mapper layout
    std::vector<TensorInfo> mapper_tensors_;
    int64_t mapper_count_ = 0;
    Arena arena_;
    // Weight-gradient GEMMs of the mapper backward run on a side stream: they only feed the gradient buffer, so they
    // overlap the (sub-wave, latency-bound) dgrad chain on the main stream; forked and joined with events per step.
    cudaStream_t side_ = nullptr;
    std::vector<cudaEvent_t> fork_events_;
    size_t fork_used_ = 0;
    cudaEvent_t join_event_ = nullptr;
    std::vector<cudaEvent_t> grad_events_;      // caller-owned, one per grad bucket (empty = not requested)
    cudaEvent_t bucket_sync_event_ = nullptr;
    void bucket_done(int bucket, cudaStream_t main);
    bool side_enabled_ = true;           // EAVQA_WGRAD_STREAM=0 serialises everything on the caller's stream (measurements)
    cudaStream_t fork(cudaStream_t main);   // side stream made to wait for everything enqueued on `main` so far
    void join(cudaStream_t main);           // `main` waits for the side stream
    int32_t* host_flags_ = nullptr;      // pinned: n_unfinished[max] + err flag read-back
    int host_flags_cap_ = 0;
    // The single-token steps of generate() as a CUDA graph (on unless EAVQA_DECODE_GRAPH=0): the ~1500 launches of steps 1..max_new-1
    // are captured once per (shape, arena placement) on an engine-owned stream and replayed with one cudaGraphLaunch.
    struct DecodeGraphKey {
        int B = 0, T0 = 0, max_new = 0, has_eos = 0, want_top = 0, want_lp = 0, prefetch = 0;
        int64_t pad_id = 0, eos_id = 0;
        const void* arena_base = nullptr;
        bool operator==(const DecodeGraphKey& o) const {
            return B == o.B && T0 == o.T0 && max_new == o.max_new && has_eos == o.has_eos && want_top == o.want_top && want_lp == o.want_lp &&
                   pad_id == o.pad_id && eos_id == o.eos_id && arena_base == o.arena_base && prefetch == o.prefetch;
        }
    };
    DecodeGraphKey dec_key_, dec_seen_;
    cudaGraphExec_t dec_graph_ = nullptr;
    int dec_graph_launches_ = 0;         // kernels inside the captured graph (for eavqa_launch_count)
    cudaStream_t dec_stream_ = nullptr;
    // single-token steps: the NEXT layer's KV history is pulled into L2 on a second stream while the current layer's
    // projection GEMMs (latency-bound, HBM nearly idle) run
    cudaStream_t pf_stream_ = nullptr;
    std::vector<cudaEvent_t> pf_events_;
    size_t pf_used_ = 0;
    cudaEvent_t next_pf_event();
};

}  // namespace eavqa
