// Step engine: see engine.cuh.  Reference call structure being replaced:
//   training   clipcap_exector.py:165-171 -> clipcap.py:290-342 -> HF GPT2LMHeadModel.forward -> loss.backward()
//   generation clipcap_exector.py:236-243 -> clipcap.py:344-471 (+ vct0.py:446-464,494-533 for k-shot prompts)
#include "engine.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "gemm.cuh"
#include "kernels.cuh"

namespace eavqa {

// ============================================================================================ arena
Arena::~Arena() {
    if (base_) cudaFree(base_);
}
void Arena::reserve(size_t bytes, cudaStream_t s) {
    if (bytes <= cap_) return;
    if (base_) {
        CUDA_CHECK(cudaStreamSynchronize(s));
        CUDA_CHECK(cudaFree(base_));
        base_ = nullptr;
        cap_ = 0;
    }
    const size_t want = bytes + bytes / 8 + (1u << 20);
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&base_), want));
    cap_ = want;
}
void* Arena::alloc(size_t bytes) {
    const size_t aligned = (bytes + 255) & ~static_cast<size_t>(255);
    void* p = base_ ? static_cast<void*>(base_ + off_) : nullptr;
    off_ += aligned;
    if (base_ != nullptr) EAVQA_CHECK(off_ <= cap_, "workspace arena overflow (planning bug)");
    return p;
}

// ============================================================================================ small helpers
namespace {

void gemm(const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s,
          int bn = 0) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.M = M; a.N = N; a.K = K; a.block_n = bn; a.ep = ep;
    gemm_bf16_tn(a, s);
}
// dW[M, N] (fp32) = At^T Bt, At [K, M], Bt [K, N] bf16 row-major: MN-major UMMA operands, no transposed copies
void gemm_wgrad(const bf16* At, int ldat, const bf16* Bt, int ldbt, int M, int N, int K, float* out, int ldo, cudaStream_t s) {
    GemmArgs a;
    a.A = At; a.lda = ldat; a.B = Bt; a.ldb = ldbt; a.M = M; a.N = N; a.K = K; a.mn_major = 1;
    a.ep.out = out; a.ep.ldo = ldo; a.ep.out_fp32 = 1;
    // weight gradients are few-tile, long-K problems (e.g. 768 x 768 x 5120: 36 tiles for 148 SMs): split K so that
    // ~2 CTAs per SM stream it, each ADDING its fp32 partial into the gradient buffer (zeroed at the top of the step)
    const int bn = N >= 128 ? 128 : 64;       // 128-wide tiles keep the main loop off the shared-memory port limit
    const int tiles = ceil_div(M, 128) * ceil_div(N, bn);
    const int kb = ceil_div(K, 64);
    int split = (2 * num_sms()) / std::max(tiles, 1);
    split = std::max(1, std::min(split, std::min(8, kb / 4)));
    a.block_n = bn;
    a.ep.split_k = split;
    gemm_bf16_tn(a, s);
}
// single-token decode: acc[M, N] (fp32, pre-zeroed) += A[M, K] * W[N, K]^T with K split so that ~all SMs stream weights
// (tools/decode_gemm_bench.py on B200: 1024 x 4096 weights for 128 rows 20.1 -> 6.6 us, 1024 x 1024 8.0 -> 5.4 us)
void gemm_decode(const bf16* A, int lda, const bf16* W, int ldb, int M, int N, int K, float* acc, cudaStream_t s) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.B = W; a.ldb = ldb; a.M = M; a.N = N; a.K = K; a.block_n = 64; a.cluster = 1;
    a.ep.out = acc; a.ep.ldo = N; a.ep.out_fp32 = 1;
    const int tiles = ceil_div(M, 128) * ceil_div(N, 64), kb = ceil_div(K, 64);
    int split = (num_sms() * 7 / 8 + tiles / 2) / std::max(tiles, 1);
    a.ep.split_k = std::max(1, std::min(split, std::max(1, kb / 2)));
    gemm_bf16_tn(a, s);
}
GemmEpilogue ep_bf16(bf16* out, int ldo, const float* bias = nullptr) {
    GemmEpilogue e;
    e.out = out; e.ldo = ldo; e.out_fp32 = 0; e.bias = bias;
    return e;
}
GemmEpilogue ep_f32(float* out, int ldo, const float* bias = nullptr, const float* residual = nullptr, int ld_res = 0) {
    GemmEpilogue e;
    e.out = out; e.ldo = ldo; e.out_fp32 = 1; e.bias = bias; e.residual = residual; e.ld_res = ld_res;
    return e;
}
int pad8(int n) { return (n + 7) & ~7; }

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int64_t n) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[i] = __bfloat162float(src[i]);
}

// dx[n, s, :] = s >= cl ? dprefix[n, s - cl, :] : 0     (gradient entering the mapper's last layer)
__global__ void scatter_prefix_grad_kernel(const float4* __restrict__ dprefix, int64_t batch_stride4, float4* __restrict__ dx,
                                           int N, int S, int cl, int d4) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = static_cast<int64_t>(N) * S * d4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % d4);
        const int64_t row = i / d4;
        const int n = static_cast<int>(row / S), sidx = static_cast<int>(row % S);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sidx >= cl) v = dprefix[n * batch_stride4 + static_cast<int64_t>(sidx - cl) * d4 + c];
        dx[i] = v;
    }
}

__global__ void last_row_index_kernel(int* row_index, int B, int T) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) row_index[b] = b * T + T - 1;
}
__global__ void fill_int_kernel(int* p, int n, int v) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

// ============================================================================================ construction
Engine::Engine(const eavqa_config& cfg) : cfg_(cfg) {
    d_ = cfg.d_model; L_ = cfg.n_layer; H_ = cfg.n_head; V_ = cfg.vocab; P_ = cfg.prefix_length; D_ = cfg.clip_dim;
    EAVQA_CHECK(d_ > 0 && L_ > 0 && H_ > 0 && V_ > 0 && P_ > 0 && D_ > 0 && cfg.n_positions > 0, "config fields must be positive");
    EAVQA_CHECK(d_ == H_ * 64, "GPT-2 head_dim must be 64 (d_model == 64 * n_head)");
    EAVQA_CHECK(d_ <= 2048, "d_model > 2048 is not supported by the LayerNorm kernels");
    EAVQA_CHECK(D_ % 8 == 0, "clip_dim must be a multiple of 8");
    EAVQA_CHECK(cfg.mapper_type == EAVQA_MAPPER_MLP || cfg.mapper_type == EAVQA_MAPPER_TRANSFORMER, "unknown mapper_type");
    int dev = 0, major = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    EAVQA_CHECK(major == 10, "eavqa_b200 needs an sm_100 (Blackwell B200) device; there is no fallback path");
    {
        // one device per process (one process per GPU, as under Lightning DDP / torchrun): kernel attributes, the RICES
        // workspace, the zero-bias vector and the SM count are cached process-wide
        static int first_device = -1;
        if (first_device < 0) first_device = dev;
        EAVQA_CHECK(first_device == dev, "eavqa_b200 is used on one CUDA device per process; this process already runs on another device");
    }
    Vpad_ = static_cast<int>(round_up64(V_, 64));
    S_ = 0;
    auto add = [&](const std::string& name, int64_t rows, int64_t cols) {
        TensorInfo t;
        t.name = name; t.offset = mapper_count_; t.rows = rows; t.cols = cols;
        mapper_tensors_.push_back(t);
        mapper_count_ += rows * cols;
        // keep every tensor 16-byte aligned inside the flat buffer (all sizes here are multiples of 4 floats)
        EAVQA_CHECK((rows * cols) % 4 == 0, "mapper tensor size must be a multiple of 4");
    };
    if (cfg.mapper_type == EAVQA_MAPPER_MLP) {
        const int64_t hdim = (static_cast<int64_t>(d_) * P_) / 2;
        EAVQA_CHECK(hdim % 8 == 0, "MLP mapper hidden width must be a multiple of 8");
        add("model.0.weight", hdim, D_);
        add("model.0.bias", hdim, 1);
        add("model.2.weight", static_cast<int64_t>(d_) * P_, hdim);
        add("model.2.bias", static_cast<int64_t>(d_) * P_, 1);
    } else {
        EAVQA_CHECK(cfg.clip_length > 0 && cfg.mapper_layers > 0, "transformer mapper needs clip_length and num_layers");
        EAVQA_CHECK(d_ % 8 == 0, "transformer mapper: d_model must be divisible by its 8 heads");
        S_ = cfg.clip_length + P_;
        add("prefix_const", P_, d_);
        for (int i = 0; i < cfg.mapper_layers; ++i) {
            const std::string p = "transformer.layers." + std::to_string(i) + ".";
            add(p + "norm1.weight", d_, 1);
            add(p + "norm1.bias", d_, 1);
            add(p + "attn.to_queries.weight", d_, d_);
            add(p + "attn.to_keys_values.weight", 2 * d_, d_);
            add(p + "attn.project.weight", d_, d_);
            add(p + "attn.project.bias", d_, 1);
            add(p + "norm2.weight", d_, 1);
            add(p + "norm2.bias", d_, 1);
            add(p + "mlp.fc1.weight", 2 * d_, d_);
            add(p + "mlp.fc1.bias", 2 * d_, 1);
            add(p + "mlp.fc2.weight", d_, 2 * d_);
            add(p + "mlp.fc2.bias", d_, 1);
        }
        add("linear.weight", static_cast<int64_t>(cfg.clip_length) * d_, D_);
        add("linear.bias", static_cast<int64_t>(cfg.clip_length) * d_, 1);
    }
    layers_.resize(L_);
    for (auto& l : layers_) std::memset(&l, 0, sizeof(LmLayer));
    const char* e = getenv("EAVQA_WGRAD_STREAM");
    side_enabled_ = !(e != nullptr && e[0] == '0');
    {
        // EAVQA_WGRAD_PRIO=high: the side stream of the mapper's weight-gradient GEMMs above the caller's stream (A/B knob)
        const char* pr = getenv("EAVQA_WGRAD_PRIO");
        int least = 0, greatest = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const int prio = (pr != nullptr && pr[0] == 'h') ? greatest : least;
        CUDA_CHECK(cudaStreamCreateWithPriority(&side_, cudaStreamNonBlocking, prio));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&join_event_, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&bucket_sync_event_, cudaEventDisableTiming));
}

std::vector<std::pair<int64_t, int64_t>> Engine::grad_buckets() const {
    std::vector<std::pair<int64_t, int64_t>> out;
    if (cfg_.mapper_type != EAVQA_MAPPER_TRANSFORMER) return out;
    const int n = cfg_.mapper_layers;
    auto layer_begin = [&](int l) { return pofs("transformer.layers." + std::to_string(l) + ".norm1.weight"); };
    const int64_t layers_end = pofs("linear.weight");
    for (int hi = n; hi > 0; hi -= 2) {                     // layers [lo, hi): the backward walks layers downwards
        const int lo = std::max(0, hi - 2);
        out.push_back(std::make_pair(layer_begin(lo), hi == n ? layers_end : layer_begin(hi)));
    }
    return out;
}

void Engine::set_grad_events(void* const* events, int n) {
    grad_events_.clear();
    if (events == nullptr || n == 0) return;
    EAVQA_CHECK(n == static_cast<int>(grad_buckets().size()), "set_grad_events: one event per gradient bucket");
    for (int i = 0; i < n; ++i) {
        EAVQA_CHECK(events[i] != nullptr, "set_grad_events: null event");
        grad_events_.push_back(static_cast<cudaEvent_t>(events[i]));
    }
}

// every gradient of `bucket` has been enqueued (main stream: bias / LayerNorm grads; side stream: weight gradients):
// the caller's event fires when both streams get there
void Engine::bucket_done(int bucket, cudaStream_t main) {
    if (grad_events_.empty()) return;
    if (fork_used_ == 0) {                                  // nothing was forked (profiling / EAVQA_WGRAD_STREAM=0)
        CUDA_CHECK(cudaEventRecord(grad_events_[bucket], main));
        return;
    }
    CUDA_CHECK(cudaEventRecord(bucket_sync_event_, main));
    CUDA_CHECK(cudaStreamWaitEvent(side_, bucket_sync_event_, 0));
    CUDA_CHECK(cudaEventRecord(grad_events_[bucket], side_));
}

cudaEvent_t Engine::next_pf_event() {
    if (pf_used_ == pf_events_.size()) {
        cudaEvent_t ev;
        CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        pf_events_.push_back(ev);
    }
    return pf_events_[pf_used_++];
}

cudaStream_t Engine::fork(cudaStream_t main) {
    if (!side_enabled_ || gemm_profile_active()) return main;
    if (fork_used_ == fork_events_.size()) {
        cudaEvent_t ev;
        CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        fork_events_.push_back(ev);
    }
    cudaEvent_t ev = fork_events_[fork_used_++];
    CUDA_CHECK(cudaEventRecord(ev, main));
    CUDA_CHECK(cudaStreamWaitEvent(side_, ev, 0));
    return side_;
}

void Engine::join(cudaStream_t main) {
    const bool forked = fork_used_ > 0;
    fork_used_ = 0;
    if (!forked) return;
    CUDA_CHECK(cudaEventRecord(join_event_, side_));
    CUDA_CHECK(cudaStreamWaitEvent(main, join_event_, 0));
}

Engine::~Engine() {
    if (side_) cudaStreamSynchronize(side_);
    for (cudaEvent_t ev : fork_events_) cudaEventDestroy(ev);
    if (join_event_) cudaEventDestroy(join_event_);
    if (bucket_sync_event_) cudaEventDestroy(bucket_sync_event_);
    if (side_) cudaStreamDestroy(side_);
    for (void* p : owned_) cudaFree(p);
    if (host_flags_) cudaFreeHost(host_flags_);
    if (dec_graph_) cudaGraphExecDestroy(dec_graph_);
    for (cudaEvent_t ev : pf_events_) cudaEventDestroy(ev);
    if (pf_stream_) cudaStreamDestroy(pf_stream_);
    if (dec_stream_) cudaStreamDestroy(dec_stream_);
}

int64_t Engine::pofs(const std::string& name) const {
    for (const auto& t : mapper_tensors_)
        if (t.name == name) return t.offset;
    throw Error("unknown mapper tensor " + name);
}

// ============================================================================================ LM weights
void Engine::load_lm_weight(const std::string& name, const void* dev_ptr, int dtype, int64_t numel, cudaStream_t s) {
    EAVQA_CHECK(dev_ptr != nullptr, "null weight pointer");
    EAVQA_CHECK(dtype == EAVQA_F32 || dtype == EAVQA_BF16, "weight dtype must be fp32 or bf16");
    auto dmalloc = [&](size_t bytes) {
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, bytes));
        owned_.push_back(p);
        return p;
    };
    const float* src = static_cast<const float*>(dev_ptr);
    if (dtype == EAVQA_BF16) {
        float* tmp = static_cast<float*>(dmalloc(sizeof(float) * numel));
        launch_kernel(bf16_to_f32_kernel, dim3(static_cast<int>(std::min<int64_t>(ceil_div64(numel, 256), 4096))), dim3(256), 0, s, 
            static_cast<const bf16*>(dev_ptr), tmp, numel);
        KERNEL_CHECK();
        src = tmp;
    }
    auto keep_f32 = [&](float*& dst, int64_t n) {
        EAVQA_CHECK(numel == n, "unexpected size for " + name);
        if (!dst) dst = static_cast<float*>(dmalloc(sizeof(float) * n));
        CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    };
    auto keep_matrix = [&](bf16*& nat, bf16*& tr, int rows, int cols) {   // src [rows, cols]
        EAVQA_CHECK(numel == static_cast<int64_t>(rows) * cols, "unexpected size for " + name);
        if (!nat) nat = static_cast<bf16*>(dmalloc(sizeof(bf16) * numel));
        if (!tr) tr = static_cast<bf16*>(dmalloc(sizeof(bf16) * numel));
        convert_transpose_f32(src, cols, rows, cols, nat, cols, tr, rows, nullptr, s);
    };
    const std::string pre = "transformer.";
    if (name == "lm_head.weight" || name.find(".attn.bias") != std::string::npos ||
        name.find(".attn.masked_bias") != std::string::npos)
        return;   // tied head / causal-mask buffers of old HF versions
    EAVQA_CHECK(name.compare(0, pre.size(), pre) == 0, "unexpected weight name " + name);
    const std::string rest = name.substr(pre.size());
    if (rest == "wte.weight") {
        keep_f32(wte_f32_, static_cast<int64_t>(V_) * d_);
        if (!wte_bf16_) wte_bf16_ = static_cast<bf16*>(dmalloc(sizeof(bf16) * static_cast<size_t>(Vpad_) * d_));
        if (!wte_t_bf16_) wte_t_bf16_ = static_cast<bf16*>(dmalloc(sizeof(bf16) * static_cast<size_t>(Vpad_) * d_));
        CUDA_CHECK(cudaMemsetAsync(wte_bf16_, 0, sizeof(bf16) * static_cast<size_t>(Vpad_) * d_, s));
        CUDA_CHECK(cudaMemsetAsync(wte_t_bf16_, 0, sizeof(bf16) * static_cast<size_t>(Vpad_) * d_, s));
        convert_transpose_f32(src, d_, V_, d_, wte_bf16_, d_, wte_t_bf16_, Vpad_, nullptr, s);
    } else if (rest == "wpe.weight") {
        keep_f32(wpe_f32_, static_cast<int64_t>(cfg_.n_positions) * d_);
    } else if (rest == "ln_f.weight") {
        keep_f32(lnf_g_, d_);
    } else if (rest == "ln_f.bias") {
        keep_f32(lnf_b_, d_);
    } else {
        EAVQA_CHECK(rest.compare(0, 2, "h.") == 0, "unexpected weight name " + name);
        const size_t dot = rest.find('.', 2);
        EAVQA_CHECK(dot != std::string::npos, "unexpected weight name " + name);
        const int li = std::stoi(rest.substr(2, dot - 2));
        EAVQA_CHECK(li >= 0 && li < L_, "layer index out of range in " + name);
        LmLayer& l = layers_[li];
        const std::string leaf = rest.substr(dot + 1);
        if (leaf == "ln_1.weight") keep_f32(l.ln1_g, d_);
        else if (leaf == "ln_1.bias") keep_f32(l.ln1_b, d_);
        else if (leaf == "ln_2.weight") keep_f32(l.ln2_g, d_);
        else if (leaf == "ln_2.bias") keep_f32(l.ln2_b, d_);
        else if (leaf == "attn.c_attn.bias") keep_f32(l.b_qkv, 3 * d_);
        else if (leaf == "attn.c_proj.bias") keep_f32(l.b_o, d_);
        else if (leaf == "mlp.c_fc.bias") keep_f32(l.b_fc, 4 * d_);
        else if (leaf == "mlp.c_proj.bias") keep_f32(l.b_pr, d_);
        else if (leaf == "attn.c_attn.weight") keep_matrix(l.w_qkv, l.w_qkv_t, d_, 3 * d_);
        else if (leaf == "attn.c_proj.weight") keep_matrix(l.w_o, l.w_o_t, d_, d_);
        else if (leaf == "mlp.c_fc.weight") keep_matrix(l.w_fc, l.w_fc_t, d_, 4 * d_);
        else if (leaf == "mlp.c_proj.weight") keep_matrix(l.w_pr, l.w_pr_t, 4 * d_, d_);
        else throw Error("unexpected weight name " + name);
    }
    loaded_[name] = true;
    finalized_ = false;
}

void Engine::finalize_lm(cudaStream_t s) {
    EAVQA_CHECK(wte_f32_ && wpe_f32_ && lnf_g_ && lnf_b_, "missing LM weights: wte / wpe / ln_f");
    for (int i = 0; i < L_; ++i) {
        const LmLayer& l = layers_[i];
        EAVQA_CHECK(l.ln1_g && l.ln1_b && l.ln2_g && l.ln2_b && l.b_qkv && l.b_o && l.b_fc && l.b_pr && l.w_qkv && l.w_o &&
                        l.w_fc && l.w_pr,
                    "missing LM weights in layer " + std::to_string(i));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    finalized_ = true;
}

// ============================================================================================ mapper
struct Engine::MapperW {
    // transformer
    std::vector<bf16*> wqkv, wqkv_t, wp, wp_t, w1, w1_t, w2, w2_t;
    bf16* wl = nullptr;
    // mlp
    bf16 *m1 = nullptr, *m2 = nullptr, *m2_t = nullptr;
    void plan(Arena& a, const Engine& e, bool bwd) {
        const int d = e.d_;
        if (e.cfg_.mapper_type == EAVQA_MAPPER_MLP) {
            const size_t hdim = static_cast<size_t>(d) * e.P_ / 2, out = static_cast<size_t>(d) * e.P_;
            m1 = a.get<bf16>(hdim * e.D_);
            m2 = a.get<bf16>(out * hdim);
            m2_t = bwd ? a.get<bf16>(out * hdim) : nullptr;
            return;
        }
        const int n = e.cfg_.mapper_layers;
        wqkv.assign(n, nullptr); wqkv_t.assign(n, nullptr); wp.assign(n, nullptr); wp_t.assign(n, nullptr);
        w1.assign(n, nullptr); w1_t.assign(n, nullptr); w2.assign(n, nullptr); w2_t.assign(n, nullptr);
        const size_t dd = static_cast<size_t>(d) * d;
        for (int i = 0; i < n; ++i) {
            wqkv[i] = a.get<bf16>(3 * dd); wp[i] = a.get<bf16>(dd); w1[i] = a.get<bf16>(2 * dd); w2[i] = a.get<bf16>(2 * dd);
            if (bwd) {
                wqkv_t[i] = a.get<bf16>(3 * dd); wp_t[i] = a.get<bf16>(dd); w1_t[i] = a.get<bf16>(2 * dd); w2_t[i] = a.get<bf16>(2 * dd);
            }
        }
        wl = a.get<bf16>(static_cast<size_t>(e.cfg_.clip_length) * d * e.D_);
    }
};

struct Engine::MapperFwd {
    bf16 *clip_bf16 = nullptr, *clip_t = nullptr;
    // transformer
    std::vector<float*> x;                    // residual stream snapshots [M2, d]
    std::vector<bf16*> a, qkv, o, g, m1;
    std::vector<float*> mean1, rstd1, mean2, rstd2;
    // mlp
    bf16* y1 = nullptr;
    float* y2 = nullptr;
    // result
    const float* prefix = nullptr;
    int64_t prefix_batch_stride = 0;
    void plan(Arena& ar, const Engine& e, int N, bool save) {
        const int d = e.d_;
        clip_bf16 = ar.get<bf16>(static_cast<size_t>(N) * e.D_);
        clip_t = nullptr;      // (wgrad reads clip_bf16 directly as an MN-major operand)
        if (e.cfg_.mapper_type == EAVQA_MAPPER_MLP) {
            y1 = ar.get<bf16>(static_cast<size_t>(N) * d * e.P_ / 2);
            y2 = ar.get<float>(static_cast<size_t>(N) * d * e.P_);
            return;
        }
        const int n = e.cfg_.mapper_layers;
        const size_t M2 = static_cast<size_t>(N) * e.S_;
        const int nx = save ? 2 * n + 1 : 3, na = save ? n : 1;
        x.assign(nx, nullptr);
        for (auto& p : x) p = ar.get<float>(M2 * d);
        a.assign(na, nullptr); qkv.assign(na, nullptr); o.assign(na, nullptr); g.assign(na, nullptr); m1.assign(na, nullptr);
        mean1.assign(na, nullptr); rstd1.assign(na, nullptr); mean2.assign(na, nullptr); rstd2.assign(na, nullptr);
        for (int i = 0; i < na; ++i) {
            a[i] = ar.get<bf16>(M2 * d); qkv[i] = ar.get<bf16>(M2 * 3 * d); o[i] = ar.get<bf16>(M2 * d);
            g[i] = ar.get<bf16>(M2 * d); m1[i] = ar.get<bf16>(M2 * 2 * d);
            mean1[i] = ar.get<float>(M2); rstd1[i] = ar.get<float>(M2); mean2[i] = ar.get<float>(M2); rstd2[i] = ar.get<float>(M2);
        }
    }
};

void Engine::pack_mapper_weights(const float* params, bool bwd, MapperW& w, cudaStream_t s) {
    const int d = d_;
    std::vector<PackJob> jobs;
    auto pack = [&](const std::string& name, int rows, int cols, bf16* nat, bf16* tr) {
        PackJob j;
        j.src = params + pofs(name); j.dst = nat; j.dst_t = tr; j.rows = rows; j.cols = cols; j.tile_begin = 0;
        jobs.push_back(j);
    };
    if (cfg_.mapper_type == EAVQA_MAPPER_MLP) {
        const int hdim = d * P_ / 2;
        pack("model.0.weight", hdim, D_, w.m1, nullptr);
        pack("model.2.weight", d * P_, hdim, w.m2, bwd ? w.m2_t : nullptr);
    } else {
        for (int i = 0; i < cfg_.mapper_layers; ++i) {
            const std::string p = "transformer.layers." + std::to_string(i) + ".";
            // to_queries [d, d] and to_keys_values [2d, d] are adjacent in the flat buffer: one [3d, d] matrix
            pack(p + "attn.to_queries.weight", 3 * d, d, w.wqkv[i], bwd ? w.wqkv_t[i] : nullptr);
            pack(p + "attn.project.weight", d, d, w.wp[i], bwd ? w.wp_t[i] : nullptr);
            pack(p + "mlp.fc1.weight", 2 * d, d, w.w1[i], bwd ? w.w1_t[i] : nullptr);
            pack(p + "mlp.fc2.weight", d, 2 * d, w.w2[i], bwd ? w.w2_t[i] : nullptr);
        }
        pack("linear.weight", cfg_.clip_length * d, D_, w.wl, nullptr);
    }
    pack_batch(jobs, s);      // one launch for the whole mapper
}

void Engine::mapper_forward(const float* params, const MapperW& w, const float* clip, int N, bool save, MapperFwd& f,
                            cudaStream_t s) {
    const int d = d_;
    convert_transpose_f32(clip, D_, N, D_, f.clip_bf16, D_, nullptr, 0, nullptr, s);
    if (cfg_.mapper_type == EAVQA_MAPPER_MLP) {
        // clipcap.py:35-42,256-262: Linear -> Tanh -> Linear
        const int hdim = d * P_ / 2;
        GemmEpilogue e1 = ep_bf16(f.y1, hdim, params + pofs("model.0.bias"));
        e1.act = ACT_TANH;
        gemm(f.clip_bf16, D_, w.m1, D_, N, hdim, D_, e1, s);
        gemm(f.y1, hdim, w.m2, hdim, N, d * P_, hdim, ep_f32(f.y2, d * P_, params + pofs("model.2.bias")), s);
        f.prefix = f.y2;
        f.prefix_batch_stride = static_cast<int64_t>(d) * P_;
        return;
    }
    // clipcap.py:213-221: x = cat(linear(clip).view(N, clip_length, d), prefix_const)
    const int cl = cfg_.clip_length, S = S_, M2 = N * S, n = cfg_.mapper_layers;
    auto xi = [&](int k) { return save ? k : k % 3; };
    auto ai = [&](int l) { return save ? l : 0; };
    gemm(f.clip_bf16, D_, w.wl, D_, N, cl * d, D_, ep_f32(f.x[0], S * d, params + pofs("linear.bias")), s);
    broadcast_rows_f32(params + pofs("prefix_const"), P_, d, f.x[0] + static_cast<size_t>(cl) * d, static_cast<int64_t>(S) * d, N, s);
    for (int l = 0; l < n; ++l) {
        const std::string p = "transformer.layers." + std::to_string(l) + ".";
        const int k = ai(l);
        float *x0 = f.x[xi(2 * l)], *x1 = f.x[xi(2 * l + 1)], *x2 = f.x[xi(2 * l + 2)];
        layernorm_fwd(x0, d, nullptr, params + pofs(p + "norm1.weight"), params + pofs(p + "norm1.bias"), f.a[k], d,
                      f.mean1[k], f.rstd1[k], M2, d, 1e-5f, s);
        gemm(f.a[k], d, w.wqkv[l], d, M2, 3 * d, d, ep_bf16(f.qkv[k], 3 * d), s);
        mapper_attention_fwd(f.qkv[k], f.o[k], N, S, 8, d / 8, s);
        gemm(f.o[k], d, w.wp[l], d, M2, d, d, ep_f32(x1, d, params + pofs(p + "attn.project.bias"), x0, d), s);
        layernorm_fwd(x1, d, nullptr, params + pofs(p + "norm2.weight"), params + pofs(p + "norm2.bias"), f.g[k], d,
                      f.mean2[k], f.rstd2[k], M2, d, 1e-5f, s);
        GemmEpilogue e1 = ep_bf16(f.m1[k], 2 * d, params + pofs(p + "mlp.fc1.bias"));
        e1.act = ACT_RELU;
        gemm(f.g[k], d, w.w1[l], d, M2, 2 * d, d, e1, s);
        gemm(f.m1[k], 2 * d, w.w2[l], 2 * d, M2, d, 2 * d, ep_f32(x2, d, params + pofs(p + "mlp.fc2.bias"), x1, d), s);
    }
    f.prefix = f.x[xi(2 * n)] + static_cast<size_t>(cl) * d;     // last P rows of every sequence (clipcap.py:220)
    f.prefix_batch_stride = static_cast<int64_t>(S) * d;
}

void Engine::mapper_backward(const float* params, const MapperW& w, const MapperFwd& f, const float* dprefix,
                             int64_t dprefix_batch_stride, int N, float* grads, cudaStream_t s) {
    const int d = d_;
    if (cfg_.mapper_type == EAVQA_MAPPER_MLP) {
        const int hdim = d * P_ / 2, out = d * P_;
        bf16* dy2 = arena_.get<bf16>(static_cast<size_t>(N) * out);
        bf16* dy1 = arena_.get<bf16>(static_cast<size_t>(N) * hdim);
        if (grads == nullptr) return;     // planning pass only
        // dY2 = d loss / d prefix, viewed [N, P*d] (rows of dh with the LM's batch stride); bias grads = column sums
        convert_transpose_f32(dprefix, static_cast<int>(dprefix_batch_stride), N, out, dy2, out, nullptr, 0,
                              grads + pofs("model.2.bias"), s);
        gemm_wgrad(dy2, out, f.y1, hdim, out, hdim, N, grads + pofs("model.2.weight"), hdim, fork(s));       // dW2 = dY2^T Y1
        GemmEpilogue e = ep_bf16(dy1, hdim);
        e.dact = DACT_TANH; e.aux = f.y1; e.ld_aux = hdim;
        gemm(dy2, out, w.m2_t, out, N, hdim, out, e, s);                                                     // dY1 = (dY2 W2) * tanh'
        convert_transpose_bf16(dy1, hdim, N, hdim, nullptr, 0, grads + pofs("model.0.bias"), s);
        gemm_wgrad(dy1, hdim, f.clip_bf16, D_, hdim, D_, N, grads + pofs("model.0.weight"), D_, fork(s));    // dW1 = dY1^T clip
        join(s);
        return;
    }
    const int cl = cfg_.clip_length, S = S_, M2 = N * S, n = cfg_.mapper_layers;
    const size_t m2 = static_cast<size_t>(M2);
    float* dx = arena_.get<float>(m2 * d);
    // operands of the weight-gradient GEMMs get their own buffers per layer (55 MB / layer at N = 256): those GEMMs run
    // on the side stream and must not race with the main stream reusing a buffer further down the dgrad chain
    std::vector<bf16*> dx_mlp(n), dx_att(n), dm1(n), dqkv(n);
    for (int l = 0; l < n; ++l) {
        dx_mlp[l] = arena_.get<bf16>(m2 * d); dx_att[l] = arena_.get<bf16>(m2 * d);
        dm1[l] = arena_.get<bf16>(m2 * 2 * d); dqkv[l] = arena_.get<bf16>(m2 * 3 * d);
    }
    bf16* dsmall = arena_.get<bf16>(m2 * d);                              // dg / d_o / da (main stream only)
    bf16* dlin = arena_.get<bf16>(static_cast<size_t>(N) * cl * d);
    if (grads == nullptr) return;         // planning pass only
    {
        const int64_t total = static_cast<int64_t>(M2) * (d / 4);
        launch_kernel(scatter_prefix_grad_kernel, dim3(static_cast<int>(std::min<int64_t>(ceil_div64(total, 256), 148 * 8))), dim3(256), 0, s,
                      reinterpret_cast<const float4*>(dprefix), dprefix_batch_stride / 4, reinterpret_cast<float4*>(dx), N, S, cl, d / 4);
        KERNEL_CHECK();
        count_launch();
    }
    // every weight gradient dW = dY^T X reads dY and X as stored ([rows, features]) through MN-major UMMA descriptors
    for (int l = n - 1; l >= 0; --l) {
        const std::string p = "transformer.layers." + std::to_string(l) + ".";
        // ---- MLP branch: x2 = x1 + fc2(relu(fc1(LN2(x1))))   (clipcap.py:61-67,116)
        convert_transpose_f32(dx, d, M2, d, dx_mlp[l], d, nullptr, 0, grads + pofs(p + "mlp.fc2.bias"), s);
        gemm_wgrad(dx_mlp[l], d, f.m1[l], 2 * d, d, 2 * d, M2, grads + pofs(p + "mlp.fc2.weight"), 2 * d, fork(s));  // dW2 = dx^T m1
        {
            GemmEpilogue e = ep_bf16(dm1[l], 2 * d);
            e.dact = DACT_RELU; e.aux = f.m1[l]; e.ld_aux = 2 * d;
            gemm(dx_mlp[l], d, w.w2_t[l], d, M2, 2 * d, d, e, s);                                          // dm1 = (dx W2) * relu'
        }
        convert_transpose_bf16(dm1[l], 2 * d, M2, 2 * d, nullptr, 0, grads + pofs(p + "mlp.fc1.bias"), s);
        gemm_wgrad(dm1[l], 2 * d, f.g[l], d, 2 * d, d, M2, grads + pofs(p + "mlp.fc1.weight"), d, fork(s));  // dW1 = dm1^T g
        gemm(dm1[l], 2 * d, w.w1_t[l], 2 * d, M2, d, 2 * d, ep_bf16(dsmall, d), s);                       // dg = dm1 W1
        layernorm_bwd(dsmall, d, f.x[2 * l + 1], d, nullptr, params + pofs(p + "norm2.weight"), f.mean2[l], f.rstd2[l], dx, d,
                      1, nullptr, 0, grads + pofs(p + "norm2.weight"), grads + pofs(p + "norm2.bias"), M2, d, 1e-5f, s);
        // ---- attention branch: x1 = x0 + project(attn(LN1(x0)))   (clipcap.py:81-104,115)
        convert_transpose_f32(dx, d, M2, d, dx_att[l], d, nullptr, 0, grads + pofs(p + "attn.project.bias"), s);
        gemm_wgrad(dx_att[l], d, f.o[l], d, d, d, M2, grads + pofs(p + "attn.project.weight"), d, fork(s));   // dWp = dx^T o
        gemm(dx_att[l], d, w.wp_t[l], d, M2, d, d, ep_bf16(dsmall, d), s);                                 // d_o = dx Wp
        mapper_attention_bwd(f.qkv[l], dsmall, dqkv[l], N, S, 8, d / 8, s);
        gemm_wgrad(dqkv[l], 3 * d, f.a[l], d, 3 * d, d, M2, grads + pofs(p + "attn.to_queries.weight"), d, fork(s));   // d[Wq;Wkv] = dqkv^T a
        gemm(dqkv[l], 3 * d, w.wqkv_t[l], 3 * d, M2, d, 3 * d, ep_bf16(dsmall, d), s);                    // da = dqkv [Wq;Wkv]
        layernorm_bwd(dsmall, d, f.x[2 * l], d, nullptr, params + pofs(p + "norm1.weight"), f.mean1[l], f.rstd1[l], dx, d, 1,
                      nullptr, 0, grads + pofs(p + "norm1.weight"), grads + pofs(p + "norm1.bias"), M2, d, 1e-5f, s);
        // layers [l, l+2) (or the odd one at the top) are complete: their all-reduce may start (grad_buckets())
        if ((n - l) % 2 == 0 || l == 0) bucket_done((n - l - 1) / 2, s);
    }
    // x0 = cat(linear(clip).view(N, cl, d), prefix_const)
    sum_over_batch_f32(dx + static_cast<size_t>(cl) * d, static_cast<int64_t>(S) * d, N, P_ * d, grads + pofs("prefix_const"), s);
    convert_transpose_f32(dx, S * d, N, cl * d, dlin, cl * d, nullptr, 0, grads + pofs("linear.bias"), s);
    gemm_wgrad(dlin, cl * d, f.clip_bf16, D_, cl * d, D_, N, grads + pofs("linear.weight"), D_, fork(s));  // dWl = dlin^T clip
    join(s);
}

// ============================================================================================ LM block
struct LmBlockIO {
    int l, M, B, T;
    const float* h_in;
    float *h_mid, *h_out;
    const int* valid;
    bf16 *u, *qkv, *att;
    float *lse, *mean1, *rstd1, *mean2, *rstd2;
    bf16 *fc_pre, *fc_act;
    // generation
    bf16* kv_cache = nullptr;   // this layer's cache (K block [B, H, Tmax, 64] + V block), filled during prefill
    int Tmax = 0;
};

static void lm_block(const LmLayer& w, int d, int H, const LmBlockIO& io, cudaStream_t s) {
    const int M = io.M;
    // HF modeling_gpt2.py:262-309
    layernorm_fwd(io.h_in, d, nullptr, w.ln1_g, w.ln1_b, io.u, d, io.mean1, io.rstd1, M, d, 1e-5f, s);
    gemm(io.u, d, w.w_qkv_t, d, M, 3 * d, d, ep_bf16(io.qkv, 3 * d, w.b_qkv), s);
    lm_attention_fwd(io.qkv, io.valid, io.att, io.lse, io.B, io.T, H, s, io.kv_cache, io.Tmax);      // prefill also fills the KV cache
    gemm(io.att, d, w.w_o_t, d, M, d, d, ep_f32(io.h_mid, d, w.b_o, io.h_in, d), s);
    layernorm_fwd(io.h_mid, d, nullptr, w.ln2_g, w.ln2_b, io.u, d, io.mean2, io.rstd2, M, d, 1e-5f, s);
    GemmEpilogue e = ep_bf16(io.fc_act, 4 * d, w.b_fc);
    e.act = ACT_GELU_NEW;
    e.out2 = io.fc_pre; e.ldo2 = 4 * d;
    gemm(io.u, d, w.w_fc_t, d, M, 4 * d, d, e, s);
    gemm(io.fc_act, 4 * d, w.w_pr_t, 4 * d, M, d, 4 * d, ep_f32(io.h_out, d, w.b_pr, io.h_mid, d), s);
}

// ============================================================================================ training step
void Engine::train_step(int B, int Tt, const float* clip, const int64_t* tokens, const int64_t* mask, const int64_t* labels,
                        const float* params, float* grads, float* loss_out, cudaStream_t s, float* logits_all, int64_t ld_logits) {
    EAVQA_CHECK(finalized_, "LM weights not loaded (call eavqa_finalize_lm)");
    EAVQA_CHECK(B > 0 && Tt > 0, "empty batch");
    EAVQA_CHECK(clip && tokens && params, "null argument");
    EAVQA_CHECK(logits_all != nullptr || (labels && loss_out), "null argument");
    EAVQA_CHECK(logits_all == nullptr || (grads == nullptr && ld_logits >= Vpad_ && ld_logits % 4 == 0 && ld_logits < (1ll << 31)),
                "forward_logits: forward only, row stride >= vocab rounded up to 64");
    const int d = d_, L = L_, T = P_ + Tt, M = B * T, Mh = B * Tt;
    EAVQA_CHECK(T <= cfg_.n_positions, "sequence longer than n_positions");
    const bool bwd = grads != nullptr;
    int head_bn = 0, head_cluster = 1;
    gemm_pick_config(Mh, Vpad_, d, 0, 0, &head_bn, &head_cluster);
    const int head_tiles = 2 * ceil_div(Vpad_, head_bn);   // one (max, sum-exp) pair per half N-tile

    MapperW mw;
    MapperFwd mf;
    int *plan = nullptr, *valid = nullptr, *row_index = nullptr, *label = nullptr, *n_valid = nullptr;
    std::vector<float*> h;
    std::vector<bf16*> qkv, att, fc_pre;
    std::vector<float*> lse, mean1, rstd1, mean2, rstd2;
    bf16 *u = nullptr, *fc_act = nullptr, *hc = nullptr, *logits = nullptr, *dhc = nullptr;
    float *mean_f = nullptr, *rstd_f = nullptr, *target = nullptr, *lse_h = nullptr, *loss_sum = nullptr;
    float2* partial = nullptr;
    float *dh = nullptr, *dq_scratch = nullptr;
    bf16 *dh_b = nullptr, *t4 = nullptr, *t1 = nullptr, *d_att = nullptr, *dqkv = nullptr;

    auto plan_all = [&](Arena& a) {
        const size_t m = static_cast<size_t>(M), mh = static_cast<size_t>(Mh);
        mw.plan(a, *this, bwd);
        mf.plan(a, *this, B, bwd);
        plan = a.get<int>(m); valid = a.get<int>(m);
        const int nh = bwd ? 2 * L + 1 : 3, na = bwd ? L : 1;
        h.assign(nh, nullptr);
        for (auto& p : h) p = a.get<float>(m * d);
        qkv.assign(na, nullptr); att.assign(na, nullptr); fc_pre.assign(na, nullptr); lse.assign(na, nullptr);
        mean1.assign(na, nullptr); rstd1.assign(na, nullptr); mean2.assign(na, nullptr); rstd2.assign(na, nullptr);
        for (int i = 0; i < na; ++i) {
            qkv[i] = a.get<bf16>(m * 3 * d); att[i] = a.get<bf16>(m * d);
            fc_pre[i] = bwd ? a.get<bf16>(m * 4 * d) : nullptr;
            lse[i] = a.get<float>(static_cast<size_t>(B) * H_ * T);
            mean1[i] = a.get<float>(m); rstd1[i] = a.get<float>(m); mean2[i] = a.get<float>(m); rstd2[i] = a.get<float>(m);
        }
        u = a.get<bf16>(m * d); fc_act = a.get<bf16>(m * 4 * d);
        row_index = a.get<int>(mh); label = a.get<int>(mh); n_valid = a.get<int>(1);
        hc = a.get<bf16>(mh * d); mean_f = a.get<float>(mh); rstd_f = a.get<float>(mh);
        partial = a.get<float2>(mh * head_tiles); target = a.get<float>(mh); lse_h = a.get<float>(mh); loss_sum = a.get<float>(1);
        if (bwd) {
            logits = a.get<bf16>(mh * Vpad_); dhc = a.get<bf16>(mh * d);
            dh = a.get<float>(m * d); dh_b = a.get<bf16>(m * d);
            t4 = a.get<bf16>(m * 4 * d); t1 = a.get<bf16>(m * d); d_att = a.get<bf16>(m * d); dqkv = a.get<bf16>(m * 3 * d);
            dq_scratch = T > 64 ? a.get<float>(m * d) : nullptr;
        }
    };
    {   // pass 1: measure (incl. mapper-backward scratch), pass 2: allocate
        arena_.begin_measure();
        plan_all(arena_);
        if (bwd) mapper_backward(params, mw, mf, nullptr, 0, B, nullptr, s);
        const size_t need = arena_.end_measure();
        arena_.reserve(need, s);
        arena_.reset();
        plan_all(arena_);
    }

    // ---- mapper: clip embedding -> P prefix rows (clipcap.py:318-320)
    pack_mapper_weights(params, bwd, mw, s);
    mapper_forward(params, mw, clip, B, bwd, mf, s);

    // ---- inputs_embeds = cat(prefix, wte[tokens]) + wpe (clipcap.py:317-321, HF modeling_gpt2.py:576-585)
    prepend_plan(tokens, mask, B, Tt, P_, plan, valid, s);
    embed_rows(plan, B, T, d, wte_f32_, V_, mf.prefix, mf.prefix_batch_stride, d, wpe_f32_, h[0], s);

    auto hi = [&](int k) { return bwd ? k : k % 3; };
    auto ai = [&](int l) { return bwd ? l : 0; };
    for (int l = 0; l < L; ++l) {
        LmBlockIO io;
        io.l = l; io.M = M; io.B = B; io.T = T;
        io.h_in = h[hi(2 * l)]; io.h_mid = h[hi(2 * l + 1)]; io.h_out = h[hi(2 * l + 2)];
        io.valid = valid; io.u = u; io.qkv = qkv[ai(l)]; io.att = att[ai(l)]; io.lse = lse[ai(l)];
        io.mean1 = mean1[ai(l)]; io.rstd1 = rstd1[ai(l)]; io.mean2 = mean2[ai(l)]; io.rstd2 = rstd2[ai(l)];
        io.fc_pre = fc_pre[ai(l)]; io.fc_act = fc_act;
        lm_block(layers_[l], d, H_, io, s);
    }
    const float* h_last = h[hi(2 * L)];
    if (logits_all != nullptr) {
        // HF modeling_gpt2.py:628,706: ln_f and the tied head on EVERY position, fp32 logits (clipcap.py:337-342 `.logits`)
        layernorm_fwd(h_last, d, nullptr, lnf_g_, lnf_b_, u, d, nullptr, nullptr, M, d, 1e-5f, s);
        gemm(u, d, wte_bf16_, d, M, Vpad_, d, ep_f32(logits_all, static_cast<int>(ld_logits)), s);
        return;
    }

    // ---- tied LM head + shifted cross-entropy on the rows that carry a target
    //      (HF modeling_gpt2.py:703-716, loss_utils.py:28-67); [B*T, V] logits are never materialised in fp32
    ce_plan(labels, B, Tt, T, P_, V_, row_index, label, n_valid, s);
    layernorm_fwd(h_last, d, row_index, lnf_g_, lnf_b_, hc, d, mean_f, rstd_f, Mh, d, 1e-5f, s);
    {
        GemmEpilogue e;
        e.out = logits; e.ldo = Vpad_; e.out_fp32 = 0;
        e.ce_partial = partial; e.ce_target = target; e.ce_label = label; e.ce_tiles = head_tiles; e.n_valid = V_;
        gemm(hc, d, wte_bf16_, d, Mh, Vpad_, d, e, s, head_bn);
    }
    ce_finalize(partial, head_tiles, target, label, lse_h, loss_sum, Mh, s);
    ce_loss(loss_sum, n_valid, loss_out, s);
    if (!bwd) return;

    // ---- backward: d logits -> d hidden (through the tied head) -> ln_f -> L blocks (dgrad only: LM is frozen)
    fill_zero(grads, sizeof(float) * mapper_count_, s);
    ce_dlogits(logits, Vpad_, Mh, V_, Vpad_, lse_h, label, n_valid, s);
    gemm(logits, Vpad_, wte_t_bf16_, Vpad_, Mh, d, Vpad_, ep_bf16(dhc, d), s);
    fill_zero(dh, sizeof(float) * static_cast<size_t>(M) * d, s);
    fill_zero(dh_b, sizeof(bf16) * static_cast<size_t>(M) * d, s);
    layernorm_bwd(dhc, d, h_last, d, row_index, lnf_g_, mean_f, rstd_f, dh, d, 0, dh_b, d, nullptr, nullptr, Mh, d, 1e-5f, s);
    for (int l = L - 1; l >= 0; --l) {
        const LmLayer& w = layers_[l];
        {   // MLP branch
            GemmEpilogue e = ep_bf16(t4, 4 * d);
            e.dact = DACT_GELU_NEW; e.aux = fc_pre[l]; e.ld_aux = 4 * d;
            gemm(dh_b, d, w.w_pr, d, M, 4 * d, d, e, s);                       // d fc_pre = (dh Wpr^T) * gelu'
            gemm(t4, 4 * d, w.w_fc, 4 * d, M, d, 4 * d, ep_bf16(t1, d), s);    // d ln2_out = d fc_pre Wfc^T
            layernorm_bwd(t1, d, h[2 * l + 1], d, nullptr, w.ln2_g, mean2[l], rstd2[l], dh, d, 1, dh_b, d, nullptr, nullptr, M, d, 1e-5f, s);
        }
        {   // attention branch
            gemm(dh_b, d, w.w_o, d, M, d, d, ep_bf16(d_att, d), s);            // d att = dh Wo^T
            lm_attention_bwd(qkv[l], valid, att[l], d_att, lse[l], dqkv, dq_scratch, B, T, H_, s);
            gemm(dqkv, 3 * d, w.w_qkv, 3 * d, M, d, 3 * d, ep_bf16(t1, d), s); // d ln1_out = dqkv Wqkv^T
            layernorm_bwd(t1, d, h[2 * l], d, nullptr, w.ln1_g, mean1[l], rstd1[l], dh, d, 1, l > 0 ? dh_b : nullptr, d, nullptr, nullptr, M, d, 1e-5f, s);
        }
    }
    // ---- d prefix = dh[:, :P] -> mapper backward (dgrad + wgrad): the only trainable parameters
    mapper_backward(params, mw, mf, dh, static_cast<int64_t>(T) * d, B, grads, s);
}

// ============================================================================================ generation
void Engine::generate(int B, int Tt, int n_images, const float* clip, const int64_t* tokens, const int64_t* mask,
                      int64_t sent_lo, int64_t sent_hi, const float* params, int max_new, int has_eos, int64_t pad_id,
                      int64_t eos_id, int64_t* tokens_out, float* top_logit, float* token_logprob, int32_t* steps_out,
                      cudaStream_t s) {
    const auto t_begin = std::chrono::steady_clock::now();
    EAVQA_CHECK(finalized_, "LM weights not loaded (call eavqa_finalize_lm)");
    EAVQA_CHECK(B > 0 && Tt > 0 && max_new > 0, "empty batch / max_length");
    EAVQA_CHECK(clip && tokens && params && tokens_out && steps_out, "null argument");
    EAVQA_CHECK(n_images >= 0, "n_images must be >= 0");
    const int d = d_, L = L_;
    const int n_img = n_images == 0 ? 1 : n_images;
    const int T0 = n_images == 0 ? P_ + Tt : Tt + (P_ - 1) * n_images;     // vct0.py:500
    const int Tmax = T0 + max_new;
    const int M = B * T0, NI = B * n_img;
    EAVQA_CHECK(Tmax - 1 <= cfg_.n_positions, "prompt + max_length exceeds n_positions");
    EAVQA_CHECK(T0 >= 1, "empty prompt");

    MapperW mw;
    MapperFwd mf;
    float *prefix = nullptr, *h0 = nullptr, *h1 = nullptr, *h2 = nullptr, *logits = nullptr, *x_a = nullptr;
    int *plan = nullptr, *valid0 = nullptr, *validD = nullptr, *row_index = nullptr, *unfinished = nullptr, *flags = nullptr;
    bf16 *u = nullptr, *qkv = nullptr, *att = nullptr, *fc_act = nullptr, *hc = nullptr, *kv = nullptr;
    float *mean = nullptr, *rstd = nullptr, *part_val = nullptr;
    float *acc_qkv = nullptr, *acc_o = nullptr, *acc_pr = nullptr;     // split-K accumulators of the decode steps
    int* part_idx = nullptr;
    unsigned* arrivals = nullptr;
    const size_t kv_layer = static_cast<size_t>(B) * Tmax * 2 * d;
    // single-token steps as ONE persistent cooperative kernel per step (decode_chain.cu) instead of ~7 launches per layer.
    // Opt-in (EAVQA_DECODE_CHAIN=1, read per call): measured on B200 (profiles/README.md, round 2) the persistent kernel does
    // NOT beat the launch-per-operation path below -- 28.9 vs 25.3 ms per 128-answer batch -- because a decode step is a chain
    // of dependent memory round trips (TMA load -> MMA -> reduce-add completion -> barrier: 6-10 us per projection whether the
    // boundary is a grid barrier or a kernel launch), and its attention phases keep fewer bytes in flight than five resident
    // CTAs per SM do.  It stays parity-tested as the starting point for overlapping two half-batches.
    const char* chain_opt = getenv("EAVQA_DECODE_CHAIN");
    const bool chain_env = chain_opt != nullptr && chain_opt[0] == '1';
    const bool use_chain = chain_env && max_new > 1 && B <= 128 && d % 64 == 0 && decode_chain_supported(Tmax);
    const int n_chain = 1 + 7 * L;
    // tokens / winning logits / picked-token log-probabilities are produced into engine-owned staging and copied out at the end:
    // the decode graph bakes its pointers, the caller's tensors move from call to call
    int64_t* tok_stage = nullptr;
    float *top_stage = nullptr, *lp_stage = nullptr;
    ChainPhase* chain_dev = nullptr;
    unsigned* chain_bar = nullptr;
    unsigned long long* chain_trace = nullptr;
    const bool trace_on = getenv("EAVQA_CHAIN_TRACE") != nullptr;     // per-phase timing of the last step, printed to stderr
    auto plan_all = [&](Arena& a) {
        const size_t m = static_cast<size_t>(M);
        mw.plan(a, *this, false);
        mf.plan(a, *this, NI, false);
        prefix = a.get<float>(static_cast<size_t>(NI) * P_ * d);
        plan = a.get<int>(m); valid0 = a.get<int>(m); validD = a.get<int>(static_cast<size_t>(B) * Tmax);
        h0 = a.get<float>(m * d); h1 = a.get<float>(m * d); h2 = a.get<float>(m * d);
        u = a.get<bf16>(m * d); qkv = a.get<bf16>(m * 3 * d); att = a.get<bf16>(m * d); fc_act = a.get<bf16>(m * 4 * d);
        mean = a.get<float>(m); rstd = a.get<float>(m);
        row_index = a.get<int>(B); hc = a.get<bf16>(static_cast<size_t>(B) * d);
        logits = a.get<float>(static_cast<size_t>(B) * Vpad_);
        x_a = a.get<float>(static_cast<size_t>(B) * d);          // fp32 residual rows of the single-token steps
        unfinished = a.get<int>(B); flags = a.get<int>(max_new + 1);
        part_val = a.get<float>(static_cast<size_t>(B) * kGreedySplitMax); part_idx = a.get<int>(static_cast<size_t>(B) * kGreedySplitMax);
        arrivals = a.get<unsigned>(B);
        acc_qkv = a.get<float>(static_cast<size_t>(B) * 3 * d); acc_o = a.get<float>(static_cast<size_t>(B) * d);
        acc_pr = a.get<float>(static_cast<size_t>(B) * d);
        kv = a.get<bf16>(kv_layer * L);
        tok_stage = a.get<int64_t>(static_cast<size_t>(B) * max_new);
        top_stage = a.get<float>(static_cast<size_t>(B) * max_new);
        lp_stage = a.get<float>(static_cast<size_t>(B) * max_new);
        chain_dev = a.get<ChainPhase>(n_chain);
        chain_bar = a.get<unsigned>(1);
        chain_trace = a.get<unsigned long long>(9 + 2 * n_chain);
    };
    {
        arena_.begin_measure();
        plan_all(arena_);
        const size_t need = arena_.end_measure();
        arena_.reserve(need, s);
        arena_.reset();
        plan_all(arena_);
    }
    if (host_flags_cap_ < max_new + 1) {
        if (host_flags_) CUDA_CHECK(cudaFreeHost(host_flags_));
        CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&host_flags_), sizeof(int32_t) * (max_new + 1)));
        host_flags_cap_ = max_new + 1;
    }
    int* n_unfinished = flags;
    int* err_flag = flags + max_new;
    fill_zero(flags, sizeof(int) * (max_new + 1), s);
    fill_zero(validD, sizeof(int) * static_cast<size_t>(B) * Tmax, s);
    fill_zero(arrivals, sizeof(unsigned) * static_cast<size_t>(B), s);
    launch_kernel(fill_int_kernel, dim3(ceil_div(B, 256)), dim3(256), 0, s, unfinished, B, 1);
    KERNEL_CHECK();
    count_launch();

    // ---- prefixes for every image (clipcap.py:378-380 / vct0.py:450-452), compacted to [B, n_img*P, d]
    pack_mapper_weights(params, false, mw, s);
    mapper_forward(params, mw, clip, NI, false, mf, s);
    CUDA_CHECK(cudaMemcpy2DAsync(prefix, sizeof(float) * P_ * d, mf.prefix, sizeof(float) * mf.prefix_batch_stride,
                                 sizeof(float) * P_ * d, NI, cudaMemcpyDeviceToDevice, s));

    // ---- prompt assembly: prepend (clipcap.py:381) or sentinel splice (vct0.py:494-533)
    if (n_images == 0) prepend_plan(tokens, mask, B, Tt, P_, plan, valid0, s);
    else splice_plan(tokens, mask, B, Tt, P_, n_images, sent_lo, sent_hi, plan, valid0, err_flag, s);
    embed_rows(plan, B, T0, d, wte_f32_, V_, prefix, static_cast<int64_t>(n_img) * P_ * d, d, wpe_f32_, h0, s);
    CUDA_CHECK(cudaMemcpy2DAsync(validD, sizeof(int) * Tmax, valid0, sizeof(int) * T0, sizeof(int) * T0, B,
                                 cudaMemcpyDeviceToDevice, s));

    // ---- prefill: full causal pass, K/V of every position written to the cache
    float *ha = h0, *hb = h1, *hcur = h2;
    for (int l = 0; l < L; ++l) {
        LmBlockIO io;
        io.l = l; io.M = M; io.B = B; io.T = T0;
        io.h_in = ha; io.h_mid = hb; io.h_out = hcur;
        io.valid = valid0; io.u = u; io.qkv = qkv; io.att = att; io.lse = nullptr;
        io.mean1 = mean; io.rstd1 = rstd; io.mean2 = mean; io.rstd2 = rstd;
        io.fc_pre = nullptr; io.fc_act = fc_act;
        io.kv_cache = kv + kv_layer * l; io.Tmax = Tmax;
        lm_block(layers_[l], d, H_, io, s);
        float* t = ha; ha = hcur; hcur = hb; hb = t;      // output becomes next input
    }
    // logits only at the LAST position of every row, pad or not (clipcap.py:420, quirk Q2)
    launch_kernel(last_row_index_kernel, dim3(ceil_div(B, 256)), dim3(256), 0, s, row_index, B, T0);
    KERNEL_CHECK();
    count_launch();
    float* const top_dst = top_logit ? top_stage : nullptr;
    float* const lp_dst = token_logprob ? lp_stage : nullptr;
    auto head_and_pick = [&](int step, cudaStream_t st) {   // hc = ln_f(last hidden) -> logits -> greedy token, next input embedding in x_a
        gemm(hc, d, wte_bf16_, d, B, Vpad_, d, ep_f32(logits, Vpad_), st);
        greedy_step(logits, Vpad_, B, V_, step, max_new, has_eos, pad_id, eos_id, unfinished, tok_stage, n_unfinished,
                    top_dst, lp_dst, wte_f32_, wpe_f32_ + static_cast<size_t>(std::min(T0 + step, cfg_.n_positions - 1)) * d, d, x_a,
                    validD + T0 + step, Tmax, part_val, part_idx, arrivals, st);
    };
    layernorm_fwd(ha, d, row_index, lnf_g_, lnf_b_, hc, d, nullptr, nullptr, B, d, 1e-5f, s);
    head_and_pick(0, s);

    // ---- decode: one token per row per step against the KV cache (the reference re-runs the whole sequence each
    //      step, clipcap.py:416-419; same function, 11x less work).  M = B rows per GEMM: every projection is split along K
    //      over all SMs into fp32 accumulators; bias / residual / LayerNorm / gelu live in the small kernels between them,
    //      each of which also zeroes the accumulator of the GEMM that follows.  x_a is the fp32 residual row.
    if (use_chain) {
        // the same sequence of operations as the loop below, as phases of the persistent kernel (identical for every step
        // except `pos`, which travels as a kernel argument)
        std::vector<ChainPhase> ph(n_chain);
        int k = 0;
        auto split_for = [&](int N, int K) { return std::max(1, std::min(num_sms() / (N / 64), ceil_div(K, 64))); };
        chain_glue_phase(ph[k++], x_a, nullptr, nullptr, layers_[0].ln1_g, layers_[0].ln1_b, u, B, d, acc_qkv, 3 * d);
        for (int l = 0; l < L; ++l) {
            const LmLayer& w = layers_[l];
            chain_gemm_phase(ph[k++], u, d, w.w_qkv_t, d, B, 3 * d, d, split_for(3 * d, d), CHAIN_REDUCE_F32, acc_qkv, 3 * d, nullptr);
            chain_attn_phase(ph[k++], acc_qkv, w.b_qkv, kv + kv_layer * l, validD, Tmax, att, acc_o, B, H_, Tmax);
            chain_gemm_phase(ph[k++], att, d, w.w_o_t, d, B, d, d, split_for(d, d), CHAIN_REDUCE_F32, acc_o, d, nullptr);
            chain_glue_phase(ph[k++], x_a, acc_o, w.b_o, w.ln2_g, w.ln2_b, u, B, d, acc_pr, d);
            chain_gemm_phase(ph[k++], u, d, w.w_fc_t, d, B, 4 * d, d, 1, CHAIN_GELU_BF16, fc_act, 4 * d, w.b_fc);
            chain_gemm_phase(ph[k++], fc_act, 4 * d, w.w_pr_t, 4 * d, B, d, 4 * d, split_for(d, 4 * d), CHAIN_REDUCE_F32, acc_pr, d, nullptr);
            if (l + 1 < L)
                chain_glue_phase(ph[k++], x_a, acc_pr, w.b_pr, layers_[l + 1].ln1_g, layers_[l + 1].ln1_b, u, B, d, acc_qkv, 3 * d);
            else
                chain_glue_phase(ph[k++], x_a, acc_pr, w.b_pr, lnf_g_, lnf_b_, hc, B, d, nullptr, 0);
        }
        EAVQA_CHECK(k == n_chain, "decode chain phase count");
        CUDA_CHECK(cudaMemcpyAsync(chain_dev, ph.data(), sizeof(ChainPhase) * n_chain, cudaMemcpyHostToDevice, s));
        fill_zero(chain_bar, sizeof(unsigned), s);
        unsigned epoch = 0;
        for (int step = 1; step < max_new; ++step) {
            launch_decode_chain(chain_dev, n_chain, T0 + step - 1, chain_bar, epoch, s, trace_on ? chain_trace : nullptr);
            epoch += static_cast<unsigned>(n_chain);
            // the head (B x Vpad x d, 100 MB of weights) runs on the wide-tile stand-alone GEMM: at 64-column tiles the chain
            // would re-read the 128 activation rows once per tile (measured 36 us against ~20)
            gemm(hc, d, wte_bf16_, d, B, Vpad_, d, ep_f32(logits, Vpad_), s);
            greedy_step(logits, Vpad_, B, V_, step, max_new, has_eos, pad_id, eos_id, unfinished, tok_stage, n_unfinished, top_dst,
                        lp_dst, wte_f32_, wpe_f32_ + static_cast<size_t>(std::min(T0 + step, cfg_.n_positions - 1)) * d, d, x_a,
                        validD + T0 + step, Tmax, part_val, part_idx, arrivals, s);
        }
    }
    // KV prefetch, opt-in (EAVQA_KV_PREFETCH=1): after layer l's attention a second stream pulls layer l + 1's KV history
    // (2 x B x H x Tmax x 128 B, one contiguous region) into L2 with cp.async.bulk.prefetch.L2 while the four projection GEMMs
    // and the glue of layer l run, so that the next attention would stream from L2 instead of HBM.  Measured on B200
    // (configs[3], parity-green): 24.2 - 24.5 ms per 128-answer batch against 23.3 - 23.4 without -- the projection GEMMs are
    // chains of dependent memory round trips, and 80 MB of prefetch traffic per layer lengthens every one of them by more
    // than the attention gains.  Off by default.
    static const bool kv_prefetch = [] { const char* e = getenv("EAVQA_KV_PREFETCH"); return e != nullptr && e[0] == '1'; }();
    const bool use_prefetch = kv_prefetch && kv_layer * sizeof(bf16) <= (96u << 20);      // one layer's history must fit in L2
    auto decode_loop = [&](cudaStream_t st) {
        if (use_prefetch && pf_stream_ == nullptr) CUDA_CHECK(cudaStreamCreateWithFlags(&pf_stream_, cudaStreamNonBlocking));
        pf_used_ = 0;
        cudaEvent_t pf_done = nullptr;                         // the prefetch of the layer about to run has been issued
        for (int step = 1; step < max_new; ++step) {
            const int pos = T0 + step - 1;
            decode_residual_ln(x_a, nullptr, nullptr, layers_[0].ln1_g, layers_[0].ln1_b, u, B, d, 1e-5f, acc_qkv, 3 * d, st);
            for (int l = 0; l < L; ++l) {
                const LmLayer& w = layers_[l];
                gemm_decode(u, d, w.w_qkv_t, d, B, 3 * d, d, acc_qkv, st);
                if (pf_done != nullptr) {
                    CUDA_CHECK(cudaStreamWaitEvent(st, pf_done, 0));
                    pf_done = nullptr;
                }
                lm_attention_decode_acc(acc_qkv, w.b_qkv, kv + kv_layer * l, validD, Tmax, att, acc_o, B, H_, pos, Tmax, st);
                if (use_prefetch && !(step == max_new - 1 && l == L - 1)) {
                    cudaEvent_t att_done = next_pf_event();
                    CUDA_CHECK(cudaEventRecord(att_done, st));
                    CUDA_CHECK(cudaStreamWaitEvent(pf_stream_, att_done, 0));
                    prefetch_l2(kv + kv_layer * ((l + 1) % L), kv_layer * sizeof(bf16), pf_stream_);
                    pf_done = next_pf_event();
                    CUDA_CHECK(cudaEventRecord(pf_done, pf_stream_));
                }
                gemm_decode(att, d, w.w_o_t, d, B, d, d, acc_o, st);
                decode_residual_ln(x_a, acc_o, w.b_o, w.ln2_g, w.ln2_b, u, B, d, 1e-5f, acc_pr, d, st);
                {   // c_fc has 4d / 64 = 64+ tiles of its own: unsplit with the fused bias + gelu epilogue beats split-K plus a
                    // separate gelu kernel (8.3 us vs 6.6 + 6.2 us per layer)
                    GemmEpilogue e = ep_bf16(fc_act, 4 * d, w.b_fc);
                    e.act = ACT_GELU_NEW;
                    gemm(u, d, w.w_fc_t, d, B, 4 * d, d, e, st, 64);
                }
                gemm_decode(fc_act, 4 * d, w.w_pr_t, 4 * d, B, d, 4 * d, acc_pr, st);
                if (l + 1 < L)
                    decode_residual_ln(x_a, acc_pr, w.b_pr, layers_[l + 1].ln1_g, layers_[l + 1].ln1_b, u, B, d, 1e-5f, acc_qkv, 3 * d, st);
                else
                    decode_residual_ln(x_a, acc_pr, w.b_pr, lnf_g_, lnf_b_, hc, B, d, 1e-5f, nullptr, 0, st);
            }
            head_and_pick(step, st);
        }
        if (pf_done != nullptr) CUDA_CHECK(cudaStreamWaitEvent(st, pf_done, 0));      // the second stream rejoins
    };
    if (!use_chain && max_new > 1) {
        // The decode loop replayed as ONE CUDA graph (EAVQA_DECODE_GRAPH=0 turns it off; measured 24.46 -> 24.08 ms per
        // 128-answer batch: the steps are bound by their chain of memory round trips, not by launch cost, so the gain is the
        // ~1.5 % of host / front-end overhead).  The first call with a new (shape, arena placement)
        // runs the launches directly (kernel attributes, tensor maps, workspace growth all settle); the second captures them
        // on the engine's own stream (the caller's may be the legacy default stream, which cannot be captured); later calls
        // only launch the instantiated graph.  Programmatic-dependent-launch attributes are kept as programmatic edges.
        const char* gopt = getenv("EAVQA_DECODE_GRAPH");
        const bool want_graph = !(gopt != nullptr && gopt[0] == '0') && !gemm_profile_active();
        DecodeGraphKey key;
        key.B = B; key.T0 = T0; key.max_new = max_new; key.has_eos = has_eos; key.want_top = top_logit != nullptr;
        key.want_lp = token_logprob != nullptr; key.pad_id = pad_id; key.eos_id = eos_id; key.arena_base = arena_.base();
        key.prefetch = use_prefetch ? 1 : 0;
        if (!want_graph || !(key == dec_seen_)) {
            dec_seen_ = key;
            decode_loop(s);
        } else {
            if (dec_graph_ == nullptr || !(key == dec_key_)) {
                if (dec_graph_ != nullptr) {
                    CUDA_CHECK(cudaGraphExecDestroy(dec_graph_));
                    dec_graph_ = nullptr;
                }
                if (dec_stream_ == nullptr) CUDA_CHECK(cudaStreamCreateWithFlags(&dec_stream_, cudaStreamNonBlocking));
                const int64_t before = kernel_launch_count() + gemm_launch_count();
                cudaGraph_t graph = nullptr;
                CUDA_CHECK(cudaStreamBeginCapture(dec_stream_, cudaStreamCaptureModeThreadLocal));
                try {
                    decode_loop(dec_stream_);
                } catch (...) {
                    cudaStreamEndCapture(dec_stream_, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    throw;
                }
                CUDA_CHECK(cudaStreamEndCapture(dec_stream_, &graph));
                CUDA_CHECK(cudaGraphInstantiate(&dec_graph_, graph, 0));
                CUDA_CHECK(cudaGraphDestroy(graph));
                dec_graph_launches_ = static_cast<int>(kernel_launch_count() + gemm_launch_count() - before);
                dec_key_ = key;
            } else {
                count_launch(dec_graph_launches_);
            }
            CUDA_CHECK(cudaGraphLaunch(dec_graph_, s));
        }
    }
    // results leave the staging buffers (the caller's tensors are not baked into anything)
    CUDA_CHECK(cudaMemcpyAsync(tokens_out, tok_stage, sizeof(int64_t) * static_cast<size_t>(B) * max_new, cudaMemcpyDeviceToDevice, s));
    if (top_logit) CUDA_CHECK(cudaMemcpyAsync(top_logit, top_stage, sizeof(float) * static_cast<size_t>(B) * max_new, cudaMemcpyDeviceToDevice, s));
    if (token_logprob)
        CUDA_CHECK(cudaMemcpyAsync(token_logprob, lp_stage, sizeof(float) * static_cast<size_t>(B) * max_new, cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(host_flags_, flags, sizeof(int) * (max_new + 1), cudaMemcpyDeviceToHost, s));
    static const bool timing = getenv("EAVQA_TIMING") != nullptr;
    const auto t_enq = std::chrono::steady_clock::now();
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (timing) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "[eavqa] generate: host enqueue %.3f ms, then waited %.3f ms for the GPU\n",
                std::chrono::duration<double, std::milli>(t_enq - t_begin).count(),
                std::chrono::duration<double, std::milli>(t_end - t_enq).count());
    }
    if (use_chain && trace_on) {
        std::vector<unsigned long long> t(9 + 2 * n_chain);
        CUDA_CHECK(cudaMemcpy(t.data(), chain_trace, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
        double work[3] = {0, 0, 0}, wait[3] = {0, 0, 0};
        int cnt[3] = {0, 0, 0};
        std::vector<ChainPhase> ph(n_chain);
        CUDA_CHECK(cudaMemcpy(ph.data(), chain_dev, sizeof(ChainPhase) * n_chain, cudaMemcpyDeviceToHost));
        double by_shape[8] = {0};
        for (int i = 0; i < n_chain; ++i) {
            const unsigned long long begin = i == 0 ? t[0] : t[2 * i], own = t[1 + 2 * i], all = t[2 + 2 * i];
            work[ph[i].type] += static_cast<double>(own - begin) * 1e-3;
            wait[ph[i].type] += static_cast<double>(all - own) * 1e-3;
            cnt[ph[i].type]++;
            if (ph[i].type == CHAIN_GEMM) by_shape[(i - 1) % 7] += static_cast<double>(all - begin) * 1e-3;
        }
        fprintf(stderr, "[eavqa] decode chain, last step: %.1f us total; per phase (CTA 0 work + wait at barrier, us): gemm %.2f + %.2f (x%d), "
                        "attention %.2f + %.2f (x%d), glue %.2f + %.2f (x%d); gemm phases per layer: qkv %.2f o %.2f fc %.2f pr %.2f\n",
                static_cast<double>(t[2 * n_chain] - t[0]) * 1e-3, work[0] / cnt[0], wait[0] / cnt[0], cnt[0], work[1] / cnt[1], wait[1] / cnt[1],
                cnt[1], work[2] / cnt[2], wait[2] / cnt[2], cnt[2], by_shape[0] / L, by_shape[2] / L, by_shape[4] / L, by_shape[5] / L);
    }
    EAVQA_CHECK(host_flags_[max_new] == 0,
                "prompt rows must each hold exactly n_images sentinel tokens (vct0.py:512 would fail its .view)");
    int steps = max_new;
    if (has_eos)
        for (int i = 0; i < max_new; ++i)
            if (host_flags_[i] == 0) {
                steps = i + 1;       // every row finished at step i: the reference breaks here (clipcap.py:463)
                break;
            }
    *steps_out = steps;
}

}  // namespace eavqa
