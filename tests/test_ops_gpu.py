"""GPU: every CUDA kernel through its C-ABI operator entry point against a plain PyTorch fp32 reference of the
same op (computed on the device).  Tolerances are written next to each check; integer / index work is bit-exact."""
import math

import pytest
import torch

from oracle import clip_prefix_lm as orc
from oracle.cases import SPLICE_GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from eavqa_b200 import lib
    return lib.load()


def _check(status):
    from eavqa_b200 import lib
    lib.check(status)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda", dtype)


def _gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def _gelu_new_grad(x):
    x = x.detach().clone().requires_grad_(True)
    _gelu_new(x).sum().backward()
    return x.grad


def gemm(L, A, B, *, out_fp32=False, bias=None, residual=None, act=0, aux=None, dact=0, want_out2=False, block_n=0, ldo=None):
    M, K = A.shape
    N = B.shape[0]
    ldo = ldo or N
    out = torch.zeros(M, ldo, device="cuda", dtype=torch.float32 if out_fp32 else torch.bfloat16)
    out2 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if want_out2 else None
    _check(L.eavqa_op_gemm(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, out.data_ptr(), ldo, int(out_fp32),
                           bias.data_ptr() if bias is not None else None,
                           residual.data_ptr() if residual is not None else None, residual.stride(0) if residual is not None else 0,
                           act, aux.data_ptr() if aux is not None else None, aux.stride(0) if aux is not None else 0, dact,
                           out2.data_ptr() if out2 is not None else None, N if want_out2 else 0, block_n, _stream()))
    torch.cuda.synchronize()
    return out[:, :N], out2


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [(128, 128, 64), (128, 256, 256), (256, 192, 128), (384, 64, 512), (300, 200, 72), (8, 3840, 512),
               (1000, 768, 768), (130, 2304, 768), (2000, 768, 3072), (777, 1000, 1000), (4096, 512, 100)]


@pytest.mark.parametrize("shape", GEMM_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("block_n", [0, 64, 128, 192, 256])
def test_gemm_plain_fp32_out(L, shape, block_n):
    """tcgen05 GEMM, fp32 output: bf16 products accumulate in fp32, so only summation order differs from torch:
    |err| <= 2e-3 * sqrt(K) * scale."""
    M, N, K = shape
    Kp = (K + 7) // 8 * 8          # row strides must be multiples of 8 elements; K itself may be ragged
    A = _rand((M, Kp), 1, dtype=torch.bfloat16)[:, :K]
    B = _rand((N, Kp), 2, dtype=torch.bfloat16)[:, :K]
    out, _ = gemm(L, A, B, out_fp32=True, block_n=block_n)
    ref = A.float() @ B.float().t()
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * math.sqrt(K), f"max err {err} (ref max {ref.abs().max().item()})"


@pytest.mark.parametrize("cluster", [8])   # 8 = CTA pair, tcgen05.mma.cta_group::2 (256 x BN tile)
@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 384, 128), (300, 200, 72), (1000, 768, 768), (130, 2304, 768),
                                   (2000, 768, 3072), (5000, 1111 // 8 * 8, 320)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("block_n", [128, 192, 256])
def test_gemm_cta_pair(L, shape, block_n, cluster):
    """CTA pairs (cluster of 2, one 256 x BN tile per pair): odd tile counts leave the second CTA of the last pair on an
    out-of-range tile that must be computed on zeros and clipped."""
    M, N, K = shape
    A = _rand((M, K), 1, dtype=torch.bfloat16)
    B = _rand((N, K), 2, dtype=torch.bfloat16)
    bias = _rand((N,), 3)
    res = _rand((M, N), 4)
    out, _ = gemm(L, A, B, out_fp32=True, bias=bias, residual=res, block_n=block_n + 1000 * cluster)
    ref = A.float() @ B.float().t() + bias + res
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * math.sqrt(K), f"max err {err}"
    out, out2 = gemm(L, A, B, bias=bias, act=1, want_out2=True, block_n=block_n + 1000 * cluster)
    pre = A.float() @ B.float().t() + bias
    assert (out2.float() - pre).abs().max().item() <= 2e-3 * math.sqrt(K) + 2 ** -8 * pre.abs().max().item()


@pytest.mark.parametrize("shape", [(128, 64, 64), (128, 128, 128), (768, 768, 5120), (2304, 768, 5120), (7680, 512, 256),
                                   (200, 136, 72), (3840, 512, 8), (768, 1536, 100)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("block_n", [0, 64, 128, 192, 256])
def test_gemm_wgrad_mn_major_operands(L, shape, block_n):
    """dW[M,N] = At^T Bt with At [K,M], Bt [K,N] row-major: MN-major UMMA descriptors, no transposed copies."""
    M, N, K = shape
    At = _rand((K, M), 1, dtype=torch.bfloat16)
    Bt = _rand((K, N), 2, dtype=torch.bfloat16)
    out = torch.zeros(M, N, device="cuda")
    _check(L.eavqa_op_gemm_wgrad(At.data_ptr(), M, Bt.data_ptr(), N, M, N, K, out.data_ptr(), N, block_n, _stream()))
    torch.cuda.synchronize()
    ref = At.float().t() @ Bt.float()
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * math.sqrt(K), f"max err {err} (ref max {ref.abs().max().item()})"


@pytest.mark.parametrize("split", [2, 3, 8])
@pytest.mark.parametrize("shape", [(768, 768, 5120), (2304, 768, 5120), (200, 136, 1000), (3840, 512, 520)], ids=lambda s: "x".join(map(str, s)))
def test_gemm_wgrad_split_k_reduce_add(L, shape, split):
    """split-K: each split ADDS its fp32 partial into the output with a TMA reduce (cp.reduce.async.bulk.tensor)."""
    M, N, K = shape
    At = _rand((K, M), 1, dtype=torch.bfloat16)
    Bt = _rand((K, N), 2, dtype=torch.bfloat16)
    init = _rand((M, N), 3)
    out = init.clone()
    _check(L.eavqa_op_gemm_wgrad(At.data_ptr(), M, Bt.data_ptr(), N, M, N, K, out.data_ptr(), N, 128 + 1000 * split, _stream()))
    torch.cuda.synchronize()
    ref = init + At.float().t() @ Bt.float()
    assert (out - ref).abs().max().item() <= 2e-3 * math.sqrt(K)


def test_gemm_identity_exposes_layout(L):
    """B = I: the output must reproduce A exactly (bf16 values are exact in fp32) -- catches any swizzle /
    descriptor / TMEM-lane mix-up as a permutation."""
    for (M, K) in [(128, 64), (256, 256), (200, 192)]:
        A = _rand((M, K), 3, dtype=torch.bfloat16)
        eye = torch.eye(K, device="cuda", dtype=torch.bfloat16)
        for bn in (64, 128, 192, 256):
            out, _ = gemm(L, A, eye, out_fp32=True, block_n=bn)
            assert torch.equal(out, A.float()), f"M={M} K={K} bn={bn}"


def test_gemm_strided_operands_and_output(L):
    M, N, K = 200, 192, 128
    Abig = _rand((M, K + 64), 4, dtype=torch.bfloat16)
    Bbig = _rand((N, K + 32), 5, dtype=torch.bfloat16)
    A, B = Abig[:, :K], Bbig[:, :K]
    out, _ = gemm(L, A, B, out_fp32=True, ldo=N + 8)
    ref = A.float() @ B.float().t()
    assert (out - ref).abs().max().item() <= 2e-3 * math.sqrt(K)


def _gemm_variant(variant):
    def run(L, A, B, **kw):
        return gemm(L, A, B, block_n=variant, **kw)
    return run


# kernel variant = block_n + 1000 * cluster: 0 = the launcher's own heuristic (single CTAs below M = 2048, the
# cta_group::2 pair above for the step's large shapes); 8xxx = the CTA-pair kernel forced with that tile width.
# M >= 2048 shapes are the ones bench.py's step actually runs on the pair kernel (12800 x {768, 2304, 3072} x {768, 3072}).
@pytest.mark.parametrize("variant", [0, 8128, 8192, 8256])
@pytest.mark.parametrize("shape", [(256, 256, 128), (300, 200, 72), (1000, 3072, 768), (2176, 776, 264), (4224, 3072, 768),
                                   (2560, 768, 3072)], ids=lambda s: "x".join(map(str, s)))
def test_gemm_epilogues(L, shape, variant):
    """Every compile-time epilogue (EpiMode) of the single-CTA AND the CTA-pair kernel against fp32 torch."""
    M, N, K = shape
    gemm = _gemm_variant(variant)       # route every call below through the requested kernel variant
    A = _rand((M, K), 1, 0.5, torch.bfloat16)
    B = _rand((N, K), 2, 0.1, torch.bfloat16)
    bias = _rand((N,), 3)
    res = _rand((M, N), 4)
    acc = A.float() @ B.float().t()
    tol = 2e-3 * math.sqrt(K)
    # bias + residual, fp32 out (attention / MLP output projections onto the residual stream)
    out, _ = gemm(L, A, B, out_fp32=True, bias=bias, residual=res)
    assert (out - (acc + bias + res)).abs().max().item() <= tol
    # bias + gelu_new, bf16 out, with the pre-activation as second output (c_fc)
    out, out2 = gemm(L, A, B, bias=bias, act=1, want_out2=True)
    pre = acc + bias
    assert (out2.float() - pre).abs().max().item() <= tol + 2 ** -8 * pre.abs().max().item()
    assert (out.float() - _gelu_new(pre)).abs().max().item() <= tol + 2 ** -8 * pre.abs().max().item()
    # tanh (MLP mapper) and relu (transformer mapper MLP), bf16 out
    out, _ = gemm(L, A, B, bias=bias, act=2)
    assert (out.float() - torch.tanh(pre)).abs().max().item() <= tol + 2 ** -8
    out, _ = gemm(L, A, B, bias=bias, act=3)
    assert (out.float() - torch.relu(pre)).abs().max().item() <= tol + 2 ** -8 * pre.abs().max().item()
    # dgrad through activations: acc * f'(aux), bf16 out (as on the step's path)
    aux = _rand((M, N), 6, 1.0, torch.bfloat16)
    out, _ = gemm(L, A, B, aux=aux, dact=1)
    ref = acc * _gelu_new_grad(aux.float())
    assert (out.float() - ref).abs().max().item() <= tol * 1.2 + (1e-4 + 2 ** -8) * acc.abs().max().item()
    t = torch.tanh(aux.float()).to(torch.bfloat16)
    out, _ = gemm(L, A, B, aux=t, dact=2)
    assert (out.float() - acc * (1 - t.float() ** 2)).abs().max().item() <= tol + 2 ** -8 * acc.abs().max().item()
    r = torch.relu(aux)
    out, _ = gemm(L, A, B, aux=r, dact=3)
    assert (out.float() - acc * (r.float() > 0)).abs().max().item() <= tol + 2 ** -8 * acc.abs().max().item()
    # epilogues without a bias (the bias-carrying modes then read a zero vector)
    out, _ = gemm(L, A, B, out_fp32=True, residual=res)
    assert (out - (acc + res)).abs().max().item() <= tol
    out, _ = gemm(L, A, B, act=3)
    assert (out.float() - torch.relu(acc)).abs().max().item() <= tol + 2 ** -8 * acc.abs().max().item()
    out, _ = gemm(L, A, B, bias=bias)
    assert (out.float() - pre).abs().max().item() <= tol + 2 ** -8 * pre.abs().max().item()
    out, _ = gemm(L, A, B, out_fp32=True, bias=bias)
    assert (out - pre).abs().max().item() <= tol
    # combinations no kernel is compiled for are rejected, not silently approximated
    from eavqa_b200 import lib
    with pytest.raises(lib.EavqaError):
        gemm(L, A, B, out_fp32=True, aux=aux, dact=1)
    with pytest.raises(lib.EavqaError):
        gemm(L, A, B, residual=res)


def test_gemm_rejects_bad_arguments(L):
    from eavqa_b200 import lib
    A = _rand((128, 60), 1, dtype=torch.bfloat16)       # row stride 60: not a multiple of 8
    B = _rand((128, 60), 2, dtype=torch.bfloat16)
    with pytest.raises(lib.EavqaError):
        gemm(L, A, B, out_fp32=True)


@pytest.mark.parametrize("M", [300, 10240], ids=["single_cta_m300", "pair_kernel_bench_head_m10240"])
def test_lmhead_ce(L, M):
    """LM-head GEMM with fused online-softmax statistics vs torch.logsumexp on fp32 logits.
    lse within 2e-3 absolute (bf16 operands, fp32 accumulate), target logit likewise.  M = 10240 is the head of bench.py's
    step (256 captions x 40 targets, 10240 x 50304 x 768): the cta_group::2 pair kernel with the EM_CE epilogue."""
    V, K = 50257, 768
    n_cols = (V + 63) // 64 * 64
    H = _rand((M, K), 1, 1.0, torch.bfloat16)
    W = torch.zeros(n_cols, K, device="cuda", dtype=torch.bfloat16)
    W[:V] = _rand((V, K), 2, 0.02, torch.bfloat16)
    g = torch.Generator().manual_seed(3)
    label = torch.randint(0, V, (M,), generator=g).int()
    label[::7] = -1
    label = label.cuda()
    logits = torch.zeros(M, n_cols, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(M, device="cuda")
    target = torch.zeros(M, device="cuda")
    loss_sum = torch.zeros(1, device="cuda")
    _check(L.eavqa_op_lmhead_ce(H.data_ptr(), W.data_ptr(), M, V, n_cols, K, label.data_ptr(), logits.data_ptr(), n_cols,
                                lse.data_ptr(), target.data_ptr(), loss_sum.data_ptr(), _stream()))
    torch.cuda.synchronize()
    ref = H.float() @ W[:V].float().t()
    ref_lse = torch.logsumexp(ref, dim=-1)
    assert (lse - ref_lse).abs().max().item() < 2e-3
    valid = label >= 0
    ref_t = ref.gather(1, label.clamp_min(0).long().unsqueeze(1)).squeeze(1)
    assert (target[valid] - ref_t[valid]).abs().max().item() < 2e-3
    # the stored logits are fp16 bit patterns in the 2-byte buffer (they only feed d logits; common.cuh: pack_f16x2)
    assert (logits.view(torch.float16)[:, :V].float() - ref).abs().max().item() < 2e-3 + 2 ** -11 * ref.abs().max().item()
    ref_loss = (ref_lse - ref_t)[valid].sum().item()
    assert abs(loss_sum.item() - ref_loss) < 1e-3 * abs(ref_loss)


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("d", [128, 768, 1024, 1600])
def test_layernorm_fwd_bwd(L, d):
    M = 333
    x = _rand((M, d), 1, 2.0) + 0.5
    gamma = 1 + 0.1 * _rand((d,), 2)
    beta = 0.1 * _rand((d,), 3)
    y = torch.zeros(M, d, device="cuda", dtype=torch.bfloat16)
    mean = torch.zeros(M, device="cuda")
    rstd = torch.zeros(M, device="cuda")
    _check(L.eavqa_op_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                    rstd.data_ptr(), M, d, _stream()))
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5)
    torch.cuda.synchronize()
    # bf16 output: half an ulp = 2^-9 relative
    assert (y.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    assert torch.allclose(mean, x.mean(-1), atol=1e-5)
    dy = _rand((M, d), 4, 1.0, torch.bfloat16)
    ref.backward(dy.float())
    dx0 = _rand((M, d), 5)
    dx = dx0.clone()
    dxb = torch.zeros(M, d, device="cuda", dtype=torch.bfloat16)
    dg = torch.zeros(d, device="cuda")
    db = torch.zeros(d, device="cuda")
    _check(L.eavqa_op_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    dx.data_ptr(), 1, dxb.data_ptr(), dg.data_ptr(), db.data_ptr(), M, d, _stream()))
    torch.cuda.synchronize()
    # fp32 math both sides: 1e-4 relative to the row scale
    assert (dx - (dx0 + xr.grad)).abs().max().item() <= 1e-4 * max(1.0, xr.grad.abs().max().item())
    assert (dxb.float() - dx).abs().max().item() <= 2 ** -8 * dx.abs().max().item()
    assert (dg - gr.grad).abs().max().item() <= 1e-3 * gr.grad.abs().max().item()
    assert (db - br.grad).abs().max().item() <= 1e-3 * br.grad.abs().max().item()
    # frozen-LM flavour: no parameter gradients, overwrite instead of accumulate
    dx2 = torch.full((M, d), 7.0, device="cuda")
    _check(L.eavqa_op_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    dx2.data_ptr(), 0, None, None, None, M, d, _stream()))
    torch.cuda.synchronize()
    assert (dx2 - xr.grad).abs().max().item() <= 1e-4 * max(1.0, xr.grad.abs().max().item())


# ------------------------------------------------------------------------------------------------ attention
def _ref_lm_attention(qkv, valid, B, T, H):
    d = H * 64
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    causal = torch.ones(T, T, dtype=torch.bool, device=qkv.device).tril()
    allowed = causal.unsqueeze(0) & valid.bool().unsqueeze(1)
    s = s.masked_fill(~allowed.unsqueeze(1), float("-inf"))
    o = s.softmax(-1) @ v
    return o.transpose(1, 2).reshape(B * T, d), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,T,H", [(3, 50, 12), (2, 64, 2), (2, 130, 4), (1, 257, 2), (4, 17, 3)])
def test_lm_attention_fwd_bwd(L, B, T, H):
    """Causal + key-padding attention vs explicit fp32 softmax.  Operands are bf16 on both sides; P is rounded to
    bf16 before P@V in the kernel: tolerance 2^-7 relative to the value scale."""
    d = H * 64
    qkv = _rand((B * T, 3 * d), 1, 1.0, torch.bfloat16)
    valid = torch.ones(B, T, dtype=torch.int32)
    for b in range(B):      # right padding, prefix-like first positions always valid
        valid[b, max(2, T - 3 * b - (b > 0) * 5):] = 0 if b > 0 else 1
    valid = valid.cuda()
    o = torch.zeros(B * T, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device="cuda")
    _check(L.eavqa_op_lm_attention_fwd(qkv.data_ptr(), valid.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, H, _stream()))
    torch.cuda.synchronize()
    qr = qkv.float().clone().requires_grad_(True)
    ref_o, ref_lse = _ref_lm_attention(qr, valid, B, T, H)
    assert (o.float() - ref_o).abs().max().item() <= 2 ** -7 * max(1.0, ref_o.abs().max().item())
    assert (lse - ref_lse).abs().max().item() <= 1e-3
    d_o = _rand((B * T, d), 2, 1.0, torch.bfloat16)
    ref_o.backward(d_o.float())
    dqkv = torch.zeros(B * T, 3 * d, device="cuda", dtype=torch.bfloat16)
    scratch = torch.zeros(B * T, d, device="cuda")
    _check(L.eavqa_op_lm_attention_bwd(qkv.data_ptr(), valid.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(),
                                       dqkv.data_ptr(), scratch.data_ptr(), B, T, H, _stream()))
    torch.cuda.synchronize()
    ref = qr.grad
    err = (dqkv.float() - ref).abs().max().item()
    assert err <= 2 ** -5 * max(1.0, ref.abs().max().item()), err
    cos = torch.nn.functional.cosine_similarity(dqkv.double().flatten(), ref.double().flatten(), dim=0).item()
    assert cos > 0.9995, cos


@pytest.mark.parametrize("B,S,hd", [(5, 20, 96), (3, 8, 16), (2, 20, 128), (2, 33, 32)])
def test_mapper_attention_fwd_bwd(L, B, S, hd):
    H = 8
    d = H * hd
    qkv = _rand((B * S, 3 * d), 1, 0.5, torch.bfloat16)
    o = torch.zeros(B * S, d, device="cuda", dtype=torch.bfloat16)
    _check(L.eavqa_op_mapper_attention_fwd(qkv.data_ptr(), o.data_ptr(), B, S, H, hd, _stream()))
    torch.cuda.synchronize()
    qr = qkv.float().clone().requires_grad_(True)
    q, k, v = qr.view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(B * S, d)
    assert (o.float() - ref).abs().max().item() <= 2 ** -8 * max(1.0, ref.abs().max().item())
    d_o = _rand((B * S, d), 2, 1.0, torch.bfloat16)
    ref.backward(d_o.float())
    dqkv = torch.zeros(B * S, 3 * d, device="cuda", dtype=torch.bfloat16)
    _check(L.eavqa_op_mapper_attention_bwd(qkv.data_ptr(), d_o.data_ptr(), dqkv.data_ptr(), B, S, H, hd, _stream()))
    torch.cuda.synchronize()
    assert (dqkv.float() - qr.grad).abs().max().item() <= 2 ** -7 * max(1.0, qr.grad.abs().max().item())


# ------------------------------------------------------------------------------------------------ packing
@pytest.mark.parametrize("R,C", [(64, 64), (100, 40), (513, 776), (20, 7680)])
def test_convert_transpose(L, R, C):
    src = _rand((R, C), 1)
    ld_t = (R + 7) // 8 * 8
    dst = torch.zeros(R, C, device="cuda", dtype=torch.bfloat16)
    dst_t = torch.zeros(C, ld_t, device="cuda", dtype=torch.bfloat16)
    colsum = torch.zeros(C, device="cuda")
    _check(L.eavqa_op_convert_transpose(src.data_ptr(), R, C, dst.data_ptr(), dst_t.data_ptr(), ld_t, colsum.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert torch.equal(dst, src.to(torch.bfloat16))                # bit-exact rounding
    assert torch.equal(dst_t[:, :R], src.to(torch.bfloat16).t())
    assert torch.allclose(colsum, src.sum(0), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ splice
def _splice(L, toks, msk, text_table, pre, P, n_img, lo, hi):
    B, Tt = toks.shape
    d = text_table.shape[1]
    T_out = Tt + (P - 1) * n_img
    emb = torch.zeros(B, T_out, d, device="cuda")
    m = torch.zeros(B, T_out, device="cuda", dtype=torch.int32)
    _check(L.eavqa_splice(B, Tt, n_img, P, d, text_table.shape[0], toks.data_ptr(), msk.data_ptr(), lo, hi,
                          text_table.data_ptr(), pre.data_ptr(), emb.data_ptr(), m.data_ptr(), _stream()))
    torch.cuda.synchronize()
    return emb, m


@pytest.mark.parametrize("golden", SPLICE_GOLDEN, ids=[g["name"] for g in SPLICE_GOLDEN])
def test_splice_vct0_golden(L, golden):
    """The reference's own golden tensors (vct0_test.py:79-211), bit-exact.  Text embeddings are supplied as a
    lookup table whose row ``b*T + j`` is the embedding of (b, j): tokens are remapped to those row ids except the
    sentinels, whose rows are never read."""
    toks = torch.tensor(golden["question_tokens"])
    B, Tt = toks.shape
    text = torch.tensor(golden["text_embeddings"])           # [B, Tt, 4-multiple?]
    d = text.shape[-1]
    dpad = 4                                                  # kernels move float4s
    table = torch.zeros(B * Tt, dpad)
    table[:, :d] = text.reshape(B * Tt, d)
    n_img = golden["num_shots"] + 1
    sent = (toks >= 32099 - golden["num_shots"]) & (toks <= 32099)
    ids = torch.arange(B * Tt).view(B, Tt)
    big = 10 ** 6
    remapped = torch.where(sent, big + (32099 - toks), ids)
    pre = torch.tensor(golden["prefix_projections"])          # [B, n_img, P, d]
    P = pre.shape[2]
    prep = torch.zeros(B, n_img * P, dpad)
    prep[..., :d] = pre.reshape(B, n_img * P, d)
    emb, m = _splice(L, remapped.cuda(), torch.tensor(golden["question_masks"]).cuda(), table.cuda(), prep.cuda(), P, n_img,
                     big, big + golden["num_shots"])
    assert torch.equal(emb[..., :d].cpu(), torch.tensor(golden["expected_embeddings"]))
    assert torch.equal(m.cpu().long(), torch.tensor(golden["expected_masks"]))


def test_splice_random_against_oracle(L):
    import eavqa_b200.synthetic as syn
    for trial in range(10):
        k, P, d, V = trial % 5, 1 + trial % 4, 8, 1000
        b = syn.make_fewshot_batch(7, k, 4, V, 990, seed=50 + trial, seg_lo=1, seg_hi=40)
        toks, msk = b["input_ids"], b["attention_mask"]
        g = torch.Generator().manual_seed(trial)
        table = torch.randn(V, d, generator=g)
        pre = torch.randn(7, (k + 1) * P, d, generator=g)
        e_ref, m_ref = orc.insert_prefix_into_input(P, k, toks, table[toks], pre.view(7, k + 1, P, d), msk, 990)
        emb, m = _splice(L, toks.cuda(), msk.cuda(), table.cuda(), pre.cuda(), P, k + 1, 990 - k, 990)
        assert torch.equal(emb.cpu(), e_ref) and torch.equal(m.cpu().long(), m_ref)


def test_splice_rejects_wrong_sentinel_count(L):
    from eavqa_b200 import lib
    toks = torch.tensor([[990, 5, 6, 7], [990, 989, 1, 2]]).cuda()
    msk = torch.ones(2, 4, dtype=torch.int64).cuda()
    with pytest.raises(lib.EavqaError):
        _splice(L, toks, msk, torch.zeros(1000, 4).cuda(), torch.zeros(2, 4, 4).cuda(), 2, 2, 989, 990)


# ------------------------------------------------------------------------------------------------ optimiser
def test_fused_adamw_matches_torch(L):
    """eavqa_adamw_step vs torch.optim.AdamW over 5 steps (fp32 both sides): 1e-6 relative."""
    n = 4096 * 3 + 4
    p0 = _rand((n,), 1)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 6):
        g = _rand((n,), 10 + step)
        ref.grad = g.clone() * 0.5
        opt.step()
        _check(L.eavqa_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 0.5,
                                  _stream()))
    torch.cuda.synchronize()
    assert torch.allclose(p, ref.data, rtol=1e-5, atol=1e-6)
