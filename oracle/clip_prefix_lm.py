"""Plain-PyTorch fp32 CPU restatement of the reference's CLIP-prefix LM step.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Parity: PINNED against the
reference by ``oracle/validate_against_reference.py``.

Every function cites the reference lines it restates.  Paths are relative to
the reference checkout (``/root/reference``); ``HF:`` means the un-vendored
third-party dependency ``transformers`` (pinned ``==4.12.5`` in the reference's
``requirements.txt:77``; 5.5.0 is what is installed and what the validation
script drives) under ``site-packages/transformers/``.

The restatement only uses elementary tensor ops (matmul, softmax, tanh, mean,
var, index) so that it is independent of both ``transformers`` and the
reference's ``nn.Module`` classes.  Weights are plain dicts of tensors using
the HF GPT-2 state-dict names and the reference's ``clip_project.*`` names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
IGNORE_INDEX = -100


# ----------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------
def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim (biased variance).

    Used by the mapper (``clipcap.py:131,135``) and by GPT-2
    (``HF:models/gpt2/modeling_gpt2.py:252,254,505``; eps = layer_norm_epsilon = 1e-5).
    """
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * gamma + beta


def gelu_new(x: Tensor) -> Tensor:
    """tanh-approximated GELU, ``HF:activations.py:59-66`` (NewGELUActivation)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x * x * x)))


# ----------------------------------------------------------------------------
# mapping networks (clipcap.py:31-237)
# ----------------------------------------------------------------------------
def mlp_mapper(p: Dict[str, Tensor], clip: Tensor) -> Tensor:
    """``MLP`` built at ``clipcap.py:256-262``: Linear -> Tanh -> Linear
    (``clipcap.py:35-42``).  nn.Linear stores ``weight[out, in]``.

    clip [B, D] -> [B, P*d]
    """
    h = torch.tanh(clip @ p["model.0.weight"].t() + p["model.0.bias"])
    return h @ p["model.2.weight"].t() + p["model.2.bias"]


def mapper_attention(x: Tensor, wq: Tensor, wkv: Tensor, wp: Tensor, bp: Tensor, num_heads: int) -> Tensor:
    """``MultiHeadAttention.forward`` (``clipcap.py:81-104``) for self-attention
    without a mask; ``to_queries`` / ``to_keys_values`` have no bias because
    ``TransformerLayer`` passes ``bias=False`` (``clipcap.py:125,132-134``), ``project``
    always has one (``clipcap.py:78``)."""
    b, n, c = x.shape
    hd = c // num_heads
    q = (x @ wq.t()).reshape(b, n, num_heads, hd)
    kv = (x @ wkv.t()).reshape(b, n, 2, num_heads, hd)
    k, v = kv[:, :, 0], kv[:, :, 1]
    att = torch.einsum("bnhd,bmhd->bnmh", q, k) * (hd ** -0.5)      # clipcap.py:94
    att = att.softmax(dim=2)                                           # clipcap.py:99
    out = torch.einsum("bnmh,bmhd->bnhd", att, v).reshape(b, n, c)    # clipcap.py:100-102
    return out @ wp.t() + bp                                           # clipcap.py:103


def transformer_mapper(p: Dict[str, Tensor], clip: Tensor, clip_length: int, num_layers: int,
                       num_heads: int = 8) -> Tensor:
    """``TransformerMapper.forward`` (``clipcap.py:213-221``) over ``Transformer``
    (``clipcap.py:141-210``, ``enc_dec=False``, ``mlp_ratio=2.0``, ReLU) of pre-LN
    ``TransformerLayer`` s (``clipcap.py:114-117``).  ``num_heads`` is fixed to 8 at
    ``clipcap.py:233``.

    clip [B, D] -> [B, P, d]   (the last P rows of the [B, clip_length+P, d] sequence)
    """
    B = clip.shape[0]
    x = (clip @ p["linear.weight"].t() + p["linear.bias"]).view(B, clip_length, -1)
    const = p["prefix_const"].unsqueeze(0).expand(B, *p["prefix_const"].shape)
    x = torch.cat((x, const), dim=1)
    for i in range(num_layers):
        L = f"transformer.layers.{i}."
        a = layer_norm(x, p[L + "norm1.weight"], p[L + "norm1.bias"])
        x = x + mapper_attention(a, p[L + "attn.to_queries.weight"], p[L + "attn.to_keys_values.weight"],
                                 p[L + "attn.project.weight"], p[L + "attn.project.bias"], num_heads)
        g = layer_norm(x, p[L + "norm2.weight"], p[L + "norm2.bias"])
        hmid = torch.relu(g @ p[L + "mlp.fc1.weight"].t() + p[L + "mlp.fc1.bias"])   # clipcap.py:61-63
        x = x + hmid @ p[L + "mlp.fc2.weight"].t() + p[L + "mlp.fc2.bias"]           # clipcap.py:65
    return x[:, clip_length:]


def clip_project(p: Dict[str, Tensor], clip: Tensor, mapping_type: str, prefix_length: int,
                 clip_length: int, num_layers: int, d_model: int) -> Tensor:
    """Mapper dispatch of ``ClipCaptionModel.__init__`` (``clipcap.py:254-271``: the
    string ``"mlp"`` selects the MLP, anything else the transformer) followed by
    the ``.view(-1, P, d)`` of ``clipcap.py:318-320``."""
    clip = clip.reshape(-1, clip.shape[-1])
    if mapping_type == "mlp":
        out = mlp_mapper(p, clip)
    else:
        out = transformer_mapper(p, clip, clip_length, num_layers)
    return out.reshape(-1, prefix_length, d_model)


# ----------------------------------------------------------------------------
# GPT-2 (HF:models/gpt2/modeling_gpt2.py)
# ----------------------------------------------------------------------------
def gpt2_hidden(w: Dict[str, Tensor], inputs_embeds: Tensor, attention_mask: Tensor, n_layer: int,
                n_head: int) -> Tensor:
    """``GPT2Model.forward`` (``HF:modeling_gpt2.py:522-636``) on ``inputs_embeds``:
    ``position_ids = arange(T)`` regardless of padding (``:579-582``), ``+wpe`` (``:584-585``),
    causal AND key-padding mask (``:591-597``; a key is masked where ``attention_mask == 0`` --
    the reference passes a *float* mask, ``clipcap.py:303-316``), L ``GPT2Block`` s
    (``:262-309``), ``ln_f`` (``:628``).  Conv1D is ``x @ W + b`` with ``W[in, out]``
    (``HF:pytorch_utils.py:97-123``).  Attention scaling is ``head_dim ** -0.5``
    (``HF:modeling_gpt2.py:96-98``)."""
    B, T, d = inputs_embeds.shape
    hd = d // n_head
    h = inputs_embeds + w["transformer.wpe.weight"][:T].unsqueeze(0)
    causal = torch.ones(T, T, dtype=torch.bool, device=h.device).tril()
    allowed = causal.unsqueeze(0) & (attention_mask != 0).unsqueeze(1)          # [B, Tq, Tk]
    bias = torch.zeros(B, 1, T, T, dtype=h.dtype, device=h.device).masked_fill(~allowed.unsqueeze(1), float("-inf"))
    for i in range(n_layer):
        L = f"transformer.h.{i}."
        u = layer_norm(h, w[L + "ln_1.weight"], w[L + "ln_1.bias"])
        qkv = u @ w[L + "attn.c_attn.weight"] + w[L + "attn.c_attn.bias"]
        q, k, v = qkv.split(d, dim=2)                                            # HF:modeling_gpt2.py:185
        q = q.view(B, T, n_head, hd).transpose(1, 2)
        k = k.view(B, T, n_head, hd).transpose(1, 2)
        v = v.view(B, T, n_head, hd).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * (hd ** -0.5) + bias
        a = s.softmax(dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, T, d)
        h = h + a @ w[L + "attn.c_proj.weight"] + w[L + "attn.c_proj.bias"]
        g = layer_norm(h, w[L + "ln_2.weight"], w[L + "ln_2.bias"])
        m = gelu_new(g @ w[L + "mlp.c_fc.weight"] + w[L + "mlp.c_fc.bias"])      # HF:modeling_gpt2.py:238-243
        h = h + m @ w[L + "mlp.c_proj.weight"] + w[L + "mlp.c_proj.bias"]
    return layer_norm(h, w["transformer.ln_f.weight"], w["transformer.ln_f.bias"])


def gpt2_logits(w: Dict[str, Tensor], hidden: Tensor) -> Tensor:
    """Tied LM head, no bias (``HF:modeling_gpt2.py:646,651,706``)."""
    return hidden @ w["transformer.wte.weight"].t()


def causal_lm_loss(logits: Tensor, labels: Tensor) -> Tensor:
    """``ForCausalLMLoss`` (``HF:loss/loss_utils.py:45-67``): fp32 logits, labels padded
    with -100 and shifted left by one, mean CE over non-ignored targets
    (``fixed_cross_entropy`` ``:28-42`` with ``num_items_in_batch=None``)."""
    logits = logits.float()
    shift = torch.nn.functional.pad(labels, (0, 1), value=IGNORE_INDEX)[..., 1:]
    lse = torch.logsumexp(logits, dim=-1)
    valid = shift != IGNORE_INDEX
    tgt = logits.gather(-1, shift.clamp_min(0).unsqueeze(-1)).squeeze(-1)
    return ((lse - tgt) * valid).sum() / valid.sum()


# ----------------------------------------------------------------------------
# the training forward (clipcap.py:290-342) and the step (autograd backward)
# ----------------------------------------------------------------------------
def caption_forward(lm: Dict[str, Tensor], mapper: Dict[str, Tensor], cfg: dict, question_tokens: Tensor,
                    prefix: Tensor, question_mask: Tensor, labels: Tensor) -> Tuple[Tensor, Tensor]:
    """``ClipCaptionModel.forward`` (``clipcap.py:290-342``).  Returns (loss, logits).

    cfg keys: n_layer, n_head, d_model, prefix_length, clip_length, mapping_type, num_layers.
    """
    P, d = cfg["prefix_length"], cfg["d_model"]
    B = question_tokens.shape[0]
    attention_mask = torch.cat((torch.ones(B, P), question_mask.float()), dim=1)            # :303-316
    embedding_text = lm["transformer.wte.weight"][question_tokens]                             # :317
    pre = clip_project(mapper, prefix, cfg["mapping_type"], P, cfg["clip_length"], cfg["num_layers"], d)
    embedding_cat = torch.cat((pre, embedding_text), dim=1)                                   # :321
    full_labels = torch.cat((torch.full((B, P), IGNORE_INDEX, dtype=torch.int64), labels), dim=1)  # :323-335
    hidden = gpt2_hidden(lm, embedding_cat, attention_mask, cfg["n_layer"], cfg["n_head"])
    logits = gpt2_logits(lm, hidden)
    return causal_lm_loss(logits, full_labels), logits


def train_step(lm: Dict[str, Tensor], mapper: Dict[str, Tensor], cfg: dict, question_tokens: Tensor,
               prefix: Tensor, question_mask: Tensor, labels: Tensor) -> Tuple[float, Dict[str, Tensor]]:
    """Forward + ``loss.backward()`` with only the mapper trainable
    (``ClipCaptionPrefix``, ``clipcap.py:590-599``).  Returns (loss, {name: grad})."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in mapper.items()}
    loss, _ = caption_forward(lm, leaf, cfg, question_tokens, prefix, question_mask, labels)
    loss.backward()
    return float(loss.detach()), {k: v.grad.detach() for k, v in leaf.items()}


# ----------------------------------------------------------------------------
# greedy generation (clipcap.py:344-471)
# ----------------------------------------------------------------------------
@torch.no_grad()
def generate_from_embeddings(lm: Dict[str, Tensor], cfg: dict, embedding_cat: Tensor, attention_mask: Tensor,
                             max_length: int = 10, pad_token_id: Optional[int] = None,
                             eos_token_id: Optional[int] = None, return_margins: bool = False,
                             return_logprobs: bool = False, top_logits_out: Optional[list] = None):
    """``ClipCaptionModel._generate_from_embeddings`` (``clipcap.py:387-471``).

    No KV cache: the whole sequence is re-run each step (``:416-419``); the next
    token is ``argmax(logits[:, -1])`` -- the last *position* even when it is a
    right-pad (``:420-421``, quirk Q2); the raw argmax embedding is appended even
    after EOS (``:423`` precedes ``:431``, quirk Q4); outputs of finished rows are
    ``pad_token_id`` (``:431-434``); a row finishes when its *output* token equals EOS
    (``:458-461``); loop stops when every row is finished (``:463``).  Token
    bookkeeping is done in int64 here; the reference does it in the embedding
    dtype (fp32: exact for ids < 2**24; quirk Q3).
    """
    B = embedding_cat.shape[0]
    if eos_token_id is not None and pad_token_id is None:
        raise ValueError("If `eos_token_id` is defined, make sure that `pad_token_id` is defined.")  # :427-430
    unfinished = torch.ones(B, 1, dtype=torch.int64)
    attention_mask = attention_mask.float()
    tokens = None
    margins = []
    logprobs = []          # log softmax(last)[argmax]: what few_shot_vqa_executor.py:316-323 reads from outputs.scores
    emb = embedding_cat
    for _ in range(max_length):
        hidden = gpt2_hidden(lm, emb, attention_mask, cfg["n_layer"], cfg["n_head"])
        last = gpt2_logits(lm, hidden[:, -1, :])
        nxt = torch.argmax(last, -1).unsqueeze(1)
        if return_margins:
            top2 = last.topk(2, dim=-1).values
            margins.append((top2[:, 0] - top2[:, 1]).clone())
        if top_logits_out is not None:       # the winning logit of every step, for value-level parity of the decode
            top_logits_out.append(last.max(dim=-1).values.clone())
        if return_logprobs:
            logprobs.append(torch.log_softmax(last.double(), dim=-1).gather(1, nxt).squeeze(1).float())
        nxt_embed = lm["transformer.wte.weight"][nxt]
        out = nxt
        if eos_token_id is not None:
            out = nxt * unfinished + pad_token_id * (1 - unfinished)
        tokens = out if tokens is None else torch.cat((tokens, out), dim=1)
        emb = torch.cat((emb, nxt_embed), dim=1)
        attention_mask = torch.cat((attention_mask, torch.ones(B, 1)), dim=-1)
        if eos_token_id is not None:
            unfinished = unfinished * (out != eos_token_id).long()
        if unfinished.max() == 0:
            break
    token_list = tokens.cpu().numpy().astype(int).tolist()
    if return_logprobs:
        return token_list, torch.stack(margins, dim=1) if return_margins else None, torch.stack(logprobs, dim=1)
    if return_margins:
        return token_list, torch.stack(margins, dim=1)
    return token_list


@torch.no_grad()
def generate(lm: Dict[str, Tensor], mapper: Dict[str, Tensor], cfg: dict, question_tokens: Tensor, prefix: Tensor,
             question_mask: Tensor, **generation_kwargs):
    """``ClipCaptionModel.generate`` (``clipcap.py:344-385``): one prefix prepended."""
    P, d = cfg["prefix_length"], cfg["d_model"]
    B = question_tokens.shape[0]
    attention_mask = torch.cat((torch.ones(B, P), question_mask.float()), dim=1)
    embedding_text = lm["transformer.wte.weight"][question_tokens]
    pre = clip_project(mapper, prefix, cfg["mapping_type"], P, cfg["clip_length"], cfg["num_layers"], d)
    embedding_cat = torch.cat((pre, embedding_text), dim=1)
    return generate_from_embeddings(lm, cfg, embedding_cat, attention_mask, **generation_kwargs)


# ----------------------------------------------------------------------------
# in-context prefix splice (vct0.py:494-533) and few-shot generation
# ----------------------------------------------------------------------------
def insert_prefix_into_input(prefix_length: int, num_shots: int, question_tokens: Tensor, text_embeddings: Tensor,
                             prefix_projections: Tensor, question_masks: Tensor,
                             special_token_id: int = 32099) -> Tuple[Tensor, Tensor]:
    """``VCT0Model.insert_prefix_into_input`` (``vct0.py:494-533``), restated as
    explicit index arithmetic instead of boolean-mask scatter.

    Sentinel ``i`` has id ``special_token_id - i`` (``:508-509``).  Each sentinel is
    replaced by ``prefix_length`` prefix rows; a text token at position ``j`` with
    ``c`` sentinels before it lands at ``(j - c) + prefix_length * c`` (``:511-515``);
    prefix rows fill the remaining destinations in order (``:528``); the mask is the
    text mask at text rows and 1 at prefix rows (``:530-531``).  Every row must hold
    exactly ``num_shots + 1`` sentinels (the ``.view`` at ``:512,518`` fails otherwise).

    text_embeddings [B, T, d]; prefix_projections [B, k+1, P, d] -> ([B, T+(P-1)(k+1), d], [B, same] int64)
    """
    B, T = question_tokens.shape
    n_img = num_shots + 1
    P = prefix_length
    d = text_embeddings.shape[-1]
    out_len = T + (P - 1) * n_img
    lo, hi = special_token_id - num_shots, special_token_id
    emb = torch.empty(B, out_len, d, dtype=text_embeddings.dtype)
    msk = torch.empty(B, out_len, dtype=torch.int64)
    pre = prefix_projections.reshape(B, n_img * P, d)
    for b in range(B):
        c = 0
        for j in range(T):
            tok = int(question_tokens[b, j])
            if lo <= tok <= hi:
                if c >= n_img:
                    raise ValueError("row %d holds more than %d sentinels" % (b, n_img))
                dst = (j - c) + P * c
                emb[b, dst:dst + P] = pre[b, c * P:(c + 1) * P]
                msk[b, dst:dst + P] = 1
                c += 1
            else:
                dst = (j - c) + P * c
                emb[b, dst] = text_embeddings[b, j]
                msk[b, dst] = question_masks[b, j]
        if c != n_img:
            raise ValueError("row %d holds %d sentinels, expected %d" % (b, c, n_img))
    return emb, msk


@torch.no_grad()
def generate_few_shot(lm: Dict[str, Tensor], mapper: Dict[str, Tensor], cfg: dict, question_tokens: Tensor,
                      prefix: Tensor, question_mask: Tensor, special_token_id: int, **generation_kwargs):
    """Few-shot VQA generation on the GPT-2 path: the prompt assembly of
    ``VCT0Model.generate`` (``vct0.py:446-464``: ``prefix [B, k+1, 1, D]`` ->
    ``clip_project(...).view(B, -1, P, d)`` -> ``insert_prefix_into_input``) followed by
    the causal-LM greedy loop of ``clipcap.py:387-471``."""
    P, d = cfg["prefix_length"], cfg["d_model"]
    B = question_tokens.shape[0]
    n_img = prefix.shape[1]
    # sentinel rows are dropped by the splice, so clamp their ids for the lookup
    V = lm["transformer.wte.weight"].shape[0]
    embedding_text = lm["transformer.wte.weight"][question_tokens.clamp_max(V - 1)]
    pre = clip_project(mapper, prefix, cfg["mapping_type"], P, cfg["clip_length"], cfg["num_layers"], d)
    pre = pre.view(B, n_img, P, d)
    emb, msk = insert_prefix_into_input(P, n_img - 1, question_tokens, embedding_text, pre, question_mask,
                                        special_token_id)
    return generate_from_embeddings(lm, cfg, emb, msk, **generation_kwargs)
