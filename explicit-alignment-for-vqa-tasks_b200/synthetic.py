"""Seeded synthetic weights and batches for the CLIP-prefix LM step.

There is no network on the build or GPU boxes, so neither pretrained GPT-2
weights nor Conceptual Captions / VQA2 data exist; every measurement and parity
run uses the generators below.  They use the CPU RNG
only, so the same seed yields the same tensors on every box.

Names follow the HF GPT-2 state dict (``transformer.h.{i}.attn.c_attn.weight`` ...)
and the reference's mapper modules (``clipcap.py:31-237``).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch

GPT2_CONFIGS = {
    # name: n_layer, n_head, d_model (head_dim is 64 for every released GPT-2)
    "gpt2": dict(n_layer=12, n_head=12, d_model=768),
    "gpt2-medium": dict(n_layer=24, n_head=16, d_model=1024),
    "gpt2-large": dict(n_layer=36, n_head=20, d_model=1280),
    "gpt2-xl": dict(n_layer=48, n_head=25, d_model=1600),
    # test-only shapes (not released models)
    "gpt2-tiny": dict(n_layer=2, n_head=2, d_model=128),
    "gpt2-mini": dict(n_layer=3, n_head=4, d_model=256),
}
GPT2_VOCAB = 50257
GPT2_POSITIONS = 1024


def lm_config(model_version: str, vocab: Optional[int] = None, n_positions: Optional[int] = None) -> dict:
    if model_version not in GPT2_CONFIGS:
        raise ValueError("unknown GPT-2 model_version %r (known: %s)" % (model_version, sorted(GPT2_CONFIGS)))
    cfg = dict(GPT2_CONFIGS[model_version])
    small = model_version in ("gpt2-tiny", "gpt2-mini")
    cfg["vocab"] = vocab if vocab is not None else (1000 if small else GPT2_VOCAB)
    cfg["n_positions"] = n_positions if n_positions is not None else (256 if small else GPT2_POSITIONS)
    return cfg


def make_lm_weights(cfg: dict, seed: int = 0, hot_rows: int = 0, successor: Optional[dict] = None) -> "OrderedDict[str, torch.Tensor]":
    """HF-default-like init (every matrix ~ N(0, 0.02^2)) with *non-trivial* biases and
    LayerNorm affines (HF's are 0 / 1 / 0, which would hide bias / gamma / beta bugs).

    ``hot_rows > 0`` multiplies that many seeded ``wte`` rows by 8 ("sharpened" LM, SURVEY.md
    section 7.3) so that greedy decoding has non-degenerate top-2 margins.  With tied embeddings a hot token mostly
    predicts itself, so those decodes repeat one token.

    ``successor`` (a dict, see ``plant_successor_table``) instead plants a bigram table in the first block's MLP:
    every hot token robustly predicts ANOTHER hot token (one seeded cycle through the hot set), so greedy answers
    are chains of distinct tokens whose top-2 margins sit a few logit standard deviations above bf16 noise.
    """
    g = torch.Generator().manual_seed(seed)
    d, L, V, NP = cfg["d_model"], cfg["n_layer"], cfg["vocab"], cfg["n_positions"]

    def mat(*shape):
        return torch.randn(*shape, generator=g) * 0.02

    def ln(prefix, out):
        out[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        out[prefix + ".bias"] = 0.1 * torch.randn(d, generator=g)

    w: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    w["transformer.wte.weight"] = mat(V, d)
    w["transformer.wpe.weight"] = mat(NP, d)
    for i in range(L):
        p = "transformer.h.%d." % i
        ln(p + "ln_1", w)
        w[p + "attn.c_attn.weight"] = mat(d, 3 * d)
        w[p + "attn.c_attn.bias"] = mat(3 * d)
        w[p + "attn.c_proj.weight"] = mat(d, d)
        w[p + "attn.c_proj.bias"] = mat(d)
        ln(p + "ln_2", w)
        w[p + "mlp.c_fc.weight"] = mat(d, 4 * d)
        w[p + "mlp.c_fc.bias"] = mat(4 * d)
        w[p + "mlp.c_proj.weight"] = mat(4 * d, d)
        w[p + "mlp.c_proj.bias"] = mat(d)
    ln("transformer.ln_f", w)
    if successor is not None:
        plant_successor_table(w, cfg, **successor)
    elif hot_rows:
        gh = torch.Generator().manual_seed(7)
        idx = torch.randperm(V, generator=gh)[:hot_rows]
        w["transformer.wte.weight"][idx] *= 8.0
    return w


def successor_hot_ids(n_hot: int, text_vocab: int) -> torch.Tensor:
    """The seeded hot-token set of the successor-planted LM: ``n_hot`` ids below ``text_vocab`` (so that neither the
    pad / eos id nor an added sentinel is ever hot).  Token ``ids[i]`` is followed by ``ids[(i + 1) % n_hot]``."""
    gh = torch.Generator().manual_seed(7)
    return torch.randperm(text_vocab, generator=gh)[:n_hot]


def plant_successor_table(w: "OrderedDict[str, torch.Tensor]", cfg: dict, n_hot: int, text_vocab: int, hot_scale: float,
                          alpha: float, theta: float, kappa: float, z_ref: float) -> None:
    """Make greedy decoding of the synthetic LM diverse AND numerically well-conditioned (in place).

    A random-init GPT-2 has near-Gaussian logits over 50 k tokens: the top-2 margin is a fraction of the logit
    standard deviation, so fp32 and bf16 decodes of the very same code disagree on ~30 % of 10-token answers
     -- and scaling a few ``wte`` rows only makes every hot token predict itself.  Here:

    * ``n_hot`` seeded ``wte`` rows are scaled by ``hot_scale`` (the candidates);
    * hidden unit ``j`` of block 0's MLP detects hot token ``j`` at the current position: its pre-activation is
      ``z = alpha * <LN2-normalised residual, unit(w_j)> - theta``, at least +6 for the matching token and at most -6 for
      every other one, so ``gelu_new(z)`` is ~z for a match and ~0 otherwise (no difference of large activations: bf16
      implementations round ``gelu_new(z)`` to 2^-9 relative, like everything else);
    * its ``c_proj`` row writes ``kappa * w_succ(j) - max(kappa, 1.3 / z_ref) * w_j`` (per unit of activation) into the
      residual stream, which the remaining blocks carry to ``ln_f``: the successor's logit gets a structural lead over the
      best of the other hot rows (the token's own tied-embedding lead is erased), while the randomly initialised rest of
      the network still moves every logit by about one standard deviation, so which margin a step ends up with depends
      on the whole context (attention, positions, KV history).

    ``alpha / theta / kappa / z_ref`` (``z_ref`` = the matching unit's typical activation) come from ``oracle/calibrate_successor_lm.py`` (probe forwards of this very LM) and are
    recorded as literals in ``oracle/cases.py``: building the weights involves no data-dependent branch.
    """
    d = cfg["d_model"]
    assert n_hot <= 4 * d, "the table needs one hidden unit per hot token"
    ids = successor_hot_ids(n_hot, text_vocab)
    wte = w["transformer.wte.weight"]
    wte[ids] *= hot_scale
    hot = wte[ids]                                           # [n, d]
    unit = hot / hot.norm(dim=1, keepdim=True)
    succ = hot[torch.roll(torch.arange(n_hot), -1)]          # row j -> embedding of the token after ids[j]
    g2, b2 = w["transformer.h.0.ln_2.weight"], w["transformer.h.0.ln_2.bias"]
    det = alpha * unit / g2                                  # LN2's gain divided out: the unit sees the normalised residual
    shift = det @ b2                                         # ... and LN2's bias folded into the unit's bias
    fc_w, fc_b = w["transformer.h.0.mlp.c_fc.weight"], w["transformer.h.0.mlp.c_fc.bias"]       # [d, 4d], [4d]
    pr_w = w["transformer.h.0.mlp.c_proj.weight"]                                                  # [4d, d]
    fc_w[:, :n_hot] = det.t()
    fc_b[:n_hot] = -theta - shift
    pr_w[:n_hot] = kappa * succ - max(kappa, 1.3 / z_ref) * hot


def mapper_param_shapes(mapping_type: str, clip_dim: int, d_model: int, prefix_length: int, clip_length: int,
                        num_layers: int) -> "OrderedDict[str, tuple]":
    """Names/shapes in the ``named_parameters()`` order of the reference's ``clip_project``
    (``clipcap.py:256-271``).  This order is also the layout of the flat parameter / gradient
    buffers the C-ABI takes (``include/eavqa_b200.h``)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    d = d_model
    if mapping_type == "mlp":
        h = (d * prefix_length) // 2
        s["model.0.weight"] = (h, clip_dim)
        s["model.0.bias"] = (h,)
        s["model.2.weight"] = (d * prefix_length, h)
        s["model.2.bias"] = (d * prefix_length,)
        return s
    s["prefix_const"] = (prefix_length, d)
    for i in range(num_layers):
        p = "transformer.layers.%d." % i
        s[p + "norm1.weight"] = (d,)
        s[p + "norm1.bias"] = (d,)
        s[p + "attn.to_queries.weight"] = (d, d)
        s[p + "attn.to_keys_values.weight"] = (2 * d, d)
        s[p + "attn.project.weight"] = (d, d)
        s[p + "attn.project.bias"] = (d,)
        s[p + "norm2.weight"] = (d,)
        s[p + "norm2.bias"] = (d,)
        s[p + "mlp.fc1.weight"] = (2 * d, d)
        s[p + "mlp.fc1.bias"] = (2 * d,)
        s[p + "mlp.fc2.weight"] = (d, 2 * d)
        s[p + "mlp.fc2.bias"] = (d,)
    s["linear.weight"] = (clip_length * d, clip_dim)
    s["linear.bias"] = (clip_length * d,)
    return s


def make_mapper_params(mapping_type: str, clip_dim: int, d_model: int, prefix_length: int, clip_length: int,
                       num_layers: int, seed: int = 1, perturb_norm: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """PyTorch-default init: nn.Linear weight/bias ~ U(+-1/sqrt(fan_in)), LayerNorm 1/0,
    ``prefix_const ~ N(0,1)`` (``clipcap.py:235-237``).  ``perturb_norm`` randomises the LayerNorm
    affines so that parity runs see them."""
    g = torch.Generator().manual_seed(seed)
    shapes = mapper_param_shapes(mapping_type, clip_dim, d_model, prefix_length, clip_length, num_layers)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in shapes.items():
        if name == "prefix_const":
            out[name] = torch.randn(*shape, generator=g)
        elif "norm" in name:
            base = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            if perturb_norm:
                base = base + 0.1 * torch.randn(*shape, generator=g)
            out[name] = base
        else:
            wname = name.rsplit(".", 1)[0] + ".weight"
            fan_in = shapes[wname][1]
            bound = 1.0 / math.sqrt(fan_in)
            out[name] = (torch.rand(*shape, generator=g) * 2.0 - 1.0) * bound
    return out


def make_caption_batch(batch: int, text_len: int, clip_dim: int, vocab: int, seed: int = 2021, ragged: bool = False,
                       pad_token_id: Optional[int] = None, final_token_ids: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Conceptual-Captions-shaped batch (``data_loader_conceptual_captions.py:78-104``):
    ``clip_embeddings [B, D]`` fp32 (raw CLIP features, un-normalised), ``input_ids`` /
    ``attention_mask`` / ``labels`` ``[B, T_text]`` int64 with pad -> -100 labels.
    Throughput runs use full rows (``labels = tokens``); parity runs use ``ragged=True``
    (lengths ~ U{8..T_text}, right-padded)."""
    g = torch.Generator().manual_seed(seed)
    pad = pad_token_id if pad_token_id is not None else min(50256, vocab - 1)
    clip = 0.5 * torch.randn(batch, clip_dim, generator=g)
    tokens = torch.randint(0, vocab, (batch, text_len), generator=g, dtype=torch.int64)
    mask = torch.ones(batch, text_len, dtype=torch.int64)
    if ragged:
        lo = min(8, text_len)
        lens = torch.randint(lo, text_len + 1, (batch,), generator=g)
        lens[0] = text_len
        ar = torch.arange(text_len).unsqueeze(0)
        mask = (ar < lens.unsqueeze(1)).long()
        tokens = torch.where(mask.bool(), tokens, torch.full_like(tokens, pad))
    if final_token_ids is not None:        # successor-planted LM: the last VALID token of every row is a hot token
        pick = torch.randint(0, len(final_token_ids), (batch,), generator=g)
        last = mask.sum(dim=1) - 1
        tokens[torch.arange(batch), last] = final_token_ids[pick]
    labels = torch.where(mask.bool(), tokens, torch.full_like(tokens, -100))
    return {"clip_embeddings": clip, "input_ids": tokens, "attention_mask": mask, "labels": labels}


def make_fewshot_batch(batch: int, num_shots: int, clip_dim: int, vocab: int, special_token_id: int, seed: int = 2021,
                       seg_lo: int = 10, seg_hi: int = 20, pad_token_id: Optional[int] = None,
                       final_token_ids: Optional[torch.Tensor] = None, pad_fraction: float = 1.0) -> Dict[str, torch.Tensor]:
    """Few-shot VQA2-shaped batch (``vqa2_datasets.py:65-181``, ``module_parser.py:68-93,466-478``):
    ``clip_embeddings [B, k+1, 1, D]``; each of the k+1 segments is one sentinel id
    (``special_token_id - i``, ``vct0.py:508-509``) followed by U{seg_lo..seg_hi} text ids; rows are
    right-padded to the batch maximum (``module_parser.py:424``).  Text ids are drawn below
    ``special_token_id - num_shots`` so that they never collide with a sentinel.

    ``final_token_ids`` (successor-planted LM): the last text id of every row is drawn from this set -- the analogue of
    every VQA prompt ending in the same few "answer:" tokens -- and only ``pad_fraction`` of the rows keep their ragged
    length; the others get their last segment stretched to the batch maximum, so that the first generated token is
    read at a real token, not at a right-pad position (quirk Q2 stays covered by the ragged rows)."""
    g = torch.Generator().manual_seed(seed)
    pad = pad_token_id if pad_token_id is not None else min(50256, vocab - 1)
    n_img = num_shots + 1
    text_hi = min(vocab, special_token_id - num_shots)
    clip = 0.5 * torch.randn(batch, n_img, 1, clip_dim, generator=g)
    rows = []
    for _ in range(batch):
        row = []
        for i in range(n_img):
            n = int(torch.randint(seg_lo, seg_hi + 1, (1,), generator=g))
            row.append(special_token_id - i)
            row.extend(torch.randint(0, text_hi, (n,), generator=g).tolist())
        rows.append(row)
    T = max(len(r) for r in rows)
    if final_token_ids is not None:
        keep = torch.rand(batch, generator=g) < pad_fraction
        pick = torch.randint(0, len(final_token_ids), (batch,), generator=g)
        for b, r in enumerate(rows):
            if not bool(keep[b]):
                r.extend(torch.randint(0, text_hi, (T - len(r),), generator=g).tolist())
            r[-1] = int(final_token_ids[pick[b]])
    tokens = torch.full((batch, T), pad, dtype=torch.int64)
    mask = torch.zeros(batch, T, dtype=torch.int64)
    for b, r in enumerate(rows):
        tokens[b, :len(r)] = torch.tensor(r)
        mask[b, :len(r)] = 1
    return {"clip_embeddings": clip, "input_ids": tokens, "attention_mask": mask}
