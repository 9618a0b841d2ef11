"""Calibrate the successor-planted synthetic LM (TEST INFRASTRUCTURE; ``eavqa_b200.synthetic.plant_successor_table``).

Finds, by probe forwards of the oracle on the very LM a case uses, the three constants the planted bigram table needs:

* ``alpha`` / ``theta`` / ``z_ref``: detector gain and threshold such that the matching unit's pre-activation is >= +6
  (positions >= 8; ``z_ref`` = its median) and every non-matching one <= -6 on the probe (hot tokens at random positions
  in random context);
* ``kappa``: strength of the written ``w_succ - w_tok`` vector such that the successor's logit leads the best other hot
  row by ``target`` logit standard deviations (the std of a hot row's logit is ``|w_hot|``).

The constants are printed and then pasted, as literals, into ``oracle/cases.py`` so that weight construction on any box is
free of data-dependent decisions.

    python oracle/calibrate_successor_lm.py gpt2-medium 512 50256 4.0
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import eavqa_b200.synthetic as syn            # noqa: E402
from oracle import clip_prefix_lm as orc      # noqa: E402


def probe_tokens(ids, text_vocab, rows=4, T=64, seed=3):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, text_vocab, (rows, T), generator=g)
    hot_pos = torch.rand(rows, T, generator=g) < 0.5
    pick = ids[torch.randint(0, len(ids), (rows, T), generator=g)]
    return torch.where(hot_pos, pick, tok), hot_pos


def residual_after_attention0(w, cfg, tok):
    """Normalised input of block 0's MLP (LN2 without its affine) for every probe position."""
    B, T = tok.shape
    d, H = cfg["d_model"], cfg["n_head"]
    h = w["transformer.wte.weight"][tok] + w["transformer.wpe.weight"][:T]
    u = orc.layer_norm(h, w["transformer.h.0.ln_1.weight"], w["transformer.h.0.ln_1.bias"])
    qkv = u @ w["transformer.h.0.attn.c_attn.weight"] + w["transformer.h.0.attn.c_attn.bias"]
    q, k, v = (t.view(B, T, H, d // H).transpose(1, 2) for t in qkv.split(d, dim=2))
    s = (q @ k.transpose(-1, -2)) * ((d // H) ** -0.5)
    s = s.masked_fill(~torch.ones(T, T, dtype=torch.bool).tril(), float("-inf"))
    a = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, T, d)
    h = h + a @ w["transformer.h.0.attn.c_proj.weight"] + w["transformer.h.0.attn.c_proj.bias"]
    mu = h.mean(-1, keepdim=True)
    return (h - mu) * torch.rsqrt(((h - mu) ** 2).mean(-1, keepdim=True) + 1e-5)


def calibrate(model_version, n_hot, text_vocab, hot_scale, vocab=None, target=6.5, n_positions=None):
    cfg = syn.lm_config(model_version, vocab=vocab, n_positions=n_positions)
    ids = syn.successor_hot_ids(n_hot, text_vocab)
    base = dict(n_hot=n_hot, text_vocab=text_vocab, hot_scale=hot_scale)
    w = syn.make_lm_weights(cfg, seed=0, successor=dict(base, alpha=0.0, theta=0.0, kappa=0.0, z_ref=1.0))
    tok, hot_pos = probe_tokens(ids, text_vocab)
    xhat = residual_after_attention0(w, cfg, tok)                                   # [B, T, d]
    hot = w["transformer.wte.weight"][ids]
    unit = hot / hot.norm(dim=1, keepdim=True)
    G = xhat @ unit.t()                                                             # [B, T, n]
    match = (tok.unsqueeze(-1) == ids.view(1, 1, -1))
    m_lo = float(G[:, 8:][match[:, 8:]].min())   # the first few positions attend over too few keys to average block 0's
                                                 # attention output down; answers are generated far behind them
    n_hi = float(G[~match].max()) + 1.5          # the probe sees 192 positions, a real batch 20 000
    assert m_lo > n_hi + 1.0, ("hot tokens are not separable in block 0", m_lo, n_hi)
    alpha = float("%.3g" % (12.0 / (m_lo - n_hi)))
    theta = float("%.3g" % (alpha * (m_lo + n_hi) / 2))
    z_ref = float("%.3g" % (alpha * float(G[:, 8:][match[:, 8:]].median()) - theta))     # typical activation of a matching unit
    print(f"detector: min match {m_lo:.2f}, max non-match (+1.5) {n_hi:.2f} -> alpha {alpha}, theta {theta}, z_ref {z_ref}")

    sigma = float(hot.norm(dim=1).mean())        # std of a hot row's logit (ln_f output has norm sqrt(d))
    kappa = 1.0 / z_ref

    def lead(kappa):
        w = syn.make_lm_weights(cfg, seed=0, successor=dict(base, alpha=alpha, theta=theta, kappa=kappa, z_ref=z_ref))
        hidden = orc.gpt2_hidden(w, w["transformer.wte.weight"][tok], torch.ones_like(tok), cfg["n_layer"], cfg["n_head"])
        sel = hot_pos.clone()
        sel[:, :8] = False
        logits = hidden[sel] @ w["transformer.wte.weight"][ids].t()                 # hot rows only: [n_pos, n]
        j = (tok[sel].unsqueeze(-1) == ids.view(1, -1)).float().argmax(-1)
        succ = (j + 1) % n_hot
        s = logits.gather(1, succ.unsqueeze(1)).squeeze(1)
        other = logits.scatter(1, succ.unsqueeze(1), float("-inf")).max(dim=1).values
        return (s - other) / sigma, s / sigma

    for it in range(4):                          # fixed-point: the structural lead is close to linear in kappa
        m, s = lead(kappa)
        print(f"kappa {kappa:.4f}: successor logit {float(s.median()):.2f} sigma, lead over the best other hot row median "
              f"{float(m.median()):.2f} / min {float(m.min()):.2f} sigma (sigma = {sigma:.3f})")
        kappa = float("%.3g" % (kappa * target / float(s.median())))
    m, s = lead(kappa)
    print(f"final kappa {kappa}: lead median {float(m.median()):.2f} min {float(m.min()):.2f} sigma; absolute "
          f"{float(m.median()) * sigma:.2f} / {float(m.min()) * sigma:.2f}")
    print("successor=dict(n_hot=%d, text_vocab=%d, hot_scale=%s, alpha=%s, theta=%s, kappa=%s, z_ref=%s)"
          % (n_hot, text_vocab, hot_scale, alpha, theta, kappa, z_ref))


if __name__ == "__main__":
    mv, n_hot, tv, hs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    kw = {}
    if len(sys.argv) > 5:
        kw["vocab"] = int(sys.argv[5])
    if len(sys.argv) > 6:
        kw["target"] = float(sys.argv[6])
    calibrate(mv, n_hot, tv, hs, **kw)
