"""Pin the oracle against the UNMODIFIED reference and write ``tests/golden/*.json``.

Runs only where ``/root/reference`` exists (the dev container).  It

1. imports ``/root/reference/src/models/clipcap.py`` and ``vct0.py`` behind the shim of
   SURVEY.md Appendix B (``transformers.AdamW`` alias; ``GPT2LMHeadModel.from_pretrained`` ->
   random-init ``GPT2Config``; ``flamingo_pytorch`` stub) -- no reference source is copied;
2. loads the seeded synthetic weights (``eavqa_b200.synthetic``) into the reference modules;
3. checks ``oracle/clip_prefix_lm.py`` against the reference's loss, mapper gradients,
   greedy tokens and prefix splice (incl. the golden tensors of ``vct0_test.py:79-211``);
4. records the REFERENCE's outputs as fixtures under ``tests/golden/``.

    python oracle/validate_against_reference.py            # all cases
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("EAVQA_REFERENCE", "/root/reference")

import eavqa_b200.synthetic as syn            # noqa: E402
from oracle import clip_prefix_lm as orc      # noqa: E402
from oracle.cases import CASES, SPLICE_GOLDEN, build_case   # noqa: E402


def import_reference():
    from transformers import GPT2Config, GPT2LMHeadModel    # heavy import first (lazy module is replaced)
    sys.modules["transformers"].AdamW = torch.optim.AdamW     # clipcap.py:10 imports a removed, unused symbol
    holder = {}

    def fake_from_pretrained(cls, name, *a, **k):
        return GPT2LMHeadModel(holder["cfg"])
    GPT2LMHeadModel.from_pretrained = classmethod(fake_from_pretrained)
    fl = types.ModuleType("flamingo_pytorch")
    fl.PerceiverResampler = object                            # vct0.py:17 import only
    sys.modules.setdefault("flamingo_pytorch", fl)
    sys.path.insert(0, os.path.join(REF, "src", "models"))
    import clipcap
    import vct0
    return clipcap, vct0, GPT2Config, holder


def build_reference_model(clipcap, GPT2Config, holder, case, lm_w, mapper_w):
    c = case["lm"]
    holder["cfg"] = GPT2Config(vocab_size=c["vocab"], n_positions=c["n_positions"], n_embd=c["d_model"],
                               n_layer=c["n_layer"], n_head=c["n_head"])
    m = clipcap.ClipCaptionPrefix(prefix_length=case["prefix_length"], clip_length=case["clip_length"],
                                  prefix_size=case["clip_dim"], num_layers=case["num_layers"],
                                  mapping_type=case["mapping_type"], model_version="synthetic")
    sd = {k: v for k, v in lm_w.items()}
    sd["lm_head.weight"] = lm_w["transformer.wte.weight"]
    missing, unexpected = m.gpt.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(("attn.bias" in k or "masked_bias" in k) for k in missing), missing
    m.clip_project.load_state_dict(mapper_w, strict=True)
    assert [n for n, _ in m.clip_project.named_parameters()] == list(mapper_w.keys()), "flat layout order"
    return m.train()


def grad_summary(g: torch.Tensor) -> dict:
    g = g.double().flatten()
    probe = torch.Generator().manual_seed(11)
    sign = (torch.randint(0, 2, (g.numel(),), generator=probe) * 2 - 1).double()
    return {"norm": float(g.norm()), "sum": float(g.sum()), "probe": float((g * sign).sum()),
            "head": g[:6].tolist(), "tail": g[-6:].tolist()}


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def run_train_case(name, case, clipcap, GPT2Config, holder):
    lm_w, mapper_w, batch, cfg = build_case(case)
    ref = build_reference_model(clipcap, GPT2Config, holder, case, lm_w, mapper_w)
    t0 = time.time()
    out = ref(question_tokens=batch["input_ids"], labels=batch["labels"], prefix=batch["clip_embeddings"],
              question_mask=batch["attention_mask"], pad_token_id=case["pad_token_id"])
    out.loss.backward()
    t_ref = time.time() - t0
    assert all(p.grad is None for p in ref.gpt.parameters()), "frozen LM got grads"
    ref_grads = {n: p.grad.detach() for n, p in ref.clip_project.named_parameters()}
    loss_o, grads_o = orc.train_step(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"],
                                     batch["attention_mask"], batch["labels"])
    rel = abs(loss_o - float(out.loss)) / abs(float(out.loss))
    worst = min(cosine(ref_grads[k], grads_o[k]) for k in ref_grads)
    flat_r = torch.cat([ref_grads[k].flatten() for k in ref_grads])
    flat_o = torch.cat([grads_o[k].flatten() for k in ref_grads])
    relg = float((flat_r - flat_o).double().norm() / flat_r.double().norm())
    print(f"[train {name}] ref loss {float(out.loss):.6f} oracle {loss_o:.6f} rel {rel:.2e}; "
          f"grad cos(min over params) {worst:.8f} rel-l2 {relg:.2e}; ref step {t_ref:.2f}s")
    assert rel < 2e-5 and worst > 0.99999 and relg < 1e-3, "oracle does not match the reference"
    res = {"kind": "train", "case": case, "loss": float(out.loss),
           "n_valid": int((torch.nn.functional.pad(batch["labels"], (1, 0), value=-100)[:, 1:] != -100).sum()),
           "grads": {k: grad_summary(v) for k, v in ref_grads.items()},
           "grad_total_norm": float(flat_r.double().norm())}
    if case["batch"] >= 64:
        # full-size case: the gradient itself (167 MB) cannot be committed, so keep a seeded sample of it -- the cosine
        # over 16384 random coordinates estimates the full one to ~1e-3
        idx = torch.randperm(flat_r.numel(), generator=torch.Generator().manual_seed(13))[:16384]
        res["grad_sample"] = {"seed": 13, "n": 16384, "values": [float("%.6g" % v) for v in flat_r[idx].tolist()]}
    return res


def run_generate_case(name, case, clipcap, vct0, GPT2Config, holder):
    lm_w, mapper_w, batch, cfg = build_case(case)
    ref = build_reference_model(clipcap, GPT2Config, holder, case, lm_w, mapper_w).eval()
    kw = dict(max_length=case["max_length"], pad_token_id=case["pad_token_id"], eos_token_id=case["eos_token_id"])
    tops = []
    t0 = time.time()
    with torch.no_grad():
        if case["num_shots"] is None:
            ref_tokens = ref.generate(question_tokens=batch["input_ids"], prefix=batch["clip_embeddings"],
                                      question_mask=batch["attention_mask"], **kw)
            got, margins = orc.generate(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"],
                                        batch["attention_mask"], return_margins=True, top_logits_out=tops, **kw)
        else:
            # GPT-2 few-shot = the T0 path's prompt assembly (vct0.py:446-464) + clipcap's greedy loop
            B, n_img = batch["clip_embeddings"].shape[:2]
            P, d = case["prefix_length"], cfg["d_model"]
            emb_text = ref.gpt.transformer.wte(batch["input_ids"])
            pre = ref.clip_project(batch["clip_embeddings"].reshape(-1, case["clip_dim"])).reshape(B, n_img, P, d).contiguous()
            stub = types.SimpleNamespace(prefix_length=P, lm_embedding_size=d)
            emb, msk = vct0.VCT0Model.insert_prefix_into_input(stub, B, n_img - 1, batch["input_ids"], emb_text, pre,
                                                               batch["attention_mask"], case["special_token_id"])
            ref_tokens = ref._generate_from_embeddings(emb, msk, **kw)
            got, margins = orc.generate_few_shot(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"],
                                                 batch["attention_mask"], case["special_token_id"],
                                                 return_margins=True, top_logits_out=tops, **kw)
    same = sum(int(a == b) for a, b in zip(ref_tokens, got))
    tops = torch.stack(tops, dim=1)
    # how degenerate are the answers?  (round-1 review: hot-row sharpening made most rows repeat ONE token)
    flat = [t for row in ref_tokens for t in row if t != case["pad_token_id"]]
    one_token_rows = sum(int(len(set(row)) == 1) for row in ref_tokens if len(row) > 1)
    mixed = sum(int(a != b) for a, b in zip(ref_tokens, got))
    print(f"[generate {name}] identical answers {same}/{len(ref_tokens)}; top-2 margin median {float(margins.median()):.4f} "
          f"min {float(margins.min()):.5f} (logit std over the vocabulary: see fixture); distinct tokens {len(set(flat))} in "
          f"{len(flat)} outputs, rows repeating one token {one_token_rows}/{len(ref_tokens)}; {time.time() - t0:.0f}s; "
          f"first {ref_tokens[0]}")
    # fp32 summation order differs between HF's fused kernels and the restatement: a row may legitimately part ways at a
    # step whose top-2 margin is at fp32 rounding level (seen only on the un-sharpened 128-row case)
    for a, b, mg in zip(ref_tokens, got, margins):
        if a != b:
            first = next(i for i, (x, y) in enumerate(zip(a, b)) if x != y)
            assert float(mg[first]) < 2e-4, "oracle greedy tokens differ from the reference beyond an fp32-level tie"
    assert mixed <= max(1, len(ref_tokens) // 64), "oracle greedy tokens differ from the reference"
    return {"kind": "generate", "case": case, "tokens": ref_tokens,
            "margins": [[round(float(x), 6) for x in row] for row in margins],
            "top_logits": [[round(float(x), 5) for x in row] for row in tops],
            "stats": {"distinct_tokens": len(set(flat)), "outputs": len(flat), "rows_repeating_one_token": one_token_rows,
                      "margin_median": float(margins.median()), "margin_min": float(margins.min()),
                      "top_logit_median": float(tops.median())}}


def run_splice(vct0):
    results = []
    for g in SPLICE_GOLDEN:
        toks = torch.tensor(g["question_tokens"], dtype=torch.int64)
        msk = torch.tensor(g["question_masks"], dtype=torch.int64)
        text = torch.tensor(g["text_embeddings"])
        pre = torch.tensor(g["prefix_projections"])
        stub = types.SimpleNamespace(prefix_length=g["prefix_length"], lm_embedding_size=text.shape[-1])
        e_ref, m_ref = vct0.VCT0Model.insert_prefix_into_input(stub, toks.shape[0], g["num_shots"], toks, text, pre, msk)
        e_o, m_o = orc.insert_prefix_into_input(g["prefix_length"], g["num_shots"], toks, text, pre, msk)
        exp_e, exp_m = torch.tensor(g["expected_embeddings"]), torch.tensor(g["expected_masks"])
        assert torch.equal(e_ref, exp_e) and torch.equal(m_ref, exp_m), "reference vs its own golden"
        assert torch.equal(e_o, exp_e) and torch.equal(m_o, exp_m), "oracle vs vct0_test golden"
        results.append(g["name"])
    # random cases, larger shapes
    gen = torch.Generator().manual_seed(5)
    for trial in range(20):
        B, k, P, d = 3, int(torch.randint(0, 5, (1,), generator=gen)), int(torch.randint(1, 6, (1,), generator=gen)), 5
        b = syn.make_fewshot_batch(B, k, 4, 1000, 990, seed=100 + trial, seg_lo=1, seg_hi=6)
        toks, msk = b["input_ids"], b["attention_mask"]
        text = torch.randn(B, toks.shape[1], d, generator=gen)
        pre = torch.randn(B, k + 1, P, d, generator=gen)
        stub = types.SimpleNamespace(prefix_length=P, lm_embedding_size=d)
        e_ref, m_ref = vct0.VCT0Model.insert_prefix_into_input(stub, B, k, toks, text, pre, msk, 990)
        e_o, m_o = orc.insert_prefix_into_input(P, k, toks, text, pre, msk, 990)
        assert torch.equal(e_ref, e_o) and torch.equal(m_ref, m_o), "oracle splice differs (random case %d)" % trial
    print(f"[splice] {len(results)} vct0_test golden cases + 20 random cases identical")
    return results


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    clipcap, vct0, GPT2Config, holder = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])
    run_splice(vct0)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        if case["kind"] == "train":
            res = run_train_case(name, case, clipcap, GPT2Config, holder)
        else:
            res = run_generate_case(name, case, clipcap, vct0, GPT2Config, holder)
        res["generated_by"] = "oracle/validate_against_reference.py (reference outputs, fp32 CPU)"
        res["torch"] = torch.__version__
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(res, f, indent=1)
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main()
