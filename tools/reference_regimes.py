"""Greedy-answer agreement between three runs of the SAME prompts ("two oracle regimes"):

  (i)   the unmodified reference module in fp32 on the CPU           = the committed fixtures (tests/golden/gen_*.json)
  (ii)  the unmodified reference module on THIS GPU under torch.autocast("cuda", bfloat16)   (Lightning --precision bf16)
  (iii) this repository's CUDA path (bf16 operands, fp32 accumulate)

Needs a B200 and oracle/_ref/clipcap.py (placed by oracle/install_reference.py; travels with the gpurun snapshot).
The k-shot prompt assembly uses the oracle's insert_prefix_into_input (pinned bit-exact against vct0.py:494-533); everything
after it -- the LM, the greedy loop -- is the reference's own code.

    python tools/reference_regimes.py [case ...]          # default: the two 128-prompt configs[3] cases
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import eavqa_b200
from oracle import clip_prefix_lm as orc
from oracle import reference_shim
from oracle.cases import CASES, build_case


def rows_equal(a, b):
    return sum(int(list(x) == list(y)) for x, y in zip(a, b))


def main():
    names = sys.argv[1:] or ["gen_c4_medium_fewshot_128", "gen_c4_medium_fewshot_128_flat", "gen_gpt2_prepend_chain", "gen_gpt2_prepend"]
    real_stdout = os.dup(1)
    os.dup2(2, 1)                      # the reference prints its architecture
    clipcap, _, GPT2Config, holder = reference_shim.import_reference()
    assert str(clipcap.device).startswith("cuda"), "the reference resolves its device at import: run this on the GPU box"
    lines = []
    for name in names:
        case = CASES[name]
        with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
            fx = json.load(f)
        lm_w, mapper_w, batch, cfg = build_case(case)
        kw = dict(max_length=case["max_length"], pad_token_id=case["pad_token_id"], eos_token_id=case["eos_token_id"])
        # (iii) this repository
        m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=case["prefix_length"], clip_length=case["clip_length"],
                                             prefix_size=case["clip_dim"], num_layers=case["num_layers"], mapping_type=case["mapping_type"],
                                             model_version=case["model_version"], lm_state_dict=lm_w,
                                             special_token_id=case.get("special_token_id"))
        m.clip_project.load_state_dict(mapper_w)
        m = m.cuda().eval()
        ours = m.generate(question_tokens=batch["input_ids"], prefix=batch["clip_embeddings"], question_mask=batch["attention_mask"], **kw)
        ours = ours.cpu().tolist() if torch.is_tensor(ours) else ours
        del m
        torch.cuda.empty_cache()
        # (ii) the reference module on the GPU under bf16 autocast
        ref = reference_shim.build_reference_model(clipcap, GPT2Config, holder, case["lm"], lm_w, mapper_w,
                                                   prefix_length=case["prefix_length"], clip_length=case["clip_length"],
                                                   clip_dim=case["clip_dim"], num_layers=case["num_layers"],
                                                   mapping_type=case["mapping_type"]).cuda().eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            if case["num_shots"] is None:
                auto = ref.generate(question_tokens=batch["input_ids"].cuda(), prefix=batch["clip_embeddings"].cuda(),
                                    question_mask=batch["attention_mask"].cuda(), **kw)
            else:
                B, n_img = batch["clip_embeddings"].shape[:2]
                P, d = case["prefix_length"], cfg["d_model"]
                emb_text = ref.gpt.transformer.wte(batch["input_ids"].cuda().clamp_max(case["lm"]["vocab"] - 1)).float()
                pre = ref.clip_project(batch["clip_embeddings"].cuda().reshape(-1, case["clip_dim"])).float().reshape(B, n_img, P, d)
                emb, msk = orc.insert_prefix_into_input(P, n_img - 1, batch["input_ids"], emb_text.cpu(), pre.cpu(), batch["attention_mask"],
                                                        case["special_token_id"])
                auto = ref._generate_from_embeddings(emb.cuda(), msk.cuda(), **kw)
        del ref
        torch.cuda.empty_cache()
        fp32 = fx["tokens"]
        n = len(fp32)
        clear = [i for i, mg in enumerate(fx["margins"]) if min(mg) >= (0.05 if "top_logits" not in fx else
                                                                          max(0.05, 0.01 * abs(fx["stats"]["top_logit_median"])))]
        line = ("%-34s %3d prompts | this repo vs reference fp32: %5.1f %% | reference bf16-autocast vs reference fp32: %5.1f %% | "
                "this repo vs reference bf16-autocast: %5.1f %% | rows without a near-tie (%d): this repo %5.1f %%, autocast %5.1f %%"
                % (name, n, 100.0 * rows_equal(ours, fp32) / n, 100.0 * rows_equal(auto, fp32) / n, 100.0 * rows_equal(ours, auto) / n,
                   len(clear), 100.0 * sum(int(list(ours[i]) == list(fp32[i])) for i in clear) / max(len(clear), 1),
                   100.0 * sum(int(list(auto[i]) == list(fp32[i])) for i in clear) / max(len(clear), 1)))
        lines.append(line)
        print(line, file=sys.stderr, flush=True)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
