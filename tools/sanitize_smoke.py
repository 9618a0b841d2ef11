"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck): one training step (T <= 64: the persistent
attention forward, the single-block backward), one few-shot generation (prefill with resident K/V, decode kernels, CUDA-graph
replay off and on), and one generation through the persistent decode kernel.

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import eavqa_b200
from oracle.cases import CASES, _case, build_case


def model_for(case, lm_w, mapper_w):
    m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=case["prefix_length"], clip_length=case["clip_length"], prefix_size=case["clip_dim"],
                                         num_layers=case["num_layers"], mapping_type=case["mapping_type"], model_version=case["model_version"],
                                         lm_state_dict=lm_w, special_token_id=case.get("special_token_id"))
    m.clip_project.load_state_dict(mapper_w)
    return m.cuda()


case = CASES["train_tiny_transformer"]
lm_w, mapper_w, batch, cfg = build_case(case)
m = model_for(case, lm_w, mapper_w).train()
out = m(question_tokens=batch["input_ids"], labels=batch["labels"], prefix=batch["clip_embeddings"], question_mask=batch["attention_mask"])
out.loss.backward()
torch.cuda.synchronize()
print("train step ok", float(out.loss))

# a prompt of ~100 positions: two key blocks in the prefill kernel
case = _case("generate", "gpt2-tiny", "mlp", 3, 90, 64, 4, 4, 2, ragged=True, hot_rows=64, max_length=4, n_positions=256)
lm_w, mapper_w, batch, cfg = build_case(case)
m = model_for(case, lm_w, mapper_w).eval()
kw = dict(question_tokens=batch["input_ids"], prefix=batch["clip_embeddings"], question_mask=batch["attention_mask"],
          max_length=case["max_length"], pad_token_id=case["pad_token_id"], eos_token_id=None)
m.gpt.config.eos_token_id = None
a = m.generate(**kw)
b = m.generate(**kw)          # captures the decode graph
c = m.generate(**kw)          # replays it
os.environ["EAVQA_DECODE_CHAIN"] = "1"
d = m.generate(**kw)
torch.cuda.synchronize()
print("generate ok", a == b == c, a == d)
