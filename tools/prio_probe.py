"""Does stream priority matter for the step?  The mapper's weight-gradient GEMMs run on the engine's side stream (default
priority) next to the dgrad chain on the caller's stream; with the caller's stream at a HIGHER priority the critical chain
should win the SMs and the weight gradients fill the gaps.  Alternates 40-step timings on a default and a high-priority stream.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import eavqa_b200
import eavqa_b200.synthetic as syn
from eavqa_b200.optim import FlatAdamW

dev = torch.device("cuda", 0)
lm_cfg = syn.lm_config("gpt2", vocab=50257)
model = eavqa_b200.ClipCaptionPrefixB200(prefix_length=10, clip_length=10, prefix_size=512, num_layers=8, mapping_type="transformer",
                                         model_version="gpt2", lm_state_dict=syn.make_lm_weights(lm_cfg, seed=0))
model.clip_project.load_state_dict(syn.make_mapper_params("transformer", 512, lm_cfg["d_model"], 10, 10, 8, seed=1, perturb_norm=True))
model = model.to(dev).train()
b = {k: v.to(dev) for k, v in syn.make_caption_batch(256, 40, 512, 50257, seed=2021).items()}
opt = FlatAdamW(model, lr=1e-4)


def step():
    out = model(question_tokens=b["input_ids"], labels=b["labels"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"])
    out.loss.backward()
    opt.step(model.last_flat_grads)
    opt.zero_grad()


def timed(stream, n=40):
    with torch.cuda.stream(stream):
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


lo, hi = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
for r in range(2):
    print("EAVQA_WGRAD_PRIO=%s: default-priority stream %.3f ms / step, high-priority stream %.3f ms / step" %
          (os.environ.get("EAVQA_WGRAD_PRIO", "-"), timed(lo), timed(hi)), flush=True)
