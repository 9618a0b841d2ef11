"""In-pipeline GPU timeline of the training step (CUPTI through torch.profiler; nsys is not in the image): per kernel
family the summed duration INSIDE the warm, back-to-back pipeline (not ncu's cold, serialised replay), and the idle gaps
between consecutive kernels.  Not a bench value (the profiler adds overhead to every launch).

    python tools/step_timeline.py [--steps 2] [--json out.json]
"""
import argparse
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import eavqa_b200
import eavqa_b200.synthetic as syn
from eavqa_b200.optim import FlatAdamW

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--json", default="")
ap.add_argument("--dump", default="", help="write every kernel record (start us, duration us, stream, name)")
args = ap.parse_args()

dev = torch.device("cuda", 0)
W = dict(prefix_length=10, clip_length=10, clip_dim=512, num_layers=8, mapping_type="transformer", model_version="gpt2", vocab=50257)
lm_cfg = syn.lm_config(W["model_version"], vocab=W["vocab"])
model = eavqa_b200.ClipCaptionPrefixB200(prefix_length=10, clip_length=10, prefix_size=512, num_layers=8, mapping_type="transformer",
                                         model_version="gpt2", lm_state_dict=syn.make_lm_weights(lm_cfg, seed=0))
model.clip_project.load_state_dict(syn.make_mapper_params("transformer", 512, lm_cfg["d_model"], 10, 10, 8, seed=1, perturb_norm=True))
model = model.to(dev).train()
b = {k: v.to(dev) for k, v in syn.make_caption_batch(args.batch, 40, 512, W["vocab"], seed=2021).items()}
opt = FlatAdamW(model, lr=1e-4)


def step():
    out = model(question_tokens=b["input_ids"], labels=b["labels"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"])
    out.loss.backward()
    opt.step(model.last_flat_grads)
    opt.zero_grad()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()

recs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        recs.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name))
recs.sort()
if args.dump:
    with open(args.dump, "w") as f:
        for s, d, n in recs:
            f.write("%.3f\t%.3f\t%s\n" % (s, d, n[:120]))


def family(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"eavqa::|\(anonymous namespace\)::|<unnamed>::", "", n)
    m = re.match(r"([A-Za-z_0-9]+)(<[^(]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n[:60]


t0, t1 = recs[0][0], max(s + d for s, d, _ in recs)
span = (t1 - t0) / args.steps
by = collections.OrderedDict()
busy_until, gap_total, gaps = t0, 0.0, []
for s, d, n in recs:
    f = family(n)
    a = by.setdefault(f, [0, 0.0])
    a[0] += 1
    a[1] += d
    if s > busy_until:
        gap_total += s - busy_until
        gaps.append(s - busy_until)
    busy_until = max(busy_until, s + d)
rows = sorted(by.items(), key=lambda kv: -kv[1][1])
print("training step in the pipeline: %.3f ms per step over %d steps (profiler attached), %d kernels per step; GPU idle between kernels %.3f ms per step (%d gaps)"
      % (span / 1e3, args.steps, len(recs) // args.steps, gap_total / args.steps / 1e3, len(gaps) // args.steps))
print("%10s %7s %9s %9s  %s" % ("ms/step", "share", "launches", "avg us", "kernel"))
for f, (c, d) in rows:
    print("%10.3f %6.1f%% %9d %9.1f  %s" % (d / args.steps / 1e3, 100 * d / args.steps / span, c // args.steps, d / c, f))
if args.json:
    json.dump({"ms_per_step": span / 1e3, "idle_ms_per_step": gap_total / args.steps / 1e3,
               "kernels": [{"kernel": f, "launches": c // args.steps, "ms_per_step": d / args.steps / 1e3} for f, (c, d) in rows]},
              open(args.json, "w"), indent=1)
