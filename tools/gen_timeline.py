"""In-pipeline GPU timeline of one few-shot generate call (BASELINE configs[3]) through torch.profiler / CUPTI: what each
kernel ADDS to the timeline (end_k - end_{k-1}; with programmatic dependent launch a kernel's own duration includes waiting for
its predecessor), split into prefill and the single-token steps.  Not a bench value.

    python tools/gen_timeline.py [--dump file.tsv]
"""
import argparse
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import eavqa_b200
import eavqa_b200.synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--dump", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
c = bench.C4
k = c["num_shots"]
lm_cfg = syn.lm_config(c["model_version"], vocab=50257 + k + 1)
lm_w = syn.make_lm_weights(lm_cfg, seed=0, hot_rows=512)
host = syn.make_fewshot_batch(c["batch"], k, c["clip_dim"], lm_cfg["vocab"], 50257 + k, seed=2021, pad_token_id=50256)
b = {kk: v.to(dev) for kk, v in host.items()}
torch.manual_seed(1)
m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=c["prefix_length"], clip_length=c["clip_length"], prefix_size=c["clip_dim"],
                                     num_layers=c["num_layers"], mapping_type=c["mapping_type"], model_version=c["model_version"],
                                     lm_state_dict=lm_w, special_token_id=50257 + k).to(dev).eval()
m.gpt.config.eos_token_id = None


def gen():
    return m.generate(question_tokens=b["input_ids"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"],
                      max_length=c["max_length"], pad_token_id=50256, eos_token_id=None)


for _ in range(4):
    gen()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gen()
    torch.cuda.synchronize()
recs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        recs.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name))
recs.sort(key=lambda r: r[0] + r[1])
if args.dump:
    with open(args.dump, "w") as f:
        for s, d, n in recs:
            f.write("%.3f\t%.3f\t%s\n" % (s, d, n[:140]))


def fam(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"eavqa::|\(anonymous namespace\)::|gk::", "", n)
    mm = re.match(r"([A-Za-z_0-9 ]+)(<[^(]*>)?", n)
    return (mm.group(1) + (mm.group(2) or "")) if mm else n[:50]


by = collections.OrderedDict()
phase_t = collections.OrderedDict()
prev_end = recs[0][0]
t0 = prev_end
n_greedy = 0
for s, d, n in recs:
    phase = "prefill + first pick" if n_greedy == 0 else "single-token steps"
    end = s + d
    delta = max(0.0, end - prev_end)
    prev_end = max(prev_end, end)
    a = by.setdefault((phase, fam(n)), [0, 0.0, 0.0]); a[0] += 1; a[1] += delta; a[2] += d
    p = phase_t.setdefault(phase, [0, 0.0]); p[0] += 1; p[1] += delta
    if "greedy_step" in n:
        n_greedy += 1
total = prev_end - t0
print("one generate call in the pipeline: %.3f ms, %d kernels (profiler attached)" % (total / 1e3, len(recs)))
for ph, (cnt, t) in phase_t.items():
    print("  %-22s %5d launches %8.3f ms  %5.1f%%" % (ph, cnt, t / 1e3, 100 * t / total))
print("%-22s %-44s %5s %9s %8s %9s" % ("phase", "kernel", "n", "added ms", "avg us", "cupti us"))
for (ph, f), (cnt, t, d) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    if t / 1e3 >= 0.03:
        print("%-22s %-44s %5d %9.3f %8.2f %9.2f" % (ph, f[:44], cnt, t / 1e3, t / cnt, d / cnt))
