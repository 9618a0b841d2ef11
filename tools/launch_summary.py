"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for ONE step
(the launches between the last two fused-AdamW kernels).  Usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
recs = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in recs]
idx = [i for i, n in enumerate(names) if "adamw" in n]
start, end = idx[-2] + 1, idx[-1] + 1
step = recs[start:end]
agg = collections.defaultdict(lambda: [0.0, 0])
for x in step:
    n = re.sub(r"\(.*", "", x["Kernel Name"])
    n = re.sub(r"^void |eavqa::|\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", n)
    v = float(x["Metric Value"].replace(",", ""))
    unit = x["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    agg[n][0] += v
    agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"one training step: {len(step)} launches, {tot / 1e3:.3f} ms summed kernel time (ncu: serialised, cold caches)")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  x{c:4d}  {n[:100]}")
