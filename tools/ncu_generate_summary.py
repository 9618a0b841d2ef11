"""Per-kernel table of ONE few-shot `generate` call from an ncu launch list (gpu__time_duration.sum per launch):

    ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file launches.csv python tools/gen_bench.py 1
    python tools/ncu_generate_summary.py launches.csv

A call starts at a mapper weight packing (`pack_batch_kernel`); the last call with the most launches is summarised
(bench_generate ends with prefill-only calls).  Run with EAVQA_DECODE_GRAPH=0 so that the steps are plain launches.  ncu serialises launches with cold caches: compare shares, not absolute times.
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
recs = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in recs]
starts = [i for i, n in enumerate(names) if "pack_batch" in n]
# bench_generate ends with prefill-only calls (max_length = 1, its roofline split): take the last FULL call
segs = [recs[a:b] for a, b in zip(starts, starts[1:] + [len(recs)])] if starts else [recs]
longest = max(len(x) for x in segs)
call = [x for x in segs if len(x) == longest][-1]
agg = collections.defaultdict(lambda: [0.0, 0])
for x in call:
    n = re.sub(r"\(.*", "", x["Kernel Name"])
    n = re.sub(r"^void |eavqa::|gk::|dc::|<unnamed>::|\(anonymous namespace\)::", "", n)
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[n][0] += v
    agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"one few-shot generate call (BASELINE configs[3]: GPT-2 medium, batch 128, 10 new tokens): {len(call)} launches, "
      f"{tot / 1e3:.3f} ms summed kernel time (ncu: serialised, cold caches)")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  x{c:5d}  avg {t / c:8.1f} us  {n[:100]}")
