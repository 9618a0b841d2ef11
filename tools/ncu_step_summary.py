"""Per-kernel table of ONE training step from an ncu launch list that carries, per launch,
gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c N --csv \
        --log-file launches.csv python bench.py --steps 2 --warmup 1 --only-timed ...

The step = the launches between the last two fused-AdamW kernels.  Prints the table and, with --json, writes the GEMM
kernels' DRAM traffic (bench.py's roofline.traffic reads it).  ncu serialises launches with cold caches: compare
shares, not absolute times.
"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
by = collections.OrderedDict()
for x in csv.DictReader(lines):
    d = by.setdefault(x["ID"], {"name": x["Kernel Name"]})
    v = float(x["Metric Value"].replace(",", ""))
    u, m = x["Metric Unit"], x["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # -> us
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    d[m] = v
L = list(by.values())
idx = [i for i, d in enumerate(L) if "adamw" in d["name"]]
step = L[idx[-2] + 1: idx[-1] + 1]
agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for d in step:
    n = re.sub(r"\(.*", "", d["name"])
    n = re.sub(r"^void |eavqa::|gk::|<unnamed>::|\(anonymous namespace\)::", "", n)
    if "gemm" in n:
        n = re.sub(r"<.*", "", n)
    a = agg[n]
    a[0] += d["gpu__time_duration.sum"]
    a[1] += 1
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[0] for a in agg.values())
print(f"one training step: {len(step)} launches, {tot / 1e3:.3f} ms summed kernel time (ncu: serialised, cold caches)")
print("      time  share  launches   DRAM read   DRAM write   DRAM GB/s  kernel")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0] / 1e3:7.3f} ms {100 * a[0] / tot:5.1f}%  x{a[1]:4d}  {a[2] / 1e6:9.1f} MB {a[3] / 1e6:9.1f} MB  {(a[2] + a[3]) / a[0] / 1e3:8.0f}   {n[:90]}")
if "--json" in sys.argv:
    out = sys.argv[sys.argv.index("--json") + 1]
    g = [a for n, a in agg.items() if "gemm" in n]
    rec = {
        "source": path,
        "gemm_launches_per_step": sum(a[1] for a in g),
        "gemm_dram_read_bytes_per_step": sum(a[2] for a in g),
        "gemm_dram_write_bytes_per_step": sum(a[3] for a in g),
        "gemm_time_share_of_step": sum(a[0] for a in g) / tot,
        "step_launches": len(step),
    }
    rec["gemm_dram_bytes_per_launch"] = (rec["gemm_dram_read_bytes_per_step"] + rec["gemm_dram_write_bytes_per_step"]) / rec["gemm_launches_per_step"]
    json.dump(rec, open(out, "w"), indent=1)
