"""Attribution of the step time from a tools/step_timeline.py --dump file: with programmatic dependent launch a kernel's
CUPTI duration includes the time it waits for its predecessor, so the cost of kernel k on the serial main-stream chain is
taken as end_k - end_{k-1} (what it ADDS to the timeline).  Side-stream weight-gradient GEMMs (MN-major instantiations) are
left out of the chain; phases are cut at the head GEMM / ce_dlogits / the mapper backward's first kernel."""
import collections
import re
import sys

recs = []
for line in open(sys.argv[1]):
    s, d, n = line.rstrip("\n").split("\t")
    recs.append((float(s), float(d), n))
# second step only: from the last pack_batch_kernel on
starts = [i for i, r in enumerate(recs) if "pack_batch_kernel" in r[2]]
recs = recs[starts[-1]:]


def fam(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"eavqa::|\(anonymous namespace\)::|gk::", "", n)
    m = re.match(r"([A-Za-z_0-9 ]+)(<[^(]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n[:50]


main = [r for r in recs if "true>" not in r[2]]
side = [r for r in recs if "true>" in r[2]]
main.sort(key=lambda r: r[0] + r[1])
by = collections.OrderedDict()
phase_by = collections.OrderedDict()
phase = "mapper fwd"
prev_end = main[0][0]
t_begin = prev_end
for s, d, n in main:
    f = fam(n)
    if "embed_rows" in n: phase = "LM fwd"
    if "ce_plan" in n: phase = "head + CE"
    if "layernorm_bwd_lean" in n and phase == "head + CE": phase = "LM bwd"
    if "scatter_prefix_grad" in n: phase = "mapper bwd"
    if "adamw" in n: phase = "adamw"
    end = s + d
    delta = max(0.0, end - prev_end)
    prev_end = max(prev_end, end)
    a = by.setdefault((phase, f), [0, 0.0, 0.0]); a[0] += 1; a[1] += delta; a[2] += d
    p = phase_by.setdefault(phase, [0, 0.0]); p[0] += 1; p[1] += delta
total = prev_end - t_begin
print("main-stream chain of one step: %.3f ms; side-stream weight-gradient GEMMs: %d launches, %.3f ms summed" % (total / 1e3, len(side), sum(r[1] for r in side) / 1e3))
for ph, (c, t) in phase_by.items():
    print("  %-12s %4d launches %8.3f ms  %5.1f%%" % (ph, c, t / 1e3, 100 * t / total))
print("%-12s %-46s %5s %9s %8s %9s" % ("phase", "kernel", "n", "added ms", "avg us", "cupti us"))
for (ph, f), (c, t, d) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    if t / 1e3 >= 0.02:
        print("%-12s %-46s %5d %9.3f %8.1f %9.1f" % (ph, f[:46], c, t / 1e3, t / c, d / c))
