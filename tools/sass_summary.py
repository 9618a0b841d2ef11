"""Opcode census of the shipped library: `cuobjdump -sass libeavqa_b200.so`, grouped by kernel family.

Proves which hardware paths the kernels use (profiles/README.md, B200_PROFILING.md mnemonics):
  UTCHMMA / UTCQMMA = tcgen05.mma        LDTM / STTM     = tcgen05.ld / st (TMEM)         UTCBAR = tcgen05.commit
  UTMALDG / UTMASTG / UTMAREDG / UBLKCP  = TMA tensor load / store / reduce-add / 1-D bulk copy
  HMMA = mma.sync (legacy tensor path)   LDGSTS = cp.async      SYNCS = mbarrier      UCGABAR = cluster barrier

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "explicit-alignment-for-vqa-tasks_b200", "libeavqa_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF", "HMMA", "LDGSTS", "SYNCS",
         "UCGABAR", "REDG", "ATOMG", "FFMA2", "FMUL2", "FADD2", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "ACQBULK", "CCTL"]


def family(name: str) -> str:
    name = re.sub(r"^void |eavqa::|\(anonymous namespace\)::|<unnamed>::|dc::|gk::", "", name)
    m = re.match(r"([A-Za-z_0-9]+)", name)
    base = m.group(1) if m else name
    if base.startswith("gemm_bf16_tn_2cta"):
        return "gemm_bf16_tn_2cta_kernel (cta_group::2 pair, all tile widths / epilogues)"
    if base.startswith("gemm_bf16_tn"):
        return "gemm_bf16_tn_kernel (single CTA, all tile widths / epilogues, MN-major wgrad)"
    return base


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dm):
        demangle[n] = d
    fam = collections.OrderedDict()
    count = collections.Counter()
    for n, c in kernels.items():
        f = family(demangle.get(n, n))
        fam.setdefault(f, collections.Counter()).update(c)
        count[f] += 1
    total = collections.Counter()
    for c in fam.values():
        total.update(c)
    print("SASS opcode census of %s (sm_100a; cuobjdump -sass), %d kernels in %d families" % (os.path.basename(LIB), len(kernels), len(fam)))
    print("whole library: " + ", ".join("%s %d" % (k, total[k]) for k in WATCH if total[k]))
    print()
    for f, c in sorted(fam.items(), key=lambda kv: -sum(kv[1].values())):
        hits = ", ".join("%s %d" % (k, c[k]) for k in WATCH if c[k])
        print("%-90s x%-3d %7d instr | %s" % (f[:90], count[f], sum(c.values()), hits))


if __name__ == "__main__":
    main()
