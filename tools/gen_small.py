"""generate() at a tiny batch: separates host launch cost from GPU time (EAVQA_TIMING=1 prints the host enqueue time)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import types, torch, bench, eavqa_b200, eavqa_b200.synthetic as syn
bench.C4["batch"] = int(sys.argv[1]) if len(sys.argv) > 1 else 8
print(bench.bench_generate(torch.device("cuda", 0), eavqa_b200, syn, types.SimpleNamespace(steps=3))["ms_per_batch"])
