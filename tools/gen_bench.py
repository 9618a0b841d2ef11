"""Few-shot generation benchmark (BASELINE configs[3]) on its own: python tools/gen_bench.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import types
import torch
import bench
import eavqa_b200
import eavqa_b200.synthetic as syn
args = types.SimpleNamespace(steps=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
print(bench.bench_generate(torch.device("cuda", 0), eavqa_b200, syn, args))
