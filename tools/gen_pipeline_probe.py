"""Few-shot generation throughput with TWO generate calls in flight (two engines, two host threads, two CUDA streams): the
single-token steps of one batch are a latency chain that leaves the SMs mostly idle, the prefill of the next batch is
tensor-bound -- do they fill each other's gaps?

    python tools/gen_pipeline_probe.py [batches_per_thread]
"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import eavqa_b200
import eavqa_b200.synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
c = bench.C4
k = c["num_shots"]
lm_cfg = syn.lm_config(c["model_version"], vocab=50257 + k + 1)
lm_w = syn.make_lm_weights(lm_cfg, seed=0, hot_rows=512)
host = syn.make_fewshot_batch(c["batch"], k, c["clip_dim"], lm_cfg["vocab"], 50257 + k, seed=2021, pad_token_id=50256)
host = {kk: v.pin_memory() for kk, v in host.items()}


def make():
    torch.manual_seed(1)
    m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=c["prefix_length"], clip_length=c["clip_length"], prefix_size=c["clip_dim"],
                                         num_layers=c["num_layers"], mapping_type=c["mapping_type"], model_version=c["model_version"],
                                         lm_state_dict=lm_w, special_token_id=50257 + k).to(dev).eval()
    m.gpt.config.eos_token_id = None
    return m


def gen(m):
    b = {kk: v.to(dev, non_blocking=True) for kk, v in host.items()}
    return m.generate(question_tokens=b["input_ids"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"],
                      max_length=c["max_length"], pad_token_id=50256, eos_token_id=None)


def run(n_threads):
    models = [make() for _ in range(n_threads)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_threads)]
    outs = [None] * n_threads

    def work(i):
        with torch.cuda.stream(streams[i]):
            for _ in range(n):
                outs[i] = gen(models[i])
    for i in range(n_threads):          # warm-up (arena, tensor maps, kernel attributes)
        with torch.cuda.stream(streams[i]):
            for _ in range(2):
                gen(models[i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    same = all(torch.equal(outs[0], o) for o in outs)
    return n_threads * n * c["batch"] / sec, sec * 1e3 / n, same


for nt in (1, 2, 3):
    a, ms, same = run(nt)
    print(f"{nt} generate call(s) in flight: {a:8.0f} answers/s ({ms:.2f} ms per round of {nt} x {c['batch']} answers; identical outputs: {same})", flush=True)
