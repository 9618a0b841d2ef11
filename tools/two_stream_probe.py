"""Probe: does running two half-batches (2 x 128 samples) as independent chains on two CUDA streams beat one chain of 256?
Two engines (each with its own frozen-LM copy) stand in for the two chains; the step = forward + backward (no optimiser).
With cluster-launch-control scheduling a kernel uses whatever SMs are free, so the chains fill each other's launch ramps
and tails."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eavqa_b200
import eavqa_b200.synthetic as syn

cfg = syn.lm_config("gpt2", vocab=50257)
lm_w = syn.make_lm_weights(cfg, seed=0)


def make():
    torch.manual_seed(1)
    return eavqa_b200.ClipCaptionPrefixB200(prefix_length=10, clip_length=10, prefix_size=512, num_layers=8, mapping_type="transformer",
                                            model_version="gpt2", lm_state_dict=lm_w).cuda().train()


def step(m, b):
    m.zero_grad(set_to_none=True)
    m(question_tokens=b["input_ids"], labels=b["labels"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"]).loss.backward()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


full = {k: v.cuda() for k, v in syn.make_caption_batch(256, 40, 512, 50257, seed=2021).items()}
halves = [{k: v[i * 128:(i + 1) * 128].contiguous() for k, v in full.items()} for i in range(2)]
m0, m1 = make(), make()
t_full = timed(lambda: step(m0, full))
t_half = timed(lambda: step(m0, halves[0]))
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    cur = torch.cuda.current_stream()
    s0.wait_stream(cur); s1.wait_stream(cur)
    with torch.cuda.stream(s0):
        step(m0, halves[0])
    with torch.cuda.stream(s1):
        step(m1, halves[1])
    cur.wait_stream(s0); cur.wait_stream(s1)


t_both = timed(both)
print(f"one chain of 256: {t_full:.3f} ms | one chain of 128: {t_half:.3f} ms (x2 = {2 * t_half:.3f}) | two concurrent chains of 128: {t_both:.3f} ms")
