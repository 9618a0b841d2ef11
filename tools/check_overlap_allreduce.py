"""N-rank check (torchrun): the bucketed, overlapped gradient all-reduce gives the same summed gradients as one all-reduce
of the whole flat buffer after the step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_overlap_allreduce.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import eavqa_b200
import eavqa_b200.synthetic as syn
from eavqa_b200.parallel import OverlappedGradReducer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = syn.lm_config("gpt2", vocab=50257)
lm_w = syn.make_lm_weights(cfg, seed=0)
torch.manual_seed(1)
m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=10, clip_length=10, prefix_size=512, num_layers=8, mapping_type="transformer",
                                     model_version="gpt2", lm_state_dict=lm_w).cuda().train()
b = {k: v.cuda() for k, v in syn.make_caption_batch(64, 40, 512, 50257, seed=100 + rank, ragged=True).items()}


def step():
    m.zero_grad(set_to_none=True)
    out = m(question_tokens=b["input_ids"], labels=b["labels"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"])
    out.loss.backward()
    return m.last_flat_grads


g = step().clone()
local_norm = float(g.double().norm())
dist.all_reduce(g)
red = OverlappedGradReducer(m)
assert len(red.buckets) == 4 and sum(e - s for s, e in red.buckets) + sum(e - s for s, e in red.rest) == g.numel()
for it in range(3):
    g2 = red.reduce(step())
    torch.cuda.synchronize()
    rel = float((g2.double() - g.double()).norm() / g.double().norm())
    assert rel < 1e-5, rel                                  # only the step's own atomics reorder between runs
red.close()
g3 = step()
assert abs(float(g3.double().norm()) - local_norm) < 1e-4 * local_norm        # events uninstalled: plain local gradients again
if rank == 0:
    print("overlapped all-reduce == plain all-reduce on %d ranks (rel diff %.1e), buckets %s + rest %s" % (world, rel, red.buckets, red.rest))
dist.destroy_process_group()
