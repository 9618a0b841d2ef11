"""N-rank check (torchrun): the module wrapped in a REAL ``torch.nn.parallel.DistributedDataParallel`` -- what Lightning's
DDP strategy does with ``ClipCapExecutor.model`` (main.py:133-138) -- yields the same mean-over-ranks mapper gradients as
bench.py's hand-rolled path (one all-reduce of the flat gradient buffer, 1/W folded into AdamW).  Under DDP the reducer's
hooks see the mapper parameters as separate leaves (views into the flat buffer) and nothing else: the frozen LM holds no
``nn.Parameter``.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_wrap.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

import eavqa_b200
import eavqa_b200.synthetic as syn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mapping in ("transformer", "mlp"):
    cfg = syn.lm_config("gpt2", vocab=50257)
    lm_w = syn.make_lm_weights(cfg, seed=0)
    m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=10, clip_length=10, prefix_size=512, num_layers=8, mapping_type=mapping,
                                         model_version="gpt2", lm_state_dict=lm_w)
    m.clip_project.load_state_dict(syn.make_mapper_params(mapping, 512, 768, 10, 10, 8, seed=1, perturb_norm=True))
    m = m.cuda().train()
    b = {k: v.cuda() for k, v in syn.make_caption_batch(32, 40, 512, 50257, seed=100 + rank, ragged=True).items()}
    kw = dict(question_tokens=b["input_ids"], labels=b["labels"], prefix=b["clip_embeddings"], question_mask=b["attention_mask"])

    # (1) bench.py's path: flat buffer, one all-reduce, mean
    m.zero_grad(set_to_none=True)
    m(**kw).loss.backward()
    flat = m.last_flat_grads.clone()
    dist.all_reduce(flat)
    flat /= world

    # (2) real DDP around the same module
    ddp = DDP(m, device_ids=[local])
    n_params = len(list(ddp.parameters()))
    m.zero_grad(set_to_none=True)
    out = ddp(**kw)
    out.loss.backward()
    torch.cuda.synchronize()
    got = torch.cat([p.grad.flatten() for p in m.clip_project.parameters()])
    rel = float((got.double() - flat.double()).norm() / flat.double().norm())
    # every rank must hold the same reduced gradient
    chk = got.double().norm().reshape(1).clone()
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert rel < 1e-5, (mapping, rel)
    assert float(hi - lo) <= 1e-9 * float(hi), (mapping, float(lo), float(hi))
    # a second DDP step after an optimiser update still works (parameters stay views into the flat buffer)
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4)
    opt.step()
    m.zero_grad(set_to_none=True)
    l2 = ddp(**kw).loss
    l2.backward()
    assert m._params_are_flat() and torch.isfinite(l2)
    if rank == 0:
        print("DDP-wrapped %s mapper on %d ranks: %d parameter leaves reduced, gradients equal the flat all-reduce path "
              "(rel diff %.1e)" % (mapping, world, n_params, rel))
    del ddp, m
    torch.cuda.empty_cache()
dist.destroy_process_group()
