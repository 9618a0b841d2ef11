"""Data-parallel plumbing for the mapper training step.

One process per GPU (``torchrun`` / Lightning DDP).  Samples are independent through mapper and LM, so the batch is
sharded by rank with no data-path collective; the frozen LM is replicated; the ONE exchange step is the all-reduce
of the flat mapper-gradient buffer.  The reference gets this implicitly from Lightning's DDP wrapper
(``main.py:138``): gradients are averaged over ranks, each rank's loss being the mean over its LOCAL valid tokens
(quirk Q5: mean of per-rank means).  ``all_reduce_mean_`` reproduces that; the 1/W can instead be folded into the
optimiser (``FlatAdamW.step(grad_scale=1/W)``) to save a pass over the buffer.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_batch(batch: Dict[str, torch.Tensor], rank_: int, world_: int) -> Dict[str, torch.Tensor]:
    """Rank r takes rows [r*B/W, (r+1)*B/W) (DistributedSampler-like contiguous split of a global batch)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        if n % world_ != 0:
            raise ValueError("global batch %d is not divisible by world size %d" % (n, world_))
        per = n // world_
        out[k] = v[rank_ * per:(rank_ + 1) * per]
    return out


def all_reduce_sum_(flat_grads: torch.Tensor) -> torch.Tensor:
    """Sum the flat mapper-gradient buffer over ranks in place (one NCCL call over NVLink; no-op for one rank)."""
    if world() > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def all_reduce_mean_(flat_grads: torch.Tensor) -> torch.Tensor:
    """DDP semantics: average over ranks (mean of per-rank token-mean gradients)."""
    w = world()
    if w > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
        flat_grads.mul_(1.0 / w)
    return flat_grads
