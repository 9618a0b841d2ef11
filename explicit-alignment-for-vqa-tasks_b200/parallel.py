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


def bucket_plan(n: int, buckets):
    """Ranges of ``[0, n)`` to all-reduce: the engine's buckets in completion order, then whatever they leave uncovered
    (ascending).  Pure host logic (tested on CPU)."""
    buckets = [(int(b), int(e)) for b, e in buckets]
    for b, e in buckets:
        if not (0 <= b < e <= n):
            raise ValueError("bucket [%d, %d) outside the gradient buffer of %d elements" % (b, e, n))
    covered = sorted(buckets)
    for (b0, e0), (b1, e1) in zip(covered, covered[1:]):
        if b1 < e0:
            raise ValueError("gradient buckets overlap")
    rest, pos = [], 0
    for b, e in covered:
        if b > pos:
            rest.append((pos, b))
        pos = e
    if pos < n:
        rest.append((pos, n))
    return buckets, rest


class OverlappedGradReducer:
    """All-reduce (sum) of the flat mapper gradient, bucket by bucket, overlapped with the mapper backward.

    The engine records one CUDA event per bucket as soon as that bucket's gradients are final (``eavqa_set_grad_events``);
    a communication stream waits on the event and runs the NCCL all-reduce of that range while the remaining layers'
    backward is still executing.  The ranges the buckets do not cover (prefix_const, the input linear) are reduced after
    the step.  ``reduce(grads)`` returns once the current stream is ordered after every all-reduce.

    Use with a plain ``loss.backward()`` (upstream gradient 1): a loss scale would be applied to the buffer after the
    buckets may already be in flight -- fold such factors into ``FlatAdamW.step(grad_scale=...)`` instead.
    """

    def __init__(self, model):
        from . import lib as _lib
        import ctypes as C
        model._ensure_engine()
        self.model = model
        L = _lib.load()
        h = model._handle
        n = L.eavqa_grad_bucket_count(h)
        ranges = []
        for i in range(n):
            b, e = C.c_int64(), C.c_int64()
            _lib.check(L.eavqa_grad_bucket_range(h, i, C.byref(b), C.byref(e)))
            ranges.append((b.value, e.value))
        self.buckets, self.rest = bucket_plan(model._flat.numel(), ranges)
        self.events = [torch.cuda.Event(enable_timing=False) for _ in self.buckets]
        for ev in self.events:
            ev.record()                                    # materialises the cudaEvent_t the engine will re-record
        self.comm_stream = torch.cuda.Stream(device=model._flat.device)
        if self.events:
            arr = (C.c_void_p * len(self.events))(*[C.c_void_p(ev.cuda_event) for ev in self.events])
            _lib.check(L.eavqa_set_grad_events(h, arr, len(self.events)))

    def reduce(self, flat_grads: torch.Tensor) -> torch.Tensor:
        if not (dist.is_available() and dist.is_initialized()):
            return flat_grads
        cur = torch.cuda.current_stream(flat_grads.device)
        works = []
        with torch.cuda.stream(self.comm_stream):
            for (b, e), ev in zip(self.buckets, self.events):
                self.comm_stream.wait_event(ev)
                works.append(dist.all_reduce(flat_grads[b:e], op=dist.ReduceOp.SUM, async_op=True))
        for (b, e) in self.rest:                           # final when the step's own stream gets here
            works.append(dist.all_reduce(flat_grads[b:e], op=dist.ReduceOp.SUM, async_op=True))
        for w in works:
            w.wait()                                       # orders the current stream after the collective
        flat_grads.record_stream(self.comm_stream)
        cur.wait_stream(self.comm_stream)
        return flat_grads

    def close(self):
        """Uninstall the events from the engine (it only borrows them: they must outlive their installation)."""
        from . import lib as _lib
        if self.events and self.model._handle is not None:
            _lib.check(_lib.load().eavqa_set_grad_events(self.model._handle, None, 0))
        self.events = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
