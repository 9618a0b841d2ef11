"""Data-parallel plumbing for the mapper training step.

One process per GPU (``torchrun`` / Lightning DDP).  Samples are independent through mapper and LM, so the batch is
sharded by rank with no data-path collective; the frozen LM is replicated; the ONE exchange step is the all-reduce
of the flat mapper-gradient buffer.  The reference gets this implicitly from Lightning's DDP wrapper
(``main.py:138``): gradients are averaged over ranks, each rank's loss being the mean over its LOCAL valid tokens
(quirk Q5: mean of per-rank means).  ``all_reduce_mean_`` reproduces that; the 1/W can instead be folded into the
optimiser (``FlatAdamW.step(grad_scale=1/W)``) to save a pass over the buffer.
"""
from __future__ import annotations

from typing import Dict, Optional

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


def shard_range(n: int, rank_: int, world_: int):
    """Elements ``[begin, end)`` of a flat buffer of ``n`` floats that ``rank_`` owns in ``NvlinkShardedAdamW`` (the host-side
    mirror of ``eavqa_sharded_adamw_range``: contiguous shards of ``ceil(n / 4 / world)`` float4)."""
    if n % 4 != 0 or not (0 <= rank_ < world_):
        raise ValueError("shard_range: n must be a multiple of 4 and 0 <= rank < world")
    n4 = n // 4
    per = (n4 + world_ - 1) // world_
    return min(n4, per * rank_) * 4, min(n4, per * (rank_ + 1)) * 4


def owned_ranges(ranges, rank_: int, world_: int):
    """What ``rank_`` owns when a flat buffer is exchanged range by range (one ``eavqa_sharded_adamw_step`` call per
    ``(begin, end)`` in ``ranges``): its shard of every range.  Over all ranks these tile the union of ``ranges``."""
    out = []
    for b, e in ranges:
        sb, se = shard_range(e - b, rank_, world_)
        if se > sb:
            out.append((b + sb, b + se))
    return out


class NvlinkShardedAdamW(torch.optim.Optimizer):
    """The data-parallel exchange step and the optimiser as ONE kernel per rank (reduce-scatter + AdamW + all-gather).

    Replaces ``dist.all_reduce(flat_grads)`` + ``FlatAdamW.step`` (the reference: Lightning DDP's NCCL all-reduce,
    ``main.py:133-138``, then ``torch.optim.AdamW`` on every rank, ``clipcap_exector.py:79-81``).  The flat parameter buffer
    and the flat gradient buffer are re-homed into symmetric memory (``torch.distributed._symmetric_memory``: every rank can
    address every rank's copy, and the NVSwitch exposes one multicast address for all copies).  Rank ``r`` owns the shard
    ``shard_range(n, r, W)``: its kernel reads the sum of all ranks' gradients of that shard (``multimem.ld_reduce``: reduced
    inside the switch), updates the shard with its slice of the AdamW moments and stores the new parameters into every
    rank's buffer (``multimem.st``).  Same update as ``FlatAdamW`` on the averaged gradient (``grad_scale = 1 / W`` folded
    in); each element is computed by exactly one rank, so all replicas stay bit-identical.

    Use: ``opt = NvlinkShardedAdamW(model)`` after ``dist.init_process_group("nccl")``; then per step ``loss.backward()``,
    ``opt.step()``, ``opt.zero_grad()``.  ``p.grad`` holds the LOCAL (un-reduced) gradient in this mode.  One backward per
    step: every backward overwrites the symmetric gradient buffer, so gradient accumulation over micro-batches needs
    ``FlatAdamW`` (``accumulate()``) with an NCCL all-reduce instead.  ``timed_out()`` reports a barrier spin that gave up
    (a rank that never made the call); check it wherever a silent desynchronisation would matter.

    ``overlap=True``: the buffer is exchanged bucket by bucket (the engine's gradient buckets: pairs of mapper layers, final
    long before the mapper backward ends; ``eavqa_set_grad_events``) on a communication stream, each call on at most
    ``overlap_ctas`` CTAs -- the kernel is bound by NVLink, not by SMs -- while the backward still running keeps the other
    SMs; only the last bucket and the ranges no bucket covers are exchanged after the step.  A bucket's parameters are
    rewritten while the backward of the layers below still runs: safe, because a bucket's start barrier waits until EVERY rank
    has finished the backward of those layers, after which no rank reads them again in this step.  Use with a plain
    ``loss.backward()`` (upstream gradient 1), as for ``OverlappedGradReducer``.
    """

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01, group=None,
                 use_multicast: Optional[bool] = None, inkernel_barrier: bool = True, overlap: bool = False, overlap_ctas: int = 24):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import lib as _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NvlinkShardedAdamW needs an initialised process group (one process per GPU)")
        model._ensure_engine()
        self.model = model
        super().__init__(list(model._param_list), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = model._flat.device
        n = model._flat.numel()
        if n % 4 != 0:
            raise ValueError("the flat parameter buffer must hold a multiple of 4 floats")
        self.n = n
        self.params = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.grads = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=dev)
        self._h_params = symm_mem.rendezvous(self.params, group.group_name)
        self._h_grads = symm_mem.rendezvous(self.grads, group.group_name)
        self._h_flags = symm_mem.rendezvous(self.flags, group.group_name)
        self.flags.zero_()
        self.grads.zero_()
        model._adopt_flat(self.params, self.grads)
        dist.broadcast(self.params, src=dist.get_global_rank(group, 0), group=group)      # DDP: replicas start from rank 0's
        self.exp_avg = torch.zeros_like(self.params)      # only [begin, end) is ever touched
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.begin, self.end = shard_range(n, self.rank, self.world)
        self.steps = 0
        self._token = 0                                   # barrier generation: only ever grows (unlike `steps`, which a checkpoint may rewind)
        self.inkernel_barrier = inkernel_barrier

        def peer_array(hdl, t):
            off = t.data_ptr() - hdl.buffer_ptrs[hdl.rank]                  # the tensor's offset inside its allocation
            return (C.c_void_p * self.world)(*[C.c_void_p(p + off) for p in hdl.buffer_ptrs]), off
        self._g_ptrs, g_off = peer_array(self._h_grads, self.grads)
        self._p_ptrs, p_off = peer_array(self._h_params, self.params)
        self._f_ptrs, _ = peer_array(self._h_flags, self.flags)
        mc_g, mc_p = self._h_grads.multicast_ptr, self._h_params.multicast_ptr
        # multicast moves n * (1 + 1/W) bytes per direction and GPU (a rank's own copy also travels to the switch), peer loads /
        # stores 2 * n * (W - 1) / W: peer pointers win for W = 2 (measured: 0.277 vs 0.46 ms for 167 MB), multicast for W >= 4
        if use_multicast is None:
            use_multicast = self.world >= 4
        self.multicast = bool(use_multicast and mc_g and mc_p)
        self._mc_g = C.c_void_p(mc_g + g_off) if self.multicast else None
        self._mc_p = C.c_void_p(mc_p + p_off) if self.multicast else None
        self._lib = _lib
        self.overlap = bool(overlap and inkernel_barrier)
        self.overlap_ctas = int(overlap_ctas)
        self._events, self._buckets, self._rest = [], [], [(0, n)]
        if self.overlap:
            L = _lib.load()
            h = model._handle
            ranges = []
            for i in range(L.eavqa_grad_bucket_count(h)):
                b, e = C.c_int64(), C.c_int64()
                _lib.check(L.eavqa_grad_bucket_range(h, i, C.byref(b), C.byref(e)))
                ranges.append((b.value, e.value))
            self._buckets, self._rest = bucket_plan(n, ranges)
            if any(b % 4 or e % 4 for b, e in self._buckets + self._rest):
                raise ValueError("gradient bucket boundaries must be multiples of 4 floats")
            self._events = [torch.cuda.Event(enable_timing=False) for _ in self._buckets]
            for ev in self._events:
                ev.record()                               # materialises the cudaEvent_t the engine will re-record
            self._comm = torch.cuda.Stream(device=dev)
            if self._events:
                arr = (C.c_void_p * len(self._events))(*[C.c_void_p(ev.cuda_event) for ev in self._events])
                _lib.check(L.eavqa_set_grad_events(h, arr, len(self._events)))
        # what this rank owns: its shard of every range one call exchanges
        self.owned = owned_ranges(self._buckets + self._rest if self.overlap else [(0, n)], self.rank, self.world)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                         # every rank's flags are zero before the first kernel signals

    def zero_grad(self, set_to_none: bool = True):
        self.model.zero_grad(set_to_none=set_to_none)
        self.model.last_flat_grads = None

    def state_dict(self):
        """Full-size moments, gathered from the shards (a collective: call it on every rank)."""
        full = []
        for t in (self.exp_avg, self.exp_avg_sq):
            f = torch.zeros_like(t)
            for b, e in self.owned:
                f[b:e] = t[b:e]
            dist.all_reduce(f, group=self.group)
            full.append(f)
        return {"steps": self.steps, "exp_avg": full[0], "exp_avg_sq": full[1],
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)

    @torch.no_grad()
    def step(self, grads: torch.Tensor = None, closure=None):
        m = self.model
        g = grads if grads is not None else m.last_flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        if g.data_ptr() != self.grads.data_ptr() or m._flat is not self.params or not m._params_are_flat():
            raise RuntimeError("NvlinkShardedAdamW: the model's flat buffers are no longer the symmetric-memory ones "
                               "(the module was moved or re-flattened after the optimiser was built)")
        grp = self.param_groups[0]
        self.steps += 1
        L = self._lib.load()

        def launch(offset, count, max_ctas, stream):
            self._token = (self._token + 1) & 0xffffffff
            self._lib.check(L.eavqa_sharded_adamw_step(
                self._g_ptrs, self._p_ptrs, self._mc_g, self._mc_p, self._f_ptrs if self.inkernel_barrier else None,
                self._token, self.rank, self.world, self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), offset, count, max_ctas,
                float(grp["lr"]), grp["betas"][0], grp["betas"][1], grp["eps"], grp["weight_decay"], self.steps,
                1.0 / self.world, stream))

        with torch.cuda.device(self.params.device):
            cur = torch.cuda.current_stream(self.params.device)
            if self.overlap:
                # every bucket but the last: on the communication stream as soon as its gradients are final, on a few CTAs
                with torch.cuda.stream(self._comm):
                    for (b, e), ev in list(zip(self._buckets, self._events))[:-1]:
                        self._comm.wait_event(ev)
                        launch(b, e - b, self.overlap_ctas, self._comm.cuda_stream)
                    self._comm.wait_stream(cur)           # the backward is complete: the rest at full width
                    for (b, e) in self._buckets[-1:] + self._rest:
                        launch(b, e - b, 0, self._comm.cuda_stream)
                cur.wait_stream(self._comm)
                return
            if not self.inkernel_barrier:
                self._h_grads.barrier(channel=0)
            launch(0, self.n, 0, self._lib.current_stream())
            if not self.inkernel_barrier:
                self._h_grads.barrier(channel=0)

    def close(self):
        """Uninstall the bucket events from the engine (it only borrows them)."""
        if self._events and self.model._handle is not None:
            self._lib.check(self._lib.load().eavqa_set_grad_events(self.model._handle, None, 0))
        self._events = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def link_bytes_per_direction(self) -> int:
        """Bytes one step's exchange moves over this GPU's NVLink ports in EACH direction: through the multicast mapping the
        gradient buffer once out (a rank's own copy travels to the switch too) plus its shard of the parameters, and the
        mirror image in; through peer pointers (W - 1) / W of the buffer for the gradients plus as much for the parameters."""
        nbytes = 4 * self.n
        if self.multicast:
            return nbytes + nbytes // self.world
        return 2 * nbytes * (self.world - 1) // self.world

    def timed_out(self) -> bool:
        """True when a barrier spin inside the kernel gave up after 10 s (a peer never made the call)."""
        return bool(int(self.flags[33]))
