"""Fused AdamW over the mapper's flat parameter buffer (``eavqa_adamw_step``).

Same update as ``torch.optim.AdamW(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)``, which is what
``ClipCapExecutor.configure_optimizers`` builds (``clipcap_exector.py:79-81``); one kernel, one pass over
params / grads / moments.  ``grad_scale`` folds the data-parallel 1/world_size in.
"""
from __future__ import annotations

import torch

from . import lib as _lib


class FlatAdamW(torch.optim.Optimizer):
    """A real ``torch.optim.Optimizer`` (torch's and transformers' schedulers insist on one): ``param_groups`` (one group;
    ``lr`` is read from it at every step) and ``state_dict`` / ``load_state_dict`` make it drivable by the schedulers ``ClipCapExecutor.configure_optimizers`` attaches (``get_constant_schedule_with_warmup``,
    ``get_linear_schedule_with_warmup``, ``CosineAnnealingLR``: ``clipcap_exector.py:83-109``); they only read and write
    ``optimizer.param_groups[i]["lr"]`` / ``["initial_lr"]``."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01):
        model._ensure_engine()
        self.model = model
        super().__init__(list(model._param_list), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.exp_avg = torch.zeros_like(model._flat)
        self.exp_avg_sq = torch.zeros_like(model._flat)
        self.steps = 0
        self._accum = None              # flat gradient accumulated over micro-batches (accumulate_grad_batches > 1)
        self._accum_src = None

    # attributes kept for callers of the first version
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = v

    def zero_grad(self, set_to_none: bool = True):
        self.model.zero_grad(set_to_none=set_to_none)
        self.model.last_flat_grads = None
        self._accum = None
        self._accum_src = None

    def accumulate(self):
        """Call after every ``loss.backward()`` of a micro-batch when more than one precedes ``step()``: the engine
        hands back only the LATEST backward's flat gradient (``model.last_flat_grads``), so the sum is kept here."""
        g = self.model.last_flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        if g is self._accum_src:
            return                      # already counted
        self._accum_src = g
        self._accum = g.clone() if self._accum is None else self._accum.add_(g)

    def state_dict(self):
        return {"steps": self.steps, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)

    @torch.no_grad()
    def step(self, grads: torch.Tensor = None, grad_scale: float = 1.0, closure=None):
        m = self.model
        if grads is None and self._accum is not None:
            self.accumulate()           # the last micro-batch, if the caller did not add it
            grads = self._accum
        g = grads if grads is not None else m.last_flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        if not m._params_are_flat():
            raise RuntimeError("mapper parameters were moved; call the model once to re-flatten before stepping")
        grp = self.param_groups[0]
        self.steps += 1
        with torch.cuda.device(m._flat.device):
            _lib.check(_lib.load().eavqa_adamw_step(m._flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                                    self.exp_avg_sq.data_ptr(), m._flat.numel(), float(grp["lr"]), grp["betas"][0],
                                                    grp["betas"][1], grp["eps"], grp["weight_decay"], self.steps, grad_scale,
                                                    _lib.current_stream()))
