"""Fused AdamW over the mapper's flat parameter buffer (``eavqa_adamw_step``).

Same update as ``torch.optim.AdamW(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)``, which is what
``ClipCapExecutor.configure_optimizers`` builds (``clipcap_exector.py:79-81``); one kernel, one pass over
params / grads / moments.  ``grad_scale`` folds the data-parallel 1/world_size in.
"""
from __future__ import annotations

import torch

from . import lib as _lib


class FlatAdamW:
    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01):
        model._ensure_engine()
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(model._flat)
        self.exp_avg_sq = torch.zeros_like(model._flat)
        self.steps = 0

    def zero_grad(self, set_to_none: bool = True):
        self.model.zero_grad(set_to_none=set_to_none)
        self.model.last_flat_grads = None

    @torch.no_grad()
    def step(self, grads: torch.Tensor = None, grad_scale: float = 1.0):
        m = self.model
        g = grads if grads is not None else m.last_flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        if not m._params_are_flat():
            raise RuntimeError("mapper parameters were moved; call the model once to re-flatten before stepping")
        self.steps += 1
        with torch.cuda.device(m._flat.device):
            _lib.check(_lib.load().eavqa_adamw_step(m._flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                                    self.exp_avg_sq.data_ptr(), m._flat.numel(), self.lr, self.betas[0],
                                                    self.betas[1], self.eps, self.weight_decay, self.steps, grad_scale,
                                                    _lib.current_stream()))
