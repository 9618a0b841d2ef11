"""``ClipCaptionPrefixB200``: host-side mirror of the reference's ``ClipCaptionPrefix``.

Same constructor kwargs, same ``forward`` / ``generate`` surface, same ``clip_project.*`` parameter names and
shapes as ``src/models/clipcap.py:240-471,590-599``, so ``ClipCapExecutor`` (``clipcap_exector.py:52-56,165-171,
236-243``) can build it by name (``ModelClass``) and AdamW / DDP / checkpoints keep working.  All arithmetic runs
in the C-ABI CUDA library (``lib.py`` -> ``libeavqa_b200.so``); the ``nn.Module`` below only owns the trainable
mapper parameters (as views into one flat fp32 buffer) and an opaque engine handle.  There is no PyTorch
fallback: without the library or without a B200 every call raises.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import types
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import lib as _lib
from . import synthetic

logger = logging.getLogger(__name__)


# ------------------------------------------------------------------------------------------------
# parameter containers: same module tree / registration order as clipcap.py so that names, shapes and
# default initialisation (under the same torch seed) are the reference's.  They are never called.
# ------------------------------------------------------------------------------------------------
class _MLPParams(nn.Module):                                   # clipcap.py:31-42
    def __init__(self, sizes):
        super().__init__()
        layers = []
        for i in range(len(sizes) - 1):
            layers.append(nn.Linear(sizes[i], sizes[i + 1], bias=True))
            if i < len(sizes) - 2:
                layers.append(nn.Tanh())
        self.model = nn.Sequential(*layers)


class _AttnParams(nn.Module):                                  # clipcap.py:70-79 (bias=False from TransformerLayer)
    def __init__(self, d):
        super().__init__()
        self.to_queries = nn.Linear(d, d, bias=False)
        self.to_keys_values = nn.Linear(d, 2 * d, bias=False)
        self.project = nn.Linear(d, d)


class _MlpTransformerParams(nn.Module):                        # clipcap.py:45-59, mlp_ratio = 2.0 (clipcap.py:165)
    def __init__(self, d):
        super().__init__()
        self.fc1 = nn.Linear(d, 2 * d)
        self.fc2 = nn.Linear(2 * d, d)


class _LayerParams(nn.Module):                                 # clipcap.py:119-138
    def __init__(self, d):
        super().__init__()
        self.norm1 = nn.LayerNorm(d)
        self.attn = _AttnParams(d)
        self.norm2 = nn.LayerNorm(d)
        self.mlp = _MlpTransformerParams(d)


class _TransformerParams(nn.Module):                           # clipcap.py:159-210
    def __init__(self, d, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_LayerParams(d) for _ in range(num_layers)])


class _TransformerMapperParams(nn.Module):                     # clipcap.py:223-237
    def __init__(self, dim_clip, d, prefix_length, clip_length, num_layers):
        super().__init__()
        self.clip_length = clip_length
        self.transformer = _TransformerParams(d, num_layers)
        self.linear = nn.Linear(dim_clip, clip_length * d)
        self.prefix_const = nn.Parameter(torch.randn(prefix_length, d), requires_grad=True)


class _StepOutput:
    """What ``ClipCaptionModel.forward`` returns to the executor: ``.loss`` (0-d, differentiable w.r.t. the
    mapper) and ``.logits`` -- ``None`` unless ``forward(..., return_logits=True)`` asked for them (the step never
    materialises [B, T, V] fp32: 2.6 GB at B=256; the executors only read ``.loss``)."""

    def __init__(self, loss, logits=None):
        self.loss = loss
        self.logits = logits

    def __getitem__(self, i):
        return (self.loss, self.logits)[i]


class _TrainStepFn(torch.autograd.Function):
    """One fused forward+backward C call; autograd only routes the mapper gradients."""

    @staticmethod
    def forward(ctx, model, need_grad, clip, tokens, mask, labels, *params):
        loss, grads = model._run_step(clip, tokens, mask, labels, need_grad)
        ctx.model = model
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, g):
        grads, model = ctx.grads, ctx.model
        if getattr(ctx, "consumed", False):
            # the flat gradient buffer was handed over by the first backward (and possibly rescaled in place since)
            raise RuntimeError("ClipCaptionPrefixB200: backward through the same step a second time; run forward again "
                               "(the reference would raise 'Trying to backward through the graph a second time' too)")
        ctx.consumed = True
        ctx.grads = None
        if grads is None:
            return (None,) * (6 + len(model._param_list))
        # grads *= g on the device; the kernel returns at once when g == 1 (plain loss.backward()): no host sync, and no
        # extra pass over the 167 MB buffer
        with torch.cuda.device(grads.device):
            _lib.check(_lib.load().eavqa_scale_grads(grads.data_ptr(), grads.numel(),
                                                     g.to(device=grads.device, dtype=torch.float32).contiguous().data_ptr(),
                                                     _lib.current_stream()))
        model.last_flat_grads = grads
        outs = [grads[o:o + n].view(shape) for (o, n, shape) in model._slices]
        return (None, None, None, None, None, None, *outs)


class _GPTShim:
    """The ``.gpt`` attribute the executor touches: ``resize_token_embeddings`` (clipcap_exector.py:56) and
    ``config``.  The LM itself lives, packed, inside the engine."""

    def __init__(self, owner):
        self._owner = owner
        cfg = owner._lm_cfg
        self.config = types.SimpleNamespace(vocab_size=cfg["vocab"], n_positions=cfg["n_positions"], n_embd=cfg["d_model"],
                                            n_layer=cfg["n_layer"], n_head=cfg["n_head"], pad_token_id=None,
                                            eos_token_id=50256, bos_token_id=50256)

    def resize_token_embeddings(self, new_num_tokens: Optional[int] = None):
        if new_num_tokens is not None and new_num_tokens != self.config.vocab_size:
            self._owner._resize_vocab(int(new_num_tokens))
            self.config.vocab_size = int(new_num_tokens)
        return self

    def parameters(self):
        return iter(())

    def eval(self):
        return self


class ClipCaptionModelB200(nn.Module):
    def __init__(self, prefix_length: int, clip_length: Optional[int] = None, prefix_size: int = 512, num_layers: int = 8,
                 mapping_type: str = "mlp", model_version: str = "gpt2", lm_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 lm_config: Optional[dict] = None, special_token_id: Optional[int] = None):
        super().__init__()
        self.prefix_length = prefix_length
        self.clip_length = clip_length
        self.prefix_size = prefix_size
        self.num_layers = num_layers
        self.mapping_type = "mlp" if mapping_type == "mlp" else "transformer"     # clipcap.py:254-271
        self.special_token_id = special_token_id
        self._lm_cfg, self._lm_weights = self._resolve_lm(model_version, lm_state_dict, lm_config)
        self.gpt_embedding_size = self._lm_cfg["d_model"]                          # clipcap.py:253
        d = self.gpt_embedding_size
        if self.mapping_type == "mlp":
            self.clip_project = _MLPParams((prefix_size, (d * prefix_length) // 2, d * prefix_length))
        else:
            if clip_length is None:
                raise ValueError("the transformer mapper needs clip_length")
            self.clip_project = _TransformerMapperParams(prefix_size, d, prefix_length, clip_length, num_layers)
        expected = synthetic.mapper_param_shapes(self.mapping_type, prefix_size, d, prefix_length, clip_length or 0, num_layers)
        got = OrderedDict((n, tuple(p.shape)) for n, p in self.clip_project.named_parameters())
        assert list(got.items()) == list(expected.items()), "mapper parameter layout drifted from the reference's"
        self.gpt = _GPTShim(self)
        self._handle = None
        self._flat = None
        self._slices: List[tuple] = []
        self._param_list: List[nn.Parameter] = []
        self.last_flat_grads = None
        self._grad_buffer = None        # persistent flat gradient buffer (symmetric memory), see _adopt_flat
        self.save_lm = False           # state_dict() also emits the frozen LM (reference checkpoint layout) when True

    # ------------------------------------------------------------------ LM weights
    @staticmethod
    def _resolve_lm(model_version, lm_state_dict, lm_config):
        if lm_state_dict is not None:
            sd = {k[4:] if k.startswith("gpt.") else k: v for k, v in lm_state_dict.items()}
            wte, wpe = sd["transformer.wte.weight"], sd["transformer.wpe.weight"]
            n_layer = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.h."))
            cfg = dict(n_layer=n_layer, d_model=wte.shape[1], n_head=wte.shape[1] // 64, vocab=wte.shape[0], n_positions=wpe.shape[0])
            if lm_config:
                cfg.update(lm_config)
            return cfg, OrderedDict((k, v.detach().float().cpu()) for k, v in sd.items())
        # Synthetic (seeded random-init) weights are an explicit opt-in: the test-only shapes "gpt2-tiny" / "gpt2-mini",
        # a "synthetic:<name>" model_version, or EAVQA_SYNTHETIC_LM=1 in the environment (offline boxes).  Anything else is
        # a real checkpoint name and resolves exactly like the reference's GPT2LMHeadModel.from_pretrained (clipcap.py:252):
        # local cache or download, and an error when neither works -- never a silent random LM.
        name = model_version or ""
        synthetic_name = name[len("synthetic:"):] if name.startswith("synthetic:") else None
        if synthetic_name is None and (name in ("gpt2-tiny", "gpt2-mini") or os.environ.get("EAVQA_SYNTHETIC_LM") == "1"):
            synthetic_name = name
        if synthetic_name is not None:
            cfg = synthetic.lm_config(synthetic_name) if lm_config is None else dict(lm_config)
            logger.warning("using seeded SYNTHETIC GPT-2 weights for %r (%s)", model_version, cfg)
            return cfg, synthetic.make_lm_weights(cfg, seed=0)
        from transformers import GPT2LMHeadModel
        try:
            hf = GPT2LMHeadModel.from_pretrained(model_version)
        except OSError as e:        # no cached checkpoint and no way to fetch one
            raise OSError("cannot load the GPT-2 checkpoint %r (%s).  Pass lm_state_dict=..., or opt in to seeded synthetic "
                          "weights with model_version='synthetic:%s' / EAVQA_SYNTHETIC_LM=1" % (model_version, e, model_version)) from e
        return ClipCaptionModelB200._resolve_lm(model_version, hf.state_dict(), None)

    def _resize_vocab(self, n: int):
        wte = self._lm_weights["transformer.wte.weight"]
        old = wte.shape[0]
        if n > old:     # HF >= 4.46 default: new rows start at the mean of the old embeddings
            extra = wte.mean(dim=0, keepdim=True).expand(n - old, -1)
            wte = torch.cat([wte, extra], dim=0)
        else:
            wte = wte[:n]
        self._lm_weights["transformer.wte.weight"] = wte.contiguous()
        self._lm_cfg["vocab"] = n
        self._destroy_engine()

    def load_lm_state_dict(self, sd: Dict[str, torch.Tensor]):
        cfg, w = self._resolve_lm(None, sd, None)
        self._lm_cfg.update(cfg)
        self._lm_weights = w
        self.gpt.config.vocab_size = cfg["vocab"]
        self._destroy_engine()

    # ------------------------------------------------------------------ engine / flat parameters
    def _destroy_engine(self):
        if self._handle is not None:
            _lib.check(_lib.load().eavqa_destroy(self._handle))
            self._handle = None

    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:
            pass

    def _flatten(self):
        """Re-home every mapper parameter as a view into one flat fp32 device buffer (C-ABI layout)."""
        params = list(self.clip_project.named_parameters())
        dev = params[0][1].device
        total = sum(p.numel() for _, p in params)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        self._slices, self._param_list = [], []
        off = 0
        for _, p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            self._slices.append((off, n, tuple(p.shape)))
            self._param_list.append(p)
            off += n
        self._flat = flat
        self._grad_buffer = None        # a persistent gradient buffer belongs to the flat buffer it was adopted with

    def _adopt_flat(self, new_flat: torch.Tensor, grad_buffer: Optional[torch.Tensor] = None):
        """Move the flat parameter buffer into caller-provided storage of the same size (symmetric memory that peers can
        address: ``parallel.NvlinkShardedAdamW``) and, optionally, make every backward write its flat gradient into
        ``grad_buffer`` instead of a fresh allocation."""
        self._ensure_engine()
        if new_flat.shape != self._flat.shape or new_flat.dtype != torch.float32 or new_flat.device != self._flat.device:
            raise ValueError("the new flat buffer must match the mapper's: %s fp32 on %s" % (tuple(self._flat.shape), self._flat.device))
        new_flat.copy_(self._flat)
        for p_, (o, n, shape) in zip(self._param_list, self._slices):
            p_.data = new_flat[o:o + n].view(shape)
        self._flat = new_flat
        if grad_buffer is not None and (grad_buffer.shape != new_flat.shape or grad_buffer.dtype != torch.float32):
            raise ValueError("the gradient buffer must have the flat parameter buffer's shape and dtype")
        self._grad_buffer = grad_buffer

    def _params_are_flat(self) -> bool:
        if self._flat is None:
            return False
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * o and p.device == self._flat.device
                   for p, (o, _, _) in zip(self._param_list, self._slices))

    def _ensure_engine(self):
        dev = next(self.clip_project.parameters()).device
        if dev.type != "cuda":
            raise _lib.EavqaError("ClipCaptionPrefixB200 has no CPU path: move the module to a B200 (`.cuda()`) first")
        if not self._params_are_flat():
            self._flatten()
        if self._handle is not None:
            return
        L = _lib.load()
        c = self._lm_cfg
        cfg = _lib.EavqaConfig(n_layer=c["n_layer"], n_head=c["n_head"], d_model=c["d_model"], vocab=c["vocab"],
                               n_positions=c["n_positions"], prefix_length=self.prefix_length,
                               clip_length=self.clip_length or 0, clip_dim=self.prefix_size,
                               mapper_type=_lib.MAPPER_MLP if self.mapping_type == "mlp" else _lib.MAPPER_TRANSFORMER,
                               mapper_layers=self.num_layers)
        with torch.cuda.device(dev):
            h = C.c_void_p()
            _lib.check(L.eavqa_create(C.byref(cfg), C.byref(h)))
            self._handle = h
            stream = _lib.current_stream()
            for name, w in self._lm_weights.items():
                if name == "lm_head.weight" or name.endswith(".attn.bias") or name.endswith(".attn.masked_bias"):
                    continue
                t = w.to(device=dev, dtype=torch.float32).contiguous()
                _lib.check(L.eavqa_load_lm_weight(h, name.encode(), t.data_ptr(), _lib.F32, t.numel(), stream))
                torch.cuda.current_stream().synchronize()      # `t` is a temporary
            _lib.check(L.eavqa_finalize_lm(h, stream))
            # the C side and this module must agree on the flat layout
            assert L.eavqa_mapper_param_count(h) == self._flat.numel()
            names = [n for n, _ in self.clip_project.named_parameters()]
            buf = C.create_string_buffer(256)
            off, rows, cols = C.c_int64(), C.c_int64(), C.c_int64()
            assert L.eavqa_mapper_num_tensors(h) == len(names)
            for i, (n, (o, cnt, _)) in enumerate(zip(names, self._slices)):
                _lib.check(L.eavqa_mapper_tensor_info(h, i, buf, 256, C.byref(off), C.byref(rows), C.byref(cols)))
                assert buf.value.decode() == n and off.value == o and rows.value * cols.value == cnt, (n, buf.value)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat = None          # parameters were re-created: re-flatten lazily
        return out

    # ------------------------------------------------------------------ reference surface
    def get_dummy_token(self, batch_size: int, num_question_tokens: int, device: torch.device) -> torch.Tensor:
        return torch.ones(batch_size, self.prefix_length + num_question_tokens, dtype=torch.int64, device=device) * -100

    def _prep(self, t, dtype):
        dev = self._flat.device
        return t.to(device=dev, dtype=dtype, non_blocking=True).contiguous()

    def _run_step(self, clip, tokens, mask, labels, need_grad):
        L = _lib.load()
        B, Tt = tokens.shape
        loss = torch.empty((), dtype=torch.float32, device=self._flat.device)
        grads = None
        if need_grad:
            grads = self._grad_buffer if self._grad_buffer is not None else torch.empty_like(self._flat)
        with torch.cuda.device(self._flat.device):
            _lib.check(L.eavqa_train_step(self._handle, B, Tt, clip.data_ptr(), tokens.data_ptr(), _lib.ptr(mask),
                                          labels.data_ptr(), self._flat.data_ptr(), _lib.ptr(grads), loss.data_ptr(),
                                          _lib.current_stream()))
        return loss, grads

    def forward(self, question_tokens: torch.Tensor, prefix: torch.Tensor, question_mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, pad_token_id=None, return_logits: bool = False):
        """clipcap.py:290-342.  ``labels`` are the un-shifted text labels (-100 = ignore); the shift happens inside.
        ``return_logits=True`` additionally fills ``.logits`` ([B, T, V] fp32, detached; one more forward pass through
        ``eavqa_forward_logits`` -- meant for small batches); without ``labels`` only the logits are computed and
        ``.loss`` is ``None``, as for the reference's HF output."""
        self._ensure_engine()
        tokens = self._prep(question_tokens, torch.int64)
        clip = self._prep(prefix, torch.float32).reshape(tokens.shape[0], -1)
        if clip.shape[1] != self.prefix_size:
            raise ValueError("prefix must hold one %d-d CLIP embedding per sample" % self.prefix_size)
        mask = self._prep(question_mask, torch.int64) if question_mask is not None else None
        if labels is None and not return_logits:
            raise ValueError("ClipCaptionPrefixB200.forward computes the caption loss and needs `labels` "
                             "(or return_logits=True for the logits alone)")
        logits = None
        if return_logits:
            B, Tt = tokens.shape
            V = self._lm_cfg["vocab"]
            ld = (V + 63) // 64 * 64
            buf = torch.empty(B, self.prefix_length + Tt, ld, dtype=torch.float32, device=self._flat.device)
            with torch.cuda.device(self._flat.device):
                _lib.check(_lib.load().eavqa_forward_logits(self._handle, B, Tt, clip.data_ptr(), tokens.data_ptr(), _lib.ptr(mask),
                                                            self._flat.data_ptr(), buf.data_ptr(), ld, _lib.current_stream()))
            logits = buf[:, :, :V]
        if labels is None:
            return _StepOutput(None, logits)
        labels = self._prep(labels, torch.int64)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        loss = _TrainStepFn.apply(self, need_grad, clip, tokens, mask, labels, *self._param_list)
        return _StepOutput(loss, logits)

    @torch.no_grad()
    def generate(self, question_tokens: torch.Tensor, prefix: torch.Tensor, question_mask: Optional[torch.Tensor] = None,
                 max_length: Optional[int] = 10, pad_token_id: Optional[int] = None, eos_token_id: Optional[int] = None,
                 special_token_id: Optional[int] = None, return_top_logits: bool = False, return_logprobs: bool = False,
                 **unused):
        """clipcap.py:344-471 (prepend; returns ``list[list[int]]``) and, when ``prefix`` is ``[B, k+1, 1, D]`` and a
        sentinel id is known, the k-shot assembly of vct0.py:446-464,494-533 (returns a ``LongTensor [B, steps]`` like
        ``lm.generate``).  Extra few-shot kwargs of ``FewShotVQAExecutor`` (``decoder_input_ids=None`` ...) are accepted.
        ``return_logprobs=True`` returns ``(tokens [B, steps] LongTensor, logprob [B, steps])`` on the device: the
        log-softmax of every picked token, i.e. what ``generate_from_ensembles`` gathers from ``outputs.scores``
        (few_shot_vqa_executor.py:316-323)."""
        self._ensure_engine()
        L = _lib.load()
        tokens = self._prep(question_tokens, torch.int64)
        B, Tt = tokens.shape
        mask = self._prep(question_mask, torch.int64) if question_mask is not None else None
        pad = pad_token_id if pad_token_id is not None else self.gpt.config.pad_token_id         # clipcap.py:398-407
        eos = eos_token_id if eos_token_id is not None else self.gpt.config.eos_token_id
        if eos is not None and pad is None:
            raise ValueError("If `eos_token_id` is defined, make sure that `pad_token_id` is defined.")   # clipcap.py:427-430
        sentinel = special_token_id if special_token_id is not None else self.special_token_id
        few_shot = prefix.dim() >= 3 and sentinel is not None
        clip = self._prep(prefix, torch.float32)
        if few_shot:
            n_img = clip.shape[1]
            clip = clip.reshape(B, n_img, self.prefix_size)
            lo, hi = sentinel - (n_img - 1), sentinel
        else:
            n_img, lo, hi = 0, 0, -1
            clip = clip.reshape(B, self.prefix_size)
        dev = self._flat.device
        out = torch.empty(B, max_length, dtype=torch.int64, device=dev)
        top = torch.empty(B, max_length, dtype=torch.float32, device=dev) if return_top_logits else None
        lp = torch.empty(B, max_length, dtype=torch.float32, device=dev) if return_logprobs else None
        steps = C.c_int32(0)
        with torch.cuda.device(dev):
            _lib.check(L.eavqa_generate(self._handle, B, Tt, n_img, clip.data_ptr(), tokens.data_ptr(), _lib.ptr(mask), lo, hi,
                                        self._flat.data_ptr(), max_length, 1 if eos is not None else 0,
                                        pad if pad is not None else 0, eos if eos is not None else 0, out.data_ptr(),
                                        _lib.ptr(top), _lib.ptr(lp), C.byref(steps), _lib.current_stream()))
        out = out[:, :steps.value]
        if return_logprobs:
            return out, lp[:, :steps.value]
        if return_top_logits:
            return out.cpu().tolist(), top[:, :steps.value].cpu()
        if few_shot:
            return out
        return out.cpu().numpy().astype(int).tolist()                                              # clipcap.py:469

    # ------------------------------------------------------------------ checkpoints
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """Reference checkpoints also carry the frozen LM (``gpt.transformer.*``, ``gpt.lm_head.weight``).
        ``self.gpt`` is not an ``nn.Module`` here, so those keys are taken out of the incoming dict and re-packed into the
        engine.  This is the hook ``nn.Module.load_state_dict`` calls on every sub-module, so it also works when the
        model is nested (Lightning loads a checkpoint through the executor: keys ``model.gpt.*``) and with ``strict=True``."""
        lm_prefix = prefix + "gpt."
        lm = {k[len(prefix):]: state_dict.pop(k) for k in [k for k in state_dict if k.startswith(lm_prefix)]}
        if lm:
            self.load_lm_state_dict(lm)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def state_dict(self, *args, include_lm: Optional[bool] = None, **kwargs):
        """``clip_project.*`` with the reference's names and shapes.  ``include_lm=True`` (or ``self.save_lm = True``, for
        callers such as Lightning that cannot pass the argument) adds the frozen LM under the reference's ``gpt.*`` keys,
        so the result strict-loads into the reference's ``ClipCaptionPrefix``."""
        sd = super().state_dict(*args, **kwargs)
        if include_lm if include_lm is not None else self.save_lm:
            prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
            for k, v in self._lm_weights.items():
                if k != "lm_head.weight" and not k.endswith(".attn.bias") and not k.endswith(".attn.masked_bias"):
                    sd[prefix + "gpt." + k] = v
            sd[prefix + "gpt.lm_head.weight"] = self._lm_weights["transformer.wte.weight"]      # tied head
        return sd


class ClipCaptionPrefixB200(ClipCaptionModelB200):
    """clipcap.py:590-599: only ``clip_project`` trains; the LM is frozen and in eval mode (it has no dropout here)."""

    def parameters(self, recurse: bool = True):
        return self.clip_project.parameters()

    def train(self, mode: bool = True):
        super().train(mode)
        return self
