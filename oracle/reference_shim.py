"""Import the UNMODIFIED reference model files behind the shim of SURVEY.md Appendix B (TEST INFRASTRUCTURE).

``import_reference(src_dir)`` loads ``clipcap.py`` (and ``vct0.py`` when present) from ``src_dir`` -- the reference checkout
(``/root/reference/src/models``, dev container only) or ``oracle/_ref`` (the byte-identical copies placed there by
``oracle/install_reference.py``; git-ignored, shipped to the GPU box with the snapshot) -- after

* aliasing ``transformers.AdamW`` (``clipcap.py:10`` imports a symbol that newer ``transformers`` dropped and never uses it);
* pointing ``GPT2LMHeadModel.from_pretrained`` at a random-init ``GPT2Config`` (no checkpoints exist offline);
* stubbing ``flamingo_pytorch`` (``vct0.py:17`` import only).

Nothing of the reference is modified or copied into tracked files.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF_MODELS = os.path.join(os.environ.get("EAVQA_REFERENCE", "/root/reference"), "src", "models")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(ROOT, "oracle", "_ref")


def reference_dir():
    """Where the reference's model files can be imported from on this box (None: nowhere)."""
    for d in (REF_MODELS, INSTALLED):
        if os.path.exists(os.path.join(d, "clipcap.py")):
            return d
    return None


def import_reference(src_dir=None):
    src_dir = src_dir or reference_dir()
    if src_dir is None:
        raise RuntimeError("the reference's clipcap.py is neither under /root/reference nor under oracle/_ref")
    from transformers import GPT2Config, GPT2LMHeadModel    # heavy import first (the lazy module object is replaced)
    sys.modules["transformers"].AdamW = torch.optim.AdamW
    holder = {}

    def fake_from_pretrained(cls, name, *a, **k):
        return GPT2LMHeadModel(holder["cfg"])
    GPT2LMHeadModel.from_pretrained = classmethod(fake_from_pretrained)
    fl = types.ModuleType("flamingo_pytorch")
    fl.PerceiverResampler = object
    sys.modules.setdefault("flamingo_pytorch", fl)
    if src_dir not in sys.path:
        sys.path.insert(0, src_dir)
    import clipcap
    vct0 = None
    if os.path.exists(os.path.join(src_dir, "vct0.py")):
        import vct0
    return clipcap, vct0, GPT2Config, holder


def build_reference_model(clipcap, GPT2Config, holder, lm_cfg, lm_w, mapper_w, *, prefix_length, clip_length, clip_dim,
                          num_layers, mapping_type):
    """``ClipCaptionPrefix`` (clipcap.py:590-599) holding the given synthetic LM and mapper weights."""
    holder["cfg"] = GPT2Config(vocab_size=lm_cfg["vocab"], n_positions=lm_cfg["n_positions"], n_embd=lm_cfg["d_model"],
                               n_layer=lm_cfg["n_layer"], n_head=lm_cfg["n_head"])
    m = clipcap.ClipCaptionPrefix(prefix_length=prefix_length, clip_length=clip_length, prefix_size=clip_dim,
                                  num_layers=num_layers, mapping_type=mapping_type, model_version="synthetic")
    sd = dict(lm_w)
    sd["lm_head.weight"] = lm_w["transformer.wte.weight"]
    missing, unexpected = m.gpt.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(("attn.bias" in k or "masked_bias" in k) for k in missing), missing
    m.clip_project.load_state_dict(mapper_w, strict=True)
    assert [n for n, _ in m.clip_project.named_parameters()] == list(mapper_w.keys()), "flat layout order"
    return m.train()
