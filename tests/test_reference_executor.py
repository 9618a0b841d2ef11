"""The drop-in boundary exercised from the REFERENCE's side.

``ClipCapExecutor`` (``src/trainers/clipcap_exector.py``) is imported UNMODIFIED -- from ``/root/reference`` in the dev
container, from the byte-identical copy under ``oracle/_ref/trainers`` (``oracle/install_reference.py``) on the GPU box --
with stand-ins only for packages that are not installed (``pytorch_lightning`` 1.6.3, ``easydict``) and for reference
modules that are not on the path (``utils.*`` helpers, ``models.clipcap``).  ``ClipCaptionPrefixB200`` is then put into the
executor module's globals, exactly what the one-line import of INTEGRATION.md section 1 does, and the executor's own
``__init__`` / ``configure_optimizers`` / ``training_step`` / ``_generative_step`` run against it:

* ``ModelClass = globals()[config.model_config.ModelClass]``; ``ModelClass(**model_args)``     (clipcap_exector.py:52-53)
* ``self.model.gpt.resize_token_embeddings(len(tokenizer))``                                    (:56)
* ``torch.optim.AdamW([{"params": [p for n, p in self.model.named_parameters()], ...}])``     (:62-81)
* the Python label loop, ``self.model(question_tokens=, labels=, prefix=, question_mask=, pad_token_id=).loss``  (:134-173)
* ``self.model.generate(...)`` -> rows with ``.index(bos)`` / ``in``                          (:236-264)

The loss the executor gets back is checked against the oracle on the labels the executor itself built.
"""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TRAINERS = [os.path.join(os.environ.get("EAVQA_REFERENCE", "/root/reference"), "src", "trainers"),
                os.path.join(ROOT, "oracle", "_ref", "trainers")]


def _trainers_dir():
    for d in REF_TRAINERS:
        if os.path.exists(os.path.join(d, "clipcap_exector.py")):
            return d
    return None


class _EasyDict(dict):
    """Minimal ``easydict.EasyDict``: attribute access, nested dicts converted."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _install_stand_ins():
    """Stand-ins for what is not installed / not on the path.  Returns the names added to sys.modules."""
    added = []

    def add(name, mod):
        if name not in sys.modules:
            sys.modules[name] = mod
            added.append(name)
        return sys.modules[name]

    ed = types.ModuleType("easydict")
    ed.EasyDict = _EasyDict
    add("easydict", ed)

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):          # the few members the executor touches
        def __init__(self):
            super().__init__()
            self.logged = {}
            self.global_step = -1
            self.trainer = types.SimpleNamespace(loggers=[], estimated_stepping_batches=100)

        @property
        def device(self):
            return next(self.parameters()).device

        def log(self, name, value, **kw):
            self.logged[name] = value

    pl.LightningModule = LightningModule
    pl.Trainer = object
    pl.seed_everything = lambda seed: torch.manual_seed(seed)
    add("pytorch_lightning", pl)
    loggers = types.ModuleType("pytorch_lightning.loggers")
    loggers.TensorBoardLogger = type("TensorBoardLogger", (), {})
    loggers.WandbLogger = type("WandbLogger", (), {})
    add("pytorch_lightning.loggers", loggers)
    import importlib
    try:
        importlib.import_module("wandb")
    except Exception:
        add("wandb", types.ModuleType("wandb"))
    try:
        importlib.import_module("torch.utils.tensorboard")
    except Exception:
        tb = types.ModuleType("torch.utils.tensorboard")
        tb.SummaryWriter = object
        add("torch.utils.tensorboard", tb)

    utils = types.ModuleType("utils")
    utils.__path__ = []
    add("utils", utils)
    for name, attrs in (("utils.dirs", {}), ("utils.metrics_log_callback", {"MetricsHistoryLogger": type("MetricsHistoryLogger", (), {})}),
                        ("utils.vqaEval", {"VQAEval": object}), ("utils.text_cleaner", {"TextCleaner": object})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__all__ = list(attrs)
        add(name, m)
    # `from models.clipcap import ClipCaptionModel, ClipCaptionPrefix` (clipcap_exector.py:37): the executor only needs
    # the names to exist; the class it instantiates is looked up by config string
    models = types.ModuleType("models")
    models.__path__ = []
    add("models", models)
    mc = types.ModuleType("models.clipcap")
    mc.ClipCaptionModel = mc.ClipCaptionPrefix = object
    add("models.clipcap", mc)
    return added


@pytest.fixture()
def executor_module():
    d = _trainers_dir()
    if d is None:
        pytest.skip("the reference's executor is neither under /root/reference nor under oracle/_ref")
    added = _install_stand_ins()
    # a bare package object for `trainers` (its own __init__.py imports EVERY executor, T0 path included)
    pkg = types.ModuleType("trainers")
    pkg.__path__ = [d]
    saved = {k: sys.modules.get(k) for k in ("trainers", "trainers.clipcap_exector", "trainers.base_executor", "trainers.metrics_processors")}
    sys.modules["trainers"] = pkg
    for k in list(saved)[1:]:
        sys.modules.pop(k, None)
    import importlib
    mod = importlib.import_module("trainers.clipcap_exector")
    import eavqa_b200
    mod.ClipCaptionPrefixB200 = eavqa_b200.ClipCaptionPrefixB200          # the one import
    yield mod
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    for k in added:
        sys.modules.pop(k, None)


class _Tokenizer:
    """GPT-2 tokenizer + the added <BOS> of the reference's data loader: len 50258 in the real run, scaled to the tiny LM."""

    def __init__(self, vocab):
        self.vocab = vocab
        self.eos_token, self.pad_token = "<eos>", None
        self.eos_token_id = vocab - 2
        self.bos_token_id = vocab - 1

    @property
    def pad_token_id(self):
        return self.eos_token_id if self.pad_token == self.eos_token else None

    def __len__(self):
        return self.vocab

    def decode(self, ids, skip_special_tokens=True):
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        return " ".join(str(i) for i in ids if not (skip_special_tokens and i >= self.vocab - 2))


def _build(executor_module, mapping_type="transformer"):
    tok = _Tokenizer(1002)            # gpt2-tiny has 1000 text ids: <eos> = 1000, <BOS> = 1001 -> resize to 1002
    config = _EasyDict({
        "model_config": {"ModelClass": "ClipCaptionPrefixB200",
                         "model_args": {"prefix_length": 4, "clip_length": 4, "prefix_size": 64, "mapping_type": mapping_type,
                                        "num_layers": 2, "model_version": "gpt2-tiny"}},
        "train": {"lr": 1e-3, "scheduler": "none", "epochs": 1, "additional": {"warmup_steps": 0}},
        "data_loader": {"additional": {"max_target_length": 5}},
        "test": {"evaluation_name": "t"},
    })
    lookup = {str(i): {"img_key": i, "question": "q", "answers": ["a"], "gold_answer": "a"} for i in range(16)}
    data_loader = types.SimpleNamespace(train_dataloader=None, test_dataloader=None, tokenizer=tok, decoder_tokenizer=tok,
                                        data=types.SimpleNamespace(vqa_data=types.SimpleNamespace(lookup=lookup)))
    ex = executor_module.ClipCapExecutor(config, data_loader)
    return ex, tok


def _caption_batch(tok, B=6, T=14, seed=5):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 1000, (B, T), generator=g)
    for b in range(B):                                   # question tokens, <BOS>, answer tokens, right padding
        n = int(torch.randint(6, T + 1, (1,), generator=g))
        ids[b, 3] = tok.bos_token_id
        ids[b, n:] = tok.eos_token_id
    mask = (ids != tok.eos_token_id).long()
    clip = 0.5 * torch.randn(B, 64, generator=g)
    return {"input_ids": ids, "attention_mask": mask, "clip_embeddings": clip}


def test_executor_module_imports_and_resolves_the_b200_class_by_name(executor_module):
    """No GPU needed: construction goes through ``globals()[ModelClass]``, resizes the embedding for the added token and
    exposes only the mapper to the optimiser."""
    ex, tok = _build(executor_module)
    import eavqa_b200
    assert isinstance(ex.model, eavqa_b200.ClipCaptionPrefixB200)
    assert ex.model.gpt.config.vocab_size == len(tok) == 1002
    out = ex.configure_optimizers()
    params = [p for g in out["optimizer"].param_groups for p in g["params"]]
    assert len(params) == len(list(ex.model.clip_project.parameters()))
    assert all(n.startswith("model.clip_project.") for n in ex.state_dict())


@pytest.mark.gpu
@pytest.mark.parametrize("mapping_type", ["mlp", "transformer"])
def test_reference_executor_trains_and_generates_through_the_b200_class(executor_module, mapping_type):
    from oracle import clip_prefix_lm as orc
    from oracle import executor_steps as orc_exec
    ex, tok = _build(executor_module, mapping_type)
    lm_w = dict(ex.model._lm_weights)                    # after resize_token_embeddings: 1002 rows
    cfg = dict(n_layer=2, n_head=2, d_model=128, prefix_length=4, clip_length=4, mapping_type=mapping_type, num_layers=2)
    ex = ex.cuda()
    opt = ex.configure_optimizers()["optimizer"]
    batch = _caption_batch(tok)
    # the labels the reference's loop produces (oracle/executor_steps.py restates it, pinned against the reference's lines)
    labels = torch.tensor(orc_exec.caption_labels(batch["input_ids"].tolist(), tok.pad_token_id, tok.bos_token_id))
    losses = []
    for it in range(6):
        opt.zero_grad(set_to_none=True)
        mapper_w = {n: p.detach().cpu().clone() for n, p in ex.model.clip_project.named_parameters()}
        out = ex.training_step(batch, it)                # the reference's own label loop + forward
        out["loss"].backward()                           # what Lightning does with the returned loss
        if it in (0, 5):                                 # the executor's step against the oracle, before and after updates
            loss_o, grads_o = orc.train_step(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"],
                                             batch["attention_mask"], labels)
            got = torch.cat([p.grad.flatten() for p in ex.model.clip_project.parameters()]).double().cpu()
            ref = torch.cat([grads_o[k].flatten() for k in grads_o]).double()
            assert abs(float(out["loss"]) - loss_o) <= 1e-3 * abs(loss_o), (it, float(out["loss"]), loss_o)
            assert float(got @ ref / (got.norm() * ref.norm())) >= 0.999, it
        opt.step()
        ex.scheduler.step()
        losses.append(float(out["loss"].detach()))
    assert "train/loss" in ex.logged and all(l == l for l in losses)
    assert losses[-1] < losses[0], losses

    # _generative_step: generate -> list[list[int]] that supports `in` / `.index` (clipcap_exector.py:236-264)
    mapper_w = {n: p.detach().cpu().clone() for n, p in ex.model.clip_project.named_parameters()}
    gen_batch = {"generative_input_ids": batch["input_ids"][:, :6].contiguous(), "generative_attention_mask": torch.ones(6, 6, dtype=torch.long),
                 "clip_embeddings": batch["clip_embeddings"], "labels": [[1, 2, -100]] * 6, "question_ids": list(range(6)),
                 "answers": [["a"]] * 6}
    res = ex.eval()._generative_step(gen_batch, batch_idx=99)
    assert len(res["predictions"]) == 6 and len(res["outputs"]) == 6
    assert all(isinstance(r, list) and 1 <= len(r) <= 5 for r in res["outputs"])
    ref_tokens, margins = orc.generate(lm_w, mapper_w, cfg, gen_batch["generative_input_ids"], batch["clip_embeddings"],
                                       gen_batch["generative_attention_mask"], max_length=5, pad_token_id=tok.pad_token_id,
                                       eos_token_id=tok.eos_token_id, return_margins=True)
    for g, r, mg in zip(res["outputs"], ref_tokens, margins.tolist()):
        n = next((i for i, m in enumerate(mg) if m < 0.05), len(r))
        assert g[:n] == r[:n]


def test_state_dict_with_lm_round_trips_into_the_reference_module():
    """``state_dict(include_lm=True)`` strict-loads into the reference's ``ClipCaptionPrefix`` and a reference checkpoint
    loads back -- also when the model is nested in a parent module (Lightning: keys ``model.gpt.*``)."""
    from oracle import reference_shim
    if reference_shim.reference_dir() is None:
        pytest.skip("the reference's clipcap.py is neither under /root/reference nor under oracle/_ref")
    import eavqa_b200
    import eavqa_b200.synthetic as syn
    clipcap, _, GPT2Config, holder = reference_shim.import_reference()
    m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=4, clip_length=4, prefix_size=64, num_layers=2, mapping_type="transformer",
                                         model_version="gpt2-tiny")
    cfg_lm = syn.lm_config("gpt2-tiny")
    mapper_w = syn.make_mapper_params("transformer", 64, 128, 4, 4, 2, seed=3, perturb_norm=True)
    ref = reference_shim.build_reference_model(clipcap, GPT2Config, holder, cfg_lm, syn.make_lm_weights(cfg_lm, seed=9), mapper_w,
                                               prefix_length=4, clip_length=4, clip_dim=64, num_layers=2, mapping_type="transformer")
    # ours -> reference, strict
    sd = m.state_dict(include_lm=True)
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith(".attn.bias") or k.endswith(".attn.masked_bias") for k in missing), missing   # HF causal-mask buffers
    assert torch.equal(ref.gpt.transformer.wte.weight, m._lm_weights["transformer.wte.weight"])
    assert torch.equal(ref.gpt.lm_head.weight, m._lm_weights["transformer.wte.weight"])
    # reference checkpoint -> ours, nested and strict: mapper AND the LM it carries arrive
    ref2 = reference_shim.build_reference_model(clipcap, GPT2Config, holder, cfg_lm, syn.make_lm_weights(cfg_lm, seed=9), mapper_w,
                                                prefix_length=4, clip_length=4, clip_dim=64, num_layers=2, mapping_type="transformer")

    class Executor(torch.nn.Module):
        def __init__(self, model):
            super().__init__()
            self.model = model
    ckpt = {"model." + k: v for k, v in ref2.state_dict().items()}
    res = Executor(m).load_state_dict(ckpt, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m._lm_weights["transformer.wte.weight"], ref2.gpt.transformer.wte.weight)
    assert torch.equal(m.clip_project.prefix_const.detach(), mapper_w["prefix_const"])
    assert set(m.state_dict().keys()) == {"clip_project." + k for k in mapper_w}         # default: mapper only
