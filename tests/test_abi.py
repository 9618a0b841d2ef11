"""CPU: the C-ABI library loads and exports every symbol ``include/eavqa_b200.h`` declares, and the host mirror
keeps the reference's names / layout.  No compute is called (there is no GPU on the build box)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "eavqa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eavqa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from eavqa_b200 import lib
    L = lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export " + s
        assert s in lib.PROTOTYPES, "lib.py has no prototype for " + s
    assert sorted(lib.PROTOTYPES) == syms
    assert L.eavqa_abi_version() == 1


def test_no_cpu_fallback():
    import eavqa_b200
    from eavqa_b200 import lib
    m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=4, clip_length=4, prefix_size=64, num_layers=2,
                                         mapping_type="mlp", model_version="gpt2-tiny")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.EavqaError):
        m(question_tokens=torch.zeros(2, 5, dtype=torch.long), prefix=torch.zeros(2, 64), labels=torch.zeros(2, 5, dtype=torch.long))
    with pytest.raises(lib.EavqaError):
        m.generate(question_tokens=torch.zeros(2, 5, dtype=torch.long), prefix=torch.zeros(2, 64), pad_token_id=0, eos_token_id=0)


@pytest.mark.parametrize("mapping_type", ["mlp", "transformer"])
def test_host_mirror_keeps_reference_parameter_names(mapping_type):
    import eavqa_b200
    from eavqa_b200 import synthetic as syn
    m = eavqa_b200.ClipCaptionPrefixB200(prefix_length=10, clip_length=10, prefix_size=512, num_layers=8,
                                         mapping_type=mapping_type, model_version="gpt2-tiny")
    names = [n for n, _ in m.named_parameters()]
    shapes = syn.mapper_param_shapes(mapping_type, 512, 128, 10, 10, 8)
    assert names == ["clip_project." + k for k in shapes]
    assert [tuple(p.shape) for _, p in m.named_parameters()] == list(shapes.values())
    # ClipCaptionPrefix.parameters() yields the mapper only (clipcap.py:591-592); the LM holds no nn.Parameters at all
    assert sum(p.numel() for p in m.parameters()) == sum(p.numel() for p in m.clip_project.parameters())
    assert set(m.state_dict().keys()) == set(names)
    # executor surface (clipcap_exector.py:52-56)
    assert m.gpt.resize_token_embeddings(1001) is m.gpt and m.gpt.config.vocab_size == 1001
    assert m._lm_weights["transformer.wte.weight"].shape[0] == 1001


def test_reference_counts_for_gpt2_small():
    """Probed on the reference: 31.47 M (MLP) and 41,745,408 (transformer) trainable parameters."""
    from eavqa_b200 import synthetic as syn
    n = lambda s: sum(int(torch.tensor(v).prod()) for v in s.values())
    assert n(syn.mapper_param_shapes("transformer", 512, 768, 10, 10, 8)) == 41745408
    assert n(syn.mapper_param_shapes("mlp", 512, 768, 10, 10, 8)) == 3840 * 512 + 3840 + 7680 * 3840 + 7680
