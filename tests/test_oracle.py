"""CPU: the oracle restatement against the committed golden fixtures.

The fixtures under tests/golden/ hold the UNMODIFIED reference's outputs
(``oracle/validate_against_reference.py``), so these tests pin the oracle on any box,
including ones where ``/root/reference`` does not exist.
"""
import json
import os

import pytest
import torch

from oracle import clip_prefix_lm as orc
from oracle.cases import CASES, SLOW_CASES, SPLICE_GOLDEN, build_case

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
# the two GPT-2-small/medium generate cases cost ~20-40 s each on 8 cores
# full-size cases (minutes of CPU each) are re-run only by oracle/validate_against_reference.py and, on the GPU box, by
# tests/test_step_gpu.py; here their fixtures are only checked for presence and case-table agreement
TRAIN = [k for k, v in CASES.items() if v["kind"] == "train" and k not in SLOW_CASES]
GEN = [k for k, v in CASES.items() if v["kind"] == "generate" and k not in SLOW_CASES]


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def test_every_case_has_a_fixture():
    for name in CASES:
        assert os.path.exists(os.path.join(GOLDEN, name + ".json")), name
        assert load(name)["case"] == json.loads(json.dumps(CASES[name])), "case table drifted from fixture: " + name


@pytest.mark.parametrize("golden", SPLICE_GOLDEN, ids=[g["name"] for g in SPLICE_GOLDEN])
def test_splice_vct0_golden(golden):
    toks = torch.tensor(golden["question_tokens"])
    e, m = orc.insert_prefix_into_input(golden["prefix_length"], golden["num_shots"], toks,
                                        torch.tensor(golden["text_embeddings"]),
                                        torch.tensor(golden["prefix_projections"]),
                                        torch.tensor(golden["question_masks"]))
    assert torch.equal(e, torch.tensor(golden["expected_embeddings"]))
    assert torch.equal(m, torch.tensor(golden["expected_masks"]))


def test_splice_rejects_wrong_sentinel_count():
    toks = torch.tensor([[32099, 5, 6, 7]])
    with pytest.raises(ValueError):
        orc.insert_prefix_into_input(2, 1, toks, torch.zeros(1, 4, 3), torch.zeros(1, 2, 2, 3), torch.ones(1, 4, dtype=torch.int64))


@pytest.mark.parametrize("name", TRAIN)
def test_train_step_matches_reference_fixture(name):
    fx = load(name)
    lm_w, mapper_w, batch, cfg = build_case(CASES[name])
    loss, grads = orc.train_step(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"],
                                 batch["attention_mask"], batch["labels"])
    assert abs(loss - fx["loss"]) / abs(fx["loss"]) < 2e-5
    assert list(grads.keys()) == list(fx["grads"].keys())
    total = 0.0
    for k, g in grads.items():
        ref = fx["grads"][k]
        g64 = g.double().flatten()
        total += float(g64.norm()) ** 2
        assert abs(float(g64.norm()) - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-9, k
        scale = max(ref["norm"] / max(g64.numel(), 1) ** 0.5, 1e-12)
        assert torch.allclose(g64[:6], torch.tensor(ref["head"], dtype=torch.float64), rtol=5e-3, atol=5e-2 * scale), k
        assert torch.allclose(g64[-6:], torch.tensor(ref["tail"], dtype=torch.float64), rtol=5e-3, atol=5e-2 * scale), k
    assert abs(total ** 0.5 - fx["grad_total_norm"]) < 1e-3 * fx["grad_total_norm"]


@pytest.mark.parametrize("name", GEN)
def test_generate_matches_reference_fixture(name):
    fx = load(name)
    case = CASES[name]
    lm_w, mapper_w, batch, cfg = build_case(case)
    kw = dict(max_length=case["max_length"], pad_token_id=case["pad_token_id"], eos_token_id=case["eos_token_id"])
    if case["num_shots"] is None:
        got = orc.generate(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"], batch["attention_mask"], **kw)
    else:
        got = orc.generate_few_shot(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"],
                                    batch["attention_mask"], case["special_token_id"], **kw)
    # rows whose reference margins are all comfortably above fp32 cross-machine noise must match exactly
    for row, (g, r, mg) in enumerate(zip(got, fx["tokens"], fx["margins"])):
        if min(mg) > 1e-3:
            assert g == r, (name, row)


def test_generate_requires_pad_when_eos_given():
    case = CASES["gen_tiny_prepend"]
    lm_w, mapper_w, batch, cfg = build_case(case)
    with pytest.raises(ValueError):
        orc.generate(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"], batch["attention_mask"],
                     max_length=2, pad_token_id=None, eos_token_id=3)


def test_eos_stops_rows_and_pads_outputs():
    """Force EOS to be the argmax: every row finishes at step 0, outputs are [eos], loop breaks (clipcap.py:458-464)."""
    case = CASES["gen_tiny_prepend"]
    lm_w, mapper_w, batch, cfg = build_case(case)
    fx = load("gen_tiny_prepend")
    eos = fx["tokens"][0][0]
    rows = [i for i, t in enumerate(fx["tokens"]) if t[0] == eos]
    got = orc.generate(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"], batch["attention_mask"],
                       max_length=4, pad_token_id=7, eos_token_id=eos)
    for i in rows:
        assert got[i][0] == eos and all(t == 7 for t in got[i][1:])
