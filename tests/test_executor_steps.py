"""Executor-side steps next to the path: caption-label construction and ensemble scoring.

CPU: the oracle against the fixtures that ``oracle/validate_executor_steps.py`` produced by executing the reference's
own lines.  GPU: the kernels, through the C ABI, against the fixtures and the oracle -- bit-exact for labels and for the
selected ensemble member, 1e-4 for the fp32 score sums."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import executor_steps as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LABELS = json.load(open(os.path.join(ROOT, "tests", "golden", "executor_labels.json")))["cases"]
ENSEMBLES = json.load(open(os.path.join(ROOT, "tests", "golden", "executor_ensembles.json")))["cases"]


# ------------------------------------------------------------------------------------------------ CPU: oracle vs reference
@pytest.mark.parametrize("i", range(len(LABELS)))
def test_oracle_labels_match_reference_fixture(i):
    c = LABELS[i]
    assert orc.caption_labels(c["input_ids"], c["pad_token_id"], c["bos_token_id"]) == c["labels"]


@pytest.mark.parametrize("i", range(len(ENSEMBLES)))
def test_oracle_ensembles_match_reference_fixture(i):
    c = ENSEMBLES[i]
    table = np.stack([orc.ensemble_scores(np.array(m["step_logits"], dtype=np.float32), np.array(m["sequences"]))
                      for m in c["members"]], axis=1)
    assert np.abs(table - np.array(c["scores"])).max() < 1e-4 * max(1.0, np.abs(np.array(c["scores"])).max())
    assert orc.ensemble_select(table).tolist() == c["best"]


def test_oracle_labels_edge_cases():
    # no pad at all / only pads / <BOS> as last token / nothing before the first pad
    assert orc.caption_labels([[7, 9, 3, 4]], 0, 9) == [[-100, -100, 3, 4]]
    assert orc.caption_labels([[0, 0, 0]], 0, 9) == [[0, -100, -100]]
    assert orc.caption_labels([[5, 6, 9]], 0, 9) == [[-100, -100, -100]]
    assert orc.caption_labels([[9, 5, 0, 6, 0]], 0, 9) == [[-100, 5, 0, 6, -100]]


# ------------------------------------------------------------------------------------------------ GPU: kernels vs oracle
@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(LABELS)))
def test_caption_labels_kernel_matches_reference_fixture(i):
    from eavqa_b200.executor_steps import build_caption_labels
    c = LABELS[i]
    got = build_caption_labels(torch.tensor(c["input_ids"], dtype=torch.int64, device="cuda"), c["pad_token_id"], c["bos_token_id"])
    assert got.dtype == torch.int64 and got.cpu().tolist() == c["labels"]


@pytest.mark.gpu
def test_caption_labels_kernel_full_size_against_oracle():
    """BASELINE batch (256 x 40) and a long ragged batch: bit-exact against the oracle."""
    from eavqa_b200.executor_steps import build_caption_labels
    g = torch.Generator().manual_seed(5)
    for B, T in ((256, 40), (1000, 77), (1, 1)):
        ids = torch.randint(0, 50258, (B, T), generator=g)
        lens = torch.randint(0, T + 1, (B,), generator=g)
        ids[torch.arange(T)[None, :] >= lens[:, None]] = 50256
        got = build_caption_labels(ids.cuda(), 50256, 50257).cpu().tolist()
        assert got == orc.caption_labels(ids.tolist(), 50256, 50257)


def _members_to_arrays(c):
    lps, toks = [], []
    for m in c["members"]:
        lp = orc.log_softmax(np.array(m["step_logits"], dtype=np.float32))          # [S, B, V]
        seq = np.array(m["sequences"])[:, 1:]                                       # drop the start token -> [B, S]
        S, B, _ = lp.shape
        lps.append(np.stack([[lp[k, b, seq[b, k]] for k in range(S)] for b in range(B)]))
        toks.append(seq)
    return np.stack(lps).astype(np.float32), np.stack(toks)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(ENSEMBLES)))
def test_ensemble_select_kernel_matches_reference_fixture(i):
    from eavqa_b200.executor_steps import ensemble_select
    c = ENSEMBLES[i]
    lp, tk = _members_to_arrays(c)
    best, best_tokens, scores = ensemble_select(torch.tensor(lp).cuda(), torch.tensor(tk).cuda(), skip_ids=(0, 1, 2))
    assert best.cpu().tolist() == c["best"]
    ref = np.array(c["scores"])
    assert np.abs(scores.cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
    for b, e in enumerate(c["best"]):
        assert best_tokens[b].cpu().tolist() == tk[e, b].tolist()


@pytest.mark.gpu
def test_ensemble_select_ties_keep_the_first_member_and_empty_skip_list():
    from eavqa_b200.executor_steps import ensemble_select
    lp = torch.tensor([[[-1.0, -2.0]], [[-2.0, -1.0]], [[-0.5, -0.5]]]).cuda()          # [E=3, B=1, S=2]; member 2 wins
    tk = torch.tensor([[[5, 6]], [[7, 8]], [[9, 10]]]).cuda()
    best, toks, scores = ensemble_select(lp, tk, skip_ids=())
    assert best.cpu().tolist() == [2] and toks.cpu().tolist() == [[9, 10]]
    best, _, scores = ensemble_select(lp[:2], tk[:2], skip_ids=())                       # -3 == -3: np.argmax keeps member 0
    assert best.cpu().tolist() == [0] and scores.cpu().tolist() == [[-3.0, -3.0]]


@pytest.mark.gpu
def test_generate_logprobs_and_ensembles_against_oracle():
    """``generate(return_logprobs=True)``: log softmax of the picked token per step against the fp32 oracle (bf16
    tensor-core logits: 0.05 absolute), then the whole ``generate_from_ensembles`` flow: same winner as the oracle for
    every question whose score gap exceeds the bf16 noise."""
    import eavqa_b200
    from eavqa_b200.executor_steps import generate_from_ensembles
    from oracle import clip_prefix_lm as ref
    from oracle.cases import CASES, build_case
    case = CASES["gen_tiny_prepend"]
    lm_w, mapper_w, batch, cfg = build_case(case)
    model = eavqa_b200.ClipCaptionPrefixB200(prefix_length=case["prefix_length"], clip_length=case["clip_length"],
                                             prefix_size=case["clip_dim"], num_layers=case["num_layers"],
                                             mapping_type=case["mapping_type"], model_version="synthetic", lm_state_dict=lm_w)
    model.clip_project.load_state_dict(mapper_w)
    model = model.cuda().eval()
    kw = dict(max_length=case["max_length"], pad_token_id=case["pad_token_id"], eos_token_id=None)
    toks, lp = model.generate(question_tokens=batch["input_ids"].cuda(), prefix=batch["clip_embeddings"].cuda(),
                              question_mask=batch["attention_mask"].cuda(), return_logprobs=True, **kw)
    r_tok, r_margin, r_lp = ref.generate(lm_w, mapper_w, cfg, batch["input_ids"], batch["clip_embeddings"], batch["attention_mask"],
                                         return_margins=True, return_logprobs=True, **kw)
    toks, lp = toks.cpu(), lp.cpu()
    checked = 0
    for b in range(toks.shape[0]):
        for k in range(toks.shape[1]):
            if int(toks[b, k]) != r_tok[b][k] or float(r_margin[b, k]) < 0.05:
                break                                  # beyond a near-tie the two decodes follow different prefixes
            assert abs(float(lp[b, k]) - float(r_lp[b, k])) < 0.05, (b, k, float(lp[b, k]), float(r_lp[b, k]))
            checked += 1
    assert checked >= toks.numel() // 2
    # ---- ensembles: E members = the same questions with permuted clip embeddings
    E, B = 3, batch["input_ids"].shape[0]
    g = torch.Generator().manual_seed(3)
    clips = torch.stack([batch["clip_embeddings"][torch.randperm(B, generator=g)] for _ in range(E)], dim=1)      # [B, E, D]
    ids = batch["input_ids"][:, None, :].expand(B, E, -1).contiguous()
    msk = batch["attention_mask"][:, None, :].expand(B, E, -1).contiguous()
    best_tokens, best, scores = generate_from_ensembles(model, ids.cuda(), msk.cuda(), clips.cuda(), E, **kw)
    table = np.zeros((B, E))
    ref_tokens = []
    for e in range(E):
        t, _, l = ref.generate(lm_w, mapper_w, cfg, ids[:, e], clips[:, e], msk[:, e], return_margins=True, return_logprobs=True, **kw)
        table[:, e] = l.double().sum(dim=1).numpy()
        ref_tokens.append(t)
    ref_best = orc.ensemble_select(table)
    srt = np.sort(table, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 0.5            # score gaps the bf16 path cannot flip
    assert clear.sum() >= 1
    assert (best.cpu().numpy()[clear] == ref_best[clear]).all()
    assert np.abs(scores.cpu().numpy() - table)[clear].max() < 0.5
