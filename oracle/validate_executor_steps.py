"""Pin ``oracle/executor_steps.py`` against the reference's own source lines and write the golden fixtures.

Runs only where ``/root/reference`` exists.  The reference's executors cannot be imported here (pytorch-lightning 1.6.3,
wandb, easydict ... are absent), so the relevant statements are read from the reference files at run time, dedented and
executed verbatim on seeded inputs with a stand-in ``self`` -- the reference's code computes the expected values; none
of it is copied into this repository.

    python oracle/validate_executor_steps.py
"""
from __future__ import annotations

import json
import os
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("EAVQA_REFERENCE", "/root/reference")
from oracle import executor_steps as orc      # noqa: E402


def ref_lines(rel, first, last):
    with open(os.path.join(REF, rel)) as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def label_cases():
    g = torch.Generator().manual_seed(2021)
    pad, bos = 50256, 50257
    cases = []
    for B, T in ((4, 12), (8, 40), (3, 7), (16, 33)):
        ids = torch.randint(0, 50256, (B, T), generator=g)
        for b in range(B):
            L = int(torch.randint(1, T + 1, (1,), generator=g))
            ids[b, L:] = pad                                       # right padding (module_parser.py:424)
            if L >= 2 and b % 5 != 4:                              # every fifth row has no <BOS> at all
                ids[b, int(torch.randint(0, L - 1, (1,), generator=g))] = bos
            if b % 7 == 3 and L >= 4:                              # an EOS (= pad id) in the middle of the text
                ids[b, L // 2] = pad
            if b % 6 == 2 and L >= 6:                              # a second <BOS> after the first
                ids[b, L - 2] = bos
        cases.append(dict(pad_token_id=pad, bos_token_id=bos, input_ids=ids.tolist()))
    cases.append(dict(pad_token_id=0, bos_token_id=0, input_ids=[[5, 0, 7, 0], [0, 0, 0, 0], [3, 4, 5, 6]]))   # pad == bos
    return cases


def run_reference_labels(code, case):
    self = types.SimpleNamespace(tokenizer=types.SimpleNamespace(pad_token_id=case["pad_token_id"], bos_token_id=case["bos_token_id"]))
    ns = dict(self=self, sample_batched={"input_ids": torch.tensor(case["input_ids"], dtype=torch.int64)}, torch=torch)
    exec(code, ns)
    return ns["labels"].tolist()


def ensemble_cases():
    g = torch.Generator().manual_seed(7)
    cases = []
    for E, B, S, V in ((3, 4, 5, 11), (5, 6, 10, 37), (2, 3, 1, 8)):
        members = []
        for e in range(E):
            logits = torch.randn(S, B, V, generator=g) * 2.0
            seq = torch.zeros(B, S + 1, dtype=torch.int64)
            seq[:, 1:] = logits.argmax(dim=-1).t()                 # greedy sequences behind a start token (id 0)
            for b in range(B):                                     # finished rows are padded with 0 after an eos (id 1)
                stop = int(torch.randint(1, S + 2, (1,), generator=g))
                if stop <= S:
                    seq[b, stop] = 1
                    seq[b, stop + 1:] = 0
            members.append(dict(step_logits=logits.tolist(), sequences=seq.tolist()))
        cases.append(dict(E=E, B=B, S=S, V=V, members=members))
    return cases


def run_reference_ensembles(code, case):
    """Executes few_shot_vqa_executor.py:316-323 once per member and :328 on the table, as the method body does."""
    B, E = case["B"], case["E"]
    batch_sequence_scores = np.zeros((B, E))
    for i, m in enumerate(case["members"]):
        outputs = types.SimpleNamespace(scores=[torch.tensor(s, dtype=torch.float32) for s in m["step_logits"]],
                                        sequences=torch.tensor(m["sequences"], dtype=torch.int64))
        ns = dict(outputs=outputs, torch=torch, np=np, batch_sequence_scores=batch_sequence_scores, i=i)
        exec(code, ns)
    best = np.argmax(batch_sequence_scores, axis=1)
    return batch_sequence_scores, best


def main():
    label_code = ref_lines("src/trainers/clipcap_exector.py", 134, 150)
    assert label_code.lstrip().startswith("labels = sample_batched") and "labels[i, j] = -100" in label_code, label_code
    lab_out = []
    for c in label_cases():
        ref = run_reference_labels(label_code, c)
        mine = orc.caption_labels(c["input_ids"], c["pad_token_id"], c["bos_token_id"])
        assert ref == mine, "label oracle differs from the reference"
        lab_out.append(dict(c, labels=ref))
    ens_code = ref_lines("src/trainers/few_shot_vqa_executor.py", 316, 324)
    assert "outputs_scores = torch.log" in ens_code and "batch_sequence_scores[j, i] = sequence_score" in ens_code, ens_code
    ens_out = []
    for c in ensemble_cases():
        table, best = run_reference_ensembles(ens_code, c)
        mine = np.stack([orc.ensemble_scores(np.array(m["step_logits"], dtype=np.float32), np.array(m["sequences"]))
                         for m in c["members"]], axis=1)
        assert np.abs(mine - table).max() < 1e-4 * max(1.0, np.abs(table).max()), (mine, table)
        assert (orc.ensemble_select(mine) == best).all()
        ens_out.append(dict(c, scores=table.tolist(), best=best.tolist()))
    gd = os.path.join(ROOT, "tests", "golden")
    json.dump(dict(source="reference src/trainers/clipcap_exector.py:134-150 executed by oracle/validate_executor_steps.py",
                   cases=lab_out), open(os.path.join(gd, "executor_labels.json"), "w"))
    json.dump(dict(source="reference src/trainers/few_shot_vqa_executor.py:316-328 executed by oracle/validate_executor_steps.py",
                   cases=ens_out), open(os.path.join(gd, "executor_ensembles.json"), "w"))
    print("executor-step oracle pinned: %d label cases, %d ensemble cases" % (len(lab_out), len(ens_out)))


if __name__ == "__main__":
    main()
