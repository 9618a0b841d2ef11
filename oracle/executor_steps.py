"""CPU oracle for the executor-side steps next to the CLIP-prefix LM path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pure-Python / numpy restatements, small cases only.

Parity status: PINNED.  ``oracle/validate_executor_steps.py`` executes the reference's OWN source lines
(``src/trainers/clipcap_exector.py:134-150`` and ``src/trainers/few_shot_vqa_executor.py:316-329``, read from
``/root/reference`` at run time, never copied) on seeded inputs, checks these functions against them bit-for-bit
(labels, selected member) / to 1e-6 (scores), and writes ``tests/golden/executor_labels.json`` and
``tests/golden/executor_ensembles.json``.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def caption_labels(input_ids: Sequence[Sequence[int]], pad_token_id: int, bos_token_id: int) -> List[List[int]]:
    """``ClipCapExecutor.training_step`` (clipcap_exector.py:134-150).

    :134-135  labels = input_ids.clone(); labels[labels == pad] = -100
    :137-150  per row, left to right: the first -100 (= first pad) is set back to ``pad_token_id`` and ends the scan
              (everything after it keeps the cloned value: pads stay -100, other ids stay themselves); a ``<BOS>``
              switches "answer tokens" on and is itself ignored; tokens before the first ``<BOS>`` are ignored.
    """
    out = []
    for row in input_ids:
        lab = [-100 if t == pad_token_id else int(t) for t in row]
        answer = False
        for j, t in enumerate(lab):
            if t == -100:
                lab[j] = pad_token_id
                break
            if t == bos_token_id:
                answer = True
                lab[j] = -100
                continue
            if not answer:
                lab[j] = -100
        out.append(lab)
    return out


def log_softmax(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.float64)
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def ensemble_scores(step_logits: np.ndarray, sequences: np.ndarray, skip_ids: Sequence[int] = (0, 1, 2)) -> np.ndarray:
    """One ensemble member of ``generate_from_ensembles`` (few_shot_vqa_executor.py:316-323).

    ``step_logits`` [S, B, V]: ``outputs.scores`` stacked; ``sequences`` [B, S+1]: ``outputs.sequences`` (position 0 is
    the decoder start token).  score[b] = sum_k log_softmax(step_logits)[k-1, b, sequences[b, k]] over positions whose
    token is not in ``skip_ids`` (k = 0 would index step -1, but the start token is always in the skip set)."""
    lp = log_softmax(step_logits)
    B = sequences.shape[0]
    scores = np.zeros(B, dtype=np.float64)
    for j in range(B):
        for k, tok in enumerate(sequences[j]):
            if int(tok) not in skip_ids:
                scores[j] += lp[k - 1, j, int(tok)]
    return scores


def ensemble_select(member_scores: np.ndarray) -> np.ndarray:
    """few_shot_vqa_executor.py:328: ``np.argmax(batch_sequence_scores, axis=1)`` on the [B, E] score table."""
    return np.argmax(member_scores, axis=1)
