"""CPU oracle for the RICES retrieval step.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Parity status: UNPINNED.  The arithmetic lives in faiss (``import faiss`` in
``src/in_context_example_selection/get_question_knn.py:6`` and ``get_image_knn_from_text_knn.py:6``; un-pinned in the
reference's ``requirements.txt``, not vendored under ``/root/reference`` and not installed in this image), and the
reference holds no test or golden vector for this step.  This file restates faiss' published semantics for the calls the
reference makes:

* ``faiss.normalize_L2(x)`` -- every row is scaled in place by ``1 / ||row||_2`` computed in float32; rows of norm 0 are
  left untouched (``fvec_renorm_L2``);
* ``IndexFlatIP.add(db)``; ``search(q, k)`` -- exact inner products, the ``k`` largest per query in descending order,
  ``I`` = database row numbers; with fewer than ``k`` database rows the tail is ``-FLT_MAX`` / ``-1``.

Scores are accumulated in float64 here; ties are ordered by ascending row number (faiss leaves the order of equal scores
unspecified).  The anchors available are the reference's call sites: ``get_question_knn.py:64-76`` (k = 2048 over the
train questions) and ``get_image_knn_from_text_knn.py:79-92`` (per-question index, k = all candidates).
"""
from __future__ import annotations

import numpy as np

FLT_MAX = np.float32(3.402823466e38)


def normalize_l2(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    nr = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32)).astype(np.float32)
    scale = np.where(nr > 0, np.float32(1.0) / np.where(nr > 0, nr, 1), np.float32(1.0)).astype(np.float32)
    return (x * scale[:, None]).astype(np.float32)


def knn_inner_product(queries: np.ndarray, database: np.ndarray, k: int):
    """get_question_knn.py:64-76.  Returns ``(D [M, k] float64, I [M, k] int64)``."""
    q = normalize_l2(queries).astype(np.float64)
    db = normalize_l2(database).astype(np.float64)
    scores = q @ db.T
    M, N = scores.shape
    D = np.full((M, k), -float(FLT_MAX))
    I = np.full((M, k), -1, dtype=np.int64)
    idx = np.arange(N)
    for m in range(M):
        order = np.lexsort((idx, -scores[m]))[:k]          # score descending, then row ascending
        D[m, :len(order)] = scores[m, order]
        I[m, :len(order)] = order
    return D, I


def rerank_candidates(query: np.ndarray, table: np.ndarray, candidates: np.ndarray):
    """get_image_knn_from_text_knn.py:79-92 for a batch of questions: ``candidates[q]`` are rows of ``table`` (-1 pads).
    Returns ``(sims [M, C] float64 descending, positions [M, C] int64)``; padding sorts last as ``-FLT_MAX`` / ``-1``."""
    M, C = candidates.shape
    qn = normalize_l2(query).astype(np.float64)
    sims = np.full((M, C), -float(FLT_MAX))
    pos = np.full((M, C), -1, dtype=np.int64)
    for m in range(M):
        valid = np.nonzero(candidates[m] >= 0)[0]
        if len(valid) == 0:
            continue
        x = normalize_l2(table[candidates[m, valid]]).astype(np.float64)
        s = x @ qn[m]
        order = np.lexsort((valid, -s))
        sims[m, :len(order)] = s[order]
        pos[m, :len(order)] = valid[order]
    return sims, pos
