"""RICES in-context example retrieval on the device: the step before the few-shot path.

The reference builds its in-context example lists offline with faiss (``src/in_context_example_selection/``):

* ``get_question_knn.py:64-76`` -- ``faiss.normalize_L2`` on the train / val CLIP text embeddings, ``IndexFlatIP`` on the
  GPU, ``search(val, k=2048)``  ->  :func:`knn_inner_product`;
* ``get_image_knn_from_text_knn.py:79-92`` -- per test question, the images of its 2048 text neighbours ranked by the same
  normalised inner product with the question's image  ->  :func:`rerank_candidates`.

Both call the C ABI (``eavqa_rices_search`` / ``eavqa_rices_rerank``); there is no CPU path.
"""
from __future__ import annotations

import torch

from . import lib as _lib


def _f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.device.type != "cuda":
        raise _lib.EavqaError("%s has no CPU path: pass CUDA tensors" % what)
    return t.to(torch.float32).contiguous()


def knn_inner_product(queries: torch.Tensor, database: torch.Tensor, k: int):
    """``faiss.normalize_L2(database); faiss.normalize_L2(queries); D, I = IndexFlatIP(d).search(queries, k)``.
    Returns ``(D [M, k] fp32 descending, I [M, k] int64)``; the inputs are not modified."""
    q, db = _f32_cuda(queries, "knn_inner_product"), _f32_cuda(database, "knn_inner_product")
    if q.dim() != 2 or db.dim() != 2 or q.shape[1] != db.shape[1]:
        raise ValueError("queries [M, D] and database [N, D] must share D")
    M, D = q.shape
    scores = torch.empty(M, k, dtype=torch.float32, device=q.device)
    index = torch.empty(M, k, dtype=torch.int64, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().eavqa_rices_search(q.data_ptr(), db.data_ptr(), M, db.shape[0], D, int(k), scores.data_ptr(),
                                                  index.data_ptr(), _lib.current_stream()))
    return scores, index


def rerank_candidates(query: torch.Tensor, table: torch.Tensor, candidates: torch.Tensor):
    """Per question ``q``: cosine similarity of ``query[q]`` with ``table[candidates[q, j]]`` (``-1`` = padding), all
    candidates sorted by similarity.  Returns ``(sims [M, C] fp32 descending, positions [M, C] int32)`` -- ``positions``
    index the candidate list of the question, like ``I`` of the per-question faiss index of the reference."""
    qv, tb = _f32_cuda(query, "rerank_candidates"), _f32_cuda(table, "rerank_candidates")
    cand = candidates.to(device=qv.device, dtype=torch.int32).contiguous()
    M, C = cand.shape
    sims = torch.empty(M, C, dtype=torch.float32, device=qv.device)
    pos = torch.empty(M, C, dtype=torch.int32, device=qv.device)
    with torch.cuda.device(qv.device):
        _lib.check(_lib.load().eavqa_rices_rerank(qv.data_ptr(), tb.data_ptr(), M, qv.shape[1], cand.data_ptr(), C, sims.data_ptr(),
                                                  pos.data_ptr(), _lib.current_stream()))
    return sims, pos
