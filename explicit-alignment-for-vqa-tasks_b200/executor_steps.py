"""The steps right around the CLIP-prefix LM path inside the reference's executors, on the device.

* :func:`build_caption_labels` -- ``ClipCapExecutor.training_step``'s label construction
  (``src/trainers/clipcap_exector.py:134-150``): a Python double loop over ``[B, T]`` tensor elements there, one kernel here.
* :func:`generate_from_ensembles` -- ``FewShotVQAExecutor.generate_from_ensembles``
  (``src/trainers/few_shot_vqa_executor.py:293-332``): one greedy generation per ensemble member (permutation of the
  in-context examples), a sequence score = sum of the generated tokens' log-probabilities, the best member per question.

Both call the C ABI (``eavqa_build_caption_labels``, ``eavqa_generate`` + ``eavqa_ensemble_select``); there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import lib as _lib


def _need_cuda(t: torch.Tensor, what: str):
    if t.device.type != "cuda":
        raise _lib.EavqaError("%s has no CPU path: pass CUDA tensors" % what)


def build_caption_labels(input_ids: torch.Tensor, pad_token_id: int, bos_token_id: int) -> torch.Tensor:
    """labels for a caption batch (clipcap_exector.py:134-150): ``-100`` on pads and on everything up to and including
    ``<BOS>``; the first pad position keeps ``pad_token_id`` (the EOS target).  ``input_ids`` ``[B, T]`` int64, CUDA."""
    _need_cuda(input_ids, "build_caption_labels")
    tok = input_ids.to(torch.int64).contiguous()
    labels = torch.empty_like(tok)
    with torch.cuda.device(tok.device):
        _lib.check(_lib.load().eavqa_build_caption_labels(tok.data_ptr(), tok.shape[0], tok.shape[1], int(pad_token_id),
                                                          int(bos_token_id), labels.data_ptr(), _lib.current_stream()))
    return labels


def ensemble_select(logprob: torch.Tensor, tokens: torch.Tensor, skip_ids: Sequence[int] = (0, 1, 2)):
    """``logprob`` / ``tokens`` ``[E, B, S]`` -> ``(best [B] int32, best_tokens [B, S] int64, scores [B, E] fp32)``
    (few_shot_vqa_executor.py:316-331; the reference skips token ids 0, 1, 2 when summing)."""
    _need_cuda(logprob, "ensemble_select")
    lp = logprob.to(torch.float32).contiguous()
    tk = tokens.to(device=lp.device, dtype=torch.int64).contiguous()
    E, B, S = tk.shape
    skip = torch.tensor(list(skip_ids), dtype=torch.int64, device=lp.device)
    scores = torch.empty(B, E, dtype=torch.float32, device=lp.device)
    best = torch.empty(B, dtype=torch.int32, device=lp.device)
    best_tokens = torch.empty(B, S, dtype=torch.int64, device=lp.device)
    with torch.cuda.device(lp.device):
        _lib.check(_lib.load().eavqa_ensemble_select(lp.data_ptr(), tk.data_ptr(), E, B, S, _lib.ptr(skip) if len(skip_ids) else None,
                                                     len(skip_ids), scores.data_ptr(), best.data_ptr(), best_tokens.data_ptr(),
                                                     _lib.current_stream()))
    return best, best_tokens, scores


@torch.no_grad()
def generate_from_ensembles(model, input_ids: torch.Tensor, attention_mask: torch.Tensor, clip_embeddings: torch.Tensor,
                            num_ensembles: int, max_length: int = 10, pad_token_id: Optional[int] = None,
                            eos_token_id: Optional[int] = None, skip_ids: Optional[Sequence[int]] = None,
                            ensemble_one_shots: bool = False):
    """few_shot_vqa_executor.py:293-332 for the GPT-2 prefix model.  ``input_ids`` / ``attention_mask`` ``[B, E, T]``;
    ``clip_embeddings`` ``[B, E, k+1, 1, D]`` (one prompt permutation per member) or, with ``ensemble_one_shots``,
    ``[B, k+1, 1, D]`` from which member i takes images ``[i, -1]`` (``:300-301``).  Returns the winning member's tokens
    per question (``[B, steps]``), the member index and the score table.  ``skip_ids`` defaults to the pad id
    (finished rows are padded, like the T5 pad the reference skips)."""
    E = num_ensembles
    toks, lps = [], []
    for i in range(E):
        clip = clip_embeddings[:, [i, -1]] if ensemble_one_shots else clip_embeddings[:, i]
        t, lp = model.generate(question_tokens=input_ids[:, i], question_mask=attention_mask[:, i], prefix=clip,
                               max_length=max_length, pad_token_id=pad_token_id, eos_token_id=eos_token_id, return_logprobs=True)
        toks.append(t)
        lps.append(lp)
    steps = max(t.shape[1] for t in toks)          # members may stop at different steps (all rows finished): pad them
    pad = pad_token_id if pad_token_id is not None else 0
    tk = torch.full((E, toks[0].shape[0], steps), pad, dtype=torch.int64, device=toks[0].device)
    lp = torch.zeros(E, toks[0].shape[0], steps, dtype=torch.float32, device=toks[0].device)
    for i in range(E):
        tk[i, :, :toks[i].shape[1]] = toks[i]
        lp[i, :, :lps[i].shape[1]] = lps[i]
    if skip_ids is None:
        skip_ids = (pad,) if pad_token_id is not None else ()
    best, best_tokens, scores = ensemble_select(lp, tk, skip_ids)
    return best_tokens, best, scores
