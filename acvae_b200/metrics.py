"""Caption post-processing and diversity metrics of the K captions decoded per clip, computed from the
id tensor the sampling loop leaves on the device.

* `diversity_stats`  -- Div1 / Div2 / gDiv1: `utils/div_utils.py:11-44` as called by `utils/diverse_mutil.py:25-29`.
* `mbleu`            -- mBLEU-1..4: `utils/diverse_mutil.py:35-51` (every caption scored against the clip's other captions
                        with pycocoevalcap's `Bleu(4)`, averaged over the K candidate positions).
* `ids_to_sentences`, `predictions_json` -- ids -> words -> the prediction file of the runner
                        (`runners/base_runner.py:146-157` `_convert_idx2sentence`, `:243-293`).
The n-gram statistics run on the device (`acvae_diversity_stats`, `acvae_mbleu_stats`); the string building is host
Python, as in the reference.  The reference applies pycocoevalcap's PTB tokenizer (Java) to the sentences first; on
captions that are already vocabulary words it only lower-cases and drops punctuation tokens, so the metrics here are
defined on the word ids.
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .functional import _dev, _guard, _stream


def _check_seqs(seqs: torch.Tensor) -> torch.Tensor:
    if seqs.dim() != 3 or seqs.dtype != torch.int64:
        raise ValueError("seqs must be an int64 tensor [clips, K, L]")
    if not seqs.is_cuda:
        raise RuntimeError("acvae_b200 runs on CUDA tensors only (no CPU path)")
    return seqs.contiguous()


@_guard
def diversity_stats(seqs: torch.Tensor, vocab_size: int, start_idx: int = 1, end_idx: int = 2) -> dict:
    """seqs [clips, K, L] int64 on the device (e.g. `model(..., method="sample", n_captions=K)["seqs"].view(clips, K, L)`
    or the `[clips, beam, L]` output of `method="dbs"`).  Returns Div1 / Div2 (means over clips, `compute_div_n`),
    gDiv1 (`compute_global_div_n(caps, 1)`: number of distinct words) and the per-clip arrays div1 / div2 (fp64)."""
    seqs = _check_seqs(seqs)
    clips, K, L = seqs.shape
    dev = seqs.device
    div1 = torch.empty(clips, dtype=torch.float64, device=dev)
    div2 = torch.empty(clips, dtype=torch.float64, device=dev)
    flags = torch.zeros(int(vocab_size), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().acvae_diversity_stats(clips, K, L, int(vocab_size), _dev(seqs, torch.int64), int(start_idx), int(end_idx),
                                               div1.data_ptr(), div2.data_ptr(), flags.data_ptr(), _stream()),
               "acvae_diversity_stats")
    return {"Div1": float(div1.mean()), "Div2": float(div2.mean()), "gDiv1": float(flags.sum()), "div1": div1, "div2": div2}


def bleu_from_stats(testlen: int, reflen: int, guess: Sequence[int], correct: Sequence[int]) -> List[float]:
    """Corpus BLEU-1..4 from summed statistics, pycocoevalcap BleuScorer.compute_score (tiny = 1e-15, small = 1e-9,
    brevity penalty exp(1 - 1/ratio) when the candidates are shorter than the closest references)."""
    tiny, small = 1e-15, 1e-9
    bleus, bleu = [], 1.0
    for k in range(4):
        bleu *= (float(correct[k]) + tiny) / (float(guess[k]) + small)
        bleus.append(bleu ** (1.0 / (k + 1)))
    ratio = (testlen + tiny) / (reflen + small)
    if ratio < 1:
        bleus = [b * math.exp(1 - 1 / ratio) for b in bleus]
    return bleus


@_guard
def mbleu(seqs: torch.Tensor, start_idx: int = 1, end_idx: int = 2) -> dict:
    """mBLEU of `eval_div_stats` (utils/diverse_mutil.py:35-51): for i in range(K): candidates = caption i of every clip,
    references = the clip's other K-1 captions, `Bleu(4).compute_score` over all clips; the K score vectors are averaged.
    Lower = more diverse.  seqs [clips, K, L] int64 on the device, K >= 2.  Returns {"mBLeu_1".."mBLeu_4", "per_candidate"}."""
    seqs = _check_seqs(seqs)
    clips, K, L = seqs.shape
    stats = torch.empty(clips, K, 10, dtype=torch.int32, device=seqs.device)
    _lib.check(_lib.lib().acvae_mbleu_stats(clips, K, L, _dev(seqs, torch.int64), int(start_idx), int(end_idx), stats.data_ptr(),
                                            _stream()), "acvae_mbleu_stats")
    tot = stats.to(torch.int64).sum(0).cpu().tolist()              # [K][10]: corpus sums per candidate position
    per = [bleu_from_stats(t[0], t[1], t[2:6], t[6:10]) for t in tot]
    out = {f"mBLeu_{n + 1}": sum(p[n] for p in per) / K for n in range(4)}
    out["per_candidate"] = per
    out["stats"] = stats
    return out


def _one_sentence(row, idx2word, zh, start_word, end_word):
    cand = []
    for w in row:
        word = idx2word[int(w)]
        if word == end_word:
            break
        if word == start_word:
            continue
        cand.append(word)
    return cand if zh else " ".join(cand)


def ids_to_sentences(seqs, idx2word, zh: bool = False, start_word: str = "<start>", end_word: str = "<end>"):
    """`_convert_idx2sentence` (runners/base_runner.py:146-157) over a batch: stop at <end>, skip <start>, join with spaces
    (a token list when `zh`).  seqs: [N, L] or [N, K, L] ids (tensor, array or nested lists); returns a list (of lists, for
    [N, K, L]) of sentences."""
    if torch.is_tensor(seqs):
        seqs = seqs.detach().cpu().tolist()
    elif hasattr(seqs, "tolist"):
        seqs = seqs.tolist()
    if seqs and seqs[0] and isinstance(seqs[0][0], (list, tuple)):
        return [[_one_sentence(r, idx2word, zh, start_word, end_word) for r in clip] for clip in seqs]
    return [_one_sentence(r, idx2word, zh, start_word, end_word) for r in seqs]


def predictions_json(keys: Sequence[str], seqs, idx2word, zh: bool = False, path: Optional[str] = None) -> Dict:
    """The prediction file of `BaseRunner.evaluate` (runners/base_runner.py:243-293): {"predictions": [...]} with one entry
    per clip -- {"filename", "caption", "tokens"} for a single caption ([N, L] ids, or K = 1), {"filename", "captions":
    [{"caption", "cap_id", "tokens"}, ...]} for several ([N, K, L]: K-caption sampling, `group_nbest` diverse beam search).
    Written to `path` when given (json.dump(..., indent=4) as the reference)."""
    if torch.is_tensor(seqs):
        seqs = seqs.detach().cpu().tolist()
    elif hasattr(seqs, "tolist"):
        seqs = seqs.tolist()
    multi = bool(seqs) and bool(seqs[0]) and isinstance(seqs[0][0], (list, tuple))
    sents = ids_to_sentences(seqs, idx2word, zh)
    pred_data = []
    for key, pred in zip(keys, sents):
        preds = pred if multi else [pred]
        if len(preds) > 1:                                                         # base_runner.py:272-283
            caps = [{"caption": "".join(p) if zh else p, "cap_id": i, "tokens": " ".join(p) if zh else p} for i, p in enumerate(preds)]
            pred_data.append({"filename": key, "captions": caps})
        else:                                                                      # :284-289
            p0 = preds[0]
            pred_data.append({"filename": key, "caption": "".join(p0) if zh else p0, "tokens": " ".join(p0) if zh else p0})
    doc = {"predictions": pred_data}
    if path is not None:
        with open(path, "w") as f:
            json.dump(doc, f, indent=4)
    return doc
