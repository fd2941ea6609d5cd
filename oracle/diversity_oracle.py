"""CPU restatement (test infrastructure, never imported by the product) of the reference's diversity statistics over
the K captions decoded per clip: Div-1 / Div-2 (`utils/div_utils.py:11-29`, compute_div_n) and the global distinct
unigram count gDiv-1 (`utils/div_utils.py:31-44`, compute_global_div_n with n = 1), as used by
`utils/diverse_mutil.py:25-29`.  Works on token ids instead of words: `_convert_idx2sentence`
(`runners/base_runner.py:146-157`) maps ids to words one to one, stops at `<end>` and skips `<start>`, and
`c.split()` recovers the words, so n-gram sets over ids equal n-gram sets over words.  (The PTB tokenizer the
reference applies first -- pycocoevalcap, Java, absent here -- only lower-cases and drops punctuation tokens; the
synthetic vocabularies of this repository have none.)  Pinned on the reference functions themselves:
tests/golden/make_golden.py div -> tests/golden/div_stats.npz.
"""
import numpy as np


def caption_tokens(row, start_idx=1, end_idx=2):
    """runners/base_runner.py:146-157: ids up to the first <end>, <start> skipped."""
    out = []
    for w in row:
        w = int(w)
        if w == end_idx:
            break
        if w == start_idx:
            continue
        out.append(w)
    return out


def compute_div_n(seqs, n=1, start_idx=1, end_idx=2):
    """utils/div_utils.py:11-29.  seqs [clips, K, L] ids -> (mean, per-clip array) of |distinct n-grams| / (1e-6 + #tokens)."""
    aggr = []
    for clip in seqs:
        grams, len_t = set(), 0.0
        for cap in clip:
            tk = caption_tokens(cap, start_idx, end_idx)
            len_t += len(tk)                                              # :19
            grams.update(zip(*[tk[i:] for i in range(n)]))               # find_ngrams, :8-9, :21-22
        aggr.append(float(len(grams)) / (1e-6 + float(len_t)))           # :26
    aggr = np.array(aggr)
    return aggr.mean(), aggr


def compute_global_div_1(seqs, start_idx=1, end_idx=2):
    """utils/div_utils.py:31-44 with n = 1: the number of distinct words over all clips and captions."""
    words = set()
    for clip in seqs:
        for cap in clip:
            words.update(caption_tokens(cap, start_idx, end_idx))
    return float(len(words))


def diversity_stats(seqs, start_idx=1, end_idx=2):
    d1, a1 = compute_div_n(seqs, 1, start_idx, end_idx)
    d2, a2 = compute_div_n(seqs, 2, start_idx, end_idx)
    return {"Div1": d1, "Div2": d2, "gDiv1": compute_global_div_1(seqs, start_idx, end_idx), "div1": a1, "div2": a2}
