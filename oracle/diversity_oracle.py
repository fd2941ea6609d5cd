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


# ---- mBLEU (utils/diverse_mutil.py:35-51) ---------------------------------------------------------------------------
# The scorer is pycocoevalcap's Bleu(4) (pycocoevalcap/bleu/bleu_scorer.py, BleuScorer with option="closest"); the package
# is a third-party dependency that is neither vendored in the reference nor installed here (no version is pinned: the
# reference ships no requirements file), so this restates its published algorithm and is NOT pinned on its output
# ("parity unpinned" for this function; the device kernel is checked against this restatement and against hand-worked cases).
def _ngram_counts(tokens, n=4):
    counts = {}
    for k in range(1, n + 1):
        for i in range(len(tokens) - k + 1):
            g = tuple(tokens[i:i + k])
            counts[g] = counts.get(g, 0) + 1
    return counts


def bleu_stats(cand, refs):
    """cook_refs + cook_test: {testlen, reflen (closest), guess[4], correct[4]} for one candidate and its references."""
    maxcounts, reflens = {}, []
    for r in refs:
        reflens.append(len(r))
        for g, c in _ngram_counts(r).items():
            maxcounts[g] = max(maxcounts.get(g, 0), c)
    testlen = len(cand)
    reflen = min((abs(l - testlen), l) for l in reflens)[1]
    guess = [max(0, testlen - k + 1) for k in range(1, 5)]
    correct = [0, 0, 0, 0]
    for g, c in _ngram_counts(cand).items():
        correct[len(g) - 1] += min(maxcounts.get(g, 0), c)
    return testlen, reflen, guess, correct


def corpus_bleu(stats):
    """BleuScorer.compute_score over a list of per-candidate statistics -> [BLEU-1..4]."""
    import math
    tiny, small = 1e-15, 1e-9
    testlen = sum(s[0] for s in stats); reflen = sum(s[1] for s in stats)
    guess = [sum(s[2][k] for s in stats) for k in range(4)]
    correct = [sum(s[3][k] for s in stats) for k in range(4)]
    bleus, bleu = [], 1.0
    for k in range(4):
        bleu *= (float(correct[k]) + tiny) / (float(guess[k]) + small)
        bleus.append(bleu ** (1.0 / (k + 1)))
    ratio = (testlen + tiny) / (reflen + small)
    if ratio < 1:
        bleus = [b * math.exp(1 - 1 / ratio) for b in bleus]
    return bleus


def mbleu(seqs, start_idx=1, end_idx=2):
    """utils/diverse_mutil.py:35-51: candidate position i against the clip's other captions, averaged over i."""
    caps = [[caption_tokens(c, start_idx, end_idx) for c in clip] for clip in seqs]
    K = len(caps[0])
    all_scrs = []
    for i in range(K):
        stats = [bleu_stats(clip[i], clip[:i] + clip[i + 1:]) for clip in caps]      # :40-45
        all_scrs.append(corpus_bleu(stats))
    return {f"mBLeu_{n + 1}": float(np.mean([s[n] for s in all_scrs])) for n in range(4)}, all_scrs
