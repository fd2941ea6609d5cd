"""Deterministic synthetic Clotho-shaped inputs and hot-path weights.

Everything here is generated with `numpy.random.RandomState` (a frozen legacy
stream, identical on every platform) so that the build container -- where the
reference runs and the golden vectors are minted -- and the GPU box -- where
only this repo exists -- see bit-identical inputs and weights from a seed.

Shapes follow SURVEY.md section 8d: the hot-path-only form
(`audio_embeds = |N(0,1)| [N,Te,Eenc]`, `audio_embeds_lens`, captions
`[1, U{4..V-1}.., 2, 0..]` stored float32 and sorted by length descending as
`datasets/caption_dataset.py:278-318` does) and the noise tensors the
reference draws from the CPU generator.
Weight shapes/keys are the reference's `state_dict`.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import numpy as np


@dataclasses.dataclass(frozen=True)
class Dims:
    N: int = 32        # clips (or sampled sequences) in the batch
    Te: int = 62       # encoder frames after 16x downsampling (1000 // 16)
    L: int = 20        # padded caption length incl. <start>/<end>; T = L-1 steps
    E: int = 256       # word-embedding = latent z = prior hidden size
    H: int = 256       # decoder GRU hidden size (== E, decoder.py:171)
    A: int = 256       # attention size
    Hq: int = 256      # posterior GRU hidden size per direction
    V: int = 4400      # vocabulary
    Eenc: int = 512    # audio encoder output width (Cnn10)

    @property
    def T(self) -> int:
        return self.L - 1


TINY = Dims(N=4, Te=9, L=7, E=32, H=32, A=32, Hq=32, V=50, Eenc=48)
CFG0 = Dims(N=4, Te=62, L=20)                       # BASELINE.json configs[0] hot path
CFG1 = Dims(N=32, Te=62, L=20)                      # configs[1]: default training batch
STRESS = Dims(N=128, Te=187, L=30, V=5000)          # configs[4]


def _uniform(rs, shape, bound):
    return rs.uniform(-bound, bound, size=shape).astype(np.float32)


def _xavier(rs, out_f, in_f):
    return _uniform(rs, (out_f, in_f), math.sqrt(6.0 / (in_f + out_f)))


def _kaiming(rs, out_f, in_f):
    # nn.init.kaiming_uniform_ default (a=0): bound = sqrt(6 / fan_in)
    return _uniform(rs, (out_f, in_f), math.sqrt(6.0 / in_f))


def make_params(d: Dims, seed: int = 1, variant: str = "hybrid") -> Dict[str, np.ndarray]:
    """Weights with the reference's state_dict keys/shapes and initialiser
    scales.  `variant="vae"` swaps the posterior head
    for PosteriorRNN's `qnet.mean_log_out` and drops the global head."""
    rs = np.random.RandomState(seed)
    E, H, A, Hq, V, Eenc = d.E, d.H, d.A, d.Hq, d.V, d.Eenc
    p: Dict[str, np.ndarray] = {}
    p["ln.weight"] = _xavier(rs, E, Eenc)
    p["ln.bias"] = _uniform(rs, (E,), 1.0 / math.sqrt(Eenc))
    # posterior
    p["qnet.word_embedding.weight"] = rs.standard_normal((V, E)).astype(np.float32)
    k = 1.0 / math.sqrt(Hq)
    for suf in ("", "_reverse"):
        p[f"qnet.network.weight_ih_l0{suf}"] = _uniform(rs, (3 * Hq, E), k)
        p[f"qnet.network.weight_hh_l0{suf}"] = _uniform(rs, (3 * Hq, Hq), k)
        p[f"qnet.network.bias_ih_l0{suf}"] = _uniform(rs, (3 * Hq,), k)
        p[f"qnet.network.bias_hh_l0{suf}"] = _uniform(rs, (3 * Hq,), k)
    if variant == "hybrid":
        p["qnet.token_mean_log.weight"] = _xavier(rs, 2 * E, 2 * Hq)
        p["qnet.token_mean_log.bias"] = _uniform(rs, (2 * E,), 0.05)
    else:
        p["qnet.mean_log_out.weight"] = _xavier(rs, 2 * E, 2 * Hq + E)
        p["qnet.mean_log_out.bias"] = _uniform(rs, (2 * E,), 0.05)
    # prior
    p["pnet.word_embedding.weight"] = rs.standard_normal((V, E)).astype(np.float32)
    # the prior's attention width is E, not attn_size (text_encoder.py:225)
    p["pnet.word_attn.h2attn.weight"] = _xavier(rs, E, 2 * E)
    p["pnet.word_attn.h2attn.bias"] = _uniform(rs, (E,), 0.05)
    p["pnet.word_attn.v"] = rs.standard_normal((E,)).astype(np.float32)
    k = 1.0 / math.sqrt(E)
    p["pnet.network.weight_ih_l0"] = _uniform(rs, (4 * E, 3 * E), k)
    p["pnet.network.weight_hh_l0"] = _uniform(rs, (4 * E, E), k)
    p["pnet.network.bias_ih_l0"] = _uniform(rs, (4 * E,), k)
    p["pnet.network.bias_hh_l0"] = _uniform(rs, (4 * E,), k)
    p["pnet.mean_log_out.weight"] = _xavier(rs, 2 * E, E)
    p["pnet.mean_log_out.bias"] = _uniform(rs, (2 * E,), 0.05)
    # decoder
    p["decoder.word_embeddings.weight"] = _kaiming(rs, V, E)
    p["decoder.attn.h2attn.weight"] = _kaiming(rs, A, H + E)
    p["decoder.attn.h2attn.bias"] = _uniform(rs, (A,), 1.0 / math.sqrt(H + E))
    p["decoder.attn.v"] = rs.standard_normal((A,)).astype(np.float32)
    k = 1.0 / math.sqrt(H)
    p["decoder.model.weight_ih_l0"] = _uniform(rs, (3 * H, 3 * E), k)
    p["decoder.model.weight_hh_l0"] = _uniform(rs, (3 * H, H), k)
    p["decoder.model.bias_ih_l0"] = _uniform(rs, (3 * H,), k)
    p["decoder.model.bias_hh_l0"] = _uniform(rs, (3 * H,), k)
    p["decoder.classifier.weight"] = _kaiming(rs, V, H)
    p["decoder.classifier.bias"] = _uniform(rs, (V,), 1.0 / math.sqrt(H))
    if variant == "hybrid":
        p["mean_log_out.weight"] = _xavier(rs, 2 * E, E)
        p["mean_log_out.bias"] = _uniform(rs, (2 * E,), 1.0 / math.sqrt(E))
    return p


def make_batch(d: Dims, seed: int = 1, min_cap_len: Optional[int] = None,
               sample_steps: int = 0, beam: int = 0, cap_lens_override: Optional[np.ndarray] = None,
               dbs_groups: int = 0) -> Dict[str, np.ndarray]:
    """One synthetic batch in hot-path-only form plus all injected noise.

    Returns float32 `audio_embeds [N,Te,Eenc]`, int64 `mem_lens [N]`,
    float32 `caps [N,L]`, int64 `cap_lens [N]` (sorted descending, first == L),
    `eps_q [N,T,E]`, `eps_p [T,N,E]`, `eps_q_steps [T,N,E]` (AR posterior),
    `u_tf [T]`, `u_dis [T]` (the uniforms behind the per-step teacher-forcing
    / dis_ratio decisions), and for sampling `eps_s [S,N,E]`, `u_s [S,N,V]`,
    for beam search `eps_b [N,S,beam,E]`, for diverse beam search (`beam` hypotheses in `dbs_groups` groups)
    `eps_dbs [N, S+G-1, G, beam//G, E]` (global step, group; only the active (t, g) pairs are consumed).
    """
    rs = np.random.RandomState(seed + 1000)
    N, Te, L, E, V = d.N, d.Te, d.L, d.E, d.V
    T = L - 1
    out: Dict[str, np.ndarray] = {}
    out["audio_embeds"] = np.abs(rs.standard_normal((N, Te, d.Eenc))).astype(np.float32)
    lo = max(1, (Te * 480) // 1000)
    mem_lens = rs.randint(lo, Te + 1, size=N).astype(np.int64)
    mem_lens[rs.randint(0, N)] = Te
    out["mem_lens"] = mem_lens
    lo_c = min(L, min_cap_len if min_cap_len is not None else min(8, max(3, L // 2)))
    cap_lens = rs.randint(lo_c, L + 1, size=N).astype(np.int64)
    cap_lens[0] = L
    cap_lens = np.sort(cap_lens)[::-1].copy()
    if cap_lens_override is not None:       # share one length profile across a pool of batches
        cap_lens = np.asarray(cap_lens_override, dtype=np.int64).copy()
    caps = np.zeros((N, L), dtype=np.float32)
    for n in range(N):
        ln = int(cap_lens[n])
        caps[n, 0] = 1.0
        caps[n, 1:ln - 1] = rs.randint(4, V, size=ln - 2).astype(np.float32)
        caps[n, ln - 1] = 2.0
    out["caps"] = caps
    out["cap_lens"] = cap_lens
    out["eps_q"] = rs.standard_normal((N, T, E)).astype(np.float32)
    out["eps_p"] = rs.standard_normal((T, N, E)).astype(np.float32)
    out["eps_q_steps"] = rs.standard_normal((T, N, E)).astype(np.float32)
    out["u_tf"] = rs.uniform(size=T).astype(np.float64)
    out["u_dis"] = rs.uniform(size=T).astype(np.float32)
    if sample_steps:
        out["eps_s"] = rs.standard_normal((sample_steps, N, E)).astype(np.float32)
        out["u_s"] = rs.uniform(size=(sample_steps, N, V)).astype(np.float32)
    if beam:
        out["eps_b"] = rs.standard_normal((N, sample_steps, beam, E)).astype(np.float32)
    if dbs_groups:
        G = dbs_groups
        out["eps_dbs"] = np.random.RandomState(seed + 7000).standard_normal(
            (N, sample_steps + G - 1, G, beam // G, E)).astype(np.float32)
    return out
