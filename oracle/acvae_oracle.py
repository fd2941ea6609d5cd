"""CPU oracle for the AC-VAE per-token latent word-decoding hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this file, and
only as the checker or the timed CPU baseline -- never as the product path.
The product (`acvae_b200`) never imports `oracle/`.

What this is
------------
A plain restatement, in explicit tensor arithmetic on the CPU, of the
reference's algorithm for the hot path named by BASELINE.json `north_star`
(SURVEY.md section 8a rows H1..H11).  It follows the reference's own op
structure (un-factorised attention, one decode step at a time) so that it can
also stand in as the reference's CPU cost model in `bench.py --impl reference`.
Every function cites the reference file:line it follows (paths relative to
the upstream repo root).

Why torch-on-CPU rather than numpy/C: the path is floating point and the
parity bar includes *gradients* (SURVEY.md 8c / A.9); the oracle's backward is
torch autograd over this explicit forward, in fp32 or fp64.

Third-party arithmetic: the reference calls `torch.nn.GRU`, `torch.nn.LSTM`,
`torch.nn.Linear`, `softmax`, `log_softmax` from PyTorch (no version pinned by
the reference; this image has torch 2.11.0).  Their published cell equations
(PyTorch docs: GRU gates r,z,n with n = tanh(W_in x + b_in + r*(W_hn h + b_hn));
LSTM gates i,f,g,o) are restated here explicitly (`gru_cell`, `lstm_cell`).

Pinning
-------
The reference ships no tests and no golden vectors, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build
container by `tests/golden/make_golden.py` (which imports /root/reference via
`oracle/ref_loader.py`, injects identical noise, and writes
`tests/golden/*.npz`).  `tests/test_oracle_golden.py` checks this file against
those fixtures on every CPU test run.

Noise / RNG: the reference draws `torch.randn(N,T,E)` once for
the posterior, then per step `random.random()` (teacher forcing), `randn(N,E)`
(prior) and optionally `torch.rand(1)` (dis_ratio).  The oracle takes all of
those as explicit inputs: `eps_q [N,T,E]`, `eps_p [T,N,E]`, `tf_flags [T]`
(True = feed caps[:,t]), `dis_flags [T]` (True = feed the prior's z).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

PAD_IDX, START_IDX, END_IDX = 0, 1, 2  # models/word_model.py:19-21
NEG_FILL = -1e10                       # models/attn_model.py:41

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------
# cells (PyTorch nn.GRU / nn.LSTM published equations)
# ----------------------------------------------------------------------------
def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """One nn.GRU step.  Call sites: models/text_encoder.py:166-172,189 and
    models/decoder.py:39-44,194."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    H = h.shape[-1]
    r = torch.sigmoid(gi[..., :H] + gh[..., :H])
    z = torch.sigmoid(gi[..., H:2 * H] + gh[..., H:2 * H])
    n = torch.tanh(gi[..., 2 * H:] + r * gh[..., 2 * H:])
    return (1.0 - z) * n + z * h


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """One nn.LSTM step.  Call site: models/text_encoder.py:229-235,253."""
    g = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    H = h.shape[-1]
    i = torch.sigmoid(g[..., :H])
    f = torch.sigmoid(g[..., H:2 * H])
    gg = torch.tanh(g[..., 2 * H:3 * H])
    o = torch.sigmoid(g[..., 3 * H:])
    c2 = f * c + i * gg
    h2 = o * torch.tanh(c2)
    return h2, c2


# ----------------------------------------------------------------------------
# attention (models/attn_model.py:20-46)
# ----------------------------------------------------------------------------
def attention(h_dec, h_enc, src_lens, w, b, v):
    """Seq2SeqAttention.forward, in the reference's un-factorised form.

    h_dec [N,Dq], h_enc [N,Te,E], src_lens [N] -> ctx [N,E], weights [N,Te].
    `cat((h_dec, h_enc))` puts the QUERY columns first (attn_model.py:31).
    """
    N, Te, _ = h_enc.shape
    q = h_dec.unsqueeze(1).expand(N, Te, h_dec.shape[-1])        # :29
    attn_in = torch.cat((q, h_enc), dim=-1)                      # :31
    attn_out = torch.tanh(attn_in @ w.t() + b)                   # :32
    score = attn_out @ v                                         # :34-36
    idx = torch.arange(Te).unsqueeze(0)
    mask = idx < torch.as_tensor(src_lens).view(-1, 1)           # :38-39
    score = score.masked_fill(~mask, NEG_FILL)                   # :41
    weights = torch.softmax(score, dim=-1)                       # :42
    ctx = (weights.unsqueeze(1) @ h_enc).squeeze(1)              # :44
    return ctx, weights


# ----------------------------------------------------------------------------
# pooling helpers (utils/train_util.py:198-231)
# ----------------------------------------------------------------------------
def length_mask(lens, T):
    return torch.arange(T).unsqueeze(0) < torch.as_tensor(lens).view(-1, 1)


def mean_with_lens(x, lens):
    """utils/train_util.py:207-217."""
    lens_t = torch.as_tensor(lens)
    m = length_mask(lens_t, x.shape[1]).to(x.dtype)
    return (x * m.unsqueeze(-1)).sum(1) / lens_t.to(x.dtype).unsqueeze(1)


def max_with_lens(x, lens):
    """utils/train_util.py:220-231."""
    m = length_mask(lens, x.shape[1])
    y = x.masked_fill(~m.unsqueeze(-1), float("-inf"))
    return y.max(1).values


# ----------------------------------------------------------------------------
# posterior q(z_t | caption)   (models/text_encoder.py:156-216)
# ----------------------------------------------------------------------------
def bigru_packed(x, lens, p: Params, prefix: str):
    """Packed bidirectional GRU with zero initial state; padded outputs are 0
    (pack_padded_sequence / pad_packed_sequence, text_encoder.py:188-191)."""
    N, T, _ = x.shape
    Hq = p[prefix + "weight_hh_l0"].shape[1]
    lens_t = torch.as_tensor(lens)
    out_f: List[torch.Tensor] = []
    h = x.new_zeros(N, Hq)
    for t in range(T):
        hn = gru_cell(x[:, t], h, p[prefix + "weight_ih_l0"], p[prefix + "weight_hh_l0"],
                      p[prefix + "bias_ih_l0"], p[prefix + "bias_hh_l0"])
        act = (t < lens_t).to(x.dtype).unsqueeze(1)
        h = act * hn + (1 - act) * h
        out_f.append(act * hn)
    out_b: List[Optional[torch.Tensor]] = [None] * T
    h = x.new_zeros(N, Hq)
    for t in range(T - 1, -1, -1):
        hn = gru_cell(x[:, t], h, p[prefix + "weight_ih_l0_reverse"], p[prefix + "weight_hh_l0_reverse"],
                      p[prefix + "bias_ih_l0_reverse"], p[prefix + "bias_hh_l0_reverse"])
        act = (t < lens_t).to(x.dtype).unsqueeze(1)
        h = act * hn + (1 - act) * h   # stays 0 until the row's last valid token
        out_b[t] = act * hn
    return torch.cat([torch.stack(out_f, 1), torch.stack(out_b, 1)], dim=-1)  # [N,T,2Hq]


def posterior_hybrid(p: Params, caps, cap_lens, eps_q):
    """PosteriorRNN_hybrid.forward (text_encoder.py:182-216)."""
    E = p["qnet.token_mean_log.weight"].shape[0] // 2
    ids = caps[:, :-1].long()                                    # :184
    lens = torch.as_tensor(cap_lens) - 1                         # :186
    T = int(lens.max())
    x = p["qnet.word_embedding.weight"][ids][:, :T]
    ho = bigru_packed(x, lens, p, "qnet.network.")               # :188-191  [N,T,2Hq]
    ml = ho @ p["qnet.token_mean_log.weight"].t() + p["qnet.token_mean_log.bias"]  # :193
    q_means, q_logs = ml[..., :E], ml[..., E:]                   # :194-195
    q_z = eps_q * torch.exp(0.5 * q_logs) + q_means              # :196-197
    utt = mean_with_lens(ho, lens) + max_with_lens(ho, lens)     # :199-201
    return {"q_means": q_means, "q_logs": q_logs, "q_z": q_z,
            "q_means_utt": utt, "ho": ho}


def posterior_ar(p: Params, caps, cap_lens, eps_steps):
    """PosteriorRNN.forward (text_encoder.py:121-154): autoregressive posterior
    used by the secondary `VAEModel` variant.
    eps_steps [T,N,E]: one draw per step (text_encoder.py:143)."""
    E = p["qnet.mean_log_out.weight"].shape[0] // 2
    ids = caps[:, :-1].long()
    lens = torch.as_tensor(cap_lens) - 1
    T = int(lens.max())
    x = p["qnet.word_embedding.weight"][ids]
    L1 = x.shape[1]
    ho = bigru_packed(x[:, :T], lens, p, "qnet.network.")
    N = x.shape[0]
    means = x.new_zeros(N, L1, E)
    logs = x.new_zeros(N, L1, E)
    zs = x.new_zeros(N, L1, E)
    z_prev = x.new_zeros(N, E)
    ms, ls, zz = [], [], []
    for t in range(T):                                           # :137
        ml = torch.cat([ho[:, t], z_prev], 1) @ p["qnet.mean_log_out.weight"].t() + p["qnet.mean_log_out.bias"]
        mean, log = ml[:, :E], ml[:, E:]
        z_t = eps_steps[t] * torch.exp(0.5 * log) + mean         # :143-144
        ms.append(mean); ls.append(log); zz.append(z_t)
        z_prev = z_t
    means = torch.cat([torch.stack(ms, 1), means[:, T:]], 1)
    logs = torch.cat([torch.stack(ls, 1), logs[:, T:]], 1)
    zs = torch.cat([torch.stack(zz, 1), zs[:, T:]], 1)
    return {"q_means": means, "q_logs": logs, "q_z": zs}


# ----------------------------------------------------------------------------
# prior step  (models/text_encoder.py:247-268)
# ----------------------------------------------------------------------------
def prior_step(p: Params, word, mem, mem_lens, h, c, last_z, eps):
    E = p["pnet.mean_log_out.weight"].shape[0] // 2
    xe = p["pnet.word_embedding.weight"][word]                   # :249
    ctx, attn_w = attention(xe, mem, mem_lens, p["pnet.word_attn.h2attn.weight"],
                            p["pnet.word_attn.h2attn.bias"], p["pnet.word_attn.v"])  # :251
    u = torch.cat([xe, ctx, last_z], dim=-1)                     # :253
    h2, c2 = lstm_cell(u, h, c, p["pnet.network.weight_ih_l0"], p["pnet.network.weight_hh_l0"],
                       p["pnet.network.bias_ih_l0"], p["pnet.network.bias_hh_l0"])
    ml = h2 @ p["pnet.mean_log_out.weight"].t() + p["pnet.mean_log_out.bias"]  # :255
    mean, log = ml[:, :E], ml[:, E:]                             # :257-258
    z = eps * torch.exp(0.5 * log) + mean                        # :259-262
    return {"mean": mean, "log": log, "z": z, "h": h2, "c": c2, "attn_w": attn_w}


# ----------------------------------------------------------------------------
# decoder step  (models/decoder.py:175-203)
# ----------------------------------------------------------------------------
def decoder_step(p: Params, word, mem, mem_lens, h, z):
    de = p["decoder.word_embeddings.weight"][word]               # :183 (dropout p=0)
    ctx, attn_w = attention(h, mem, mem_lens, p["decoder.attn.h2attn.weight"],
                            p["decoder.attn.h2attn.bias"], p["decoder.attn.v"])    # :186
    x = torch.cat([de, ctx, z], dim=-1)                          # :188
    h2 = gru_cell(x, h, p["decoder.model.weight_ih_l0"], p["decoder.model.weight_hh_l0"],
                  p["decoder.model.bias_ih_l0"], p["decoder.model.bias_hh_l0"])   # :194
    logits = h2 @ p["decoder.classifier.weight"].t() + p["decoder.classifier.bias"]  # :199
    return {"h": h2, "logits": logits, "attn_w": attn_w, "rnn_input": x}


# ----------------------------------------------------------------------------
# next-word selection  (models/word_model.py:173-207)
# ----------------------------------------------------------------------------
def gumbel_from_uniform(u, eps=1e-20):
    """word_model.py:188-190."""
    return -torch.log(-torch.log(u + eps) + eps)


def sample_next_word(logits, method="greedy", temp=1.0, u=None):
    """Returns (w_t [N] int64, logprob [N]).
    greedy: word_model.py:178-179.  gumbel: :187-195 with injected uniform `u`.
    sample: the reference calls torch.multinomial(exp(logp/temp)) (:197-198),
    which has no reproducible stream; the oracle DEFINES it as the equivalent
    Gumbel-max draw argmax(logp/temp + G(u)) with injected uniform `u`
    (same distribution; `make_golden.py` patches torch.multinomial likewise).
    """
    logp = torch.log_softmax(logits, dim=1)                      # :177
    if method == "greedy":
        lp, w = torch.max(logp, 1)
    elif method == "gumbel":
        y = torch.log_softmax((logp + gumbel_from_uniform(u)) / temp, dim=-1)
        w = torch.max(y, 1).indices
        lp = logp.gather(1, w.unsqueeze(-1)).squeeze(1)
    else:
        w = torch.max(logp / temp + gumbel_from_uniform(u), 1).indices
        lp = logp.gather(1, w.unsqueeze(-1)).squeeze(1)
    return w.long(), lp


# ----------------------------------------------------------------------------
# memory projection H1 (models/vae_model.py:743-744)
# ----------------------------------------------------------------------------
def project_memory(p: Params, audio_embeds):
    if "ln.weight" in p:
        return audio_embeds @ p["ln.weight"].t() + p["ln.bias"]
    return audio_embeds


# ----------------------------------------------------------------------------
# training forward (models/vae_model.py:700-869, 871-878)
# ----------------------------------------------------------------------------
def train_forward(p: Params, audio_embeds, mem_lens, caps, cap_lens, eps_q, eps_p,
                  tf_flags: Optional[Sequence[bool]] = None,
                  dis_flags: Optional[Sequence[bool]] = None,
                  variant: str = "hybrid", eps_q_steps=None):
    """Hybrid_VAEModel.forward 4-input branch with the encoder output given.

    variant="vae": VAEModel (vae_model.py:12-365) with the AR posterior
    (`eps_q_steps [T,N,E]`), no global head, extra `rnn_input` output.
    """
    mem = project_memory(p, audio_embeds)                        # :743-744
    if variant == "hybrid":
        q = posterior_hybrid(p, caps, cap_lens, eps_q)           # :745
    else:
        q = posterior_ar(p, caps, cap_lens, eps_q_steps)
    lens = torch.as_tensor(cap_lens) - 1
    T = int(lens.max())                                          # :703
    N = mem.shape[0]
    E = p["pnet.mean_log_out.weight"].shape[0] // 2
    H = p["decoder.model.weight_hh_l0"].shape[1]
    tf_flags = [True] * T if tf_flags is None else list(tf_flags)
    dis_flags = [False] * T if dis_flags is None else list(dis_flags)
    h_d = mem.new_zeros(N, H)                                    # decoder.py:94-98
    h_p = mem.new_zeros(N, E); c_p = mem.new_zeros(N, E)         # text_encoder.py:240-245
    last_z = mem.new_zeros(N, E)                                 # vae_model.py:839-842
    ids = caps.long()
    seqs, logits, outs, pm, pl, pz, aw, lps, rin = [], [], [], [], [], [], [], [], []
    for t in range(T):                                           # :710
        if tf_flags[t]:                                          # :826-827
            word = ids[:, t]
        elif t == 0:
            word = torch.full((N,), START_IDX, dtype=torch.long)
        else:
            word = seqs[-1]                                      # :832
        pr = prior_step(p, word, mem, mem_lens, h_p, c_p, last_z, eps_p[t])  # :797
        z = pr["z"] if dis_flags[t] else q["q_z"][:, t]          # :800-806
        de = decoder_step(p, word, mem, mem_lens, h_d, z)        # :810
        w_t, lp = sample_next_word(de["logits"], "greedy")       # :814
        h_d = de["h"]; h_p, c_p = pr["h"], pr["c"]
        last_z = pr["z"]                                         # :869 (prior's own sample)
        seqs.append(w_t); logits.append(de["logits"]); outs.append(de["h"])
        pm.append(pr["mean"]); pl.append(pr["log"]); pz.append(pr["z"])
        aw.append(de["attn_w"]); lps.append(lp); rin.append(de["rnn_input"])
    out = {
        "seqs": torch.stack(seqs, 1), "logits": torch.stack(logits, 1),
        "outputs": torch.stack(outs, 1), "sampled_logprobs": torch.stack(lps, 1),
        "attn_weights": torch.stack(aw, 2)[:, :int(max(mem_lens))],
        "p_means": torch.stack(pm, 1), "p_logs": torch.stack(pl, 1), "p_z": torch.stack(pz, 1),
        "q_means": q["q_means"], "q_logs": q["q_logs"], "q_z": q["q_z"],
        "state": h_d.unsqueeze(0), "hiddens_state": (h_p.unsqueeze(0), c_p.unsqueeze(0)),
        "last_z": last_z, "mem": mem,
    }
    if variant == "hybrid":
        pool = mean_with_lens(out["outputs"], lens) + max_with_lens(out["outputs"], lens)  # :722-724
        out["p_means_utt"] = pool @ p["mean_log_out.weight"].t() + p["mean_log_out.bias"]   # :726
        out["q_means_utt"] = q["q_means_utt"]
        out["p_logs_utt"] = None; out["q_logs_utt"] = None
    else:
        out["rnn_input"] = torch.stack(rin, 1)                   # vae_model.py:187
    return out


# ----------------------------------------------------------------------------
# losses  (utils/train_util.py:234-266, runners/pytorch_runner_vae.py:89-98,315-320)
# ----------------------------------------------------------------------------
def pack_rows(x, lens):
    """`pack_padded_sequence(x, lens, batch_first=True).data` for lens sorted
    descending: time-major concatenation of the valid rows."""
    lens_t = torch.as_tensor(lens)
    rows = []
    for t in range(int(lens_t.max())):
        n_valid = int((lens_t > t).sum())
        rows.append(x[:n_valid, t])
    return torch.cat(rows, 0)


def label_smoothing_loss(logit, target, classes, smoothing):
    """LabelSmoothingLoss.forward (train_util.py:244-251)."""
    pred = torch.log_softmax(logit, dim=-1)
    true = torch.full_like(pred, smoothing / (classes - 1))
    true.scatter_(1, target.long().unsqueeze(1), 1.0 - smoothing)
    return torch.mean(torch.sum(-true * pred, dim=-1))


def normal_kl_loss(mu1, lv1, mu2, lv2):
    """Normal_kl_loss.forward (train_util.py:259-266): sum over d, mean over
    ALL N*T positions, padding included."""
    v1, v2 = torch.exp(lv1), torch.exp(lv2)
    kl = lv2 / 2.0 - lv1 / 2.0 + (v1 + (mu1 - mu2) ** 2.0) / (2.0 * v2) - 0.5
    return kl.sum(-1).mean()


def train_loss(out, caps, cap_lens, vocab_size, smoothing=0.1, kl_weight=0.5, alpha=1.0,
               global_loss: Optional[str] = "MSE"):
    """Loss composition of Runner.train (pytorch_runner_vae.py:89-98, 315-320)."""
    lens = torch.as_tensor(cap_lens) - 1
    targets = pack_rows(caps[:, 1:], lens)                       # :89-90
    packed = pack_rows(out["logits"], lens)                      # :94-95
    ce = label_smoothing_loss(packed, targets, vocab_size, smoothing)
    kl = normal_kl_loss(out["q_means"], out["q_logs"], out["p_means"], out["p_logs"])
    terms = {"ce": ce, "kl": kl}
    loss = ce + kl_weight * kl                                   # :315
    if global_loss == "MSE" and "p_means_utt" in out:            # :316-318
        g = torch.mean((out["q_means_utt"] - out["p_means_utt"]) ** 2)
        terms["global"] = g
        loss = loss + alpha * g
    terms["loss"] = loss
    return terms


# ----------------------------------------------------------------------------
# inference: greedy / sample / gumbel with prior latents
# (models/vae_model.py:700-720, 880-894)
# ----------------------------------------------------------------------------
def inference_forward(p: Params, audio_embeds, mem_lens, eps_p, method="greedy",
                      max_length=20, temp=1.0, u_steps=None):
    """Returns seqs [N,max_length] (END-filled after stop), sampled_logprobs,
    p_means/p_logs/p_z for the executed steps, and n_steps executed."""
    mem = project_memory(p, audio_embeds)
    N = mem.shape[0]
    E = p["pnet.mean_log_out.weight"].shape[0] // 2
    H = p["decoder.model.weight_hh_l0"].shape[1]
    h_d = mem.new_zeros(N, H); h_p = mem.new_zeros(N, E); c_p = mem.new_zeros(N, E)
    last_z = mem.new_zeros(N, E)
    seqs = torch.full((N, max_length), END_IDX, dtype=torch.long)
    lps = mem.new_zeros(N, max_length)
    pm, pl, pz, outs = [], [], [], []
    unfinished = torch.ones(N, dtype=torch.bool)
    word = torch.full((N,), START_IDX, dtype=torch.long)
    n_steps = 0
    for t in range(max_length):
        pr = prior_step(p, word, mem, mem_lens, h_p, c_p, last_z, eps_p[t])
        de = decoder_step(p, word, mem, mem_lens, h_d, pr["z"])  # :808
        w_t, lp = sample_next_word(de["logits"], method, temp,
                                   None if u_steps is None else u_steps[t])
        h_d = de["h"]; h_p, c_p = pr["h"], pr["c"]; last_z = pr["z"]
        pm.append(pr["mean"]); pl.append(pr["log"]); pz.append(pr["z"]); outs.append(de["h"])
        lps[:, t] = lp
        unfinished = unfinished & (w_t != END_IDX)               # :713-717
        w_t = torch.where(unfinished, w_t, torch.full_like(w_t, END_IDX))  # :718
        seqs[:, t] = w_t
        word = w_t
        n_steps = t + 1
        if int(unfinished.sum()) == 0:                           # :719-720
            break
    return {"seqs": seqs, "sampled_logprobs": lps, "n_steps": n_steps,
            "p_means": torch.stack(pm, 1), "p_logs": torch.stack(pl, 1),
            "p_z": torch.stack(pz, 1), "outputs": torch.stack(outs, 1)}


# ----------------------------------------------------------------------------
# beam search with prior latents (models/vae_model.py:896-995)
# ----------------------------------------------------------------------------
def beam_search(p: Params, audio_embeds, mem_lens, eps_beam, beam_size=3, max_length=20):
    """Hybrid_VAEModel.beam_search: per clip, `beam_size` hypotheses, every
    step draws prior noise per beam (eps_beam [N_clips, max_length, beam, E]),
    always runs `max_length` steps (no done-beam bookkeeping in this override)
    and returns the top beam (vae_model.py:986)."""
    mem_all = project_memory(p, audio_embeds)
    V = p["decoder.classifier.weight"].shape[0]
    E = p["pnet.mean_log_out.weight"].shape[0] // 2
    H = p["decoder.model.weight_hh_l0"].shape[1]
    N = mem_all.shape[0]
    out_seqs = torch.full((N, max_length), END_IDX, dtype=torch.long)
    for i in range(N):                                           # :901
        mem = mem_all[i].unsqueeze(0).repeat(beam_size, 1, 1)    # :954
        lens = torch.as_tensor(mem_lens)[i].repeat(beam_size)
        h_d = mem.new_zeros(beam_size, H); h_p = mem.new_zeros(beam_size, E)
        c_p = mem.new_zeros(beam_size, E); last_z = mem.new_zeros(beam_size, E)
        word = torch.full((beam_size,), START_IDX, dtype=torch.long)
        top_lp = mem.new_zeros(beam_size)
        seqs = None
        for t in range(max_length):
            pr = prior_step(p, word, mem, lens, h_p, c_p, last_z, eps_beam[i, t])
            de = decoder_step(p, word, mem, lens, h_d, pr["z"])
            logp = torch.log_softmax(de["logits"], dim=1)        # :909
            logp = top_lp.unsqueeze(1) + logp                    # :911
            top_lp, top_w = logp.view(-1).topk(beam_size, 0, True, True)  # :912
            prev = torch.div(top_w, V, rounding_mode="trunc")    # :915
            word = top_w % V                                     # :916
            seqs = word.unsqueeze(1) if t == 0 else torch.cat([seqs[prev], word.unsqueeze(1)], 1)
            h_d = de["h"][prev]; h_p = pr["h"][prev]; c_p = pr["c"][prev]  # :963-969
            last_z = pr["z"][prev]
        out_seqs[i] = seqs[0]                                    # :986
    return {"seqs": out_seqs}


# ----------------------------------------------------------------------------
# diverse beam search with prior latents
# (models/word_model.py:297-394 driven by models/vae_model.py:997-1048)
# ----------------------------------------------------------------------------
def diverse_beam_search(p: Params, audio_embeds, mem_lens, eps_dbs, beam_size=5, group_size=5, diversity_lambda=0.5,
                        temperature=1.0, group_nbest=True, max_length=20):
    """CaptionModel.diverse_beam_search with Hybrid_VAEModel.dbs_step: per clip, `group_size` groups of
    bdash = beam_size // group_size hypotheses; group g runs g steps behind group 0 (word_model.py:333-335) and is
    pushed away from the words the earlier groups hold at the same position (add_diversity, :298-313).  Every
    dbs_step draws prior noise for its bdash rows: eps_dbs[i][(t, g)] is a [bdash, E] tensor, in the reference's
    draw order (clip, global step t, group g active at t).  Returns seqs [N, beam_size | group_size, max_length]."""
    mem_all = project_memory(p, audio_embeds)
    V = p["decoder.classifier.weight"].shape[0]
    E = p["pnet.mean_log_out.weight"].shape[0] // 2
    H = p["decoder.model.weight_hh_l0"].shape[1]
    N = mem_all.shape[0]
    G, bdash = group_size, beam_size // group_size
    out_seqs = torch.full((N, beam_size if group_nbest else G, max_length), END_IDX, dtype=torch.long)   # :322-327
    for i in range(N):                                                                          # :329
        mem = mem_all[i].unsqueeze(0).repeat(bdash, 1, 1)                                       # vae_model.py:1001
        lens = torch.as_tensor(mem_lens)[i].repeat(bdash)
        seq_table = [torch.zeros(bdash, 0, dtype=torch.long) for _ in range(G)]                 # :330
        lp_table = [mem.new_zeros(bdash) for _ in range(G)]                                     # :331
        done = [[] for _ in range(G)]
        st = [None] * G                                                                         # (h_d, h_p, c_p, last_z)
        nxt, prevb = [None] * G, [None] * G
        for t in range(max_length + G - 1):                                                     # :336
            for g in range(G):
                if not (g <= t <= max_length + g - 1):                                          # :338
                    continue
                lt = t - g
                if lt == 0:                                                                     # vae_model.py:1008-1012
                    word = torch.full((bdash,), START_IDX, dtype=torch.long)
                    h_d = mem.new_zeros(bdash, H); h_p = mem.new_zeros(bdash, E)
                    c_p = mem.new_zeros(bdash, E); last_z = mem.new_zeros(bdash, E)
                else:                                                                           # :1013-1024
                    word = nxt[g]
                    h_d, h_p, c_p, last_z = (x[prevb[g]] for x in st[g])
                pr = prior_step(p, word, mem, lens, h_p, c_p, last_z, eps_dbs[i][(t, g)])
                de = decoder_step(p, word, mem, lens, h_d, pr["z"])
                logp = torch.log_softmax(de["logits"], dim=1)                                   # :353
                logp = torch.log_softmax(logp / temperature, dim=1)                             # :354
                if g > 0:                                                                       # add_diversity :302-311
                    change = torch.zeros(V, dtype=logp.dtype)
                    for pc in range(g):
                        dec = seq_table[pc][..., lt]
                        for b in range(bdash):
                            change[dec[b]] += 1
                    logp = logp - change.unsqueeze(0) * diversity_lambda
                logp = lp_table[g].unsqueeze(-1) + logp                                         # :356
                if lt == 0:
                    top_lp, top_w = logp[0].topk(bdash, 0, True, True)                          # :358-359
                else:
                    top_lp, top_w = logp.view(-1).topk(bdash, 0, True, True)                    # :361-362
                lp_table[g] = top_lp
                prevb[g] = torch.div(top_w, V, rounding_mode="floor")                           # :365
                nxt[g] = top_w % V                                                              # :366
                if lt > 0:
                    seq_table[g] = seq_table[g][prevb[g]]                                       # :368
                seq_table[g] = torch.cat([seq_table[g], nxt[g].unsqueeze(-1)], -1)              # :369-371
                is_end = seq_table[g][:, lt] == END_IDX                                         # :373
                if t == max_length + g - 1:
                    is_end = torch.ones_like(is_end)                                            # :375-376
                for b in range(bdash):                                                          # :377-384
                    if is_end[b]:
                        done[g].append({"seq": seq_table[g][b].clone(), "score": lp_table[g][b].item() / (lt + 1)})
                lp_table[g] = lp_table[g].clone()
                lp_table[g][is_end] -= 1000                                                     # :385
                st[g] = (de["h"], pr["h"], pr["c"], pr["z"])                                    # dbs_process_step, vae_model.py:1026-1030
        done = [sorted(done[g], key=lambda x: -x["score"])[:bdash] for g in range(G)]           # :387
        beams = sum(done, []) if group_nbest else [gb[0] for gb in done]                        # :388-391
        for k, bm in enumerate(beams):
            out_seqs[i, k, :len(bm["seq"])] = bm["seq"]                                         # :394
    return {"seqs": out_seqs}


def algorithmic_flops_train_fwd(N, Te, T, E, H, A, Hq, V, Eenc):
    """factorised-attention forward FLOPs of one train step."""
    f = 2 * N * Te * Eenc * E + 2 * 2 * N * Te * E * A
    f += 2 * T * 2 * N * (E + Hq) * 3 * Hq + 2 * N * T * 2 * Hq * 2 * E
    f += T * (2 * N * E * A + 3 * N * Te * A + 2 * N * Te * E + 2 * N * (3 * E + E) * 4 * E + 2 * N * E * 2 * E)
    f += T * (2 * N * H * A + 3 * N * Te * A + 2 * N * Te * E + 2 * N * (3 * E + H) * 3 * H)
    f += 2 * N * T * H * V + 2 * N * H * 2 * E
    return f
