"""Shared test harness: replays the reference runner's train/eval call sequence
(`runners/pytorch_runner_vae.py:76-108, 315-321`) against (a) the CUDA product
through its public, reference-shaped API and (b) the CPU oracle, on the same
seeded synthetic inputs with the same injected noise."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from acvae_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def dims_from_golden(g):
    N, Te, L, E, H, A, Hq, V, Eenc = [int(x) for x in g["meta_dims"]]
    return synthetic.Dims(N=N, Te=Te, L=L, E=E, H=H, A=A, Hq=Hq, V=V, Eenc=Eenc)


def flags_for(b, T, ss_ratio, dis_ratio):
    tf = [bool(u < ss_ratio) for u in b["u_tf"][:T]]
    dis = [bool(dis_ratio != 0 and u <= dis_ratio) for u in b["u_dis"][:T]]
    return tf, dis


# ------------------------------------------------------------------ oracle side
def oracle_params(d, seed, variant="hybrid", dtype=torch.float32, grad=False):
    p = synthetic.make_params(d, seed, variant)
    return {k: torch.from_numpy(v).to(dtype).requires_grad_(grad) for k, v in p.items()}


def run_oracle_train(d, seed, ss_ratio=1.0, dis_ratio=0.0, variant="hybrid", smoothing=0.1, kl_weight=0.5,
                     alpha=1.0, dtype=torch.float32, backward=True):
    import acvae_oracle as oracle
    b = synthetic.make_batch(d, seed)
    T = int(b["cap_lens"].max()) - 1
    tf, dis = flags_for(b, T, ss_ratio, dis_ratio)
    p = oracle_params(d, seed, variant, dtype, grad=backward)
    feats = torch.from_numpy(b["audio_embeds"]).to(dtype).requires_grad_(backward)
    caps = torch.from_numpy(b["caps"])
    out = oracle.train_forward(p, feats, b["mem_lens"], caps, b["cap_lens"],
                               torch.from_numpy(b["eps_q"][:, :T]).to(dtype), torch.from_numpy(b["eps_p"][:T]).to(dtype),
                               tf, dis, variant=variant, eps_q_steps=torch.from_numpy(b["eps_q_steps"][:T]).to(dtype))
    terms = oracle.train_loss(out, caps, b["cap_lens"], d.V, smoothing, kl_weight, alpha,
                              "MSE" if variant == "hybrid" else None)
    grads = {}
    if backward:
        terms["loss"].backward()
        grads = {k: v.grad.detach() for k, v in p.items() if v.grad is not None}
        grads["audio_embeds"] = feats.grad.detach()
    if "global" not in terms:
        terms["global"] = torch.tensor(float("nan"))
    return {"out": out, "terms": {k: v.detach() for k, v in terms.items()}, "grads": grads}


# ------------------------------------------------------------------ CUDA side
def build_model(d, seed, variant="hybrid", device="cuda"):
    import acvae_b200 as models
    params = synthetic.make_params(d, seed, variant)
    dec = models.decoder.VAERNNBahdanauAttnDecoder(vocab_size=d.V, enc_mem_size=d.E, embed_size=d.E,
                                                   hidden_size=d.H, dropout=0.0, attn_size=d.A)
    enc = models.PrecomputedEncoder(d.Eenc)
    if variant == "hybrid":
        m = models.Hybrid_VAEModel(enc, dec, posterior_model="PosteriorRNN_hybrid",
                                   posterior_args={"hidden_size": d.Hq},
                                   prior_model="PriorRNN", prior_args={"hidden_size": d.E})
    else:
        m = models.VAEModel(enc, dec, posterior_model="PosteriorRNN", posterior_args={"hidden_size": d.Hq},
                            prior_model="PriorRNN", prior_args={"hidden_size": d.E})
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return m.to(device)


def run_cuda_train(d, seed, ss_ratio=1.0, dis_ratio=0.0, variant="hybrid", smoothing=0.1, kl_weight=0.5, alpha=1.0,
                   dense_logits=False, backward=True, model=None, keep_grads=False, fused_loss=False):
    """Runner._forward(mode="train") + loss composition + backward on the product."""
    import acvae_b200 as models
    dev = "cuda"
    b = synthetic.make_batch(d, seed)
    T = int(b["cap_lens"].max()) - 1
    tf, dis = flags_for(b, T, ss_ratio, dis_ratio)
    m = model if model is not None else build_model(d, seed, variant, dev)
    m.train()
    if not keep_grads:
        m.zero_grad(set_to_none=True)
    m.materialize_logits = dense_logits
    feats = torch.from_numpy(b["audio_embeds"]).to(dev).requires_grad_(backward)
    caps = torch.from_numpy(b["caps"])                     # float32 on the CPU, as the collate_fn leaves it
    cap_lens = b["cap_lens"].copy()
    eps_q = torch.from_numpy(b["eps_q"][:, :T].copy() if variant == "hybrid" else b["eps_q_steps"][:T].copy())
    lens1 = torch.as_tensor(cap_lens) - 1
    targets = torch.nn.utils.rnn.pack_padded_sequence(caps[:, 1:], lens1, batch_first=True).data      # :89-90
    out = m(feats, torch.from_numpy(b["mem_lens"].copy()), caps, cap_lens, ss_ratio=ss_ratio, dis_ratio=dis_ratio,
            eps_q=eps_q, eps_p=torch.from_numpy(b["eps_p"][:T].copy()), tf_flags=tf, dis_flags=dis)     # :92
    packed = torch.nn.utils.rnn.pack_padded_sequence(out["logits"], lens1, batch_first=True).data      # :94-95
    crit = models.LabelSmoothingLoss(d.V, smoothing=smoothing, device=dev)
    klf = models.Normal_kl_loss(device=dev)
    if fused_loss:       # opt-in: the same composition as one autograd node
        fl = models.FusedVAELoss(d.V, smoothing=smoothing, alpha=alpha if variant == "hybrid" else None)
        loss = fl(out, packed, targets, kl_weight)
        ce, kl, g = fl.terms[1], fl.terms[2], fl.terms[3]
    else:
        ce = crit(packed, targets)
        kl = klf(out["q_means"], out["q_logs"], out["p_means"], out["p_logs"])
        loss = ce + kl_weight * kl                                                                      # :315
        g = torch.tensor(float("nan"))
        if variant == "hybrid":
            g = torch.nn.MSELoss()(out["q_means_utt"], out["p_means_utt"])                             # :318
            loss = loss + alpha * g
    grads = {}
    if backward:
        loss.backward()                                                                                # :321
        grads = {k: v.grad.detach().cpu() for k, v in m.named_parameters() if v.grad is not None}
        grads["audio_embeds"] = feats.grad.detach().cpu()
    torch.cuda.synchronize()
    return {"out": out, "terms": {"loss": loss.detach().cpu(), "ce": ce.detach().cpu(), "kl": kl.detach().cpu(),
                                  "global": g.detach().cpu()}, "grads": grads, "model": m, "packed": packed}


def rel_err(a, b):
    a, b = _as_double(a), _as_double(b)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _as_double(x):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).detach().double().cpu()


def linf_rel_err(a, b):
    """max |a - b| / max |b|: one wrong ELEMENT of typical size shows up here even when the Frobenius error of a
    4400 x 256 gradient hides it."""
    a, b = _as_double(a), _as_double(b)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rowwise_rel_err(a, b):
    """max over rows (last dimension = a row) of |a_r - b_r| / (|b_r| + floor), floor = 1e-3 of the RMS row norm: a wrong
    ROW (one clip, one vocabulary entry, one hidden unit) cannot hide behind the other rows."""
    a, b = _as_double(a), _as_double(b)
    if a.dim() < 2:
        a, b = a.reshape(1, -1), b.reshape(1, -1)
    a, b = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    rn = b.norm(dim=1)
    floor = 1e-3 * float(rn.pow(2).mean().sqrt()) + 1e-30
    return float(((a - b).norm(dim=1) / (rn + floor)).max())


def assert_close(a, b, tol, what=""):
    """The parity bar three ways: norm-wise (Frobenius) < tol, element-wise (infinity norm, relative to the largest entry)
    < tol, and row-wise < 10 tol (rows whose own norm is tiny are measured against 1e-3 of the RMS row norm)."""
    e_f, e_inf, e_row = rel_err(a, b), linf_rel_err(a, b), rowwise_rel_err(a, b)
    assert e_f < tol and e_inf < tol and e_row < 10 * tol, (what, "fro", e_f, "linf", e_inf, "row", e_row)


def max_grad_rel_err(got, ref):
    worst = 0.0
    for k, r in ref.items():
        assert k in got, f"missing gradient for {k}"
        worst = max(worst, rel_err(got[k], r))
    return worst
