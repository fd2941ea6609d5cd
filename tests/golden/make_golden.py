"""Mint golden vectors by running THE REFERENCE ITSELF (read-only at
/root/reference) on CPU in the build container.

    python tests/golden/make_golden.py

The reference ships no tests or fixtures, so parity is
pinned on its own outputs: this script builds deterministic synthetic inputs
and weights (`acvae_b200.synthetic`), loads the weights into the reference's
`Hybrid_VAEModel` / `VAEModel` (`models/vae_model.py:674,12`), replaces the
reference's RNG draws (`torch.randn`, `random.random`,
`torch.rand`, `torch.multinomial`) with the same injected noise our kernels
receive, runs forward + the runner's loss composition
(`runners/pytorch_runner_vae.py:89-98,315-320`) + backward, and stores the
results in `tests/golden/*.npz`.  Inputs/weights are NOT stored for the
full-size cases: tests regenerate them from the seed.

It also cross-checks `oracle/acvae_oracle.py` against the reference on every
case and aborts if they disagree, so a committed fixture implies a pinned
oracle.  /root/reference does not exist on the GPU box; tests only read the
.npz files.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
import acvae_oracle as oracle  # noqa: E402
from acvae_b200 import synthetic  # noqa: E402

models, train_util = ref_loader.load_reference()


class _StubEncoder(torch.nn.Module):
    """Stands in for Cnn10: returns the precomputed frame memory
    (output contract of models/encoder.py:672-707)."""

    def __init__(self, embed_size):
        super().__init__()
        self.embed_size = embed_size

    def forward(self, feats, feat_lens):
        return {"audio_embeds": feats, "audio_embeds_pooled": feats.mean(1),
                "audio_embeds_lens": torch.as_tensor(feat_lens), "state": None}


class _Patched:
    """Replace the reference's RNG draws with queued, injected values."""

    def __init__(self, randn=(), py_random=(), rand=(), gumbel_u=()):
        self.randn = list(randn); self.py_random = list(py_random)
        self.rand = list(rand); self.gumbel_u = list(gumbel_u)

    def __enter__(self):
        self._o = (torch.randn, random.random, torch.rand, torch.multinomial)

        def randn(*shape, **kw):
            t = self.randn.pop(0)
            shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert tuple(t.shape) == shp, (t.shape, shp)
            return t.clone()

        def py_random():
            return float(self.py_random.pop(0))

        def rand(*shape, **kw):
            shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            if shp == (1,):
                return torch.tensor([float(self.rand.pop(0))])
            t = self.gumbel_u.pop(0)
            assert tuple(t.shape) == shp, (t.shape, shp)
            return t.clone()

        def multinomial(prob, n, *a, **k):
            # same-distribution Gumbel-max draw with the injected uniforms
            u = self.gumbel_u.pop(0)
            g = oracle.gumbel_from_uniform(u)
            return torch.max(torch.log(prob) + g, 1).indices.unsqueeze(1)

        torch.randn, random.random, torch.rand, torch.multinomial = randn, py_random, rand, multinomial
        return self

    def __exit__(self, *a):
        torch.randn, random.random, torch.rand, torch.multinomial = self._o


def build_reference(d, params, variant, dtype):
    dec = models.decoder.VAERNNBahdanauAttnDecoder(
        vocab_size=d.V, enc_mem_size=d.E, embed_size=d.E, hidden_size=d.H,
        dropout=0.0, attn_size=d.A)
    enc = _StubEncoder(d.Eenc)
    if variant == "hybrid":
        m = models.Hybrid_VAEModel(enc, dec, posterior_model="PosteriorRNN_hybrid",
                                   posterior_args={"hidden_size": d.Hq},
                                   prior_model="PriorRNN", prior_args={"hidden_size": d.E})
    else:
        m = models.VAEModel(enc, dec, posterior_model="PosteriorRNN",
                            posterior_args={"hidden_size": d.Hq},
                            prior_model="PriorRNN", prior_args={"hidden_size": d.E})
        # VAEModel.forward passes 4 args to the posterior (vae_model.py:71); no
        # shipped posterior accepts them -> one-line shim.
        fwd = m.qnet.forward
        m.qnet.forward = lambda x, lengths, *_: fwd(x, lengths)
        if not hasattr(m, "ln"):
            pass
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith("encoder") for k in missing.missing_keys) or not missing.missing_keys, missing
    return m.to(dtype)


def tparams(params, dtype, grad=False):
    return {k: torch.from_numpy(v).to(dtype).requires_grad_(grad) for k, v in params.items()}


def run_train_case(name, d, seed, ss_ratio, dis_ratio, variant="hybrid", full=False,
                   dtype=torch.float32, smoothing=0.1, kl_weight=0.5, alpha=1.0):
    torch.set_default_dtype(dtype)  # the reference allocates states with the default dtype
    params = synthetic.make_params(d, seed, variant)
    b = synthetic.make_batch(d, seed)
    T = int(b["cap_lens"].max()) - 1
    tf_flags = [bool(u < ss_ratio) for u in b["u_tf"][:T]]
    dis_flags = [bool(dis_ratio != 0 and u <= dis_ratio) for u in b["u_dis"][:T]]
    m = build_reference(d, params, variant, dtype)
    m.train()
    feats = torch.from_numpy(b["audio_embeds"]).to(dtype).requires_grad_(True)
    caps = torch.from_numpy(b["caps"])
    cap_lens = b["cap_lens"].copy()
    eps_q = torch.from_numpy(b["eps_q"][:, :T]).to(dtype)
    eps_p = torch.from_numpy(b["eps_p"][:T]).to(dtype)
    eps_qs = torch.from_numpy(b["eps_q_steps"][:T]).to(dtype)
    # reference draw order
    if variant == "hybrid":
        randn_q = [eps_q]
    else:
        randn_q = [eps_qs[t] for t in range(T)]
    randn_q = randn_q + [eps_p[t] for t in range(T)]
    rand_q = [b["u_dis"][t] for t in range(T)] if dis_ratio != 0 else []
    with _Patched(randn=randn_q, py_random=list(b["u_tf"][:T]), rand=rand_q):
        out = m(feats, torch.from_numpy(b["mem_lens"].copy()), caps, cap_lens,
                ss_ratio=ss_ratio, dis_ratio=dis_ratio)
    # runner loss composition (pytorch_runner_vae.py:89-98, 315-320)
    lens1 = torch.as_tensor(cap_lens) - 1
    targets = torch.nn.utils.rnn.pack_padded_sequence(caps[:, 1:], lens1, batch_first=True).data
    packed = torch.nn.utils.rnn.pack_padded_sequence(out["logits"], lens1, batch_first=True).data
    crit = train_util.LabelSmoothingLoss(d.V, smoothing=smoothing, device="cpu")
    klf = train_util.Normal_kl_loss(device="cpu")
    ce = crit(packed, targets)
    kl = klf(out["q_means"], out["q_logs"], out["p_means"], out["p_logs"])
    loss = ce + kl_weight * kl
    g = None
    if variant == "hybrid":
        g = torch.nn.MSELoss()(out["q_means_utt"], out["p_means_utt"])
        loss = loss + alpha * g
    # With ss_ratio < 1 the reference's own backward raises ("modified by an
    # inplace operation": the word fed to nn.Embedding is a view of
    # output["seqs"], vae_model.py:832, which :854 then writes in place), so
    # scheduled sampling cannot train as shipped.  Forward outputs are still
    # the reference's; gradients for such cases come from the pinned oracle.
    ref_backward_ok = all(tf_flags[1:])
    grads = {}
    if ref_backward_ok:
        loss.backward()
        grads = {k: v.grad.detach().clone() for k, v in m.named_parameters() if v.grad is not None}
        grads["audio_embeds"] = feats.grad.detach().clone()

    # ---- pin the oracle against the reference -------------------------------
    op = tparams(params, dtype, grad=True)
    ofeats = torch.from_numpy(b["audio_embeds"]).to(dtype).requires_grad_(True)
    oo = oracle.train_forward(op, ofeats, b["mem_lens"], caps, cap_lens, eps_q, eps_p,
                              tf_flags, dis_flags, variant=variant, eps_q_steps=eps_qs)
    ol = oracle.train_loss(oo, caps, cap_lens, d.V, smoothing, kl_weight, alpha,
                           "MSE" if variant == "hybrid" else None)
    ol["loss"].backward()
    tol = 2e-5 if dtype == torch.float32 else 1e-10

    def chk(a, b_, what):
        a = a.detach().double(); b_ = b_.detach().double()
        err = float((a - b_).abs().max() / (b_.abs().max() + 1e-30))
        assert err < tol, f"{name}: oracle != reference on {what}: {err}"
        return err

    worst = 0.0
    for k in ("logits", "outputs", "p_means", "p_logs", "p_z", "q_means", "q_logs", "q_z"):
        worst = max(worst, chk(oo[k], out[k][:, :T] if out[k].shape[1] != oo[k].shape[1] else out[k], k))
    assert torch.equal(oo["seqs"], out["seqs"]), f"{name}: seqs differ"
    if variant == "hybrid":
        worst = max(worst, chk(oo["p_means_utt"], out["p_means_utt"], "p_means_utt"))
        worst = max(worst, chk(oo["q_means_utt"], out["q_means_utt"], "q_means_utt"))
    worst = max(worst, chk(ol["loss"], loss, "loss"))
    for k, gref in grads.items():
        gor = ofeats.grad if k == "audio_embeds" else op[k].grad
        worst = max(worst, chk(gor, gref, "grad " + k))
    if not ref_backward_ok:
        grads = {k: v.grad.detach().clone() for k, v in op.items() if v.grad is not None}
        grads["audio_embeds"] = ofeats.grad.detach().clone()
    print(f"[{name}] oracle==reference (max rel err {worst:.2e}; reference backward "
          f"{'ok' if ref_backward_ok else 'RAISES -> grads from oracle'}); loss={float(loss):.6f} "
          f"ce={float(ce):.6f} kl={float(kl):.6f} g={None if g is None else float(g):}")

    # ---- store ----------------------------------------------------------------
    store = {
        "meta_dims": np.array([d.N, d.Te, d.L, d.E, d.H, d.A, d.Hq, d.V, d.Eenc], dtype=np.int64),
        "meta_seed": np.array(seed), "meta_ss_ratio": np.array(ss_ratio),
        "meta_dis_ratio": np.array(dis_ratio), "meta_smoothing": np.array(smoothing),
        "meta_kl_weight": np.array(kl_weight), "meta_alpha": np.array(alpha),
        "tf_flags": np.array(tf_flags), "dis_flags": np.array(dis_flags),
        "meta_grads_from_reference": np.array(ref_backward_ok),
        "loss": np.array(float(loss)), "ce": np.array(float(ce)), "kl": np.array(float(kl)),
        "global": np.array(float(g) if g is not None else np.nan),
        "seqs": out["seqs"].numpy(),
    }
    keep = ["outputs", "p_means", "p_logs", "p_z", "q_means", "q_logs", "q_z"]
    if variant == "hybrid":
        keep += ["p_means_utt", "q_means_utt"]
    else:
        keep += ["rnn_input"]
    for k in keep:
        store["out_" + k] = out[k].detach().float().numpy()
    store["out_attn_weights"] = out["attn_weights"].detach().float().numpy()
    store["out_sampled_logprobs"] = out["sampled_logprobs"].detach().float().numpy()
    if full:
        store["out_logits"] = out["logits"].detach().float().numpy()
        for k, v in grads.items():
            store["grad_" + k] = v.float().numpy()
    else:
        lg = out["logits"].detach().float()
        store["out_logits_lse"] = torch.logsumexp(lg, -1).numpy()
        store["out_logits_sample"] = lg[:, :, ::97].contiguous().numpy()
        for k, v in grads.items():
            v = v.float()
            store["gradnorm_" + k] = np.array(float(v.norm()))
            flat = v.reshape(-1)
            stride = max(1, flat.numel() // 4096)
            store["gradsample_" + k] = flat[::stride][:4096].contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)
    torch.set_default_dtype(torch.float32)


def run_sample_case(name, d, seed, method, max_length, temp=1.0, dtype=torch.float32):
    params = synthetic.make_params(d, seed, "hybrid")
    b = synthetic.make_batch(d, seed, sample_steps=max_length)
    m = build_reference(d, params, "hybrid", dtype)
    m.eval()
    feats = torch.from_numpy(b["audio_embeds"]).to(dtype)
    eps = torch.from_numpy(b["eps_s"]).to(dtype)
    u = torch.from_numpy(b["u_s"]).to(dtype)
    with torch.no_grad(), _Patched(randn=[eps[t] for t in range(max_length)],
                                   gumbel_u=[u[t] for t in range(max_length)]):
        out = m(feats, torch.from_numpy(b["mem_lens"].copy()), method=method,
                max_length=max_length, temp=temp)
    op = tparams(params, dtype)
    with torch.no_grad():
        oo = oracle.inference_forward(op, feats, b["mem_lens"], eps, method, max_length, temp, u)
    assert torch.equal(oo["seqs"], out["seqs"]), f"{name}: oracle seqs != reference"
    n = oo["n_steps"]
    err = float((oo["p_z"] - out["p_z"][:, :n]).abs().max())
    assert err < 1e-4, err
    print(f"[{name}] oracle==reference; steps={n} seqs[0]={out['seqs'][0].tolist()}")
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        meta_dims=np.array([d.N, d.Te, d.L, d.E, d.H, d.A, d.Hq, d.V, d.Eenc], dtype=np.int64),
        meta_seed=np.array(seed), meta_max_length=np.array(max_length), meta_temp=np.array(temp),
        meta_method=np.array(method), seqs=out["seqs"].numpy(), n_steps=np.array(n),
        out_sampled_logprobs=out["sampled_logprobs"][:, :n].float().numpy(),
        out_p_z=out["p_z"][:, :n].float().numpy(), out_p_means=out["p_means"][:, :n].float().numpy())


def run_beam_case(name, d, seed, beam, max_length, dtype=torch.float32):
    params = synthetic.make_params(d, seed, "hybrid")
    b = synthetic.make_batch(d, seed, sample_steps=max_length, beam=beam)
    m = build_reference(d, params, "hybrid", dtype)
    m.eval()
    feats = torch.from_numpy(b["audio_embeds"]).to(dtype)
    eps_b = torch.from_numpy(b["eps_b"]).to(dtype)
    order = [eps_b[i, t] for i in range(d.N) for t in range(max_length)]
    with torch.no_grad(), _Patched(randn=order):
        out = m(feats, torch.from_numpy(b["mem_lens"].copy()), method="beam",
                beam_size=beam, max_length=max_length)
    op = tparams(params, dtype)
    with torch.no_grad():
        oo = oracle.beam_search(op, feats, b["mem_lens"], eps_b, beam, max_length)
    assert torch.equal(oo["seqs"], out["seqs"]), f"{name}: oracle beam seqs != reference"
    print(f"[{name}] oracle==reference; seqs[0]={out['seqs'][0].tolist()}")
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        meta_dims=np.array([d.N, d.Te, d.L, d.E, d.H, d.A, d.Hq, d.V, d.Eenc], dtype=np.int64),
        meta_seed=np.array(seed), meta_max_length=np.array(max_length), meta_beam=np.array(beam),
        seqs=out["seqs"].numpy())


def run_dbs_case(name, d, seed, beam, groups, max_length, lam=0.5, temperature=1.0, nbest=True, dtype=torch.float32):
    """Diverse beam search (word_model.py:297-394 + vae_model.py:997-1048) of the reference itself."""
    params = synthetic.make_params(d, seed, "hybrid")
    b = synthetic.make_batch(d, seed, sample_steps=max_length, beam=beam, dbs_groups=groups)
    m = build_reference(d, params, "hybrid", dtype)
    m.eval()
    feats = torch.from_numpy(b["audio_embeds"]).to(dtype)
    eps = torch.from_numpy(b["eps_dbs"]).to(dtype)
    active = lambda t, g: g <= t <= max_length + g - 1
    order = [eps[i, t, g] for i in range(d.N) for t in range(max_length + groups - 1) for g in range(groups) if active(t, g)]
    with torch.no_grad(), _Patched(randn=order):
        out = m(feats, torch.from_numpy(b["mem_lens"].copy()), method="dbs", beam_size=beam, group_size=groups,
                diversity_lambda=lam, temperature=temperature, group_nbest=nbest, max_length=max_length)
    op = tparams(params, dtype)
    eps_d = [{(t, g): eps[i, t, g] for t in range(max_length + groups - 1) for g in range(groups)} for i in range(d.N)]
    with torch.no_grad():
        oo = oracle.diverse_beam_search(op, feats, b["mem_lens"], eps_d, beam, groups, lam, temperature, nbest, max_length)
    assert torch.equal(oo["seqs"], out["seqs"]), f"{name}: oracle dbs seqs != reference"
    print(f"[{name}] oracle==reference; seqs[0]={out['seqs'][0].tolist()}")
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        meta_dims=np.array([d.N, d.Te, d.L, d.E, d.H, d.A, d.Hq, d.V, d.Eenc], dtype=np.int64),
        meta_seed=np.array(seed), meta_max_length=np.array(max_length), meta_beam=np.array(beam),
        meta_groups=np.array(groups), meta_lambda=np.array(lam), meta_temperature=np.array(temperature),
        meta_nbest=np.array(int(nbest)), seqs=out["seqs"].numpy())

def run_div_case(name="div_stats", clips=9, K=5, L=12, V=40, seed=3):
    """Diversity statistics of the reference itself (utils/div_utils.py, imported from /root/reference/utils) on synthetic
    id sequences rendered as words, against the id-based oracle restatement."""
    sys.path.insert(0, os.path.join("/root/reference", "utils"))
    import div_utils as ref_div                      # numpy-only module of the reference
    import diversity_oracle as dorc
    rs = np.random.RandomState(seed)
    seqs = rs.randint(3, V, size=(clips, K, L)).astype(np.int64)
    for c in range(clips):                           # ragged ends, an empty caption, a <start> inside, repeated captions
        for k in range(K):
            e = rs.randint(0, L + 1)
            if e < L:
                seqs[c, k, e:] = 2
    seqs[0, 0, 0] = 2
    seqs[1, 1, 0] = 1
    seqs[2, 2] = seqs[2, 1]
    def words(row):
        return " ".join(f"w{t}" for t in dorc.caption_tokens(row))
    caps = {c: [words(seqs[c, k]) for k in range(K)] for c in range(clips)}
    d1, a1 = ref_div.compute_div_n(caps, 1)
    d2, a2 = ref_div.compute_div_n(caps, 2)
    g1, _ = ref_div.compute_global_div_n(caps, 1)
    o = dorc.diversity_stats(seqs)
    assert o["Div1"] == d1 and o["Div2"] == d2 and o["gDiv1"] == g1, (o, d1, d2, g1)
    assert np.array_equal(o["div1"], a1) and np.array_equal(o["div2"], a2)
    print(f"[{name}] oracle==reference; Div1={d1:.6f} Div2={d2:.6f} gDiv1={g1}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seqs=seqs, div1=a1, div2=a2, Div1=np.array(d1), Div2=np.array(d2),
                        gDiv1=np.array(g1), meta_V=np.array(V))

if __name__ == "__main__":
    torch.set_num_threads(8)
    T, C0 = synthetic.TINY, synthetic.CFG0
    if len(sys.argv) > 1 and sys.argv[1] == "div":               # diversity-statistics fixture only
        run_div_case()
        sys.exit(0)
    only_dbs = len(sys.argv) > 1 and sys.argv[1] == "dbs"        # add the dbs fixtures without regenerating the rest
    if only_dbs:
        run_dbs_case("tiny_dbs", T, 1, 6, 3, 6, lam=0.5, temperature=1.0, nbest=True)
        run_dbs_case("tiny_dbs_best", T, 2, 4, 2, 6, lam=1.5, temperature=0.7, nbest=False)
        run_dbs_case("cfg0_dbs", C0, 1, 10, 5, 20, lam=0.5, temperature=1.0, nbest=True)
        sys.exit(0)
    # fp64 pin of the restatement (tolerance 1e-10), nothing stored from it
    run_train_case("_pin64", T, 1, 1.0, 0.0, full=True, dtype=torch.float64)
    run_train_case("_pin64ss", T, 2, 0.5, 0.5, full=True, dtype=torch.float64)
    for f in ("_pin64.npz", "_pin64ss.npz"):
        os.remove(os.path.join(HERE, f))
    run_train_case("tiny_train", T, 1, 1.0, 0.0, full=True)
    run_train_case("tiny_train_dis", T, 2, 1.0, 0.5, full=True)
    run_train_case("tiny_train_ss", T, 2, 0.5, 0.5, full=True)
    run_train_case("tiny_train_vae", T, 3, 1.0, 0.0, variant="vae", full=True)
    run_train_case("cfg0_train", C0, 1, 1.0, 0.0, full=False)
    run_train_case("cfg0_train_dis", C0, 4, 1.0, 0.3, full=False)
    run_train_case("cfg0_train_ss", C0, 4, 0.7, 0.3, full=False)
    run_sample_case("tiny_sample_greedy", T, 1, "greedy", 8)
    # method="gumbel" raises inside the reference itself (word_model.py:195
    # gathers a [N,1] logprob that vae_model.py:855 cannot store), so there is
    # no reference output to pin it on; the oracle's gumbel branch restates
    # the intended arithmetic and is exercised oracle-vs-CUDA only.
    run_sample_case("tiny_sample_multinomial", T, 3, "sample", 8, temp=1.0)
    run_sample_case("cfg0_sample_greedy", C0, 1, "greedy", 20)
    run_sample_case("cfg0_sample_multinomial", C0, 2, "sample", 20)
    run_beam_case("tiny_beam", T, 1, 3, 6)
    run_beam_case("cfg0_beam", C0, 1, 3, 20)
    run_dbs_case("tiny_dbs", T, 1, 6, 3, 6, lam=0.5, temperature=1.0, nbest=True)
    run_dbs_case("tiny_dbs_best", T, 2, 4, 2, 6, lam=1.5, temperature=0.7, nbest=False)
    run_dbs_case("cfg0_dbs", C0, 1, 10, 5, 20, lam=0.5, temperature=1.0, nbest=True)
    run_div_case()
