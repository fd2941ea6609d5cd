"""CPU: the oracle (oracle/acvae_oracle.py) reproduces the golden vectors minted from the
reference itself (tests/golden/make_golden.py).  Tolerances: fp32 both sides, 2e-5 relative."""
import numpy as np
import pytest
import torch

import harness
from harness import synthetic

import acvae_oracle as oracle

TOL = 2e-5


def _check_train(name, variant="hybrid"):
    g = harness.load_golden(name)
    d = harness.dims_from_golden(g)
    r = harness.run_oracle_train(d, int(g["meta_seed"]), float(g["meta_ss_ratio"]), float(g["meta_dis_ratio"]), variant,
                                 float(g["meta_smoothing"]), float(g["meta_kl_weight"]), float(g["meta_alpha"]))
    for k in ("loss", "ce", "kl") + (("global",) if variant == "hybrid" else ()):
        assert abs(float(r["terms"][k]) - float(g[k])) <= TOL * max(1.0, abs(float(g[k]))), k
    assert np.array_equal(r["out"]["seqs"].numpy(), g["seqs"])
    for k in [k[4:] for k in g if k.startswith("out_") and not k.startswith("out_logits")]:
        if k in ("attn_weights", "sampled_logprobs", "rnn_input"):
            continue
        assert harness.rel_err(r["out"][k], g["out_" + k]) < TOL, k
    if "out_logits" in g:
        assert harness.rel_err(r["out"]["logits"], g["out_logits"]) < TOL
        for k in [k[5:] for k in g if k.startswith("grad_")]:
            assert harness.rel_err(r["grads"][k], g["grad_" + k]) < 5e-5, k
    else:
        assert harness.rel_err(torch.logsumexp(r["out"]["logits"], -1), g["out_logits_lse"]) < TOL
        for k in [k[9:] for k in g if k.startswith("gradnorm_")]:
            gn = float(r["grads"][k].norm())
            assert abs(gn - float(g["gradnorm_" + k])) <= 5e-5 * max(1e-6, float(g["gradnorm_" + k])), k


@pytest.mark.parametrize("name", ["tiny_train", "tiny_train_dis", "tiny_train_ss"])
def test_oracle_train_tiny(name):
    _check_train(name)


def test_oracle_train_tiny_vae():
    _check_train("tiny_train_vae", "vae")


def test_oracle_train_cfg0():
    _check_train("cfg0_train")


@pytest.mark.parametrize("name", ["tiny_sample_greedy", "tiny_sample_multinomial", "cfg0_sample_greedy"])
def test_oracle_sampling(name):
    g = harness.load_golden(name)
    d = harness.dims_from_golden(g)
    seed, ml = int(g["meta_seed"]), int(g["meta_max_length"])
    b = synthetic.make_batch(d, seed, sample_steps=ml)
    p = harness.oracle_params(d, seed)
    with torch.no_grad():
        o = oracle.inference_forward(p, torch.from_numpy(b["audio_embeds"]), b["mem_lens"], torch.from_numpy(b["eps_s"]),
                                     str(g["meta_method"]), ml, float(g["meta_temp"]), torch.from_numpy(b["u_s"]))
    assert np.array_equal(o["seqs"].numpy(), g["seqs"])
    assert o["n_steps"] == int(g["n_steps"])


def test_oracle_beam_tiny():
    g = harness.load_golden("tiny_beam")
    d = harness.dims_from_golden(g)
    seed, ml, beam = int(g["meta_seed"]), int(g["meta_max_length"]), int(g["meta_beam"])
    b = synthetic.make_batch(d, seed, sample_steps=ml, beam=beam)
    p = harness.oracle_params(d, seed)
    with torch.no_grad():
        o = oracle.beam_search(p, torch.from_numpy(b["audio_embeds"]), b["mem_lens"], torch.from_numpy(b["eps_b"]), beam, ml)
    assert np.array_equal(o["seqs"].numpy(), g["seqs"])


def test_oracle_dbs_tiny():
    """The oracle's diverse beam search reproduces the reference's token ids (fixtures made by make_golden.py dbs)."""
    for name in ("tiny_dbs", "tiny_dbs_best"):
        g = harness.load_golden(name)
        d = harness.dims_from_golden(g)
        seed, ml, beam, groups = int(g["meta_seed"]), int(g["meta_max_length"]), int(g["meta_beam"]), int(g["meta_groups"])
        b = synthetic.make_batch(d, seed, sample_steps=ml, beam=beam, dbs_groups=groups)
        p = harness.oracle_params(d, seed)
        eps = torch.from_numpy(b["eps_dbs"])
        eps_d = [{(t, gg): eps[i, t, gg] for t in range(ml + groups - 1) for gg in range(groups)} for i in range(d.N)]
        with torch.no_grad():
            o = oracle.diverse_beam_search(p, torch.from_numpy(b["audio_embeds"]), b["mem_lens"], eps_d, beam, groups,
                                           float(g["meta_lambda"]), float(g["meta_temperature"]), bool(int(g["meta_nbest"])), ml)
        assert np.array_equal(o["seqs"].numpy(), g["seqs"]), name

def test_oracle_edge_min_length_and_single_frame():
    """Ragged edge cases: a caption of the minimum length (<start>,<end>) and a clip with one frame."""
    d = synthetic.Dims(N=3, Te=5, L=5, E=16, H=16, A=16, Hq=16, V=23, Eenc=20)
    b = synthetic.make_batch(d, 7, min_cap_len=2)
    b["cap_lens"][-1] = 2
    b["caps"][-1] = 0; b["caps"][-1, 0] = 1; b["caps"][-1, 1] = 2
    b["mem_lens"][1] = 1
    p = harness.oracle_params(d, 7)
    T = d.T
    out = oracle.train_forward(p, torch.from_numpy(b["audio_embeds"]), b["mem_lens"], torch.from_numpy(b["caps"]),
                               b["cap_lens"], torch.from_numpy(b["eps_q"]), torch.from_numpy(b["eps_p"]))
    terms = oracle.train_loss(out, torch.from_numpy(b["caps"]), b["cap_lens"], d.V)
    assert torch.isfinite(terms["loss"])
    # padded posterior positions see ho = 0 => head = bias
    assert torch.allclose(out["q_means"][-1, 1:], p["qnet.token_mean_log.bias"][:d.E].expand(T - 1, d.E), atol=1e-6)


def test_oracle_diversity_stats_golden():
    """The id-based restatement of utils/div_utils.py reproduces the reference's Div-1 / Div-2 / gDiv-1 (fixture made by
    `make_golden.py div` from the reference functions themselves)."""
    import diversity_oracle as dorc
    g = harness.load_golden("div_stats")
    o = dorc.diversity_stats(g["seqs"])
    assert np.array_equal(o["div1"], g["div1"]) and np.array_equal(o["div2"], g["div2"])
    assert o["Div1"] == float(g["Div1"]) and o["Div2"] == float(g["Div2"]) and o["gDiv1"] == float(g["gDiv1"])
