"""CPU: host-side logic -- state_dict parity with the reference keys, deterministic synthetic data,
the LazyLogits packing protocol, loud failure without CUDA."""
import numpy as np
import pytest
import torch

import harness
from harness import synthetic


def test_state_dict_keys_match_reference_names():
    d = synthetic.TINY
    m = harness.build_model(d, 1, "hybrid", device="cpu")
    want = set(synthetic.make_params(d, 1, "hybrid").keys())
    assert set(m.state_dict().keys()) == want
    g = harness.load_golden("tiny_train")
    ref_keys = {k[5:] for k in g if k.startswith("grad_")} - {"audio_embeds"}
    assert ref_keys == want          # the keys the REFERENCE's named_parameters() produced
    mv = harness.build_model(d, 3, "vae", device="cpu")
    gv = harness.load_golden("tiny_train_vae")
    assert {k[5:] for k in gv if k.startswith("grad_")} - {"audio_embeds"} == set(mv.state_dict().keys())


def test_synthetic_is_deterministic_and_sorted():
    a = synthetic.make_batch(synthetic.CFG0, 1)
    b = synthetic.make_batch(synthetic.CFG0, 1)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert np.all(np.diff(a["cap_lens"]) <= 0) and a["cap_lens"][0] == synthetic.CFG0.L
    assert a["caps"].dtype == np.float32 and np.all(a["caps"][:, 0] == 1)
    for n, ln in enumerate(a["cap_lens"]):
        assert a["caps"][n, ln - 1] == 2 and np.all(a["caps"][n, ln:] == 0)
    assert a["mem_lens"].max() == synthetic.CFG0.Te


def test_no_cpu_fallback():
    d = synthetic.TINY
    m = harness.build_model(d, 1, "hybrid", device="cpu")
    b = synthetic.make_batch(d, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.from_numpy(b["audio_embeds"]), torch.from_numpy(b["mem_lens"]), torch.from_numpy(b["caps"]),
          b["cap_lens"], ss_ratio=1.0, dis_ratio=0.0)
    with pytest.raises(Exception, match="Number of input should be either 4"):
        m(torch.zeros(1))


def test_lazy_logits_pack_protocol(monkeypatch):
    """pack_padded_sequence(lazy).data packs the hidden rows; other torch functions materialise."""
    from acvae_b200 import lazy as lz
    from acvae_b200 import functional as F

    class FakeFn:
        @staticmethod
        def apply(h, w, b):
            return h @ w.t() + b
    monkeypatch.setattr(F, "VocabLogitsFn", FakeFn)
    N, T, H, V = 3, 4, 8, 11
    h = torch.randn(N, T, H); w = torch.randn(V, H); b = torch.randn(V)
    lse = torch.randn(N, T); ssum = torch.randn(N, T)
    lens = torch.tensor([4, 3, 1])
    lazy = lz.LazyLogits(h, w, b, lse, ssum)
    assert tuple(lazy.shape) == (N, T, V)
    packed = torch.nn.utils.rnn.pack_padded_sequence(lazy, lens, batch_first=True).data
    assert isinstance(packed, lz.LazyLogits) and tuple(packed.shape) == (8, V)
    dense = torch.nn.utils.rnn.pack_padded_sequence(h @ w.t() + b, lens, batch_first=True).data
    assert torch.allclose(packed.materialize(), dense, atol=1e-6)
    assert torch.allclose(packed.row_lse, torch.nn.utils.rnn.pack_padded_sequence(lse, lens, batch_first=True).data)
    # generic torch function -> dense
    assert torch.allclose(torch.log_softmax(lazy, dim=-1), torch.log_softmax(h @ w.t() + b, dim=-1), atol=1e-6)
    assert torch.allclose(lazy[:, 1], (h @ w.t() + b)[:, 1], atol=1e-6)


def test_dense_label_smoothing_matches_oracle():
    import acvae_oracle as oracle
    from acvae_b200 import LabelSmoothingLoss
    x = torch.randn(7, 13); y = torch.randint(0, 13, (7,)).float()
    a = LabelSmoothingLoss(13, smoothing=0.1, device="cpu")(x, y)
    b = oracle.label_smoothing_loss(x, y, 13, 0.1)
    assert torch.allclose(a, b, atol=1e-6)


def test_bench_reference_arm_runs_without_gpu():
    """`bench.py --impl reference` (the reference algorithm's CPU path, oracle port) prints one JSON line with the
    contract's keys and needs no GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_clips_per_s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1


def test_ids_to_sentences_and_prediction_file(tmp_path):
    """runners/base_runner.py:146-157 (`_convert_idx2sentence`: stop at <end>, skip <start>) and :243-293 (prediction file)."""
    import json
    from acvae_b200 import metrics
    idx2word = {0: "<pad>", 1: "<start>", 2: "<end>", 3: "<unk>", 4: "a", 5: "dog", 6: "barks", 7: "loudly"}
    seqs = [[1, 4, 5, 6, 2, 2], [4, 5, 6, 7, 5, 6]]                     # second caption never emits <end>
    assert metrics.ids_to_sentences(seqs, idx2word) == ["a dog barks", "a dog barks loudly dog barks"]
    assert metrics.ids_to_sentences(seqs, idx2word, zh=True)[0] == ["a", "dog", "barks"]
    single = metrics.predictions_json(["x.wav", "y.wav"], seqs, idx2word)
    assert single == {"predictions": [{"filename": "x.wav", "caption": "a dog barks", "tokens": "a dog barks"},
                                      {"filename": "y.wav", "caption": "a dog barks loudly dog barks",
                                       "tokens": "a dog barks loudly dog barks"}]}
    multi = metrics.predictions_json(["x.wav"], [[[4, 5, 2, 0], [5, 6, 7, 2]]], idx2word, path=str(tmp_path / "p.json"))
    assert multi["predictions"][0]["captions"] == [{"caption": "a dog", "cap_id": 0, "tokens": "a dog"},
                                                   {"caption": "dog barks loudly", "cap_id": 1, "tokens": "dog barks loudly"}]
    assert json.load(open(tmp_path / "p.json")) == multi
    zh = metrics.predictions_json(["x.wav"], [[4, 5, 2]], idx2word, zh=True)
    assert zh["predictions"][0] == {"filename": "x.wav", "caption": "adog", "tokens": "a dog"}


def test_bleu_oracle_hand_worked_case():
    """The BLEU restatement (oracle/diversity_oracle.py, pycocoevalcap BleuScorer) on a case small enough to do by hand:
    candidate `a b a b`, references `a b c` and `a a`: unigram matches clip at max-per-reference counts (a: 2, b: 1)."""
    import math
    import diversity_oracle as dv
    from acvae_b200 import metrics
    testlen, reflen, guess, correct = dv.bleu_stats([4, 5, 4, 5], [[4, 5, 6], [4, 4]])
    assert (testlen, reflen, guess, correct) == (4, 3, [4, 3, 2, 1], [3, 1, 0, 0])
    b = dv.corpus_bleu([(testlen, reflen, guess, correct)])
    assert abs(b[0] - 0.75) < 1e-9 and abs(b[1] - math.sqrt(0.75 / 3)) < 1e-9 and b[3] < 1e-3
    assert metrics.bleu_from_stats(testlen, reflen, guess, correct) == b
    # brevity penalty: candidate shorter than the closest reference
    t2 = dv.bleu_stats([4, 5], [[4, 5, 6, 7]])
    assert abs(dv.corpus_bleu([t2])[0] - math.exp(1 - 4 / 2)) < 1e-6


def test_philox_reference_known_answers():
    """oracle/philox_ref.py (the host restatement of the device's sampling-noise generator) against the Random123 known-answer
    vectors of philox4x32-10, plus the layout properties the device relies on."""
    import philox_ref
    def hexes(t):
        return [f"{int(x):08x}" for x in t]
    assert hexes(philox_ref.philox4x32_10((0, 0), (0, 0, 0, 0))) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    assert hexes(philox_ref.philox4x32_10((0xFFFFFFFF,) * 2, (0xFFFFFFFF,) * 4)) == ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert hexes(philox_ref.philox4x32_10((0xA4093822, 0x299F31D0), (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344))) == \
        ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]
    u = philox_ref.sampling_uniforms(7, 0, 3, 5, 4400)
    assert u.shape == (3, 5, 4400) and u.dtype == np.float32 and 0.0 <= u.min() and u.max() < 1.0
    assert abs(float(u.mean()) - 0.5) < 0.01
    assert not np.array_equal(u, philox_ref.sampling_uniforms(7, 1, 3, 5, 4400))      # the call counter selects a fresh range
    assert np.array_equal(u[:, :, :300], philox_ref.sampling_uniforms(7, 0, 3, 5, 300))  # a word's draw does not depend on V
