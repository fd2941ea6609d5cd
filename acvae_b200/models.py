"""Host-side mirror of the reference's model classes for the AC-VAE hot path.

Same class names, constructor arguments, `forward` signatures, output-dict keys
and `state_dict` names as the reference, so `runners/pytorch_runner_vae.py`
(`getattr(models, config["model"])(encoder, decoder, **model_args)`,
pytorch_runner_vae.py:32-73) can drive them unchanged -- see INTEGRATION.md.
The sub-modules only OWN parameters (they are never called layer by layer):
the whole step runs inside `libacvae_b200.so`.

reference                                         here
models/attn_model.py:6-46     Seq2SeqAttention    Seq2SeqAttention      (parameter container)
models/decoder.py:164-203     VAERNNBahdanau...   VAERNNBahdanauAttnDecoder
models/text_encoder.py:156    PosteriorRNN_hybrid PosteriorRNN_hybrid
models/text_encoder.py:96     PosteriorRNN        PosteriorRNN          (AR posterior, VAEModel)
models/text_encoder.py:218    PriorRNN            PriorRNN
models/word_model.py:14       CaptionModel        CaptionModel          (ids, set_index)
models/vae_model.py:674       Hybrid_VAEModel     Hybrid_VAEModel
models/vae_model.py:12        VAEModel            VAEModel
"""
from __future__ import annotations

import random
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import functional as F
from .lazy import LazyLogits


def _fused_only(name):
    def forward(self, *a, **k):
        raise NotImplementedError(
            f"{name}.forward is fused into Hybrid_VAEModel/VAEModel.forward (libacvae_b200.so); "
            "the sub-module only owns its parameters")
    return forward


def _init_linear_xavier(module: nn.Module):
    """PosteriorBaseEncoder.init / PriorBaseEncoder.init (text_encoder.py:26-42, 66-81)."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class Seq2SeqAttention(nn.Module):
    """Parameters of the additive attention (attn_model.py:8-18): h2attn [A, hs_dec+hs_enc]
    (query columns first, :31), v [A] ~ N(0,1)."""

    def __init__(self, hs_enc, hs_dec, attn_size):
        super().__init__()
        self.h2attn = nn.Linear(hs_enc + hs_dec, attn_size)
        self.v = nn.Parameter(torch.randn(attn_size))
        nn.init.kaiming_uniform_(self.h2attn.weight)

    forward = _fused_only("Seq2SeqAttention")


class VAERNNBahdanauAttnDecoder(nn.Module):
    """decoder.py:164-173 (+ RNNDecoder.__init__ :30-48, BaseDecoder :17-22)."""

    def __init__(self, vocab_size, enc_mem_size, **kwargs):
        super().__init__()
        embed_size = kwargs.get("embed_size", 256)
        hidden_size = kwargs.get("hidden_size", 256)
        if kwargs.get("num_layers", 1) != 1 or kwargs.get("bidirectional", False) or kwargs.get("rnn_type", "GRU") != "GRU":
            raise NotImplementedError("the fused step implements the reference default: 1-layer unidirectional GRU")
        self.embed_size, self.vocab_size, self.enc_mem_size = embed_size, vocab_size, enc_mem_size * 2
        self.word_embeddings = nn.Embedding(vocab_size, embed_size)
        self.dropoutlayer = nn.Dropout(kwargs.get("dropout", 0.0))
        self.model = nn.GRU(input_size=embed_size + enc_mem_size * 2, hidden_size=hidden_size, num_layers=1,
                            batch_first=True, bidirectional=False)
        self.classifier = nn.Linear(hidden_size, vocab_size)
        nn.init.kaiming_uniform_(self.word_embeddings.weight)
        nn.init.kaiming_uniform_(self.classifier.weight)
        attn_size = kwargs.get("attn_size", hidden_size)
        self.attn = Seq2SeqAttention(enc_mem_size, hidden_size, attn_size)

    def load_word_embeddings(self, embeddings, tune=True, **kwargs):
        """decoder.py:50-64, including the `projection` form: pre-trained embeddings of another width become
        `nn.Sequential(nn.Embedding, nn.Linear(width, embed_size))` (same `state_dict` keys as the reference:
        `word_embeddings.0.weight`, `word_embeddings.1.weight / .bias`).  The fused step then runs on the EFFECTIVE table
        `E_pre . P^T + b` [V, embed_size], formed once per call on the device (`effective_word_embeddings`); gradients
        reach P, b (and E_pre when `tune`) through that contraction."""
        assert embeddings.shape[0] == self.vocab_size, "vocabulary size mismatch!"
        embeddings = torch.as_tensor(embeddings).float()
        self.word_embeddings = nn.Embedding(self.vocab_size, embeddings.shape[1])
        self.word_embeddings.weight = nn.Parameter(embeddings)
        for para in self.word_embeddings.parameters():
            para.requires_grad = tune
        if embeddings.shape[1] != self.embed_size:
            assert "projection" in kwargs, "embedding size mismatch!"
            if kwargs["projection"]:
                self.word_embeddings = nn.Sequential(self.word_embeddings, nn.Linear(embeddings.shape[1], self.embed_size))

    def effective_word_embeddings(self):
        """The [V, embed_size] table the step gathers from: the embedding itself, or `Linear(Embedding)` evaluated for every
        vocabulary entry when `load_word_embeddings(..., projection=True)` installed a projection (decoder.py:58-64)."""
        we = self.word_embeddings
        if isinstance(we, nn.Sequential):
            return F.VocabLogitsFn.apply(we[0].weight, we[1].weight, we[1].bias)
        return we.weight

    def init_hidden(self, bs):
        return torch.zeros(1, bs, self.model.hidden_size)

    forward = _fused_only("VAERNNBahdanauAttnDecoder")


class _PosteriorBase(nn.Module):
    def __init__(self, word_dim, embed_size, vocab_size, **kwargs):
        super().__init__()
        self.word_dim, self.embed_size, self.vocab_size = word_dim, embed_size, vocab_size
        self.word_embedding = nn.Embedding(vocab_size, word_dim)
        self.hidden_size = kwargs.get("hidden_size", 256)
        if (not kwargs.get("bidirectional", True)) or kwargs.get("num_layers", 1) != 1 or kwargs.get("rnn_type", "GRU") != "GRU":
            raise NotImplementedError("the fused step implements the reference default: 1-layer bidirectional GRU")
        self.network = nn.GRU(word_dim, self.hidden_size, num_layers=1, bidirectional=True, batch_first=True)


class PosteriorRNN_hybrid(_PosteriorBase):
    """text_encoder.py:156-180: contextual posterior + utterance pooling."""

    def __init__(self, word_dim, embed_size, vocab_size, **kwargs):
        super().__init__(word_dim, embed_size, vocab_size, **kwargs)
        self.token_mean_log = nn.Linear(2 * self.hidden_size, 2 * embed_size)
        _init_linear_xavier(self)

    forward = _fused_only("PosteriorRNN_hybrid")


class PosteriorRNN(_PosteriorBase):
    """text_encoder.py:96-119: autoregressive posterior (VAEModel)."""

    def __init__(self, word_dim, embed_size, vocab_size, **kwargs):
        super().__init__(word_dim, embed_size, vocab_size, **kwargs)
        self.mean_log_out = nn.Linear(embed_size + 2 * self.hidden_size, 2 * embed_size)
        _init_linear_xavier(self)

    forward = _fused_only("PosteriorRNN")


class PriorRNN(nn.Module):
    """text_encoder.py:218-238: autoregressive prior (word attention + LSTM + Gaussian head)."""

    def __init__(self, word_dim, audiofeats_size, embed_size, vocab_size, **kwargs):
        super().__init__()
        self.word_dim, self.embed_size, self.vocab_size, self.audiofeats_size = word_dim, embed_size, vocab_size, audiofeats_size
        self.word_embedding = nn.Embedding(vocab_size, word_dim)
        self.hidden_size = kwargs.get("hidden_size", 256)
        if kwargs.get("bidirectional", False) or kwargs.get("num_layers", 1) != 1 or kwargs.get("rnn_type", "LSTM") != "LSTM":
            raise NotImplementedError("the fused step implements the reference default: 1-layer unidirectional LSTM")
        self.word_attn = Seq2SeqAttention(audiofeats_size, word_dim, audiofeats_size)
        self.network = nn.LSTM(word_dim + audiofeats_size + embed_size, self.hidden_size, num_layers=1,
                               bidirectional=False, batch_first=True)
        self.mean_log_out = nn.Linear(self.hidden_size, 2 * embed_size)
        _init_linear_xavier(self)

    forward = _fused_only("PriorRNN")


_text_encoders = {"PosteriorRNN_hybrid": PosteriorRNN_hybrid, "PosteriorRNN": PosteriorRNN, "PriorRNN": PriorRNN}


class CaptionModel(nn.Module):
    """word_model.py:14-44: special-token ids and the encoder/decoder pair."""

    pad_idx = 0
    start_idx = 1
    end_idx = 2
    max_length = 20

    def __init__(self, encoder: nn.Module, decoder: nn.Module, **kwargs):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.vocab_size = decoder.vocab_size
        if kwargs.get("freeze_encoder"):
            for param in self.encoder.parameters():
                param.requires_grad = False

    @classmethod
    def set_index(cls, start_idx, end_idx):
        cls.start_idx = start_idx
        cls.end_idx = end_idx


class PreparedBatch:
    """Caption-side inputs of a training step already on the device (ids int32, lengths int32) plus
    the host-known step count T.  Lets a step run without any host->device copy (CUDA-graph capture,
    device-resident benchmarking); `Hybrid_VAEModel.prepare_batch` builds it."""

    def __init__(self, caps_ids, cap_lens_dev, T, targets=None, flat=None):
        self.caps_ids, self.cap_lens_dev, self.T, self.targets = caps_ids, cap_lens_dev, int(T), targets
        self.flat = flat        # the one int32 device buffer the views live in (prepare_batch), or None
        #: un-packed criterion inputs (FusedVAELoss.forward_padded): caps[:, 1:T+1] as int32 [N,T] and the row weights
        #: [N,T] float32 (1 where t < cap_lens[n] - 1); views of `flat` behind the packed targets
        self.targets_padded = self.row_w = None
        if flat is not None:
            N, L = caps_ids.shape
            M = flat.numel() - N * L - N - 2 * N * self.T
            base = N * L + N + M
            self.targets_padded = flat[base:base + N * self.T].view(N, self.T)
            self.row_w = flat[base + N * self.T:base + 2 * N * self.T].view(torch.float32).view(N, self.T)

    def clone(self) -> "PreparedBatch":
        """A copy with its own storage (static buffers of a captured step)."""
        if self.flat is None:
            return PreparedBatch(self.caps_ids.clone(), self.cap_lens_dev.clone(), self.T,
                                 None if self.targets is None else self.targets.clone())
        f = self.flat.clone()
        N, L = self.caps_ids.shape
        M = self.targets.numel()
        return PreparedBatch(f[:N * L].view(N, L), f[N * L:N * L + N], self.T, f[N * L + N:N * L + N + M], f)


class _FusedVAEBase(CaptionModel):
    variant = 0
    #: "device": noise from the CUDA generator (fast path); "reference_cpu": the reference's CPU
    #: generator stream in its exact draw order, for seed-for-seed comparisons
    noise_source = "device"
    #: False: output["logits"] is a LazyLogits handle; True: a dense [N,T,V] tensor
    materialize_logits = False
    #: optional `parallel.FlatGradBuffer`: the fused backward then writes every hot-path weight gradient straight
    #: into the flat all-reduce buffer (OVERWRITE semantics: one backward per optimiser step) instead of handing
    #: ~35 tensors to autograd for accumulation
    grad_sink = None

    def __init__(self, Audioencoder, Textdecoder, **kwargs):
        super().__init__(Audioencoder, Textdecoder, **kwargs)
        E = Textdecoder.embed_size
        if Textdecoder.model.hidden_size != E:
            raise ValueError("decoder hidden_size must equal embed_size (reference vae_model.py:693,726)")
        self.qnet = _text_encoders[kwargs["posterior_model"]](
            word_dim=E, embed_size=E, vocab_size=Textdecoder.vocab_size, **kwargs["posterior_args"])
        self.pnet = _text_encoders[kwargs["prior_model"]](
            word_dim=E, audiofeats_size=E, embed_size=E, vocab_size=Textdecoder.vocab_size, **kwargs["prior_args"])
        if self.qnet.hidden_size != E or self.pnet.hidden_size != E:
            raise ValueError("posterior/prior hidden_size must equal embed_size (SURVEY.md section 8)")

    # ---- plumbing ------------------------------------------------------------------------
    def _hot_weights(self) -> Dict[str, torch.Tensor]:
        sd = dict(self.named_parameters())
        hot = {k: v for k, v in sd.items() if not k.startswith("encoder.")}
        if isinstance(self.decoder.word_embeddings, nn.Sequential):     # projected pre-trained embeddings (decoder.py:58-64)
            for k in [k for k in hot if k.startswith("decoder.word_embeddings.")]:
                del hot[k]
            items = list(hot.items())
            # keep the position the plain table has in the parameter order (first decoder entry)
            pos = next((i for i, (k, _) in enumerate(items) if k.startswith("decoder.")), len(items))
            items.insert(pos, ("decoder.word_embeddings.weight", self.decoder.effective_word_embeddings()))
            hot = dict(items)
        return hot

    def _encode(self, feats, feat_lens):
        encoded = self.encoder(feats, feat_lens)
        return encoded["audio_embeds"], encoded["audio_embeds_lens"]

    def _dims(self, N, Te, T, L=0, mem_rep=1):
        dec = self.decoder
        Eenc = self.encoder.embed_size if hasattr(self, "ln") else dec.embed_size
        return F.make_dims(N, Te, T, dec.embed_size, dec.attn.h2attn.out_features, dec.vocab_size, Eenc, L,
                           mem_rep, self.variant)

    def forward(self, *input, **kwargs):
        """vae_model.py:732-760 (same two call forms, same exception text)."""
        if len(input) == 4:
            feats, feat_lens, caps, cap_lens = input
            audio_embeds, mem_lens = self._encode(feats, feat_lens)
            return self.train_forward({"audio_embeds": audio_embeds, "audio_embeds_lens": mem_lens}, caps, cap_lens, **kwargs)
        elif len(input) == 2:
            feats, feat_lens = input
            audio_embeds, mem_lens = self._encode(feats, feat_lens)
            return self.inference_forward({"audio_embeds": audio_embeds, "audio_embeds_lens": mem_lens}, **kwargs)
        raise Exception("Number of input should be either 4 (feats, feat_lens, caps, cap_lens) or 2 (feats, feat_lens)")

    def prepare_batch(self, caps, cap_lens, device, out: "PreparedBatch" = None) -> PreparedBatch:
        """Host -> device staging of (caps, cap_lens) exactly as the step consumes them, plus the packed CE targets of
        pytorch_runner_vae.py:89-90 (`pack_padded_sequence(caps[:, 1:], cap_lens - 1).data`).  Everything is converted
        on the host into ONE pinned int32 staging buffer and crosses PCIe in ONE asynchronous copy
        (ids [N*L] | lens [N] | targets [M]); `out` (a PreparedBatch made by an earlier call with the same caption-length
        profile) receives the copy in place, so a captured CUDA graph keeps reading the same addresses."""
        caps_np = caps.detach().cpu().numpy() if torch.is_tensor(caps) else np.asarray(caps)   # the collate_fn's CPU output
        lens = np.asarray(cap_lens)
        if caps_np.ndim != 2 or lens.shape != (caps_np.shape[0],):
            raise ValueError("caps must be [N,L] and cap_lens [N]")
        N, L = caps_np.shape
        # everything that depends only on the length profile is cached (a training run sees few distinct profiles)
        key = (L, lens.tobytes())
        prof = getattr(self, "_len_profiles", None)
        if prof is None:
            prof = self._len_profiles = {}
        ent = prof.get(key)
        if ent is None:
            lens64 = lens.astype(np.int64)
            if np.any(np.diff(lens64) > 0):
                raise RuntimeError("`lengths` array must be sorted in decreasing order when `enforce_sorted` is True.")
            T_ = int(lens64.max()) - 1
            lens1 = lens64 - 1
            mask_ = lens1[None, :] > np.arange(T_)[:, None]           # time-major packing order of sorted lengths
            if len(prof) > 256:
                prof.clear()
            ent = prof[key] = (T_, int(lens1.sum()), mask_, lens64.astype(np.int32))
        T, M, mask, lens32 = ent
        total = N * L + N + M + 2 * N * T                               # ids | lens | packed targets | padded targets | row weights
        ring = getattr(self, "_stage_ring", None)
        if ring is None or ring[0][0].numel() < total:
            # [pinned staging tensor, CUDA event recorded right after the last asynchronous copy OUT of it]
            ring = [[torch.empty(max(total, 4096), dtype=torch.int32).pin_memory(), None] for _ in range(4)]
            self._stage_ring, self._stage_next = ring, 0
        slot = ring[self._stage_next]
        self._stage_next = (self._stage_next + 1) % len(ring)
        stage = slot[0]
        if slot[1] is not None:
            slot[1].synchronize()     # the host may run >= 4 calls ahead of the stream: never rewrite a slot whose copy is pending
        sn = stage.numpy()
        sn[:N * L] = caps_np.reshape(-1)                              # caps.long() of vae_model.py:827, as int32
        sn[N * L:N * L + N] = lens32
        sn[N * L + N:N * L + N + M] = caps_np[:, 1:T + 1].T[mask]
        sn[N * L + N + M:N * L + N + M + N * T] = caps_np[:, 1:T + 1].reshape(-1)
        sn[N * L + N + M + N * T:total].view(np.float32)[:] = mask.T.reshape(-1)       # float bits in the int32 staging buffer
        if out is not None:
            if out.flat is None or out.flat.numel() != total or out.T != T or tuple(out.caps_ids.shape) != (N, L):
                raise ValueError("`out` was prepared for a different caption-length profile")
            with torch.cuda.device(out.flat.device):
                out.flat.copy_(stage[:total], non_blocking=True)
                slot[1] = slot[1] or torch.cuda.Event()
                slot[1].record()
            return out
        with torch.cuda.device(device):
            flat = stage[:total].to(device=device, non_blocking=True)
            slot[1] = slot[1] or torch.cuda.Event()
            slot[1].record()
        return PreparedBatch(flat[:N * L].view(N, L), flat[N * L:N * L + N], T, flat[N * L + N:N * L + N + M], flat)

    # ---- training ------------------------------------------------------------------------
    def train_forward(self, encoded, caps, cap_lens, **kwargs):
        """vae_model.py:871-878 -> stepwise_forward :700-730, fused.

        Required kwargs as in the reference: `ss_ratio`, `dis_ratio`
        (decode_step reads them without defaults, vae_model.py:802,826).
        Optional injection (parity tests): `eps_q`, `eps_p`, `tf_flags`, `dis_flags`.
        """
        if self.training and self.decoder.dropoutlayer.p > 0:
            raise NotImplementedError("decoder dropout > 0 is not implemented in the fused step (reference default 0.0)")
        audio = encoded["audio_embeds"]
        dev = audio.device
        if not audio.is_cuda:
            raise RuntimeError("acvae_b200 needs CUDA tensors: there is no CPU path")
        N, Te = audio.shape[0], audio.shape[1]
        E = self.decoder.embed_size
        if isinstance(caps, PreparedBatch):
            caps_ids, cap_lens_dev, T = caps.caps_ids, caps.cap_lens_dev, caps.T
        else:
            cap_lens_np = np.asarray(cap_lens).astype(np.int64)
            T = int(cap_lens_np.max()) - 1                                 # vae_model.py:703
            caps_ids = torch.as_tensor(caps).to(device=dev, dtype=torch.int32).contiguous()  # caps.long(), :827
            cap_lens_dev = torch.as_tensor(cap_lens_np).to(device=dev, dtype=torch.int32)
        L = caps_ids.shape[1]
        ss_ratio, dis_ratio = kwargs["ss_ratio"], kwargs["dis_ratio"]
        eps_q, eps_p = kwargs.get("eps_q"), kwargs.get("eps_p")
        tf_flags, dis_flags = kwargs.get("tf_flags"), kwargs.get("dis_flags")
        hybrid = self.variant == 0
        if self.noise_source == "reference_cpu" and eps_q is None:
            # exact CPU-generator draw order of the reference
            if hybrid:
                eps_q = torch.randn(N, T, E)
            else:
                eps_q = torch.stack([torch.randn(N, E) for _ in range(T)])
            tf_l, eps_l, dis_l = [], [], []
            for _ in range(T):
                tf_l.append(random.random() < ss_ratio)
                eps_l.append(torch.randn(N, E))
                dis_l.append(bool(dis_ratio != 0 and float(torch.rand(1)) <= dis_ratio))
            eps_p = torch.stack(eps_l)
            tf_flags, dis_flags = tf_l, dis_l
        if eps_q is None:
            eps_q = torch.randn((N, T, E) if hybrid else (T, N, E), device=dev)
        if eps_p is None:
            eps_p = torch.randn(T, N, E, device=dev)
        if tf_flags is None:
            tf_flags = [random.random() < ss_ratio for _ in range(T)]      # vae_model.py:826
        if dis_flags is None:
            dis_flags = [bool(dis_ratio != 0 and random.random() <= dis_ratio) for _ in range(T)]  # :802-804
        eps_q = eps_q.to(device=dev, dtype=torch.float32).contiguous()
        eps_p = eps_p.to(device=dev, dtype=torch.float32).contiguous()
        mem_lens = torch.as_tensor(encoded["audio_embeds_lens"]).to(device=dev, dtype=torch.int32).contiguous()
        weights = self._hot_weights()
        keys = list(weights.keys())
        dims = self._dims(N, Te, T, L)
        sink = None
        if self.grad_sink is not None and torch.is_grad_enabled():
            sink = self.grad_sink.views_for(self, prefix_skip="encoder.")
        meta = F.TrainMeta(dims, keys, caps_ids, cap_lens_dev, mem_lens, eps_q, eps_p, tf_flags, dis_flags,
                           want_logits=False, grad_sink=sink)
        (q_means, q_logs, q_z, p_means, p_logs, p_z, outputs, q_utt, p_utt, attn_w, seqs, slp, lse, lsum, rnn_input,
         _logits) = F.LatentDecodeTrainFn.apply(meta, audio, *[weights[k] for k in keys])
        lazy = LazyLogits(outputs, self.decoder.classifier.weight, self.decoder.classifier.bias, lse, lsum)
        if sink is not None and "decoder.classifier.weight" in sink:
            lazy.grad_sink = (sink["decoder.classifier.weight"], sink["decoder.classifier.bias"])
        out = {
            "seqs": seqs, "logits": lazy.materialize() if self.materialize_logits else lazy,
            "outputs": outputs, "sampled_logprobs": slp, "attn_weights": attn_w,
            "p_means": p_means, "p_logs": p_logs, "p_z": p_z,
            "q_means": q_means, "q_logs": q_logs, "q_z": q_z,
            "state": outputs[:, -1].unsqueeze(0), "last_z": p_z[:, -1],
        }
        if hybrid:
            out.update({"q_means_utt": q_utt, "q_logs_utt": None, "p_means_utt": p_utt, "p_logs_utt": None})
        else:
            out["rnn_input"] = rnn_input
        return out

    # ---- inference -----------------------------------------------------------------------
    def sampling_rng(self, dev) -> torch.Tensor:
        """Device state {seed, calls} of the word-sampling generator (Philox4x32-10, csrc/gemm.cuh).  Created on first use with a
        seed drawn from torch's CPU generator (so `torch.manual_seed` makes sampling reproducible); every sampling call advances
        `calls` on the device, which keeps a replayed CUDA graph (GraphSampler) drawing fresh noise."""
        dev = torch.device(dev)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        st = getattr(self, "_sampling_rng", None)
        if st is None or st.device != dev:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            st = torch.tensor([seed, 0], dtype=torch.int64, device=dev)
            self._sampling_rng = st
        return st

    def seed_sampling(self, seed: int, device=None) -> None:
        """Restart the word-sampling generator from `seed` (call 0)."""
        dev = device if device is not None else next(self.parameters()).device
        self._sampling_rng = torch.tensor([int(seed), 0], dtype=torch.int64, device=dev)

    def inference_forward(self, encoded, **kwargs):
        """vae_model.py:880-894.  Extra kwargs: `n_captions` (K sequences per clip sharing the
        clip's memory), `eps_p`, `u` (noise injection)."""
        method = kwargs.get("method", "greedy")
        max_length = kwargs.get("max_length", self.max_length)
        audio = encoded["audio_embeds"]
        dev = audio.device
        if not audio.is_cuda:
            raise RuntimeError("acvae_b200 needs CUDA tensors: there is no CPU path")
        clips, Te = audio.shape[0], audio.shape[1]
        E, V = self.decoder.embed_size, self.decoder.vocab_size
        mem_lens = torch.as_tensor(encoded["audio_embeds_lens"]).to(device=dev, dtype=torch.int32).contiguous()
        weights = {k: v for k, v in self._hot_weights().items()}
        if method == "beam":
            beam = kwargs.get("beam_size", 3)
            dims = self._dims(clips, Te, max_length)
            eps_b = kwargs.get("eps_b")
            if eps_b is None:
                eps_b = torch.randn(max_length, clips * beam, E, device=dev)
            return F.beam_search(dims, weights, audio, mem_lens, eps_b.to(dev), beam, self.start_idx)
        if method == "dbs":
            # vae_model.py:887-893 (same keyword names and defaults); `eps_g` injects the prior noise
            beam = int(kwargs.get("beam_size", 5))
            groups = int(kwargs.get("group_size", 5))
            lam = kwargs.get("diversity_lambda", 0.5)
            temperature = kwargs.get("temperature", 1.0)
            nbest = kwargs.get("group_nbest", True)
            if groups < 1 or beam < groups:
                raise ValueError("diverse beam search needs 1 <= group_size <= beam_size")
            dims = self._dims(clips, Te, max_length)
            rows = clips * groups * (beam // groups)
            eps_g = kwargs.get("eps_g")
            if eps_g is None:
                eps_g = torch.randn(max_length + groups - 1, rows, E, device=dev)
            return F.diverse_beam_search(dims, weights, audio, mem_lens, eps_g.to(dev), beam, groups, lam, temperature, nbest,
                                         self.start_idx, self.end_idx)
        K = int(kwargs.get("n_captions", 1))
        N = clips * K
        dims = self._dims(N, Te, max_length, mem_rep=K)
        eps_p, u = kwargs.get("eps_p"), kwargs.get("u")
        if eps_p is None:
            eps_p = torch.randn(max_length, N, E, device=dev)
        # word-sampling noise (word_model.py:187-198 draws torch.rand_like(logits) per step): injected `u` [T,N,V] for parity
        # tests, otherwise drawn inside the vocabulary GEMM from the model's counter-based generator (no [T,N,V] tensor)
        rng = self.sampling_rng(dev) if (method != "greedy" and u is None) else None
        out = F.decode_sample(dims, weights, audio, mem_lens, eps_p.to(dev).contiguous(),
                              None if u is None else u.to(dev).contiguous(), method, kwargs.get("temp", 1),
                              self.start_idx, self.end_idx, keep_latents=kwargs.get("keep_latents", False), rng_state=rng)
        if K > 1:
            out["seqs"] = out["seqs"].view(clips, K, max_length)
        return out


class Hybrid_VAEModel(_FusedVAEBase):
    """vae_model.py:674-698: contextual posterior + autoregressive prior + global constraint."""
    variant = 0

    def __init__(self, Audioencoder, Textdecoder, **kwargs):
        super().__init__(Audioencoder, Textdecoder, **kwargs)
        E = Textdecoder.embed_size
        self.mean_log_out = nn.Linear(E, 2 * E)
        if E != Audioencoder.embed_size:
            self.ln = nn.Linear(Audioencoder.embed_size, E)
            nn.init.xavier_uniform_(self.ln.weight)
        nn.init.xavier_uniform_(self.mean_log_out.weight)


class VAEModel(_FusedVAEBase):
    """vae_model.py:12-38: same step, AR posterior, no global head, extra `rnn_input` output."""
    variant = 1

    def __init__(self, Audioencoder, Textdecoder, **kwargs):
        super().__init__(Audioencoder, Textdecoder, **kwargs)
        E = Textdecoder.embed_size
        if E != Audioencoder.embed_size:
            self.ln = nn.Linear(Audioencoder.embed_size, E)
            nn.init.xavier_uniform_(self.ln.weight)


class PrecomputedEncoder(nn.Module):
    """Stands in for an audio encoder when its output is already available
    (hot-path-only form): returns the contract of
    models/encoder.py:672-707 for `feats = audio_embeds`."""

    def __init__(self, embed_size):
        super().__init__()
        self.embed_size = embed_size

    def forward(self, feats, feat_lens):
        return {"audio_embeds": feats, "audio_embeds_pooled": None,
                "audio_embeds_lens": feat_lens, "state": None}
