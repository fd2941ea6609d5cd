"""torch.autograd bridges over the C-ABI (`include/acvae_b200.h`).

PyTorch is plumbing here: it owns device memory, the current stream and the
autograd tape.  All arithmetic of the hot path happens inside
`libacvae_b200.so`; there is no eager/CPU fallback -- a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

# order in which weights are handed to the autograd Function
HOT_KEYS_HYBRID = [
    "ln.weight", "ln.bias",
    "qnet.word_embedding.weight",
    "qnet.network.weight_ih_l0", "qnet.network.weight_hh_l0", "qnet.network.bias_ih_l0", "qnet.network.bias_hh_l0",
    "qnet.network.weight_ih_l0_reverse", "qnet.network.weight_hh_l0_reverse",
    "qnet.network.bias_ih_l0_reverse", "qnet.network.bias_hh_l0_reverse",
    "qnet.token_mean_log.weight", "qnet.token_mean_log.bias",
    "pnet.word_embedding.weight",
    "pnet.word_attn.h2attn.weight", "pnet.word_attn.h2attn.bias", "pnet.word_attn.v",
    "pnet.network.weight_ih_l0", "pnet.network.weight_hh_l0", "pnet.network.bias_ih_l0", "pnet.network.bias_hh_l0",
    "pnet.mean_log_out.weight", "pnet.mean_log_out.bias",
    "decoder.word_embeddings.weight",
    "decoder.attn.h2attn.weight", "decoder.attn.h2attn.bias", "decoder.attn.v",
    "decoder.model.weight_ih_l0", "decoder.model.weight_hh_l0", "decoder.model.bias_ih_l0", "decoder.model.bias_hh_l0",
    "decoder.classifier.weight", "decoder.classifier.bias",
    "mean_log_out.weight", "mean_log_out.bias",
]

# state_dict key -> (C struct field, index or None)
_FIELD = {
    "ln.weight": ("ln_w", None), "ln.bias": ("ln_b", None),
    "qnet.word_embedding.weight": ("q_emb", None),
    "qnet.network.weight_ih_l0": ("q_wih", 0), "qnet.network.weight_hh_l0": ("q_whh", 0),
    "qnet.network.bias_ih_l0": ("q_bih", 0), "qnet.network.bias_hh_l0": ("q_bhh", 0),
    "qnet.network.weight_ih_l0_reverse": ("q_wih", 1), "qnet.network.weight_hh_l0_reverse": ("q_whh", 1),
    "qnet.network.bias_ih_l0_reverse": ("q_bih", 1), "qnet.network.bias_hh_l0_reverse": ("q_bhh", 1),
    "qnet.token_mean_log.weight": ("q_head_w", None), "qnet.token_mean_log.bias": ("q_head_b", None),
    "qnet.mean_log_out.weight": ("q_head_w", None), "qnet.mean_log_out.bias": ("q_head_b", None),
    "pnet.word_embedding.weight": ("p_emb", None),
    "pnet.word_attn.h2attn.weight": ("p_attn_w", None), "pnet.word_attn.h2attn.bias": ("p_attn_b", None),
    "pnet.word_attn.v": ("p_attn_v", None),
    "pnet.network.weight_ih_l0": ("p_wih", None), "pnet.network.weight_hh_l0": ("p_whh", None),
    "pnet.network.bias_ih_l0": ("p_bih", None), "pnet.network.bias_hh_l0": ("p_bhh", None),
    "pnet.mean_log_out.weight": ("p_head_w", None), "pnet.mean_log_out.bias": ("p_head_b", None),
    "decoder.word_embeddings.weight": ("d_emb", None),
    "decoder.attn.h2attn.weight": ("d_attn_w", None), "decoder.attn.h2attn.bias": ("d_attn_b", None),
    "decoder.attn.v": ("d_attn_v", None),
    "decoder.model.weight_ih_l0": ("d_wih", None), "decoder.model.weight_hh_l0": ("d_whh", None),
    "decoder.model.bias_ih_l0": ("d_bih", None), "decoder.model.bias_hh_l0": ("d_bhh", None),
    "decoder.classifier.weight": ("cls_w", None), "decoder.classifier.bias": ("cls_b", None),
    "mean_log_out.weight": ("g_w", None), "mean_log_out.bias": ("g_b", None),
}


def _dev(t: torch.Tensor, dtype=torch.float32) -> int:
    if not t.is_cuda:
        raise RuntimeError("acvae_b200 runs on CUDA tensors only (no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("expected a contiguous tensor")
    return t.data_ptr()


def _opt(t: Optional[torch.Tensor], dtype=torch.float32):
    return None if t is None else _dev(t, dtype)


def _stream() -> int:
    """The current stream of the CURRENT device.  Every C-ABI call is made inside `_on(tensor)` below, which makes the
    tensors' device current first: the library's per-process helpers (side streams, events, kernel attributes) are keyed
    by the current device, so a model on cuda:1 while cuda:0 is current must not launch on cuda:0's stream."""
    return torch.cuda.current_stream().cuda_stream


def _on(t: torch.Tensor):
    """Context manager: the device of `t` is current (so `_stream()` and the library's device-keyed state match it)."""
    if not t.is_cuda:
        raise RuntimeError("acvae_b200 runs on CUDA tensors only (no CPU path)")
    return torch.cuda.device(t.device)


def _guard(fn):
    """Run `fn` with the device of its first CUDA tensor argument current (autograd backward: the device its forward ran
    on).  See `_stream`."""
    @functools.wraps(fn)
    def wrapped(*args, **kw):
        dev = None
        for a in args:
            if torch.is_tensor(a):
                if a.is_cuda:
                    dev = a.device
                    break
                continue
            if isinstance(a, dict):
                cands = [v.device for v in a.values() if torch.is_tensor(v) and v.is_cuda]
                if cands:
                    dev = cands[0]
                    break
                continue
            tagged = getattr(a, "_acvae_dev", None)
            if tagged is not None:
                dev = tagged
                break
        if dev is None:
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            out = fn(*args, **kw)
        if args and hasattr(args[0], "save_for_backward"):
            args[0]._acvae_dev = dev          # autograd ctx: the backward runs on the same device
        return out
    return wrapped


def pack_weights(weights: Dict[str, torch.Tensor], struct_cls=_lib.Weights):
    s = struct_cls()
    for k, t in weights.items():
        if t is None:
            continue
        field, idx = _FIELD[k]
        if idx is None:
            setattr(s, field, _dev(t))
        else:
            getattr(s, field)[idx] = _dev(t)
    return s


def make_dims(N, Te, T, E, A, V, Eenc, L=0, mem_rep=1, variant=0) -> _lib.Dims:
    return _lib.Dims(N, Te, T, E, A, V, Eenc, L, mem_rep, variant)


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def launch_count() -> int:
    return int(_lib.lib().acvae_launch_count())


@_guard
def gemm(a, b, a_trans=False, b_trans=False, bias=None, out=None, accumulate=False):
    """C = op(A) . op(B)^T (+bias) through `acvae_gemm`; returns (C, used_tensor_cores)."""
    l = _lib.lib()
    M = a.shape[1] if a_trans else a.shape[0]
    K = a.shape[0] if a_trans else a.shape[1]
    N = b.shape[1] if b_trans else b.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    used = C.c_int32(0)
    _lib.check(l.acvae_gemm(M, N, K, _dev(a), a.stride(0), int(a_trans), _dev(b), b.stride(0), int(b_trans),
                            _opt(bias), _dev(out), out.stride(0), int(accumulate), C.byref(used), _stream()), "acvae_gemm")
    return out, bool(used.value)


# --------------------------------------------------------------------------------
class TrainMeta:
    """Non-tensor inputs of one training forward (ids, lengths, noise, decisions)."""

    def __init__(self, dims, keys, caps_ids, cap_lens, mem_lens, eps_q, eps_p, tf_flags, dis_flags,
                 want_logits=False, grad_sink=None):
        self.dims, self.keys = dims, list(keys)
        #: optional {state_dict key: tensor}: the backward WRITES (overwrites) that weight's gradient straight
        #: into the given tensor (a view of a flat all-reduce buffer) and returns None for it to autograd
        self.grad_sink = grad_sink
        self.caps_ids, self.cap_lens, self.mem_lens = caps_ids, cap_lens, mem_lens
        self.eps_q, self.eps_p = eps_q, eps_p
        self.tf_flags = np.ascontiguousarray(np.asarray(tf_flags, dtype=np.uint8))
        self.dis_flags = np.ascontiguousarray(np.asarray(dis_flags, dtype=np.uint8))
        self.want_logits = want_logits


class LatentDecodeTrainFn(torch.autograd.Function):
    """Fused training forward/backward of the latent word-decoding step.

    Replaces autograd over `Hybrid_VAEModel.forward` (reference
    models/vae_model.py:732-750, 700-730) with `acvae_train_fwd` /
    `acvae_train_bwd`.
    outputs: q_means, q_logs, q_z, p_means, p_logs, p_z, outputs, q_means_utt,
             p_means_utt, | attn_weights, seqs, sampled_logprobs, logit_lse,
             logit_sum, rnn_input, logits (the tail is non-differentiable)
    """

    @staticmethod
    @_guard
    def forward(ctx, meta: TrainMeta, audio_embeds: torch.Tensor, *weights: torch.Tensor):
        l = _lib.lib()
        d = meta.dims
        dev = audio_embeds.device
        N, T, E, Te, V = d.N, d.T, d.E, d.Te, d.V
        audio_embeds = audio_embeds.contiguous()
        wmap = dict(zip(meta.keys, [w.detach().contiguous() for w in weights]))
        wstruct = pack_weights(wmap)
        f32 = dict(dtype=torch.float32, device=dev)
        o = {k: torch.empty(N, T, E, **f32) for k in ("q_means", "q_logs", "q_z", "p_means", "p_logs", "p_z", "outputs")}
        hybrid = d.variant == 0
        o["q_means_utt"] = torch.empty(N, 2 * E, **f32) if hybrid else torch.zeros(1, **f32)
        o["p_means_utt"] = torch.empty(N, 2 * E, **f32) if hybrid else torch.zeros(1, **f32)
        o["attn_weights"] = torch.empty(N, Te, T, **f32)
        o["seqs"] = torch.empty(N, T, dtype=torch.int64, device=dev)
        o["sampled_logprobs"] = torch.empty(N, T, **f32)
        o["logit_lse"] = torch.empty(N, T, **f32)
        o["logit_sum"] = torch.empty(N, T, **f32)
        o["rnn_input"] = torch.empty(N, T, 3 * E, **f32) if not hybrid else torch.zeros(1, **f32)
        o["logits"] = torch.empty(N, T, V, **f32) if meta.want_logits else torch.zeros(1, **f32)
        io = _lib.TrainIO()
        io.audio_embeds = _dev(audio_embeds)
        io.mem_lens = _dev(meta.mem_lens, torch.int32)
        io.caps_ids = _dev(meta.caps_ids, torch.int32)
        io.cap_lens = _dev(meta.cap_lens, torch.int32)
        io.eps_q = _dev(meta.eps_q)
        io.eps_p = _dev(meta.eps_p)
        io.tf_flags = meta.tf_flags.ctypes.data
        io.dis_flags = meta.dis_flags.ctypes.data
        for k in ("q_means", "q_logs", "q_z", "p_means", "p_logs", "p_z", "outputs", "attn_weights",
                  "sampled_logprobs", "logit_lse", "logit_sum"):
            setattr(io, k, _dev(o[k]))
        io.seqs = _dev(o["seqs"], torch.int64)
        if hybrid:
            io.q_means_utt = _dev(o["q_means_utt"]); io.p_means_utt = _dev(o["p_means_utt"])
        else:
            io.rnn_input = _dev(o["rnn_input"])
        if meta.want_logits:
            io.logits = _dev(o["logits"])
        nbytes = l.acvae_train_workspace_bytes(C.byref(d))
        ws = _workspace(nbytes, dev)
        _lib.check(l.acvae_train_fwd(C.byref(d), C.byref(wstruct), C.byref(io), ws.data_ptr(), ws.numel(), _stream()),
                   "acvae_train_fwd")
        ctx.meta, ctx.ws, ctx.io, ctx.wmap, ctx.audio = meta, ws, io, wmap, audio_embeds
        ctx.need_audio_grad = audio_embeds.requires_grad
        order = ["q_means", "q_logs", "q_z", "p_means", "p_logs", "p_z", "outputs", "q_means_utt", "p_means_utt",
                 "attn_weights", "seqs", "sampled_logprobs", "logit_lse", "logit_sum", "rnn_input", "logits"]
        res = tuple(o[k] for k in order)
        ctx.mark_non_differentiable(*res[9:])
        ctx.set_materialize_grads(False)     # outputs the loss never touches arrive as None, not as zero tensors
        # `io` holds raw pointers into these outputs; saving them keeps the storage alive for backward
        ctx.save_for_backward(*res[:9])
        return res

    @staticmethod
    @_guard
    def backward(ctx, *g):
        l = _lib.lib()
        meta, d = ctx.meta, ctx.meta.dims
        dev = ctx.audio.device
        names = ["d_q_means", "d_q_logs", "d_q_z", "d_p_means", "d_p_logs", "d_p_z", "d_outputs",
                 "d_q_means_utt", "d_p_means_utt"]
        gin = _lib.TrainGradsIn()
        keep = []
        for name, gt in zip(names, g[:9]):
            if gt is None or (d.variant == 1 and name.endswith("_utt")):
                continue
            gt = gt.contiguous()
            keep.append(gt)
            setattr(gin, name, _dev(gt))
        # one flat gradient buffer, carved per weight (a single NCCL all-reduce can cover it); weights with a
        # caller-provided sink are written in place there instead
        sink = meta.grad_sink or {}
        own = [k for k in meta.keys if k not in sink and not k.startswith("decoder.classifier")]
        sizes = [ctx.wmap[k].numel() for k in own]
        offs = np.concatenate([[0], np.cumsum([(s + 63) // 64 * 64 for s in sizes])]) if own else np.zeros(1)
        gmap = {}
        if own:
            flat = torch.zeros(int(offs[-1]), dtype=torch.float32, device=dev)
            gmap = {k: flat[int(offs[i]):int(offs[i]) + sizes[i]].view_as(ctx.wmap[k]) for i, k in enumerate(own)}
        for k, t in sink.items():
            if k in ctx.wmap:
                if t.shape != ctx.wmap[k].shape or not t.is_contiguous():
                    raise RuntimeError(f"grad sink for {k} must be a contiguous tensor of the weight's shape")
                gmap[k] = t
        # cls_w / cls_b are produced by VocabCEFn (acvae_vocab_ce_bwd), never here
        gstruct = pack_weights({k: v for k, v in gmap.items() if not k.startswith("decoder.classifier")}, _lib.WeightGrads)
        d_audio = torch.empty_like(ctx.audio) if ctx.need_audio_grad else None
        wstruct = pack_weights(ctx.wmap)
        _lib.check(l.acvae_train_bwd(C.byref(d), C.byref(wstruct), C.byref(ctx.io), C.byref(gin), C.byref(gstruct),
                                     _opt(d_audio), ctx.ws.data_ptr(), ctx.ws.numel(), _stream()), "acvae_train_bwd")
        grads = []
        for k in meta.keys:
            grads.append(None if (k.startswith("decoder.classifier") or k in sink) else gmap[k])
        return (None, d_audio, *grads)


# --------------------------------------------------------------------------------
class VocabLogitsFn(torch.autograd.Function):
    """logits = hidden @ W^T + b  (reference models/decoder.py:199), materialised."""

    @staticmethod
    @_guard
    def forward(ctx, hidden, cls_w, cls_b):
        l = _lib.lib()
        shp = hidden.shape
        h2 = hidden.contiguous().view(-1, shp[-1])
        M, E = h2.shape
        V = cls_w.shape[0]
        out = torch.empty(M, V, dtype=torch.float32, device=hidden.device)
        w, b = cls_w.detach().contiguous(), cls_b.detach().contiguous()
        _lib.check(l.acvae_vocab_logits(M, V, E, _dev(h2), _dev(w), _dev(b), _dev(out), _stream()), "acvae_vocab_logits")
        ctx.save_for_backward(h2, w)
        ctx.shp = shp
        return out.view(*shp[:-1], V)

    @staticmethod
    @_guard
    def backward(ctx, g):
        l = _lib.lib()
        h2, w = ctx.saved_tensors
        M, E = h2.shape
        V = w.shape[0]
        g2 = g.contiguous().view(M, V)
        dh = torch.empty_like(h2); dw = torch.empty_like(w)
        db = torch.empty(V, dtype=torch.float32, device=h2.device)
        _lib.check(l.acvae_vocab_logits_bwd(M, V, E, _dev(h2), _dev(w), _dev(g2), _dev(dh), _dev(dw), _dev(db), _stream()),
                   "acvae_vocab_logits_bwd")
        return dh.view(ctx.shp), dw, db


class VocabCEFn(torch.autograd.Function):
    """Label-smoothed CE over packed rows without materialising logits.

    Replaces `criterion(packed_logits, targets)` (reference
    runners/pytorch_runner_vae.py:315 with utils/train_util.py:244-251).
    """

    @staticmethod
    @_guard
    def forward(ctx, hidden, cls_w, cls_b, targets, smoothing, row_lse, row_sum, grad_sink=None):
        l = _lib.lib()
        ctx.grad_sink = grad_sink     # optional (dW, db) tensors written in place (see TrainMeta.grad_sink)
        h2 = hidden.contiguous()
        M, E = h2.shape
        V = cls_w.shape[0]
        w, b = cls_w.detach().contiguous(), cls_b.detach().contiguous()
        tg = targets.to(device=h2.device, dtype=torch.int32).contiguous()
        have = row_lse is not None and row_sum is not None
        if have:
            row_lse, row_sum = row_lse.contiguous(), row_sum.contiguous()
        else:
            row_lse = torch.empty(M, dtype=torch.float32, device=h2.device)
            row_sum = torch.empty(M, dtype=torch.float32, device=h2.device)
        loss = torch.empty((), dtype=torch.float32, device=h2.device)
        nbytes = l.acvae_vocab_workspace_bytes(M, V, E)
        ws = _workspace(nbytes, h2.device)
        _lib.check(l.acvae_vocab_ce_fwd(M, V, E, _dev(h2), _dev(w), _dev(b), _dev(tg, torch.int32), None,
                                        float(smoothing), int(have), _dev(row_lse), _dev(row_sum), _dev(loss),
                                        ws.data_ptr(), ws.numel(), _stream()), "acvae_vocab_ce_fwd")
        ctx.save_for_backward(h2, w, b, tg, row_lse)
        ctx.smoothing, ctx.ws = float(smoothing), ws
        return loss

    @staticmethod
    @_guard
    def backward(ctx, g):
        l = _lib.lib()
        h2, w, b, tg, row_lse = ctx.saved_tensors
        M, E = h2.shape
        V = w.shape[0]
        g = g.contiguous().to(torch.float32)
        dh = torch.empty_like(h2)
        if ctx.grad_sink is not None:
            dw, db = ctx.grad_sink
        else:
            dw = torch.empty_like(w)
            db = torch.empty(V, dtype=torch.float32, device=h2.device)
        _lib.check(l.acvae_vocab_ce_bwd(M, V, E, _dev(h2), _dev(w), _dev(b), _dev(tg, torch.int32), None, ctx.smoothing,
                                        _dev(row_lse), _dev(g), _dev(dh), _dev(dw), _dev(db),
                                        ctx.ws.data_ptr(), ctx.ws.numel(), _stream()), "acvae_vocab_ce_bwd")
        if ctx.grad_sink is not None:
            return dh, None, None, None, None, None, None, None
        return dh, dw, db, None, None, None, None, None


class VAELossFn(torch.autograd.Function):
    """loss = CE + kl_weight*KL + alpha*MSE as ONE autograd node (runners/pytorch_runner_vae.py:315-320):
    label-smoothed CE over packed rows from the step's vocabulary statistics (no logits), the Gaussian KL and the
    global-constraint MSE, composed on the device.  Returns (loss, terms[4] = {loss, ce, kl, mse})."""

    @staticmethod
    @_guard
    def forward(ctx, hidden, cls_w, cls_b, targets, smoothing, row_lse, row_sum, grad_sink,
                q_means, q_logs, p_means, p_logs, q_utt, p_utt, kl_weight, alpha, row_w=None):
        l = _lib.lib()
        dev = hidden.device
        h2 = hidden.contiguous()
        M, E = h2.shape
        V = cls_w.shape[0]
        w, b = cls_w.detach().contiguous(), cls_b.detach().contiguous()
        tg = targets.to(device=dev, dtype=torch.int32).contiguous()
        row_lse, row_sum = row_lse.contiguous(), row_sum.contiguous()
        rw = None if row_w is None else row_w.contiguous()             # [M] row weights: 0 drops a (padded) row from the mean
        scal = torch.empty(4, dtype=torch.float32, device=dev)          # ce, kl | g, g*kl_w
        terms = torch.empty(4, dtype=torch.float32, device=dev)
        ws = _workspace(max(l.acvae_vocab_workspace_bytes(M, V, E), 1024), dev)
        _lib.check(l.acvae_vocab_ce_fwd(M, V, E, _dev(h2), _dev(w), _dev(b), _dev(tg, torch.int32), _opt(rw), float(smoothing), 1,
                                        _dev(row_lse), _dev(row_sum), scal.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "acvae_vocab_ce_fwd")
        kls = [t.contiguous() for t in (q_means, q_logs, p_means, p_logs)]
        Ez = kls[0].shape[-1]
        rows = kls[0].numel() // Ez
        _lib.check(l.acvae_kl_fwd(rows, Ez, *[_dev(t) for t in kls], scal.data_ptr() + 4, ws.data_ptr(), ws.numel(), _stream()),
                   "acvae_kl_fwd")
        have_g = q_utt is not None and p_utt is not None and alpha
        qu = q_utt.contiguous() if have_g else None
        pu = p_utt.contiguous() if have_g else None
        _lib.check(l.acvae_loss_combine_fwd(qu.numel() if have_g else 0, _opt(qu), _opt(pu), scal.data_ptr(), scal.data_ptr() + 4,
                                            float(kl_weight), float(alpha or 0.0), terms.data_ptr(), _stream()),
                   "acvae_loss_combine_fwd")
        ctx.save_for_backward(h2, w, b, tg, row_lse, *kls, *([qu, pu] if have_g else []))
        ctx.have_g, ctx.smoothing, ctx.kl_weight, ctx.alpha = bool(have_g), float(smoothing), float(kl_weight), float(alpha or 0.0)
        ctx.ws, ctx.scal, ctx.grad_sink, ctx.rw = ws, scal, grad_sink, rw
        ctx.mark_non_differentiable(terms)
        return terms[0], terms

    @staticmethod
    @_guard
    def backward(ctx, g, _g_terms):
        l = _lib.lib()
        saved = ctx.saved_tensors
        h2, w, b, tg, row_lse = saved[:5]
        kls = saved[5:9]
        qu, pu = (saved[9], saved[10]) if ctx.have_g else (None, None)
        M, E = h2.shape
        V = w.shape[0]
        g = g.contiguous().to(torch.float32)
        dqu = torch.empty_like(qu) if ctx.have_g else None
        dpu = torch.empty_like(pu) if ctx.have_g else None
        scal = ctx.scal
        _lib.check(l.acvae_loss_combine_bwd(qu.numel() if ctx.have_g else 0, _opt(qu), _opt(pu), _dev(g), ctx.kl_weight, ctx.alpha,
                                            _opt(dqu), _opt(dpu), scal.data_ptr() + 8, _stream()), "acvae_loss_combine_bwd")
        Ez = kls[0].shape[-1]
        rows = kls[0].numel() // Ez
        dk = [torch.empty_like(t) for t in kls]
        _lib.check(l.acvae_kl_bwd(rows, Ez, *[_dev(t) for t in kls], scal.data_ptr() + 12, *[_dev(t) for t in dk], _stream()),
                   "acvae_kl_bwd")
        dh = torch.empty_like(h2)
        if ctx.grad_sink is not None:
            dw, db = ctx.grad_sink
        else:
            dw = torch.empty_like(w)
            db = torch.empty(V, dtype=torch.float32, device=h2.device)
        sink = ctx.grad_sink is not None
        # sink mode: the classifier's weight / bias gradients land in the flat buffer and nothing reads them before the optimizer
        # (or the all-reduce, behind acvae_train_bwd's join) -- off the critical stream
        l.acvae_defer_classifier_grads(1 if sink else 0)
        try:
            _lib.check(l.acvae_vocab_ce_bwd(M, V, E, _dev(h2), _dev(w), _dev(b), _dev(tg, torch.int32), _opt(ctx.rw), ctx.smoothing,
                                            _dev(row_lse), scal.data_ptr() + 8, _dev(dh), _dev(dw), _dev(db),
                                            ctx.ws.data_ptr(), ctx.ws.numel(), _stream()), "acvae_vocab_ce_bwd")
        finally:
            l.acvae_defer_classifier_grads(0)
        return (dh, None if sink else dw, None if sink else db, None, None, None, None, None,
                dk[0], dk[1], dk[2], dk[3], dqu, dpu, None, None, None)


class NormalKLFn(torch.autograd.Function):
    """KL(q||p) summed over d, mean over all positions (utils/train_util.py:259-266)."""

    @staticmethod
    @_guard
    def forward(ctx, mu1, lv1, mu2, lv2):
        l = _lib.lib()
        ts = [t.contiguous() for t in (mu1, lv1, mu2, lv2)]
        E = ts[0].shape[-1]
        rows = ts[0].numel() // E
        out = torch.empty((), dtype=torch.float32, device=ts[0].device)
        ws = _workspace(1024, ts[0].device)
        _lib.check(l.acvae_kl_fwd(rows, E, *[_dev(t) for t in ts], _dev(out), ws.data_ptr(), ws.numel(), _stream()),
                   "acvae_kl_fwd")
        ctx.save_for_backward(*ts)
        return out

    @staticmethod
    @_guard
    def backward(ctx, g):
        l = _lib.lib()
        ts = ctx.saved_tensors
        E = ts[0].shape[-1]
        rows = ts[0].numel() // E
        outs = [torch.empty_like(t) for t in ts]
        g = g.contiguous().to(torch.float32)
        _lib.check(l.acvae_kl_bwd(rows, E, *[_dev(t) for t in ts], _dev(g), *[_dev(t) for t in outs], _stream()),
                   "acvae_kl_bwd")
        return tuple(outs)


@_guard
def vocab_stats(hidden, cls_w, cls_b):
    """(lse, sum, argmax, logprob_of_argmax) per row of hidden [M,E]; no logits stored."""
    l = _lib.lib()
    h2 = hidden.contiguous().view(-1, hidden.shape[-1])
    M, E = h2.shape
    V = cls_w.shape[0]
    dev = h2.device
    lse = torch.empty(M, dtype=torch.float32, device=dev); ssum = torch.empty_like(lse); lp = torch.empty_like(lse)
    arg = torch.empty(M, dtype=torch.int64, device=dev)
    ws = _workspace(l.acvae_vocab_workspace_bytes(M, V, E), dev)
    _lib.check(l.acvae_vocab_stats(M, V, E, _dev(h2), _dev(cls_w.detach().contiguous()), _dev(cls_b.detach().contiguous()),
                                   _dev(lse), _dev(ssum), _dev(arg, torch.int64), _dev(lp), ws.data_ptr(), ws.numel(),
                                   _stream()), "acvae_vocab_stats")
    return lse, ssum, arg, lp


@_guard
def decode_sample(dims, weights: Dict[str, torch.Tensor], audio_embeds, mem_lens, eps_p, u=None, method="greedy",
                  temp=1.0, start_idx=1, end_idx=2, keep_latents=False, rng_state=None):
    """Prior-latent stepwise decoding (reference vae_model.py:880-894, 700-720).  The word-sampling noise (word_model.py:187-198)
    is either injected (`u` [T,N,V] uniforms) or drawn inside the vocabulary GEMM's epilogue from `rng_state` (int64 device
    tensor {seed, calls}; the call increments `calls`): no [T,N,V] tensor exists in that mode."""
    l = _lib.lib()
    dev = audio_embeds.device
    N, T, E = dims.N, dims.T, dims.E
    code = {"greedy": 0, "sample": 1, "gumbel": 2}.get(method, 1)   # word_model.py:196: any other string samples
    io = _lib.SampleIO()
    audio_embeds = audio_embeds.contiguous()
    io.audio_embeds = _dev(audio_embeds); io.mem_lens = _dev(mem_lens, torch.int32); io.eps_p = _dev(eps_p)
    if code != 0:
        if u is not None:
            io.u = _dev(u)
        elif rng_state is not None:
            if rng_state.dtype != torch.int64 or rng_state.numel() != 2:
                raise ValueError("rng_state must be an int64 tensor {seed, calls}")
            io.rng_state = _dev(rng_state, torch.int64)
        else:
            raise RuntimeError("sampling methods need uniform noise `u` [T,N,V] or an `rng_state` to draw it from")
    io.method, io.temp, io.start_idx, io.end_idx = code, float(temp), int(start_idx), int(end_idx)
    out = {"seqs": torch.empty(N, T, dtype=torch.int64, device=dev),
           "sampled_logprobs": torch.empty(N, T, dtype=torch.float32, device=dev),
           "n_steps": torch.zeros((), dtype=torch.int32, device=dev)}
    io.seqs = _dev(out["seqs"], torch.int64); io.sampled_logprobs = _dev(out["sampled_logprobs"])
    io.n_steps = _dev(out["n_steps"], torch.int32)
    if keep_latents:
        for k in ("p_means", "p_logs", "p_z", "outputs"):
            out[k] = torch.empty(N, T, E, dtype=torch.float32, device=dev)
            setattr(io, k, _dev(out[k]))
    wmap = {k: v.detach().contiguous() for k, v in weights.items()}
    wstruct = pack_weights(wmap)
    ws = _workspace(l.acvae_sample_workspace_bytes(C.byref(dims)), dev)
    _lib.check(l.acvae_decode_sample(C.byref(dims), C.byref(wstruct), C.byref(io), ws.data_ptr(), ws.numel(), _stream()),
               "acvae_decode_sample")
    return out


@_guard
def beam_search(dims, weights, audio_embeds, mem_lens, eps_b, beam=3, start_idx=1):
    """Beam search with per-beam prior noise (reference vae_model.py:896-995).
    eps_b: [T, N*beam, E]."""
    l = _lib.lib()
    dev = audio_embeds.device
    seqs = torch.empty(dims.N, dims.T, dtype=torch.int64, device=dev)
    wmap = {k: v.detach().contiguous() for k, v in weights.items()}
    wstruct = pack_weights(wmap)
    audio_embeds = audio_embeds.contiguous()
    ws = _workspace(l.acvae_beam_workspace_bytes(C.byref(dims), int(beam)), dev)
    _lib.check(l.acvae_beam_search(C.byref(dims), C.byref(wstruct), _dev(audio_embeds), _dev(mem_lens, torch.int32),
                                   _dev(eps_b.contiguous()), int(beam), int(start_idx), _dev(seqs, torch.int64),
                                   ws.data_ptr(), ws.numel(), _stream()), "acvae_beam_search")
    return {"seqs": seqs}


@_guard
def diverse_beam_search(dims, weights, audio_embeds, mem_lens, eps_g, beam_size=5, group_size=5, diversity_lambda=0.5,
                        temperature=1.0, group_nbest=True, start_idx=1, end_idx=2):
    """Diverse beam search with prior latents (reference word_model.py:297-394, vae_model.py:997-1048).
    eps_g: [T + group_size - 1, N*group_size*(beam_size // group_size), E]."""
    l = _lib.lib()
    dev = audio_embeds.device
    n_out = int(beam_size) if group_nbest else int(group_size)
    seqs = torch.empty(dims.N, n_out, dims.T, dtype=torch.int64, device=dev)
    wmap = {k: v.detach().contiguous() for k, v in weights.items()}
    wstruct = pack_weights(wmap)
    audio_embeds = audio_embeds.contiguous()
    ws = _workspace(l.acvae_dbs_workspace_bytes(C.byref(dims), int(beam_size), int(group_size)), dev)
    _lib.check(l.acvae_diverse_beam_search(C.byref(dims), C.byref(wstruct), _dev(audio_embeds), _dev(mem_lens, torch.int32),
                                           _dev(eps_g.contiguous()), int(beam_size), int(group_size), float(diversity_lambda),
                                           float(temperature), int(bool(group_nbest)), int(start_idx), int(end_idx),
                                           _dev(seqs, torch.int64), ws.data_ptr(), ws.numel(), _stream()),
               "acvae_diverse_beam_search")
    return {"seqs": seqs}


class EncoderHandoffFn(torch.autograd.Function):
    """audio_embeds [N,Te,C] (+ pooled [N,C]) from the last convolution block's output [N,C,Te,F] in one pass: replaces
    `torch.mean(x, dim=3)` + `x.transpose(1, 2).contiguous()` of Cnn10.forward (models/encoder.py:691-700)."""

    @staticmethod
    @_guard
    def forward(ctx, fmap, want_pooled):
        l = _lib.lib()
        fmap = fmap.contiguous()
        N, Cc, Te, Fq = fmap.shape
        out = torch.empty(N, Te, Cc, dtype=torch.float32, device=fmap.device)
        pooled = torch.empty(N, Cc, dtype=torch.float32, device=fmap.device) if want_pooled else None
        _lib.check(l.acvae_encoder_handoff_fwd(N, Cc, Te, Fq, _dev(fmap), _dev(out), _opt(pooled), _stream()),
                   "acvae_encoder_handoff_fwd")
        ctx.shape = (N, Cc, Te, Fq)
        if want_pooled:
            ctx.mark_non_differentiable(pooled)      # the VAE decoder ignores the pooled embedding (decoder.py:175-180)
            return out, pooled
        return out, None

    @staticmethod
    @_guard
    def backward(ctx, g, _gp):
        l = _lib.lib()
        N, Cc, Te, Fq = ctx.shape
        g = g.contiguous()
        d = torch.empty(N, Cc, Te, Fq, dtype=torch.float32, device=g.device)
        _lib.check(l.acvae_encoder_handoff_bwd(N, Cc, Te, Fq, _dev(g), _dev(d), _stream()), "acvae_encoder_handoff_bwd")
        return d, None


def encoder_handoff(fmap: torch.Tensor, lens, want_pooled: bool = False):
    """The output contract of the reference's encoders (models/encoder.py:702-707) from the feature map of the last
    convolution block: {'audio_embeds' [N,Te,C], 'audio_embeds_pooled' (max + mean over frames of the frame means, BEFORE
    the reference's dropout / embed_pooled / relu, or None), 'state': None, 'audio_embeds_lens'}.  `lens` must already be
    in frames (the reference divides by 16 in place, :677-678)."""
    out, pooled = EncoderHandoffFn.apply(fmap, bool(want_pooled))
    return {"audio_embeds": out, "audio_embeds_pooled": pooled, "state": None, "audio_embeds_lens": lens}


def set_precision(mode: str) -> None:
    """Arithmetic of the batched contractions: "fp32" (default; 3xTF32 on the tensor cores, the 1e-4 parity mode) or
    "tf32" (single-pass TF32 products, fp32 accumulation: the reduced-precision class BASELINE.json calls bf16, 2e-2).
    Process-wide; CUDA graphs captured before the switch keep the mode they were captured with."""
    modes = {"fp32": 0, "tf32": 1}
    if mode not in modes:
        raise ValueError(f"precision must be one of {sorted(modes)}")
    _lib.check(_lib.lib().acvae_set_precision(modes[mode]), "acvae_set_precision")


def get_precision() -> str:
    return {0: "fp32", 1: "tf32"}[_lib.lib().acvae_get_precision()]


_input_event_keepalive = None


def set_input_event(event: Optional[torch.cuda.Event]) -> None:
    """Let the host-to-device copy of a step's audio embeddings overlap the posterior chain: record `event` after that
    copy (on whatever stream carries it) before every call of the training forward; the forward waits for it only where
    the audio is first read (an external event-wait node when the step is captured in a CUDA graph).  `None` restores
    plain stream order.  The caller still orders the copy after the previous step's last read of the buffer."""
    global _input_event_keepalive
    if event is None:
        _lib.check(_lib.lib().acvae_set_input_event(None), "acvae_set_input_event")
        _input_event_keepalive = None
        return
    event.record()                              # materialises the cudaEvent_t; a complete event is a no-op to wait for
    _lib.check(_lib.lib().acvae_set_input_event(event.cuda_event), "acvae_set_input_event")
    _input_event_keepalive = event              # the library keeps the raw handle: keep the Python owner alive
