"""Drop-in loss callables of the reference runner's boundary
(`utils/train_util.py:234-266`; used at `runners/pytorch_runner_vae.py:222-227, 315-320`),
backed by the fused CUDA kernels.
"""
from __future__ import annotations

import torch

from . import functional as F
from .lazy import LazyLogits


class LabelSmoothingLoss(torch.nn.Module):
    """utils/train_util.py:234-251.  `forward(logit, target)`:
    a packed `LazyLogits` runs the fused vocabulary-projection + CE kernels and
    never materialises `[M,V]`; a dense tensor follows the reference arithmetic."""

    def __init__(self, classes, smoothing=0.0, device=0, dim=-1):
        super().__init__()
        self.confidence = 1.0 - smoothing
        self.smoothing = smoothing
        self.cls = classes
        self.dim = dim
        self.device = device

    def forward(self, logit, target):
        if isinstance(logit, LazyLogits):
            if logit.dim() != 2:
                raise ValueError("fused CE expects packed rows [M,V] (pack_padded_sequence(...).data)")
            return F.VocabCEFn.apply(logit.hidden, logit.cls_w, logit.cls_b, target, self.smoothing,
                                     logit.row_lse, logit.row_sum, logit.grad_sink)
        pred = logit.log_softmax(dim=self.dim)
        with torch.no_grad():
            true_dist = torch.zeros_like(pred)
            true_dist.fill_(self.smoothing / (self.cls - 1))
            true_dist.scatter_(1, target.unsqueeze(1).long().to(pred.device), self.confidence)
        return torch.mean(torch.sum(-true_dist * pred, dim=self.dim))


class CrossEntropyLoss(torch.nn.Module):
    """`torch.nn.CrossEntropyLoss()` replacement for the `label_smoothing: False`
    branch (pytorch_runner_vae.py:226): smoothing 0 on the fused path."""

    def forward(self, logit, target):
        if isinstance(logit, LazyLogits):
            return F.VocabCEFn.apply(logit.hidden, logit.cls_w, logit.cls_b, target, 0.0, logit.row_lse, logit.row_sum,
                                     logit.grad_sink)
        return torch.nn.functional.cross_entropy(logit, target.long().to(logit.device))


class Normal_kl_loss(torch.nn.Module):
    """utils/train_util.py:253-266: sum over d, mean over ALL positions (padding included)."""

    def __init__(self, device=0, dim=-1):
        super().__init__()
        self.dim = dim
        self.device = device

    def forward(self, mu1, lv1, mu2, lv2):
        return F.NormalKLFn.apply(mu1, lv1, mu2, lv2)


class FusedVAELoss(torch.nn.Module):
    """The runner's whole loss composition as one call (opt-in fast path; the three separate callables above stay
    the drop-in boundary): `criterion(packed_logits, targets) + kl_weight * kl_loss(q_means, q_logs, p_means, p_logs)
    + alpha * MSE(q_means_utt, p_means_utt)` (runners/pytorch_runner_vae.py:94-98, 315-320) -- one autograd node, five
    small launches instead of ~25.  `forward(output, packed_logits, targets, kl_weight)` -> loss; `.terms` holds the
    device vector {loss, ce, kl, mse} of the last call (for logging, no sync)."""

    def __init__(self, classes, smoothing=0.0, alpha=None):
        super().__init__()
        self.cls, self.smoothing, self.alpha = classes, float(smoothing), alpha
        self.terms = None

    def forward_padded(self, output, targets_padded, row_w, kl_weight):
        """The same loss WITHOUT packing: the cross-entropy runs over all N*T rows of `output["logits"]` with row weights
        (`row_w[n,t] = 1` for t < cap_lens[n] - 1, else 0 -- exactly the rows `pack_padded_sequence` keeps,
        pytorch_runner_vae.py:89-95) and the padded targets `caps[:, 1:T+1]`; `Hybrid_VAEModel.prepare_batch` stages both.
        The mean is still over the valid tokens, the gradient lands directly in the [N,T,H] layout -- no gather of the
        hidden rows before and no scatter of their gradient after the criterion (five small kernels per step)."""
        lg = output["logits"]
        if not isinstance(lg, LazyLogits) or lg.dim() != 3:
            raise ValueError("forward_padded expects the un-packed LazyLogits of the model's output dict")
        H = lg.hidden.shape[-1]
        have_g = self.alpha is not None and output.get("q_means_utt") is not None and output.get("p_means_utt") is not None
        loss, self.terms = F.VAELossFn.apply(
            lg.hidden.reshape(-1, H), lg.cls_w, lg.cls_b, targets_padded.reshape(-1), self.smoothing, lg.row_lse.reshape(-1),
            lg.row_sum.reshape(-1), lg.grad_sink, output["q_means"], output["q_logs"], output["p_means"], output["p_logs"],
            output["q_means_utt"] if have_g else None, output["p_means_utt"] if have_g else None,
            float(kl_weight), float(self.alpha) if have_g else 0.0, row_w.reshape(-1))
        return loss

    def forward(self, output, packed_logits, targets, kl_weight):
        if not isinstance(packed_logits, LazyLogits) or packed_logits.dim() != 2:
            raise ValueError("FusedVAELoss expects the packed LazyLogits rows (pack_padded_sequence(output['logits'], ...).data)")
        lg = packed_logits
        have_g = self.alpha is not None and output.get("q_means_utt") is not None and output.get("p_means_utt") is not None
        loss, self.terms = F.VAELossFn.apply(
            lg.hidden, lg.cls_w, lg.cls_b, targets, self.smoothing, lg.row_lse, lg.row_sum, lg.grad_sink,
            output["q_means"], output["q_logs"], output["p_means"], output["p_logs"],
            output["q_means_utt"] if have_g else None, output["p_means_utt"] if have_g else None,
            float(kl_weight), float(self.alpha) if have_g else 0.0)
        return loss
