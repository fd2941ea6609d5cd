"""`LazyLogits`: the `[N,T,V]` logits of the reference's output dict, never
written to HBM unless somebody really asks for them.

The reference returns `output["logits"]` and its runner packs and feeds them to
the criterion (runners/pytorch_runner_vae.py:94-95, 315).  Here the dict holds
a handle over the decoder hidden states `[N,T,H]`, the classifier weights and
the per-row log-sum-exp / sum statistics the fused step already produced:

  * `torch.nn.utils.rnn.pack_padded_sequence(lazy, lens, batch_first=True).data`
    packs the hidden rows (row-wise Linear commutes with packing) and returns a
    packed `LazyLogits [M,V]`;
  * our drop-in `LabelSmoothingLoss` / `CrossEntropyLoss` consume the packed
    handle with the fused CE kernels (`acvae_vocab_ce_fwd/bwd`);
  * any other torch function materialises it once through `acvae_vocab_logits`
    (autograd-aware) and proceeds on the dense tensor.
"""
from __future__ import annotations

import torch

from . import functional as F


_PACK_CACHE = {}


def _pack_index(lengths, d0, d1, batch_first, device):
    """Flat row indices of `pack_padded_sequence(x, lengths, batch_first)` into x.reshape(d0*d1, ...) and the
    PackedSequence batch_sizes (CPU int64).  lengths: CPU int64, sorted descending (enforce_sorted=True)."""
    lens = [int(v) for v in torch.as_tensor(lengths).tolist()]
    key = (tuple(lens), d0, d1, batch_first, str(device))
    hit = _PACK_CACHE.get(key)
    if hit is not None:
        return hit
    if any(lens[i] < lens[i + 1] for i in range(len(lens) - 1)) or (lens and lens[-1] <= 0):
        raise RuntimeError("`lengths` array must be sorted in decreasing order and positive (enforce_sorted=True)")
    T = lens[0] if lens else 0
    idx, bs = [], []
    for t in range(T):
        n_t = sum(1 for v in lens if v > t)
        bs.append(n_t)
        if batch_first:
            idx.extend(n * d1 + t for n in range(n_t))
        else:
            idx.extend(t * d1 + n for n in range(n_t))
    out = (torch.tensor(idx, dtype=torch.int64).to(device), torch.tensor(bs, dtype=torch.int64))
    if len(_PACK_CACHE) > 256:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = out
    return out


class LazyLogits:
    def __init__(self, hidden, cls_w, cls_b, row_lse=None, row_sum=None):
        self.hidden, self.cls_w, self.cls_b = hidden, cls_w, cls_b
        self.row_lse, self.row_sum = row_lse, row_sum
        self._dense = None
        self.grad_sink = None       # optional (dW, db) destination of the fused CE backward

    # ---- tensor-like surface -----------------------------------------------------
    @property
    def shape(self):
        return torch.Size((*self.hidden.shape[:-1], self.cls_w.shape[0]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return len(self.shape)

    @property
    def device(self):
        return self.hidden.device

    @property
    def dtype(self):
        return self.hidden.dtype

    @property
    def is_cuda(self):
        return self.hidden.is_cuda

    @property
    def requires_grad(self):
        return self.hidden.requires_grad or self.cls_w.requires_grad

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            self._dense = F.VocabLogitsFn.apply(self.hidden, self.cls_w, self.cls_b)
        return self._dense

    def __getattr__(self, name):
        # anything not modelled above behaves like the dense tensor
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"LazyLogits(shape={tuple(self.shape)}, materialised={self._dense is not None})"

    # ---- torch function protocol ---------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name == "_pack_padded_sequence":
            lazy, lengths, batch_first = args[0], args[1], (args[2] if len(args) > 2 else kwargs.get("batch_first", False))
            # row-wise Linear commutes with packing: gather the hidden rows (and their statistics) in packed
            # order with ONE index_select each instead of the per-time-step copies of the stock pack kernel
            idx, bs = _pack_index(lengths, lazy.hidden.shape[0], lazy.hidden.shape[1], bool(batch_first), lazy.hidden.device)
            hid = lazy.hidden.reshape(-1, lazy.hidden.shape[-1]).index_select(0, idx)
            lse = ssum = None
            if lazy.row_lse is not None:
                lse = lazy.row_lse.reshape(-1).index_select(0, idx)
                ssum = lazy.row_sum.reshape(-1).index_select(0, idx)
            packed = LazyLogits(hid, lazy.cls_w, lazy.cls_b, lse, ssum)
            packed.grad_sink = lazy.grad_sink
            return packed, bs

        def dense(x):
            return x.materialize() if isinstance(x, LazyLogits) else x

        args = tuple(dense(a) for a in args)
        kwargs = {k: dense(v) for k, v in kwargs.items()}
        return func(*args, **kwargs)
