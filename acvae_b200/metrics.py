"""Diversity statistics of the K captions decoded per clip, computed on the device from the id tensor the sampling loop
returns.  Mirrors `utils/div_utils.py:11-44` as called by `utils/diverse_mutil.py:25-29`
(`eval_div_stats`: Div1, Div2, gDiv1); mBLEU needs the reference's Java tokenizer / scorer and stays out of scope.
"""
from __future__ import annotations

import torch

from . import _lib
from .functional import _dev, _stream


def diversity_stats(seqs: torch.Tensor, vocab_size: int, start_idx: int = 1, end_idx: int = 2) -> dict:
    """seqs [clips, K, L] int64 on the device (e.g. `model(..., method="sample", n_captions=K)["seqs"].view(clips, K, L)`
    or the `[clips, beam, L]` output of `method="dbs"`).  Returns Div1 / Div2 (means over clips, `compute_div_n`),
    gDiv1 (`compute_global_div_n(caps, 1)`: number of distinct words) and the per-clip arrays div1 / div2 (fp64)."""
    if seqs.dim() != 3 or seqs.dtype != torch.int64:
        raise ValueError("seqs must be an int64 tensor [clips, K, L]")
    if not seqs.is_cuda:
        raise RuntimeError("acvae_b200 runs on CUDA tensors only (no CPU path)")
    seqs = seqs.contiguous()
    clips, K, L = seqs.shape
    dev = seqs.device
    div1 = torch.empty(clips, dtype=torch.float64, device=dev)
    div2 = torch.empty(clips, dtype=torch.float64, device=dev)
    flags = torch.zeros(int(vocab_size), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().acvae_diversity_stats(clips, K, L, int(vocab_size), _dev(seqs, torch.int64), int(start_idx), int(end_idx),
                                               div1.data_ptr(), div2.data_ptr(), flags.data_ptr(), _stream()),
               "acvae_diversity_stats")
    return {"Div1": float(div1.mean()), "Div2": float(div2.mean()), "gDiv1": float(flags.sum()), "div1": div1, "div2": div2}
