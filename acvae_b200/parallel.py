"""Data-parallel plumbing for the training step (reference: DistributedDataParallel over NCCL,
runners/pytorch_runner_vae.py:155-161, 204-207, gradient all-reduce inside loss.backward() :321).

One process per GPU.  Gradients of ALL trainable parameters live in ONE flat fp32 buffer
(`param.grad` are views into it), so the exchange step is a single NCCL all-reduce (AVG) over
NVLink/NVSwitch instead of DDP's per-bucket launches, and global-norm clipping is one pass over the
same buffer.  Diverse sampling needs no collective: clips are partitioned across ranks
(`shard_range`).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


class FlatGradBuffer:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64          # keep every view 256-byte aligned
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.offsets = offs
        self.group = process_group
        self.attach()

    def attach(self) -> None:
        """(Re)point every param.grad at its slice of the flat buffer."""
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def views_for(self, module: torch.nn.Module, prefix_skip: str = ""):
        """{state_dict key: grad view} for `module`'s parameters held by this buffer (the gradient sink the fused
        backward writes into, `Hybrid_VAEModel.grad_sink`)."""
        if getattr(self, "_views_cache", None) is None:
            by_id = {id(p): self.flat[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)}
            self._views_cache = {k: by_id[id(p)] for k, p in module.named_parameters()
                                 if id(p) in by_id and not (prefix_skip and k.startswith(prefix_skip))}
        return self._views_cache

    def zero(self) -> None:
        self.flat.zero_()

    def enable_bucketing(self, module: torch.nn.Module, early_prefix: str = "decoder.") -> None:
        """Overlap the exchange with the backward (what DDP's buckets do, pytorch_runner_vae.py:204-207): the gradients of the
        parameters whose name starts with `early_prefix` -- the decoder's word embeddings, GRU, classifier and attention:
        12.7 of the 32 MB, final ~0.2 ms before the backward ends -- are reduced as soon as the fused backward records
        "decoder gradients final" (`acvae_set_bucket_event`), under the posterior / prior tail of the backward; the rest
        follows at the end.  The early parameters must be contiguous in `parameters()` order (they are: one sub-module)."""
        from . import _lib
        names = {id(p): k for k, p in module.named_parameters()}
        idx = [i for i, p in enumerate(self.params) if names.get(id(p), "").startswith(early_prefix)]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            raise ValueError(f"parameters named {early_prefix}* are not one contiguous range of the flat buffer")
        lo = self.offsets[idx[0]]
        hi = self.offsets[idx[-1] + 1] if idx[-1] + 1 < len(self.offsets) else self.flat.numel()
        self._bucket = (lo, hi)
        with torch.cuda.device(self.flat.device):
            self._bucket_event = torch.cuda.Event()
            self._bucket_event.record()
            self._bucket_stream = torch.cuda.Stream()
        _lib.check(_lib.lib().acvae_set_bucket_event(self._bucket_event.cuda_event), "acvae_set_bucket_event")

    def _reduce(self, t: torch.Tensor, async_op: bool = False):
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)   # gloo (CPU tests) has no AVG
        if not async_op:
            t.div_(dist.get_world_size(self.group))
        return w

    def all_reduce(self) -> None:
        """Average gradients over ranks: the exchange step of data-parallel training.  One all-reduce of the flat buffer, or
        -- after `enable_bucketing` -- the early bucket behind the "decoder gradients final" event of the backward that has
        just been enqueued, then the remainder."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return
        bucket = getattr(self, "_bucket", None)
        if bucket is None or dist.get_backend(self.group) != "nccl":
            self._reduce(self.flat)
            return
        lo, hi = bucket
        cur = torch.cuda.current_stream(self.flat.device)
        side = self._bucket_stream
        side.wait_event(self._bucket_event)                    # only the early gradients, not the whole backward
        with torch.cuda.stream(side):
            w_early = self._reduce(self.flat[lo:hi], async_op=True)
        works = [w_early]
        if lo > 0:
            works.append(self._reduce(self.flat[:lo], async_op=True))
        if hi < self.flat.numel():
            works.append(self._reduce(self.flat[hi:], async_op=True))
        for w in works:
            w.wait()                                           # the current stream waits for the collectives
        cur.wait_stream(side)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ semantics (pytorch_runner_vae.py:322) on the flat buffer
        (padding between views is zero, so the flat 2-norm is the global norm)."""
        total = torch.linalg.vector_norm(self.flat)
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.flat.mul_(coef)
        return total


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous partition of `n_items` clips over `world` ranks (no communication needed)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
