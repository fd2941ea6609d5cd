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

    def all_reduce(self) -> None:
        """Average gradients over ranks: the one exchange step of data-parallel training."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:  # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(dist.get_world_size(self.group))

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
