"""Fused optimizer tail of the training step: global-norm gradient clipping
(`runners/pytorch_runner_vae.py:322`, `torch.nn.utils.clip_grad_norm_`) chained with the Adam update
(`:324`; the runner builds `getattr(torch.optim, config["optimizer"])(model.parameters(),
**config["optimizer_args"])` at `:219-220`) in two launches over flat buffers.

`FusedClipAdam(flat_grads, lr=..., max_grad_norm=...)` takes the `parallel.FlatGradBuffer` that already holds
every `param.grad`; it moves the parameters themselves into one flat buffer too (`param.data` become views, so
modules, `state_dict()` and the C-ABI calls see the same storage) and keeps both Adam moments flat.
`step()` = `clip_grad_norm_` + `optimizer.step()` of the reference loop, CUDA-graph capturable (the step
counter lives on the device).  No fallback: CPU parameters raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .parallel import FlatGradBuffer


class FusedClipAdam:
    def __init__(self, flat_grads: FlatGradBuffer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None, write_clipped_grads: bool = True):
        g = flat_grads
        if g.flat.device.type != "cuda" or g.flat.dtype != torch.float32:
            raise RuntimeError("FusedClipAdam needs fp32 CUDA parameters (no CPU fallback)")
        self.grads = g
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.write_clipped_grads = bool(write_clipped_grads)
        dev = g.flat.device
        # parameters: one flat buffer with the SAME offsets as the gradients; param.data become views
        self.flat_params = torch.zeros_like(g.flat)
        with torch.no_grad():
            for p, o in zip(g.params, g.offsets):
                view = self.flat_params[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(g.flat)
        self.exp_avg_sq = torch.zeros_like(g.flat)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.total_norm = torch.zeros((), dtype=torch.float32, device=dev)
        self._ws = torch.empty(_lib.lib().acvae_clip_adam_workspace_bytes() // 4, dtype=torch.float32, device=dev)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.grads.zero()

    @torch.no_grad()
    def step(self) -> torch.Tensor:
        """clip_grad_norm_(params, max_grad_norm) + Adam step; returns the pre-clip global norm (device scalar)."""
        g = self.grads.flat
        _lib.check(_lib.lib().acvae_clip_adam(
            g.numel(), self.flat_params.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.max_grad_norm, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
            self.step_count.data_ptr(), self.total_norm.data_ptr(), int(self.write_clipped_grads),
            self._ws.data_ptr(), self._ws.numel() * 4, torch.cuda.current_stream(g.device).cuda_stream), "clip_adam")
        return self.total_norm

    def state_dict(self):
        return {"step": self.step_count.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd) -> None:
        self.step_count.copy_(sd["step"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps, self.weight_decay = sd["lr"], tuple(sd["betas"]), sd["eps"], sd["weight_decay"]
