"""Fused optimizer tail of the training step: global-norm gradient clipping
(`runners/pytorch_runner_vae.py:322`, `torch.nn.utils.clip_grad_norm_`) chained with the Adam update
(`:324`; the runner builds `getattr(torch.optim, config["optimizer"])(model.parameters(),
**config["optimizer_args"])` at `:219-220`) in two launches over flat buffers, driven by the reference's LR schedules
(`utils/lr_scheduler.py:5-86`, `scheduler.step()` every iteration, `pytorch_runner_vae.py:239-257, 305`).

`FusedClipAdam(flat_grads, lr=..., max_grad_norm=...)` takes the `parallel.FlatGradBuffer` that already holds
every `param.grad`; it moves the parameters themselves into one flat buffer too (`param.data` become views, so
modules, `state_dict()` and the C-ABI calls see the same storage) and keeps both Adam moments flat.

It IS a `torch.optim.Optimizer`: one `param_groups` entry with `lr`, `betas`, `eps`, `weight_decay` and
`max_grad_norm`, so `torch.optim.lr_scheduler.*` and the reference's `ExponentialDecayScheduler` / `NoamScheduler` /
`WarmupLinearSchedule` attach unchanged, and `state_dict()` / `load_state_dict()` use torch.optim.Adam's layout
(`state[i] = {step, exp_avg, exp_avg_sq}`, the runner's `"optimizer"` checkpoint entry, `:382`).

The hyper-parameters are NOT baked into the launch: every assignment to the group (what a scheduler does) lands in a
pinned host vector, `step()` enqueues its 24-byte copy to the device and `clip_adam_kernel` reads the device copy.
`step()` = `clip_grad_norm_` + `optimizer.step()` of the reference loop and is CUDA-graph capturable: the step counter
lives on the device, and a replayed graph re-reads the pinned vector, so a schedule keeps working without re-capture.

Semantics notes.  (1) Every parameter of the flat buffer is updated every step; a parameter whose gradient was not
produced this step sees a ZERO gradient (its moments decay, weight decay applies), whereas torch.optim.Adam skips
`grad is None` parameters -- the hot path produces every gradient every step, frozen parameters are not in the buffer.
(2) No fallback: CPU parameters raise.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .parallel import FlatGradBuffer

_HYPER_KEYS = ("max_grad_norm", "lr", "betas", "eps", "weight_decay")


class _HyperGroup(dict):
    """The optimizer's param group: a dict whose hyper-parameter writes are mirrored into the pinned host vector the
    device copy is fed from (an LR scheduler only ever does `group["lr"] = value`)."""

    _owner = None

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        if self._owner is not None and key in _HYPER_KEYS:
            self._owner._push_hyper()

    def update(self, *a, **k):
        super().update(*a, **k)
        if self._owner is not None:
            self._owner._push_hyper()


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, flat_grads: FlatGradBuffer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None, write_clipped_grads: bool = True):
        g = flat_grads
        if g.flat.device.type != "cuda" or g.flat.dtype != torch.float32:
            raise RuntimeError("FusedClipAdam needs fp32 CUDA parameters (no CPU fallback)")
        if not (lr >= 0 and 0 <= betas[0] < 1 and 0 <= betas[1] < 1 and eps >= 0):
            raise ValueError("bad hyper-parameter")
        self.grads = g
        self.write_clipped_grads = bool(write_clipped_grads)
        self._hyper_host = None
        defaults = dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps),
                        weight_decay=float(weight_decay), max_grad_norm=float(max_grad_norm) if max_grad_norm else 0.0)
        super().__init__(g.params, defaults)
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
        # hyper-parameters: pinned host vector -> device vector {max_norm, lr, beta1, beta2, eps, weight_decay}
        self._hyper_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._hyper_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        grp = _HyperGroup(self.param_groups[0])
        grp._owner = self
        self.param_groups[0] = grp
        self._push_hyper()
        self._bind_state()

    # ---- hyper-parameters -------------------------------------------------------------------------
    def _push_hyper(self) -> None:
        if self._hyper_host is None:
            return
        g = self.param_groups[0]
        lr = g["lr"]
        h = self._hyper_host
        h[0] = float(g.get("max_grad_norm") or 0.0)
        h[1] = float(lr)                      # a tensor lr (torch's capturable mode) is read here, on the host
        h[2], h[3] = float(g["betas"][0]), float(g["betas"][1])
        h[4], h[5] = float(g["eps"]), float(g["weight_decay"])

    # convenience accessors kept from the first version of this class
    @property
    def lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    @lr.setter
    def lr(self, v: float) -> None:
        self.param_groups[0]["lr"] = float(v)

    @property
    def max_grad_norm(self) -> float:
        return float(self.param_groups[0]["max_grad_norm"])

    # ---- torch.optim.Optimizer state in torch.optim.Adam's layout (views of the flat moments) ---------------------
    def _bind_state(self) -> None:
        for p, o in zip(self.grads.params, self.grads.offsets):
            self.state[p] = {"step": self.step_count.view(()),
                             "exp_avg": self.exp_avg[o:o + p.numel()].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p)}

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Zeroes the flat buffer (param.grad stay views of it; `set_to_none` is ignored on purpose)."""
        self.grads.zero()

    @torch.no_grad()
    def step(self, closure=None) -> torch.Tensor:
        """clip_grad_norm_(params, max_grad_norm) + Adam step; returns the pre-clip global norm (device scalar)."""
        if closure is not None:
            raise RuntimeError("FusedClipAdam.step does not take a closure")
        g = self.grads.flat
        with torch.cuda.device(g.device):
            self._hyper_dev.copy_(self._hyper_host, non_blocking=True)     # 32 bytes; re-read from pinned memory on every graph replay
            _lib.check(_lib.lib().acvae_clip_adam_dev(
                g.numel(), self.flat_params.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self._hyper_dev.data_ptr(), self.step_count.data_ptr(), self.total_norm.data_ptr(),
                int(self.write_clipped_grads), self._ws.data_ptr(), self._ws.numel() * 4,
                torch.cuda.current_stream(g.device).cuda_stream), "clip_adam")
        return self.total_norm

    def state_dict(self):
        sd = super().state_dict()
        sd["param_groups"] = [dict(gr) for gr in sd["param_groups"]]
        return sd

    def load_state_dict(self, sd) -> None:
        """Accepts a torch.optim.Adam-layout checkpoint (this class's own `state_dict()` or stock Adam's over the same
        parameters in the same order): moments are copied INTO the flat buffers, hyper-parameters into the group."""
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.grads.params):
            raise ValueError("checkpoint does not match this optimizer's single parameter group")
        step = None
        with torch.no_grad():
            for idx, (p, o) in zip(groups[0]["params"], zip(self.grads.params, self.grads.offsets)):
                st = sd["state"].get(idx)
                if st is None:
                    continue
                self.exp_avg[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                step = int(st["step"]) if step is None else max(step, int(st["step"]))
            if step is not None:
                self.step_count.fill_(step)
        grp = self.param_groups[0]
        for k in _HYPER_KEYS:
            if k in groups[0]:
                grp[k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        for k, v in groups[0].items():      # scheduler bookkeeping such as initial_lr
            if k not in _HYPER_KEYS and k != "params":
                dict.__setitem__(grp, k, v)
        self._bind_state()


class PeerMemoryUnavailable(RuntimeError):
    """Raised by DistributedClipAdam ON EVERY RANK when some rank cannot export or map the flat buffers through CUDA IPC (e.g. the
    caching allocator runs with expandable segments, or the GPUs have no peer access): the caller falls back to
    `FlatGradBuffer.all_reduce()` + `FusedClipAdam` on all ranks together."""


class DistributedClipAdam(FusedClipAdam):
    """Data-parallel optimizer tail as ONE fused compute + collective over NVLink peer memory (one process per GPU, one
    node): replaces DDP's gradient all-reduce (`pytorch_runner_vae.py:204-207, 321`) + `clip_grad_norm_` (`:322`) +
    `optimizer.step()` (`:324`) -- i.e. `flat.all_reduce(); FusedClipAdam.step()` -- by `step()` alone.

    Every rank owns 1/world of the flat buffers (ZeRO-1 style): it averages its shard of the gradients reading the peers'
    gradient buffers directly over NVLink, the W partial norms give the global norm for the clip, Adam runs on the shard
    and the new parameters are written straight into every rank's parameter buffer (`csrc/dp_optim.cuh`).  No NCCL call
    on this path; per step each GPU moves (W-1)/W of the buffer in and out over NVLink and runs Adam on 1/W of the
    parameters; all ranks end with bit-identical parameters.  Buffers are exchanged once, at construction, as CUDA IPC
    handles through `torch.distributed` (any backend).

    Differences from the all-reduce form: `param.grad` keeps the rank's LOCAL gradient (the averaged, clipped gradient only
    exists shard-wise), and the moments are sharded (`state_dict()` is per rank).  world == 1 is the plain `FusedClipAdam`."""

    def __init__(self, flat_grads: FlatGradBuffer, process_group=None, **kw):
        import ctypes as C
        import torch.distributed as dist
        kw.setdefault("write_clipped_grads", False)
        super().__init__(flat_grads, **kw)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        if self.world == 1:
            return
        l = _lib.lib()
        g, dev = self.grads.flat, self.grads.flat.device
        n = g.numel()
        if n % (4 * self.world):
            raise ValueError("flat buffer length must be a multiple of 4 * world_size")
        shard = n // self.world
        with torch.cuda.device(dev):
            # sharded state (the full-size moments of the base class are dropped)
            self.exp_avg = torch.zeros(shard, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(shard, dtype=torch.float32, device=dev)
            self.grad_shard = torch.zeros(shard, dtype=torch.float32, device=dev)
            self._comm = torch.zeros(l.acvae_dp_comm_bytes() // 4, dtype=torch.int32, device=dev)
            self._dp_ws = torch.zeros(l.acvae_dp_workspace_bytes() // 4 + 1, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()

            def export(t):
                h = (C.c_ubyte * 64)()
                off = C.c_int64(0)
                _lib.check(l.acvae_ipc_export(t.data_ptr(), h, C.byref(off)), "acvae_ipc_export")
                return bytes(h), int(off.value)
            # Both phases (export, map) are agreed on by all ranks: a failure anywhere raises PeerMemoryUnavailable everywhere.
            try:
                mine = {"grads": export(g), "params": export(self.flat_params), "comm": export(self._comm), "rank": self.rank}
            except Exception as e:                   # noqa: BLE001 -- reported to the peers below
                mine = {"error": f"rank {self.rank}: {e}"}
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=process_group)
            errors = [x["error"] for x in everyone if "error" in x]
            if errors:
                raise PeerMemoryUnavailable("CUDA IPC export failed: " + "; ".join(errors))
            self._peer_ptrs = {}
            err = None
            try:
                for key, local in (("grads", g), ("params", self.flat_params), ("comm", self._comm)):
                    arr = (C.c_void_p * self.world)()
                    for q, info in enumerate(everyone):
                        if q == self.rank:
                            arr[q] = local.data_ptr()
                        else:
                            hb, off = info[key]
                            out = C.c_void_p()
                            _lib.check(l.acvae_ipc_open((C.c_ubyte * 64).from_buffer_copy(hb), off, C.byref(out)), "acvae_ipc_open")
                            arr[q] = out.value
                    self._peer_ptrs[key] = arr
            except Exception as e:                   # noqa: BLE001
                err = f"rank {self.rank}: {e}"
            outcome = [None] * self.world
            dist.all_gather_object(outcome, err, group=process_group)     # also the barrier: every rank has mapped every buffer
            errors = [x for x in outcome if x]
            if errors:
                raise PeerMemoryUnavailable("CUDA IPC mapping failed: " + "; ".join(errors))
        self.state.clear()                          # moments are sharded: no per-parameter views

    @torch.no_grad()
    def step(self, closure=None) -> torch.Tensor:
        if self.world == 1:
            return super().step(closure)
        if closure is not None:
            raise RuntimeError("DistributedClipAdam.step does not take a closure")
        g = self.grads.flat
        with torch.cuda.device(g.device):
            self._hyper_dev.copy_(self._hyper_host, non_blocking=True)
            _lib.check(_lib.lib().acvae_dp_clip_adam(
                self.world, self.rank, g.numel(), self._peer_ptrs["grads"], self._peer_ptrs["params"], self._peer_ptrs["comm"],
                self.grad_shard.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self._hyper_dev.data_ptr(),
                self.step_count.data_ptr(), self.total_norm.data_ptr(), self._dp_ws.data_ptr(), self._dp_ws.numel() * 4,
                torch.cuda.current_stream(g.device).cuda_stream), "acvae_dp_clip_adam")
        return self.total_norm

    def state_dict(self):
        if self.world == 1:
            return super().state_dict()
        return {"world": self.world, "rank": self.rank, "step": self.step_count.clone(), "exp_avg_shard": self.exp_avg.clone(),
                "exp_avg_sq_shard": self.exp_avg_sq.clone(), "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd) -> None:
        if self.world == 1:
            return super().load_state_dict(sd)
        if sd["world"] != self.world or sd["rank"] != self.rank:
            raise ValueError("sharded optimizer state belongs to another (world, rank)")
        self.step_count.copy_(sd["step"]); self.exp_avg.copy_(sd["exp_avg_shard"]); self.exp_avg_sq.copy_(sd["exp_avg_sq_shard"])
        for k, v in sd["param_groups"][0].items():
            self.param_groups[0][k] = tuple(v) if k == "betas" else v
