"""Diverse sampling (BASELINE configs[3]: K prior-sampled captions per clip, clips partitioned over ranks) as ONE CUDA
graph per shape, plus the end-of-run gather of the ids.

`Hybrid_VAEModel.inference_forward` (reference `models/vae_model.py:880-894`, `700-720`) enqueues ~17 kernels per decode
step and never synchronises (the early stop is a device flag every kernel reads), so the whole `max_length`-step loop is
capturable.  At 1310 sequences per GPU (1045 clips x 10 over 8 ranks) the eager loop is launch-bound; replaying a graph
removes the per-launch host cost.

    sampler = GraphSampler(model, clips=131, Te=62, n_captions=10, max_length=20, method="sample")
    seqs = sampler(audio_embeds, audio_embeds_lens)           # [clips, K, max_length] int64 on the device
    ids = gather_captions(seqs, n_total_clips=1045)           # rank 0: [1045, K, max_length] (CPU); other ranks: None
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .parallel import shard_range


class GraphSampler:
    def __init__(self, model, clips: int, Te: int, n_captions: int = 1, max_length: Optional[int] = None, method: str = "sample",
                 temp: float = 1.0, inject_noise: bool = False, device=None):
        p0 = next(model.parameters())
        dev = torch.device(device) if device is not None else p0.device
        if dev.type != "cuda":
            raise RuntimeError("acvae_b200 needs CUDA tensors: there is no CPU path")
        if method not in ("greedy", "sample", "gumbel"):
            raise ValueError("GraphSampler captures the stepwise loop (greedy / sample / gumbel); beam and dbs run eagerly")
        self.model, self.clips, self.Te, self.K = model, int(clips), int(Te), int(n_captions)
        self.max_length = int(max_length if max_length is not None else model.max_length)
        self.method, self.temp = method, float(temp)
        Eenc = model.encoder.embed_size if hasattr(model, "ln") else model.decoder.embed_size
        E, V = model.decoder.embed_size, model.decoder.vocab_size
        N = self.clips * self.K
        self.audio = torch.zeros(self.clips, self.Te, Eenc, device=dev)
        self.mem_lens = torch.full((self.clips,), self.Te, dtype=torch.int32, device=dev)
        #: static noise buffers the captured loop reads when `inject_noise` (parity tests); otherwise the graph draws its own
        self.eps_p = torch.zeros(self.max_length, N, E, device=dev) if inject_noise else None
        self.u = torch.zeros(self.max_length, N, V, device=dev) if inject_noise and method != "greedy" else None
        self.out = None
        was_training = model.training
        model.eval()
        with torch.cuda.device(dev), torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._run()                                      # warm-up: lazy initialisation happens outside the capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()
        model.train(was_training)

    def _run(self):
        kw = {}
        if self.eps_p is not None:
            kw["eps_p"] = self.eps_p
        if self.u is not None:
            kw["u"] = self.u
        return self.model.inference_forward({"audio_embeds": self.audio, "audio_embeds_lens": self.mem_lens}, method=self.method,
                                            max_length=self.max_length, temp=self.temp, n_captions=self.K, **kw)

    def __call__(self, audio_embeds: torch.Tensor, audio_embeds_lens) -> torch.Tensor:
        """Copies the clip memory into the graph's static buffers (asynchronously when the source is pinned), replays the
        decode loop and returns the static `seqs` tensor ([clips, K, L], or [clips, L] when K == 1): valid until the next call."""
        if tuple(audio_embeds.shape) != tuple(self.audio.shape):
            raise ValueError(f"GraphSampler was captured for audio_embeds {tuple(self.audio.shape)}, got {tuple(audio_embeds.shape)}")
        self.audio.copy_(audio_embeds, non_blocking=True)
        self.mem_lens.copy_(torch.as_tensor(audio_embeds_lens).to(torch.int32), non_blocking=True)
        self.graph.replay()
        return self.out["seqs"]


def gather_captions(seqs: torch.Tensor, n_total_clips: int, group=None) -> Optional[torch.Tensor]:
    """The optional exchange at the end of partitioned sampling: every rank holds the ids of ITS clip range
    (`parallel.shard_range`); rank 0 receives all of them in clip order as one CPU tensor [n_total_clips, ...].  One
    all_gather of the padded shards (ids are int64, a few hundred KB); without a process group the input comes back."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return seqs.cpu()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total_clips, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(seqs.shape[1:]), dtype=seqs.dtype, device=seqs.device)
    pad[:seqs.shape[0]] = seqs
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    if rank != 0:
        return None
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0).cpu()
