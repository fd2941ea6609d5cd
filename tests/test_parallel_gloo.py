"""CPU, world_size 2 over gloo: the data-parallel plumbing of the training step (one flat gradient buffer, a single
all-reduce that averages it, global-norm clip on the averaged gradient -- reference semantics:
DistributedDataParallel + clip_grad_norm_, runners/pytorch_runner_vae.py:204-207, 321-322) and the clip partition
of diverse sampling (no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness  # noqa: F401  (sys.path)
from acvae_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                      # same initial weights on every rank (DDP broadcast)
        m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        flat = parallel.FlatGradBuffer(m.parameters())
        g = torch.Generator().manual_seed(100 + rank)   # a different batch per rank
        x = torch.randn(4, 5, generator=g)
        flat.zero()
        m(x).square().mean().backward()
        local = [p.grad.clone() for p in m.parameters()]
        for p, o in zip(m.parameters(), flat.offsets):   # grads live inside the flat buffer
            assert p.grad.data_ptr() == flat.flat[o:].data_ptr()
        flat.all_reduce()
        gathered = [[torch.zeros_like(t) for _ in range(world)] for t in local]
        for t, buf in zip(local, gathered):
            dist.all_gather(buf, t)
        for p, buf in zip(m.parameters(), gathered):
            assert torch.allclose(p.grad, sum(buf) / world, atol=1e-6)
        total = flat.clip_grad_norm_(0.01)
        ref = torch.sqrt(sum((sum(buf) / world).pow(2).sum() for buf in gathered))
        assert torch.allclose(total, ref, atol=1e-6)
        assert float(torch.linalg.vector_norm(flat.flat)) <= 0.01 + 1e-6
        # sampling: contiguous partition of the clips, nothing exchanged
        lo, hi = parallel.shard_range(1045, rank, world)
        spans = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(spans, torch.tensor([lo, hi]))
        assert spans[0][0] == 0 and spans[-1][1] == 1045
        for a, b in zip(spans[:-1], spans[1:]):
            assert a[1] == b[0]
        # ... and the optional gather of the ids at the end: rank 0 gets every clip's captions in clip order
        from acvae_b200 import sampler
        total_clips = 11                                   # 6 + 5: ragged shards
        lo, hi = parallel.shard_range(total_clips, rank, world)
        mine = (torch.arange(lo, hi).view(-1, 1, 1) * 100 + torch.arange(3).view(1, 3, 1) * 10 + torch.arange(4).view(1, 1, 4)).long()
        got = sampler.gather_captions(mine, total_clips)
        if rank == 0:
            want = (torch.arange(total_clips).view(-1, 1, 1) * 100 + torch.arange(3).view(1, 3, 1) * 10 + torch.arange(4).view(1, 1, 4)).long()
            assert got.shape == (total_clips, 3, 4) and torch.equal(got, want)
        else:
            assert got is None
        out.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]


def test_shard_range_covers_everything():
    for n, w in ((1045, 8), (7, 8), (32, 3), (0, 2)):
        spans = [parallel.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
