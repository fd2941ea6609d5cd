"""GPU (-m gpu), needs >= 2 GPUs (skipped otherwise): data-parallel parity of BASELINE configs[2] over real NCCL.

Two ranks (one process per GPU, `torch.distributed` / NCCL, rendezvous on 127.0.0.1) each run the fused train step on
their OWN batch of the real model (N = 32, E = 256, V = 4400) with the fused backward writing into the flat all-reduce
buffer, then `FlatGradBuffer.all_reduce()`.  The reference semantics (DDP,
runners/pytorch_runner_vae.py:204-207, 321) are: every rank's loss is the mean over ITS batch and the gradients are
averaged over ranks -- so the oracle is run rank by rank on the same batches and its gradients are averaged on the host.
"""
import os
import socket

import numpy as np
import pytest
import torch

import harness
from harness import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-4
SEEDS = (21, 22)        # model weights: seed 21 on both ranks; batch of rank r: seed SEEDS[r]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, bucketed=False):
    import torch.distributed as dist
    from acvae_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = synthetic.CFG1
    m = harness.build_model(d, SEEDS[0], device=f"cuda:{rank}")
    flat = parallel.FlatGradBuffer(m.parameters())
    m.grad_sink = flat
    if bucketed:
        flat.enable_bucketing(m)          # decoder.* gradients reduced behind the backward's "decoder gradients final" event
    with torch.cuda.device(rank):
        r = _run_rank(d, m, SEEDS[rank])
    local = flat.flat.clone()
    flat.all_reduce()
    torch.cuda.synchronize()
    views = {k: v.detach().cpu().numpy() for k, v in flat.views_for(m).items()}
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=float(r["terms"]["loss"]), local_norm=float(local.norm()), **views)
    dist.barrier()
    dist.destroy_process_group()


def _run_rank(d, m, batch_seed):
    """harness.run_cuda_train with the model's weights from one seed and the batch from another."""
    import acvae_b200 as models
    dev = next(m.parameters()).device
    b = synthetic.make_batch(d, batch_seed)
    T = int(b["cap_lens"].max()) - 1
    m.train()
    caps = torch.from_numpy(b["caps"])
    lens1 = torch.as_tensor(b["cap_lens"]) - 1
    targets = torch.nn.utils.rnn.pack_padded_sequence(caps[:, 1:], lens1, batch_first=True).data
    out = m(torch.from_numpy(b["audio_embeds"]).to(dev), torch.from_numpy(b["mem_lens"].copy()), caps, b["cap_lens"].copy(),
            ss_ratio=1.0, dis_ratio=0.0, eps_q=torch.from_numpy(b["eps_q"][:, :T].copy()),
            eps_p=torch.from_numpy(b["eps_p"][:T].copy()), tf_flags=[True] * T, dis_flags=[False] * T)
    packed = torch.nn.utils.rnn.pack_padded_sequence(out["logits"], lens1, batch_first=True).data
    fl = models.FusedVAELoss(d.V, smoothing=0.1, alpha=1.0)
    loss = fl(out, packed, targets, 0.5)
    loss.backward()
    return {"terms": {"loss": loss.detach().cpu()}}


def _oracle_rank(d, weight_seed, batch_seed):
    import acvae_oracle as oracle
    b = synthetic.make_batch(d, batch_seed)
    T = int(b["cap_lens"].max()) - 1
    p = harness.oracle_params(d, weight_seed, grad=True)
    caps = torch.from_numpy(b["caps"])
    out = oracle.train_forward(p, torch.from_numpy(b["audio_embeds"]), b["mem_lens"], caps, b["cap_lens"],
                               torch.from_numpy(b["eps_q"][:, :T]), torch.from_numpy(b["eps_p"][:T]), [True] * T, [False] * T,
                               variant="hybrid", eps_q_steps=torch.from_numpy(b["eps_q_steps"][:T]))
    terms = oracle.train_loss(out, caps, b["cap_lens"], d.V, 0.1, 0.5, 1.0, "MSE")
    terms["loss"].backward()
    return float(terms["loss"]), {k: v.grad.detach().numpy() for k, v in p.items() if v.grad is not None}


@pytest.mark.parametrize("bucketed", [False, True])
def test_two_rank_nccl_gradient_average_vs_oracle(tmp_path, bucketed):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), bucketed), nprocs=world, join=True)
    got = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world)]
    d = synthetic.CFG1
    ref = [_oracle_rank(d, SEEDS[0], SEEDS[r]) for r in range(world)]
    for r in range(world):
        assert abs(float(got[r]["loss"]) - ref[r][0]) <= TOL * max(1.0, abs(ref[r][0])), ("per-rank loss", r)
    for k in ref[0][1]:
        avg = (ref[0][1][k] + ref[1][1][k]) / world
        for r in range(world):                                # every rank holds the same averaged gradient
            harness.assert_close(got[r][k], avg, TOL, ("rank", r, k))
    for k in got[0]:
        if k not in ("loss", "local_norm"):
            assert np.array_equal(got[0][k], got[1][k]), ("ranks disagree after the all-reduce", k)


# ---- fused reduce-scatter + clip + Adam + all-gather over NVLink peer memory (DistributedClipAdam) --------------------
def _dp_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from acvae_b200 import DistributedClipAdam, parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)                                      # identical initial parameters on every rank
    shapes = [(4400, 256), (768, 768), (256,), (1, 7), (513, 3)]
    params = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    flat = parallel.FlatGradBuffer(params)
    opt = DistributedClipAdam(flat, lr=5e-4, max_grad_norm=1.0)
    init = [p.detach().cpu().clone() for p in params]
    graph = None
    gen = torch.Generator(device=dev).manual_seed(100 + rank)  # different gradients per rank
    local_grads = []
    for it, scale in enumerate((1.0, 1e-4, 3.0, 1.0)):        # clipped, not clipped, clipped; the last step replays a CUDA graph
        gs = [torch.randn(s, device=dev, generator=gen) * scale for s in shapes]
        local_grads.append([x.cpu() for x in gs])
        flat.zero()
        for p, x in zip(params, gs):
            p.grad.copy_(x)
        if it < 3:
            opt.step()
        else:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                graph.capture_begin(); opt.step(); graph.capture_end()
            torch.cuda.current_stream().wait_stream(side)
            graph.replay()
        torch.cuda.synchronize()
        dist.barrier()
    torch.save({"init": init, "grads": local_grads, "params": [p.detach().cpu() for p in params], "norm": float(opt.total_norm)},
               os.path.join(out_dir, f"dp{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_fused_dp_optimizer_vs_torch(tmp_path, world):
    """DistributedClipAdam (peer-memory reduce-scatter + global-norm clip + Adam + all-gather in two kernels, no NCCL) ==
    average of the ranks' gradients -> clip_grad_norm_ -> torch.optim.Adam on every rank; parameters bit-identical across
    ranks; also through a captured CUDA graph."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run with gpurun --gpus {world})")
    import torch.multiprocessing as mp
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"dp{r}.pt") for r in range(world)]
    ref = [torch.nn.Parameter(p.clone()) for p in got[0]["init"]]
    ropt = torch.optim.Adam(ref, lr=5e-4)
    for it in range(4):
        for i, r in enumerate(ref):
            r.grad = sum(got[q]["grads"][it][i] for q in range(world)) / world
        norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        ropt.step()
    assert abs(got[0]["norm"] - float(norm)) <= 1e-4 * float(norm)      # fp32 sums of 2 M squares in different orders
    for i, r in enumerate(ref):
        torch.testing.assert_close(got[0]["params"][i], r.data, rtol=1e-6, atol=1e-7)
        for q in range(1, world):
            assert torch.equal(got[0]["params"][i], got[q]["params"][i]), "ranks must hold bit-identical parameters"
