"""Eager train step (forward + loss + backward, CFG1) with scheduled sampling: the segmented chain schedule against the general
launch-per-step schedule (ACVAE_DISABLE_FAST=1).   python profiles/ss_step_time.py [ss_ratio]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness
from acvae_b200 import synthetic
import acvae_b200 as models
ss = float(sys.argv[1]) if len(sys.argv) > 1 else 0.8
d = synthetic.CFG1
m = harness.build_model(d, 1).train()
b = synthetic.make_batch(d, 17)
T = int(b["cap_lens"].max()) - 1
tf, dis = harness.flags_for(b, T, ss, 0.0)
feats = torch.from_numpy(b["audio_embeds"]).cuda(); lens = torch.from_numpy(b["mem_lens"].copy())
caps = torch.from_numpy(b["caps"]); cap_lens = b["cap_lens"].copy()
lens1 = torch.as_tensor(cap_lens) - 1
targets = torch.nn.utils.rnn.pack_padded_sequence(caps[:, 1:], lens1, batch_first=True).data
fl = models.FusedVAELoss(d.V, smoothing=0.1, alpha=1.0)
eq = torch.from_numpy(b["eps_q"][:, :T].copy()); ep = torch.from_numpy(b["eps_p"][:T].copy())
def step():
    m.zero_grad(set_to_none=True)
    out = m(feats, lens, caps, cap_lens, ss_ratio=ss, dis_ratio=0.0, eps_q=eq, eps_p=ep, tf_flags=tf, dis_flags=dis)
    packed = torch.nn.utils.rnn.pack_padded_sequence(out["logits"], lens1, batch_first=True).data
    loss = fl(out, packed, targets, 0.5); loss.backward()
    return loss
for _ in range(5): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): l = step()
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 30 * 1e3
# the same step replayed as ONE CUDA graph (the flag pattern is baked into the capture): device time without launch overhead
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    step(); torch.cuda.synchronize()
    g.capture_begin(); step(); g.capture_end()
torch.cuda.current_stream().wait_stream(side)
for _ in range(5): g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): g.replay()
e1.record(); torch.cuda.synchronize()
gms = e0.elapsed_time(e1) / 50
print(f"ss_ratio={ss} free steps={sum(1 for t in range(1, T) if not tf[t])}/{T} fast={'0' if os.environ.get('ACVAE_DISABLE_FAST') else '1'}: "
      f"{ms:.3f} ms per eager step (fwd + loss + bwd), {gms:.3f} ms as a CUDA graph, loss {float(l):.5f}")
