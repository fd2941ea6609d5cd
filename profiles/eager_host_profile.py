"""Where the HOST time of an eager train step goes (cProfile over 30 steps, CFG1, all teacher-forced).  python profiles/eager_host_profile.py"""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness
from acvae_b200 import synthetic, parallel
import acvae_b200 as models
d = synthetic.CFG1
m = harness.build_model(d, 1).train()
flat = parallel.FlatGradBuffer(m.parameters()); m.grad_sink = flat
opt = models.FusedClipAdam(flat, lr=5e-4, max_grad_norm=1.0)
b = synthetic.make_batch(d, 17)
T = int(b["cap_lens"].max()) - 1
feats = torch.from_numpy(b["audio_embeds"]).cuda(); lens = torch.from_numpy(b["mem_lens"].copy())
caps = torch.from_numpy(b["caps"]); cap_lens = b["cap_lens"].copy()
lens1 = torch.as_tensor(cap_lens) - 1
targets = torch.nn.utils.rnn.pack_padded_sequence(caps[:, 1:], lens1, batch_first=True).data
fl = models.FusedVAELoss(d.V, smoothing=0.1, alpha=1.0)
def step():
    flat.zero()
    out = m(feats, lens, caps, cap_lens, ss_ratio=1.0, dis_ratio=0.0)
    packed = torch.nn.utils.rnn.pack_padded_sequence(out["logits"], lens1, batch_first=True).data
    loss = fl(out, packed, targets, 0.5); loss.backward(); opt.step()
    return loss
for _ in range(5): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): step()
torch.cuda.synchronize(); print(f"eager step: {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(30): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
