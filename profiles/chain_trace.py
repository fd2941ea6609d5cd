"""Where a step of the persistent decoder forward chain spends its time: clock64 stamps of thread 0 of CTA 0 (a clip
CTA) of dec_chain_fwd_kernel during one eager train step at CFG1.   python profiles/chain_trace.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from acvae_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ts = bench.make_train_step(dev, 1, 0, use_graph=False)
for i in range(3):
    ts.load_resident(i); ts.step_body()
torch.cuda.synchronize()
T = ts.st_prep.T
buf = torch.zeros(2 * T, 16, dtype=torch.int64, device=dev)     # forward chain | backward chain
_lib.lib().acvae_debug_set_chain_trace(buf.data_ptr())
ts.load_resident(3); ts.step_body()
torch.cuda.synchronize()
_lib.lib().acvae_debug_set_chain_trace(None)
tr = buf.cpu().numpy()
names = ["partials summed", "attention", "prior LSTM", "barrier A", "GRU product", "GRU cell + prior head", "partial stores", "barrier G"]
print("step  " + "  ".join(f"{n:>18s}" for n in names) + "   total (cycles, thread 0 of CTA 0)")
for t in range(1, T - 1):
    d = [int(tr[t, k + 1] - tr[t, k]) for k in range(8)]
    print(f"{t:4d}  " + "  ".join(f"{x:18d}" for x in d) + f"   {int(tr[t, 8] - tr[t, 0])}"
          f"   [attention: scores {int(tr[t, 9] - tr[t, 1])}, softmax {int(tr[t, 10] - tr[t, 9])}, context {int(tr[t, 2] - tr[t, 10])}]")

names = ["GRU bwd (dh product + cell)", "prior LSTM bwd", "barrier", "d ctx product", "prior head bwd", "barrier", "attention bwd", "barrier"]
print("\nbackward chain (dec_chain_bwd_kernel)")
print("step  " + "  ".join(f"{n:>18s}" for n in names) + "   total")
for t in range(T - 2, 0, -1):
    d = [int(tr[T + t, k + 1] - tr[T + t, k]) for k in range(8)]
    print(f"{t:4d}  " + "  ".join(f"{x:18d}" for x in d) + f"   {int(tr[T + t, 8] - tr[T + t, 0])}")
