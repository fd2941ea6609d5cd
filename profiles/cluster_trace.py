"""Where a step of the CLUSTER chains spends its time: clock64 stamps of thread 0 of CTA 0 of dec_cl_fwd_kernel and
post_cl_fwd_kernel during one eager train step at CFG1.   python profiles/cluster_trace.py"""
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
buf = torch.zeros(4 * T, 16, dtype=torch.int64, device=dev)     # dec fwd | dec bwd | post fwd | post bwd
_lib.lib().acvae_debug_set_chain_trace(buf.data_ptr())
ts.load_resident(3); ts.step_body()
torch.cuda.synchronize()
_lib.lib().acvae_debug_set_chain_trace(None)
tr = buf.cpu().numpy()


def table(title, base, names):
    print(title)
    print("step  " + "  ".join(f"{n:>14s}" for n in names) + "   total (cycles, thread 0 of CTA 0)")
    for t in range(1, T - 1):
        d = [int(tr[base + t, k + 1] - tr[base + t, k]) for k in range(len(names))]
        print(f"{t:4d}  " + "  ".join(f"{x:14d}" for x in d) + f"   {int(tr[base + t + 1, 0] - tr[base + t, 0])}")


table("decoder forward chain (dec_cl_fwd_kernel)", 0,
      ["arm+gx", "wait A (h)", "q-proj+sync", "scores+send", "h-part GRU", "wait RS", "sum+send+wait AG", "softmax+sync", "ctx part", "cell+send"])
table("\nposterior forward chain (post_cl_fwd_kernel)", 2 * T, ["arm+gx", "wait h", "product+reduce", "cell", "saves+send"])
