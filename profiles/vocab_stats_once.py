"""One vocabulary-statistics GEMM at the sampling size (for ncu captures).  python profiles/vocab_stats_once.py [M]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acvae_b200 import functional as F
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10450
hid = torch.randn(M, 256, device="cuda"); cw = torch.randn(4400, 256, device="cuda") * 0.05; cb = torch.randn(4400, device="cuda")
for _ in range(3):
    F.vocab_stats(hid, cw, cb)
torch.cuda.synchronize()
print("done")
