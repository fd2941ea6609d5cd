import sys, os; sys.path.insert(0,"tests"); sys.path.insert(0,"oracle")
import harness, torch
from harness import synthetic
from acvae_b200 import functional as F
for (M,N,K) in [(608,768,256),(1984,256,512),(608,256,4400)]:
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M,K,generator=g).cuda(); B = torch.randn(N,K,generator=g).cuda()
    C, used = F.gemm(A,B)
    ref = A.double() @ B.double().t()
    print("gemm", M,N,K, "used_tc", used, "rel_err %.3e" % harness.rel_err(C, ref), "torch fp32 %.3e" % harness.rel_err((A@B.t()), ref))
d = synthetic.CFG1
r = harness.run_cuda_train(d, 11); o = harness.run_oracle_train(d, 11)
o64 = harness.run_oracle_train(d, 11, dtype=torch.float64)
worst = 0
for k, ref in o64["grads"].items():
    e = harness.rel_err(r["grads"][k], ref); e32 = harness.rel_err(o["grads"][k], ref); eo = harness.rel_err(r["grads"][k], o["grads"][k])
    if e > 2e-5 or eo > 2e-5: print(f"ours-vs-fp64 {e:.2e}  oracle32-vs-fp64 {e32:.2e}  ours-vs-oracle32 {eo:.2e}  {k}")
    worst = max(worst, eo)
print("worst ours-vs-oracle32", worst)
