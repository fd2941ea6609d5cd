"""Micro-benchmark of acvae_gemm on the shapes of one train step (warm L2, CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acvae_b200 import functional as F
shapes = [  # (M, N, K, a_trans, b_trans, what)
    (1984, 256, 512, 0, 0, "ln: mem = audio.W^T"),
    (1984, 256, 256, 0, 0, "P = mem.Wm^T"),
    (608, 768, 256, 0, 0, "gx_q"),
    (608, 1024, 512, 0, 0, "gx_p (2 seg in the step)"),
    (608, 4400, 256, 0, 0, "logits / vocab stats"),
    (608, 256, 4400, 0, 1, "dH = dlogits.W"),
    (4400, 256, 608, 1, 1, "dW_cls = dlogits^T.H"),
    (768, 256, 608, 1, 1, "dW_ih block"),
    (1024, 256, 608, 1, 1, "dW prior block"),
    (608, 256, 768, 0, 1, "dX = dG.W"),
    (256, 512, 1984, 1, 1, "dW_ln"),
    (10450, 1024, 1024, 0, 0, "sampling LSTM gates (M=10450)"),
    (10450, 4400, 256, 0, 0, "sampling vocab (M=10450)"),
]
sel = os.environ.get('GEMM_SHAPES')
if sel:
    shapes = [shapes[int(i)] for i in sel.split(',')]
reps = int(os.environ.get('GEMM_REPS', '20'))
for (M, N, K, at, bt, what) in shapes:
    A = torch.randn((K, M) if at else (M, K), device="cuda")
    B = torch.randn((K, N) if bt else (N, K), device="cuda")
    C = torch.empty(M, N, device="cuda")
    for _ in range(3):
        _, used = F.gemm(A, B, bool(at), bool(bt), out=C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = reps
    # back-to-back launches replayed from a CUDA graph: the ctypes call (~20 us of CPU) must not be what is timed
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=side):
        for _ in range(n):
            F.gemm(A, B, bool(at), bool(bt), out=C)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{what:34s} M={M:5d} N={N:5d} K={K:5d} tc={int(used)} {us:8.1f} us  {2*M*N*K/us/1e6:8.2f} TFLOP/s")

# the vocabulary GEMM with its statistics epilogue (EPI_STATS: logits never reach HBM)
for M in (608, 1310, 10450):
    hid = torch.randn(M, 256, device="cuda"); cw = torch.randn(4400, 256, device="cuda") * 0.05; cb = torch.randn(4400, device="cuda")
    for _ in range(3):
        F.vocab_stats(hid, cw, cb)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            F.vocab_stats(hid, cw, cb)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{'vocab_stats (GEMM + reduce)':34s} M={M:5d} N= 4400 K=  256      {us:8.1f} us  {2*M*4400*256/us/1e6:8.2f} TFLOP/s")
