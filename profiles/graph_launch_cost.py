"""CPU cost of one replay of the train-step CUDA graph (cudaGraphLaunch) vs the GPU time of the step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ts = bench.make_train_step(dev, 1, 0, use_graph=True)
for i in range(5):
    ts.load_resident(i); ts.run_step()
torch.cuda.synchronize()
n = 30
t0 = time.perf_counter()
for i in range(n):
    ts.run_step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"graph nodes: n/a; CPU per replay {(t1 - t0) / n * 1e6:.1f} us; wall per step incl. GPU {(t2 - t0) / n * 1e6:.1f} us")
cpu = []
for i in range(10):
    torch.cuda.synchronize()
    a = time.perf_counter(); ts.run_step(); b = time.perf_counter()
    cpu.append((b - a) * 1e6)
print("CPU per replay on an idle GPU (us):", " ".join(f"{x:.0f}" for x in cpu))
