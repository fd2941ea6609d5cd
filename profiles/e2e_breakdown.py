"""Where the end-to-end step's extra ~0.1 ms over the device-resident step goes (host side)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ts = bench.make_train_step(dev, 1, 0, use_graph=True)
p = ts.pinned[0]
def feed(i):
    p = ts.pinned[i % 4]
    ts.copy_stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ts.copy_stream):
        ts.st_audio.copy_(p["audio"], non_blocking=True); ts.audio_ready.record()
    ts.st_mem_lens.copy_(p["mem_lens"], non_blocking=True)
    ts.model.prepare_batch(p["caps"], p["cap_lens"], dev, out=ts.st_prep)
for i in range(5):
    feed(i); ts.run_step()
torch.cuda.synchronize()
t_feed, t_launch, t_wait, t_total = [], [], [], []
for i in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); feed(i); t1 = time.perf_counter(); ts.run_step(); t2 = time.perf_counter()
    x = float(ts.loss_buf); t3 = time.perf_counter()
    t_feed.append(t1 - t0); t_launch.append(t2 - t1); t_wait.append(t3 - t2); t_total.append(t3 - t0)
us = lambda v: f"{np.median(v) * 1e6:.0f}"
print(f"feed_host CPU {us(t_feed)} us | graph launch CPU {us(t_launch)} us | wait for loss {us(t_wait)} us | total {us(t_total)} us")
# the copy alone
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record(); ts.st_audio.copy_(p["audio"], non_blocking=True); b.record(); torch.cuda.synchronize()
print(f"audio H2D alone: {a.elapsed_time(b) * 1e3:.0f} us for {p['audio'].numel() * 4 / 1e6:.2f} MB")
