"""Diverse sampling at the 8-GPU shard size (131 clips x 10 captions = 1310 sequences on ONE GPU): eager loop vs
GraphSampler.   python profiles/sample_small_shard.py [clips]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness
from acvae_b200 import synthetic, GraphSampler, functional as F
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 131
d = synthetic.Dims(N=clips, Te=62, L=21)
m = harness.build_model(synthetic.CFG1, 1).eval()
b = synthetic.make_batch(d, 7)
a = torch.from_numpy(b["audio_embeds"]).cuda(); l = torch.from_numpy(b["mem_lens"].astype(np.int32)).cuda()
def eager():
    with torch.no_grad():
        return m.inference_forward({"audio_embeds": a, "audio_embeds_lens": l}, method="sample", max_length=20, n_captions=10)
for _ in range(3): eager()
torch.cuda.synchronize()
if os.environ.get("PROFILE"):          # one eager call between cudaProfilerStart/Stop for `ncu --profile-from-start off`
    torch.cuda.profiler.start(); eager(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    print("profiled one eager call"); sys.exit(0)
l0 = F.launch_count(); t0 = time.perf_counter()
for _ in range(10): o = eager()
torch.cuda.synchronize(); te = (time.perf_counter() - t0) / 10
gs = GraphSampler(m, clips=clips, Te=62, n_captions=10, max_length=20, method="sample")
for _ in range(3): gs(a, l)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): gs(a, l)
torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / 10
print(f"clips={clips} sequences={clips*10}: eager {te*1e3:.3f} ms ({(F.launch_count()-l0)//10} launches, n_steps {int(o['n_steps'])}), "
      f"graph {tg*1e3:.3f} ms -> {clips*10/tg:.0f} captions/s per GPU")
