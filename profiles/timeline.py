"""Concurrent kernel timeline of ONE replay of the benchmarked train step (CUDA graph, multi-stream),
collected with torch.profiler (CUPTI activity records): name, stream, start, duration.  ncu serialises
kernels, so the critical path of the forked-stream schedule can only be read from this.

    python profiles/timeline.py [train|sample] > gpurun_out/timeline.csv
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(mode):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ts = bench.make_train_step(dev, 1, 0, use_graph=(os.environ.get("NO_GRAPH") is None))
    if mode == "train":
        for i in range(5):
            ts.load_resident(i); ts.run_step()
        torch.cuda.synchronize()

        def body():
            ts.load_resident(7); ts.run_step()
    else:
        import numpy as np
        from acvae_b200 import synthetic
        d = ts.d
        ds = synthetic.Dims(N=bench.SAMPLE_CLIPS, Te=d.Te, L=bench.SAMPLE_LEN + 1)
        sb = synthetic.make_batch(ds, 7)
        ts.model.eval()
        a = torch.from_numpy(sb["audio_embeds"]).to(dev)
        l = torch.from_numpy(sb["mem_lens"].astype(np.int32)).to(dev)

        def body():
            with torch.no_grad():
                ts.model.inference_forward({"audio_embeds": a, "audio_embeds_lens": l}, method="sample",
                                           max_length=bench.SAMPLE_LEN, n_captions=bench.SAMPLE_K)
        body(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        body()
        torch.cuda.synchronize()
    import json
    import tempfile
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    rows = []
    for e in json.load(open(path))["traceEvents"]:
        if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
            rows.append((float(e["ts"]), float(e["dur"]), e.get("args", {}).get("stream", -1), e["name"]))
    rows.sort()
    t0 = rows[0][0] if rows else 0
    print("start_us,dur_us,stream,name")
    for s, d_, st, n in rows:
        print(f"{s - t0:.2f},{d_:.2f},{st},\"{n[:100]}\"")
    end = max(s + d_ for s, d_, _, _ in rows) - t0
    busy = sum(d_ for _, d_, _, _ in rows)
    sys.stderr.write(f"kernels={len(rows)} span={end:.1f}us sum_of_durations={busy:.1f}us\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "train")
