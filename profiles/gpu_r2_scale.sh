#!/bin/bash
# scaling point: the driver's own launch line at N GPUs (fused peer-memory exchange), plus the NCCL exchange for comparison at N=8
N=${1:-8}; O=gpurun_out/r2scale; mkdir -p $O
run() { ACVAE_BENCH_EXCHANGE=$1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_n${N}_$1.json 2> $O/bench_n${N}_$1.err; echo "N=$N $1 rc=$?"; }
run fused
[ "$N" = "8" ] && [ -z "$SKIP_NCCL" ] && run nccl
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_n${N}_*.json")):
    try:
        j = json.loads([x for x in open(f) if x.startswith("{")][-1])
        print(f, j["ms_per_step"], j["value"], "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"], "sampling", j["sampling"]["value"], j["sampling"]["ms"])
    except Exception as e:
        print(f, "no line", e)
PY
