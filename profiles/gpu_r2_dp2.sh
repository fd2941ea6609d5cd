#!/bin/bash
# 2-GPU check of the fused data-parallel optimizer: parity test, then bench with each exchange mode.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu > gpurun_out/dp2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/dp2_pytest.log
tail -5 gpurun_out/dp2_pytest.log
for mode in fused nccl nccl-bucketed; do
  ACVAE_BENCH_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/dp2_bench_$mode.json 2> gpurun_out/dp2_bench_$mode.err
  echo "$mode rc=$?"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/dp2_bench_$mode.json") if x.startswith("{")][-1]; j=json.loads(l)
    print("$mode", j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["sampling"]["value"])
except Exception as e: print("no line", e)
PY
done
