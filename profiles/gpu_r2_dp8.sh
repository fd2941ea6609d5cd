#!/bin/bash
# 8-GPU check: fused data-parallel optimizer parity at world 8, then the bench with the fused and the NCCL exchange.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu -k fused_dp > gpurun_out/dp8_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/dp8_pytest.log
tail -3 gpurun_out/dp8_pytest.log
for mode in fused nccl; do
  ACVAE_BENCH_EXCHANGE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/dp8_bench_$mode.json 2> gpurun_out/dp8_bench_$mode.err
  echo "$mode rc=$?"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/dp8_bench_$mode.json") if x.startswith("{")][-1]; j=json.loads(l)
    print("$mode", j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["sampling"]["value"], j["sampling"]["ms"])
except Exception as e: print("no line", e)
PY
done
