#!/bin/bash
# full GPU test suite + the two quick timings (train step, sampling at both shard sizes)
O=gpurun_out/${1:-r2chk}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --steps 200 --warmup 20 --train-only 2>/dev/null | tail -1 | cut -c1-100; done | tee $O/train_only.log
timeout 300 python profiles/sample_small_shard.py 131 2>&1 | tail -1 | tee $O/small.log
timeout 300 python profiles/sample_small_shard.py 1045 2>&1 | tail -1 | tee $O/full.log
