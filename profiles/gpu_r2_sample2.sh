#!/bin/bash
# sampling with device-drawn noise: GPU sampling tests, then timings at the two shard sizes
O=gpurun_out/${1:-r2s2}; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -k "sampl or sampler" 2>&1 | tail -3
timeout 300 python profiles/sample_small_shard.py 131 2>&1 | tail -1 | tee $O/small.log
timeout 300 python profiles/sample_small_shard.py 1045 2>&1 | tail -1 | tee $O/full.log
