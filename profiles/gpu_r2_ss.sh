#!/bin/bash
# scheduled sampling: the benchmarked train step (CUDA graph, device-resident) with 0 / 2 / 4 / 9 free decode steps, segmented chain
# schedule against the general launch-per-step schedule
O=gpurun_out/r2ss; mkdir -p $O
for free in "" "7,13" "3,7,11,15" "2,4,6,8,10,12,14,16,18"; do
  for fast in 1 0; do
    echo -n "free=[$free] fast=$fast "
    if [ $fast = 1 ]; then ACVAE_BENCH_SS_FREE="$free" timeout 300 python bench.py --steps 100 --warmup 10 --train-only 2>/dev/null | tail -1 | cut -c1-90
    else ACVAE_DISABLE_FAST=1 ACVAE_BENCH_SS_FREE="$free" timeout 300 python bench.py --steps 100 --warmup 10 --train-only 2>/dev/null | tail -1 | cut -c1-90; fi
  done
done | tee $O/ss_bench.log
