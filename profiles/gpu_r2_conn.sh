#!/bin/bash
# does the number of hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS, default 8) limit the forked-stream graph?
O=gpurun_out/r2conn; mkdir -p $O
for c in 8 16 32; do
  for i in 1 2; do echo -n "conn=$c "; CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python bench.py --steps 200 --warmup 20 --train-only 2>/dev/null | tail -1; done
done | tee $O/conn.log
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python profiles/timeline.py train > $O/timeline32.csv 2>$O/timeline.err; wc -l $O/timeline32.csv
