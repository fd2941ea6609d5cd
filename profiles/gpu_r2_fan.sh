#!/bin/bash
# k-blocks per CTA of the weight-gradient GEMMs (split-K depth) in the present schedule
O=gpurun_out/r2fan; mkdir -p $O
for k in 6 12 24 64; do
  for i in 1 2; do echo -n "fan_min_kblk=$k "; ACVAE_FAN_MIN_KBLK=$k timeout 300 python bench.py --steps 200 --warmup 20 --train-only 2>/dev/null | tail -1 | cut -c1-100; done
done | tee $O/fan.log
ACVAE_FAN_MIN_KBLK=64 timeout 300 python profiles/timeline.py train > $O/timeline64.csv 2>$O/timeline.err; wc -l $O/timeline64.csv
