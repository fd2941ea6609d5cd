#!/bin/bash
# quick A/B of the train step: GPU train tests, 3 short train-only bench runs, one timeline
O=gpurun_out/${1:-r2ab}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "train or smoke or step" 2>&1 | tail -2
for i in 1 2 3; do timeout 300 python bench.py --steps 200 --warmup 20 --train-only 2>/dev/null | tail -1; done | tee $O/train_only.log
timeout 300 python profiles/timeline.py train > $O/timeline.csv 2>$O/timeline.err; wc -l $O/timeline.csv
