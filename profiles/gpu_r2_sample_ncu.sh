#!/bin/bash
# launch lists of one diverse-sampling call: 1045 clips x 10 (one GPU) and 131 clips x 10 (the 8-GPU shard)
O=gpurun_out/r2s; mkdir -p $O
timeout 300 python profiles/sample_small_shard.py 131 > $O/small_plain.log 2>&1 || exit 1
timeout 300 python profiles/sample_small_shard.py 1045 > $O/full_plain.log 2>&1 || exit 1
cat $O/small_plain.log $O/full_plain.log
PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_sample_131.csv python profiles/sample_small_shard.py 131 > $O/ncu_small.log 2>&1
PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_sample_1045.csv python profiles/sample_small_shard.py 1045 > $O/ncu_full.log 2>&1
python profiles/agg_launches.py $O/launches_sample_131.csv 25
python profiles/agg_launches.py $O/launches_sample_1045.csv 25
