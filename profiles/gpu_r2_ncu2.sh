#!/bin/bash
# round-2 final evidence: ncu launch lists (serialised, cold-cache per-launch durations) of one eager train step and one
# sampling call, plus a full capture of the persistent statistics GEMM inside the sampling call.  Never a bench value.
O=gpurun_out/r2ncu2; mkdir -p $O
timeout 300 python bench.py --profile train > $O/plain_train.log 2>&1 || exit 1
timeout 300 python bench.py --profile sample > $O/plain_sample.log 2>&1 || exit 1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python bench.py --profile train > $O/ncu_train.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_sample.csv python bench.py --profile sample > $O/ncu_sample.log 2>&1
python profiles/agg_launches.py $O/launches_train.csv 40 | tee $O/launches_train_summary.txt
python profiles/agg_launches.py $O/launches_sample.csv 20 | tee $O/launches_sample_summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_persist_kernel.*4 --launch-skip 5 -c 1 -o $O/vocab_persist_full -f python bench.py --profile sample > $O/ncu_vocab.log 2>&1
ls -la $O
