set -x
mkdir -p gpurun_out/r1a
python -m pytest tests -m gpu -x -q > gpurun_out/r1a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1a/pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r1a/bench.json 2> gpurun_out/r1a/bench.err; echo "bench rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1a/launches_train.csv python bench.py --profile train > gpurun_out/r1a/ncu_train.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1a/launches_sample.csv python bench.py --profile sample > gpurun_out/r1a/ncu_sample.log 2>&1
tail -3 gpurun_out/r1a/pytest.log; cat gpurun_out/r1a/bench.json | cut -c1-600
