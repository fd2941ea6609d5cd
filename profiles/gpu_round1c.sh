set -x
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python profiles/timeline.py train > $O/timeline_train.csv 2> $O/timeline.err
tail -3 $O/pytest.log; cut -c1-400 $O/bench.json; tail -2 $O/bench.err
