# Round-1 measurement pass: tests, bench (ours + reference arm), ncu launch lists, one full ncu capture of the dominant kernel.
set -x
O=gpurun_out/r1f; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python bench.py --profile train > $O/ncu_train.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_sample.csv python bench.py --profile sample > $O/ncu_sample.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dec_chain_fwd -c 1 -o $O/dec_chain_fwd_full -f python bench.py --profile train > $O/ncu_full.log 2>&1
tail -3 $O/pytest.log; cut -c1-400 $O/bench.json; cat $O/bench_reference.json | cut -c1-300; tail -2 $O/bench.err
