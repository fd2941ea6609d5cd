# Round-1 final measurement pass: tests, bench (ours + reference arm), ncu launch lists, full ncu captures of the two
# dominant kernels, concurrent timeline.  Usage: bash profiles/gpu_round1_final.sh <tag>
set -x
O=gpurun_out/${1:-r1final}; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python profiles/timeline.py train > $O/timeline_train.csv 2> $O/timeline.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python bench.py --profile train > $O/ncu_train.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_sample.csv python bench.py --profile sample > $O/ncu_sample.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dec_chain_fwd -c 1 -o $O/dec_chain_fwd_full -f python bench.py --profile train > $O/ncu_full_fwd.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dec_chain_bwd -c 1 -o $O/dec_chain_bwd_full -f python bench.py --profile train > $O/ncu_full_bwd.log 2>&1
tail -3 $O/pytest.log; cut -c1-400 $O/bench.json; cut -c1-300 $O/bench_reference.json; tail -2 $O/bench.err
