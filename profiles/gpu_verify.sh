O=gpurun_out/r2m; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2> /dev/null
./build/ubench_chain > $O/ubench_chain.log 2>&1; head -8 $O/ubench_chain.log
