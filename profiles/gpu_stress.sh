O=gpurun_out/r2e; mkdir -p $O
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
ACVAE_BENCH_CONFIG=stress timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_stress.json 2> $O/bench_stress.err; echo "stress rc=$?"
cut -c1-300 $O/bench.json; cut -c1-300 $O/bench_stress.json; tail -3 $O/bench_stress.err
