O=gpurun_out/r2n; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 3 -c 8 -o $O/tc_gemm_sample_full -f python bench.py --profile sample > $O/ncu_tc.log 2>&1
ls -la $O
