O=gpurun_out/r2p; mkdir -p $O
ACVAE_BENCH_CONFIG=stress timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_stress.json 2>/dev/null; cut -c1-200 $O/bench_stress.json
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd_multi -c 1 -o $O/attn_multi_after -f python bench.py --profile sample > $O/ncu_attn.log 2>&1
ls $O
