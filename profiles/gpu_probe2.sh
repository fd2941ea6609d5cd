O=gpurun_out/r1w; mkdir -p $O
./build/ubench_gridbar > $O/ubench_gridbar.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd_multi -c 1 -o $O/attn_multi_full -f python bench.py --profile sample > $O/ncu_attn.log 2>&1
cat $O/ubench_gridbar.log
