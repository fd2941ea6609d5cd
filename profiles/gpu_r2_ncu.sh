# round 2: ncu launch list of one eager train step + full captures of the cluster chain kernels (one GPU; never a bench value)
O=gpurun_out/${1:-r2n}; mkdir -p $O
timeout 300 python bench.py --profile train > $O/plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python bench.py --profile train > $O/ncu_train.log 2>&1
for k in dec_cl_fwd dec_cl_bwd post_cl_fwd post_cl_bwd; do
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -o $O/${k}_full -f python bench.py --profile train > $O/ncu_$k.log 2>&1
done
ls -la $O
