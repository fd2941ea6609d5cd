O=gpurun_out/r1u; mkdir -p $O
for m in 0 1; do for k in 2 4 6 10; do
  echo "== ACVAE_MERGE_BWD=$m ACVAE_FAN_MIN_KBLK=$k"; ACVAE_MERGE_BWD=$m ACVAE_FAN_MIN_KBLK=$k python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['e2e']['ms_per_step'])"
done; done > $O/sweep.log 2>&1
cat $O/sweep.log
