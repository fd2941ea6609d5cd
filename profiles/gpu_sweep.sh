O=gpurun_out/r2j; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "train" 2>&1 | tail -1
for k in 6 24; do
  echo "== ACVAE_FAN_MIN_KBLK=$k"; ACVAE_FAN_MIN_KBLK=$k python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['e2e']['ms_per_step'])"
done > $O/sweep.log 2>&1
cat $O/sweep.log
