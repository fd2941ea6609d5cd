python -m pytest tests -m gpu -x -q -k "train_cfg1 or golden" 2>&1 | tail -1
for i in 1 2 3; do python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['final_loss'])"; done
ACVAE_BENCH_NO_INPUT_EVENT=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('no event', j['ms_per_step'], j['e2e']['ms_per_step'], j['final_loss'])"
