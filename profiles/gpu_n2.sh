O=gpurun_out/r1z; mkdir -p $O
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?"
cut -c1-700 $O/bench_n$N.json; tail -3 $O/bench_n$N.err
