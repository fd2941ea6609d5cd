O=gpurun_out/r1k; mkdir -p $O
./build/ubench_bcast > $O/ubench_bcast2.log 2>&1
for m in 8 4 2 1; do echo "== ACVAE_TC_MIN_KBLK=$m"; ACVAE_TC_MIN_KBLK=$m GEMM_SHAPES=1,2,3,5,6,7,8,9,10 python profiles/gemm_bench.py; done > $O/gemm_split.log 2>&1
cat $O/ubench_bcast2.log $O/gemm_split.log
