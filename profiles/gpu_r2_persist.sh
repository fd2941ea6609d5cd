#!/bin/bash
# persistent tcgen05 GEMM: GEMM / sampling / train parity, then sampling timings with and without it
O=gpurun_out/${1:-r2p1}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -k "gemm or sampl or sampler or train or vocab" 2>&1 | tail -5
for pz in 1 0; do
  echo "ACVAE_TC_PERSIST=$pz"
  ACVAE_TC_PERSIST=$pz timeout 300 python profiles/sample_small_shard.py 131 2>&1 | tail -1
  ACVAE_TC_PERSIST=$pz timeout 300 python profiles/sample_small_shard.py 1045 2>&1 | tail -1
done | tee $O/sample.log
