for dbg in ${SWEEP:-0 1 8 9 15}; do echo "== ACVAE_TC_DEBUG=$dbg"; GEMM_SHAPES=${GEMM_SHAPES:-11,5,2} ACVAE_TC_DEBUG=$dbg python profiles/gemm_bench.py 2>&1 | tail -4; done
