#!/bin/bash
# full ncu captures of the two prior chain kernels (now the longest kernels of the step) and of the persistent gate GEMM of a
# sampling step; raw pages exported on the box
O=gpurun_out/r2ncu4; mkdir -p $O
for k in prior_chain_fwd_kernel prior_chain_bwd_kernel; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -o $O/${k}_full -f python bench.py --profile train > $O/ncu_$k.log 2>&1
  ncu -i $O/${k}_full.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_gemm_persist_kernel<\(int\)0>' --launch-skip 8 -c 1 -o $O/gate_persist_full -f python bench.py --profile sample > $O/ncu_gate.log 2>&1
ncu -i $O/gate_persist_full.ncu-rep --page raw --csv > $O/gate_persist_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O
