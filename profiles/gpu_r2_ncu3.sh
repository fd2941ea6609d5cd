#!/bin/bash
# final timeline of the train step + full ncu capture of the persistent statistics GEMM with device-drawn noise (sampling call)
O=gpurun_out/r2ncu3; mkdir -p $O
timeout 300 python profiles/timeline.py train > $O/timeline.csv 2>$O/timeline.err; wc -l $O/timeline.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_gemm_persist_kernel<\(int\)4>' --launch-skip 3 -c 1 -o $O/vocab_persist_full -f python bench.py --profile sample > $O/ncu_vocab.log 2>&1
tail -3 $O/ncu_vocab.log; ls -la $O
