#!/bin/bash
# One `ncu --set full` capture of the dominant kernel (NT GEMM on the k=5 conv shape of the left stream at B=256, F=270)
# after a plain run of the same command exited 0; exports the details page and the raw CSV (small, tracked under profiles/).
set -u
OUT=gpurun_out; TAG=${1:-r2}
CMD="python scripts/one_gemm.py 39424 270 272 5 3"
$CMD > $OUT/plain_full_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_full_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 2 -c 1 -f -o $OUT/${TAG}_gemm_nt_k5 $CMD > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/${TAG}_gemm_nt_k5.ncu-rep --page details > $OUT/${TAG}_gemm_nt_conv_k5_ncu_details.txt 2>&1
ncu -i $OUT/${TAG}_gemm_nt_k5.ncu-rep --page raw --csv > $OUT/${TAG}_gemm_nt_conv_k5_ncu_raw.csv 2>&1
tail -3 $OUT/ncu_full_$TAG.log
