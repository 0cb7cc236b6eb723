#!/bin/bash
# `ncu --set full --import-source on` of ONE launch of each kernel whose name matches a regex in "$@" (eager bench step,
# CSI_NO_GRAPH=1), after a plain run of the same command exited 0.  Exports details + source pages as text (small).
set -u
OUT=gpurun_out
export CSI_NO_GRAPH=1
CMD="python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --no-config4"
$CMD > $OUT/plain_fullk.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_fullk.log; exit 1; }
for RX in "$@"; do
  TAG=$(echo "$RX" | tr -c 'A-Za-z0-9_\n' '_')
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RX" -s 12 -c 1 -f -o $OUT/full_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
  ncu -i $OUT/full_$TAG.ncu-rep --page details > $OUT/full_${TAG}_details.txt 2>&1
  ncu -i $OUT/full_$TAG.ncu-rep --page source --csv > $OUT/full_${TAG}_source.csv 2>&1
  tail -1 $OUT/ncu_full_$TAG.log
done
