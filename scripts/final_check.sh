#!/bin/bash
# End-of-round records on one B200: GPU test-suite, smoke, the bench line, the reference arm, ncu launch list + per-kernel
# metrics of one eager step, and one `ncu --set full` capture of the dominant kernel (the fused conv trio GEMM).
set -u
OUT=gpurun_out; TAG=${1:-r2k}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $OUT/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
bash scripts/ncu_capture.sh $TAG
CMD="python scripts/one_gemm_banded.py"
$CMD > $OUT/${TAG}_banded_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 1 -c 1 -f -o $OUT/${TAG}_gemm_nt_banded $CMD > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/${TAG}_gemm_nt_banded.ncu-rep --page details > $OUT/${TAG}_gemm_nt_banded_conv_ncu_details.txt 2>&1
ncu -i $OUT/${TAG}_gemm_nt_banded.ncu-rep --page raw --csv > $OUT/${TAG}_gemm_nt_banded_conv_ncu_raw.csv 2>&1
