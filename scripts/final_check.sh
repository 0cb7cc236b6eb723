python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest16.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err
export CSI_NO_GRAPH=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-config4"
ncu --metrics gpu__time_duration.sum --clock-control none -s 660 -c 200 --csv --log-file gpurun_out/launches_r2i.csv $CMD > gpurun_out/ncu_list_r2i.log 2>&1
