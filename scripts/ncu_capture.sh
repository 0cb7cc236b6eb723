#!/bin/bash
# Small ncu captures (a few launches per kernel family) + one launch list; exports CSV summaries so that
# gpurun_out stays far below the 64 MiB copy-back limit.
set -u
OUT=gpurun_out
TAG=${1:-r1}
# CUDA-graph replay of the train body makes ncu (2025.2, driver 580) fail with LaunchFailed: profile the eager launches
export CSI_NO_GRAPH=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 720 -c 230 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_list_$TAG.log 2>&1
METRICS=${NCU_METRICS:-"gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,smsp__cycles_active.avg,lts__t_bytes.sum,l1tex__t_bytes.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum"}
ncu --metrics $METRICS --clock-control none -s 720 -c 230 --csv --log-file $OUT/kernels_$TAG.csv $CMD > $OUT/ncu_k_$TAG.log 2>&1
ls -la $OUT | tail -8
