#!/bin/bash
# per-launch counters of the kernels matching a regex during the second PPO pass:  ncu_quick_k.sh <regex> <out.csv> [count]
RX=${1:-pwg_}
OUT=${2:-gpurun_out/quick_k.csv}
CNT=${3:-60}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
    --clock-control none -k regex:"$RX" -s $CNT -c $CNT --csv --log-file $OUT \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > /dev/null 2>&1
