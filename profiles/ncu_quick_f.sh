#!/bin/bash
# per-launch counters of one forward pass for kernels matching $2 (regex)
OUT=${1:-gpurun_out/quick.csv}; KRE=${2:-pw_fwd}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread \
    --clock-control none -k regex:"$KRE" -s 66 -c 33 --csv --log-file $OUT \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > /dev/null 2>&1
