#!/bin/bash
# Round-2 evidence: (1) launch list of the bench command (per-launch durations, serialised / cold cache: shares must agree with
# the in-bench CUDA-event profile, not absolutes); (2) one full-set capture each of the dominant kernel (pw_bwd_fused) and of the
# tcgen05 GEMM family (pwg_fwd / pwg_dgrad / pwg_wgrad): DRAM bytes, tensor-pipe and DRAM throughput per launch.
TAG=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,lts__t_sector_hit_rate.pct
ncu --metrics $M --clock-control none -k regex:"pw_bwd_fused|pwg_|pw_fwd_tc|pw_wgrad_tc" -s 130 -c 130 --csv --log-file gpurun_out/${TAG}_tc_counters.csv \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > /dev/null 2>&1
ls -la gpurun_out | tail -4
