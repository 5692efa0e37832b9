#!/bin/bash
# full-set ncu capture (with source) of the fused pointwise backward kernel: skip the first pass, take 3 launches
TAG=${1:-r2}
ncu --set full --import-source on --clock-control none -k regex:"pw_bwd_fused" -s 30 -c 3 -o gpurun_out/${TAG}_fused -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu_fused.log 2>&1
ncu -i gpurun_out/${TAG}_fused.ncu-rep --page raw --csv > gpurun_out/${TAG}_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_fused.ncu-rep --page source --csv > gpurun_out/${TAG}_fused_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_fused.ncu-rep
ls -la gpurun_out/ | tail -5
