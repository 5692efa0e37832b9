#!/bin/bash
# full-set ncu capture (with source) of the tcgen05 GEMM family: one forward and one data-gradient launch of a stage-3 unit layer
TAG=${1:-r2}
ncu --set full --import-source on --clock-control none -k regex:"pwg_fwd_kernel" -s 24 -c 1 -o gpurun_out/${TAG}_pwgf -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu_pwg.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pwg_dgrad_kernel" -s 22 -c 1 -o gpurun_out/${TAG}_pwgd -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile >> gpurun_out/${TAG}_ncu_pwg.log 2>&1
for f in pwgf pwgd; do
ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_$f.ncu-rep --page source --csv > gpurun_out/${TAG}_${f}_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_$f.ncu-rep
done
