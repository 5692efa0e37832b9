#!/bin/bash
# full-set ncu capture (with source) of the stage-2 stride-1 depthwise kernels: forward <120,1> and backward <120,1>
TAG=${1:-r2}
ncu --set full --import-source on --clock-control none -k regex:"dw_fwd_kernel" -s 9 -c 1 -o gpurun_out/${TAG}_dwf -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu_dw.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"dw_bwd_kernel" -s 7 -c 1 -o gpurun_out/${TAG}_dwb -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile >> gpurun_out/${TAG}_ncu_dw.log 2>&1
for f in dwf dwb; do
ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_$f.ncu-rep --page source --csv > gpurun_out/${TAG}_${f}_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_$f.ncu-rep
done
