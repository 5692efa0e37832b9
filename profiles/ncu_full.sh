#!/bin/bash
# full-set ncu captures (with source) of a stage-2 stride-1 unit's pointwise kernels: forward (pw1, tail), backward
# (tail dgrad/wgrad, pw1 dgrad/wgrad).  Kept small: gpurun_out/ is limited to 64 MiB.
TAG=${1:-r1}
ncu --set full --import-source on --clock-control none -k regex:"pw_fwd_kernel" -s 76 -c 2 -o gpurun_out/${TAG}_pwfwd -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu_fwd.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pw_dgrad_kernel|pw_wgrad_kernel" -s 174 -c 4 -o gpurun_out/${TAG}_pwbwd -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu_bwd.log 2>&1
for f in pwfwd pwbwd; do
  ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$f.ncu-rep --page source --csv > gpurun_out/${TAG}_${f}_source.csv 2>/dev/null
done
ls -la gpurun_out/
