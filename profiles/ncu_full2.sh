#!/bin/bash
# full-set ncu captures (with SASS source counters) of a stage-2 unit's backward kernels in the current build
TAG=${1:-r1b}
ncu --set full --import-source on --clock-control none -k regex:"pw_wgrad_tc_kernel|pw_dgrad_kernel" -s 174 -c 4 -o gpurun_out/${TAG}_bwd -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_bwd.ncu-rep --page raw --csv > gpurun_out/${TAG}_bwd_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_bwd.ncu-rep --page source --csv > gpurun_out/${TAG}_bwd_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_bwd.ncu-rep
python profiles/ncu_source_hot.py gpurun_out/${TAG}_bwd_source.csv 14
