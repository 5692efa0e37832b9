#!/bin/bash
# full-set ncu capture of ONE kernel of the bench command: bash profiles/ncu_one.sh <kernel regex> <tag> [skip] [count]
K=$1; TAG=${2:-one}; S=${3:-6}; C=${4:-2}
ncu --set full --import-source on --clock-control none -k regex:"$K" -s $S -c $C -o gpurun_out/${TAG} -f \
    python bench.py --steps 1 --warmup 1 --T 8 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}.ncu-rep
