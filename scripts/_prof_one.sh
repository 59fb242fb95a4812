#!/bin/bash
# Development: full ncu capture of the kernels matching $1 (regex) in one cfg3 step -> gpurun_out/prof_one.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-4} -c ${3:-2} -f \
    -o gpurun_out/prof_one python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_one.log 2>&1
tail -2 gpurun_out/prof_one.log | cut -c1-200
