#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'raster_bwd' -s 4 -c 2 -f \
    -o gpurun_out/prof_bwd2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bwd2.log 2>&1
tail -2 gpurun_out/prof_bwd2.log | cut -c1-200
