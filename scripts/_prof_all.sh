#!/bin/bash
# Development: full ncu captures of the heavy kernels of one cfg3 step (raster fwd / bwd, sdf_pair, mano_bwd).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'raster_bwd|raster_fwd|sdf_pair|mano_bwd' -s 14 -c 7 -f \
    -o gpurun_out/prof_r02_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_r02_a.log 2>&1
tail -3 gpurun_out/prof_r02_a.log | cut -c1-300
