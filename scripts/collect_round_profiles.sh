#!/bin/bash
# Collects the numbers profiles/ records for a round on the GPU box (one B200): bench lines, ncu launch list,
# full ncu capture of the heavy kernels, comparator benchmark, pose-init benchmark. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'raster_bwd|raster_fwd|sdf_pair|mano_bwd|sil_loss_prep' -s 16 -c 9 -f \
    -o gpurun_out/prof_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_final.log 2>&1
python scripts/bench_raster_vs_nmr_style.py > gpurun_out/raster_vs_nmr_style.json 2> gpurun_out/raster_vs_nmr_style.err
python scripts/bench_iteration_vs_nmr_style.py --iters 5 --out gpurun_out/iteration_vs_nmr_style.json > /dev/null 2> gpurun_out/iteration_vs_nmr_style.err
python scripts/bench_pose_init.py --out gpurun_out/pose_init.json > /dev/null 2> gpurun_out/pose_init.err
python bench.py --workload cfg5 --steps 20 --no-cpu-baseline > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err
python bench.py --workload cfg2 --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg2_n1.json 2> gpurun_out/bench_cfg2_n1.err
python bench.py --workload cfg4 --steps 30 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err
python bench.py --hand-mesh polar --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg3_polar.json 2> gpurun_out/bench_cfg3_polar.err
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
tail -c 400 gpurun_out/bench_cfg3_n1.json; tail -c 300 gpurun_out/bench_reference.json
