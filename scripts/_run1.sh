timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_raster_stress_gpu.py -m gpu -q --no-header 2>&1 | tail -5
for lib in homan_b200/libhoman_b200.so homan_b200/_variants/*.so; do
HOMAN_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a.json")); b=d["breakdown_us"]
print("$lib", d["value"], d["ms_per_step"], {k:v["us_each"] for k,v in b.items() if "raster" in k})
PY
done
ncu --set full --clock-control none --import-source on -k regex:raster_bwd -s 4 -c 2 -f -o gpurun_out/prof_bwd_new python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bwd_new.log 2>&1
