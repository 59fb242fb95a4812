timeout 900 python -m pytest tests/test_raster_gpu.py tests/test_raster_stress_gpu.py tests/test_fullsize_gpu.py -m gpu -q --no-header -x 2>&1 | tail -15
for v in new old; do
if [ $v = old ]; then export HOMAN_B200_BWD_OLD=1; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a.json")); b=d["breakdown_us"]
print("$v", d["value"], d["ms_per_step"], {k:v["us_each"] for k,v in b.items() if "raster" in k})
PY
done
