for lib in homan_b200/_variants/*.so; do
HOMAN_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_a.json")); b=d["breakdown_us"]
print("$lib", d["value"], d["ms_per_step"], {k:v["us_each"] for k,v in b.items() if "raster" in k})
PY
done
