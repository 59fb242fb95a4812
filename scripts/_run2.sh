python bench.py --workload cfg4 --steps 30 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; tail -3 gpurun_out/bench_cfg4_n1.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg4_n1.json")); b=d["breakdown_us"]
print("cfg4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], {k:v["us_each"] for k,v in b.items() if "raster" in k or "sdf" in k})
PY
python bench.py --steps 100 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err; tail -3 gpurun_out/bench_cfg3_n1.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg3_n1.json")); b=d["breakdown_us"]
print("cfg3", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"], {k:v["us_each"] for k,v in b.items() if "raster" in k or "sdf" in k or "mano" in k})
PY
