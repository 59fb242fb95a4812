for mesh in delaunay polar; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --hand-mesh $mesh > gpurun_out/bench_$mesh.json 2> gpurun_out/bench_a.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$mesh.json")); b=d["breakdown_us"]
print("$mesh", d["value"], d["ms_per_step"], {k:v["us_each"] for k,v in b.items() if "raster" in k or "sdf" in k or "mano" in k})
PY
done
bash scripts/gpu_tests.sh 2>&1 | tail -30
