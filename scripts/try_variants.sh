#!/bin/bash
# Development helper: benchmarks every library variant under homan_b200/_variants/ (built with different -D knobs).
cp homan_b200/libhoman_b200.so /tmp/lib_orig.so
for f in homan_b200/_variants/*.so; do
    cp $f homan_b200/libhoman_b200.so
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/var_$(basename $f .so).json 2>/dev/null
    python - <<PY
import json
d=json.load(open("gpurun_out/var_$(basename $f .so).json"))
b=d["breakdown_us"]
print("$(basename $f .so)", round(d["ms_per_step"],3), "fwd", b["hm_raster_sil_fwd"]["us_each"], "bwd", b["hm_raster_sil_bwd"]["us_each"])
PY
done
cp /tmp/lib_orig.so homan_b200/libhoman_b200.so
