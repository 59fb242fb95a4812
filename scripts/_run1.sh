#!/bin/bash
# Development: A/B bench of homan_b200/_variants/*.so (and optionally a full ncu capture of $PROF)
mkdir -p gpurun_out
ab() {
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.json")); b=d["breakdown_us"]
print("$1", round(d["value"],1), round(d["ms_per_step"],3), {k:v["us_each"] for k,v in b.items() if "raster" in k or "prep" in k or "sdf" in k})
PY
}
cp homan_b200/libhoman_b200.so /tmp/lib_new.so
for f in homan_b200/_variants/*.so; do
  cp $f homan_b200/libhoman_b200.so
  ab $(basename $f .so)
done
cp /tmp/lib_new.so homan_b200/libhoman_b200.so
if [ -n "$PROF" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$PROF" -s ${PROF_SKIP:-4} -c ${PROF_N:-2} -f \
    -o gpurun_out/prof_one python bench.py --steps 1 --warmup 3 --no-cpu-baseline $PROF_ARGS > gpurun_out/prof_one.log 2>&1
tail -2 gpurun_out/prof_one.log | cut -c1-200
fi
