#!/bin/bash
# Development: engine-level A/B through environment knobs
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_properties(0).name)"
ab() {
name=$1; shift
env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print("$name", round(d["value"],1), round(d["ms_per_step"],3))
PY
}
ab default A=1
ab main2_hand1 HOMAN_B200_MAIN_PRIORITY=-2 HOMAN_B200_HAND_PRIORITY=-1
ab main1_hand2 HOMAN_B200_MAIN_PRIORITY=-1 HOMAN_B200_HAND_PRIORITY=-2
ab main3_hand2 HOMAN_B200_MAIN_PRIORITY=-3 HOMAN_B200_HAND_PRIORITY=-2
ab default2 A=1
