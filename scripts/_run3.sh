#!/bin/bash
# Development: _run2 + the iteration-level comparator
bash scripts/_run2.sh
timeout 600 python scripts/bench_iteration_vs_nmr_style.py --iters 5 --out gpurun_out/iteration_vs_nmr_style.json 2>&1 | tail -5
