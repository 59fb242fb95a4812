#!/bin/bash
# Runs the GPU test files one process each (a CUDA fault in one file cannot poison the others);
# full logs under gpurun_out/tests_*.log, failures summarised on stdout.
mkdir -p gpurun_out
python -m oracle.build > /dev/null
rc=0
for f in tests/test_*gpu*.py; do
    n=$(basename $f .py)
    timeout 600 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider "$@" > gpurun_out/tests_$n.log 2>&1 || rc=1
    tail -1 gpurun_out/tests_$n.log
    grep -E "^(E  |FAILED|ERROR)" gpurun_out/tests_$n.log | head -${GPU_TEST_LINES:-25}
done
exit $rc
