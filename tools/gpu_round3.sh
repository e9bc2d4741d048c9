#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|Error|assert|^E " gpurun_out/pytest_gpu.log | head -60
timeout 600 python tools/dev_bench_c3.py > gpurun_out/dev_c3.json 2> gpurun_out/dev_c3.err; cat gpurun_out/dev_c3.json; tail -5 gpurun_out/dev_c3.err
