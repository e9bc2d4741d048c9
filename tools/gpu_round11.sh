#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_midx.py tests/test_gpu_sampling_methods.py -q -m gpu --timeout 600 ) > gpurun_out/pytest_midx.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_midx.log
grep -E "passed|failed|FAILED|Error|error|assert|^E " gpurun_out/pytest_midx.log | head -60
