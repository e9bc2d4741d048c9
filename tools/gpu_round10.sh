#!/bin/bash
# full GPU test suite (timed) after the 8(f)-2 additions
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 -x --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
