#!/bin/bash
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu --timeout 1500 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|Error|assert|^E " gpurun_out/pytest_gpu.log | head -40
rm -f gpurun_out/dev_bench.log
for cfg in "--loss 1 --sampler popular --mode 0"; do
  echo "# $cfg" >> gpurun_out/dev_bench.log
  timeout 600 python tools/dev_bench.py $cfg >> gpurun_out/dev_bench.log 2>&1
done
cut -c1-420 gpurun_out/dev_bench.log
