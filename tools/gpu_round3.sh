#!/bin/bash
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -q -m gpu --timeout 1500 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|Error|assert|^E " gpurun_out/pytest_gpu.log | head -60
rm -f gpurun_out/dev_bench.log
for cfg in "--loss 0 --score 0 --variant 0" "--loss 0 --score 0 --variant 3" "--loss 1 --score 0 --variant 3"; do
  echo "# $cfg" >> gpurun_out/dev_bench.log
  timeout 300 python tools/dev_bench.py $cfg >> gpurun_out/dev_bench.log 2>&1
done
cut -c1-330 gpurun_out/dev_bench.log
