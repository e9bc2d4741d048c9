#!/bin/bash
# One gpurun call: GPU tests + per-phase timing at config 2.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
for cfg in "--loss 0 --score 0" "--loss 1 --score 0" "--loss 0 --score 1"; do
  timeout 300 python tools/dev_bench.py $cfg >> gpurun_out/dev_bench.log 2>&1
done
cat gpurun_out/dev_bench.log
