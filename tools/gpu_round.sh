#!/bin/bash
# One gpurun call: GPU tests + per-phase timing at config 2 + the contract bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu -x --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
rm -f gpurun_out/dev_bench.log
for cfg in "--loss 0 --score 0 --variant 0" "--loss 0 --score 0 --variant 1" "--loss 1 --score 0 --variant 0" "--loss 0 --score 1 --variant 0"; do
  echo "# $cfg" >> gpurun_out/dev_bench.log
  timeout 300 python tools/dev_bench.py $cfg >> gpurun_out/dev_bench.log 2>&1
done
cat gpurun_out/dev_bench.log
if [ "$1" == "bench" ]; then
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
