#!/bin/bash
# One gpurun call: GPU tests + per-phase timing at config 2 + the contract bench.  Logs -> gpurun_out/.
# usage: tools/gpu_round.sh [tests] [dev] [bench] [ncu] [c4]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
case $what in
tests)
  timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  grep -E "passed|failed|FAILED|ERROR|Error|assert" gpurun_out/pytest_gpu.log | head -40 ;;
dev)
  rm -f gpurun_out/dev_bench.log
  for cfg in "--loss 0 --score 0" "--loss 1 --score 0" "--loss 0 --score 1"; do
    echo "# $cfg" >> gpurun_out/dev_bench.log
    timeout 300 python tools/dev_bench.py $cfg >> gpurun_out/dev_bench.log 2>&1
  done
  cut -c1-330 gpurun_out/dev_bench.log ;;
c4)
  timeout 600 python tools/dev_bench_c4.py > gpurun_out/dev_c4.json 2> gpurun_out/dev_c4.err; cat gpurun_out/dev_c4.json; tail -3 gpurun_out/dev_c4.err ;;
bench)
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json ;;
ncu)
  bash tools/gpu_ncu.sh ;;
esac
done
