#!/bin/bash
# One gpurun call: GPU tests + per-phase timing at config 2 + the contract bench.  Logs -> gpurun_out/.
# usage: tools/gpu_round.sh [tests] [dev] [bench] [cpu]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
case $what in
tests)
  timeout 2400 python -m pytest tests -q -m gpu -x --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -30 gpurun_out/pytest_gpu.log ;;
dev)
  rm -f gpurun_out/dev_bench.log
  for cfg in "--loss 0 --score 0 --variant 0" "--loss 0 --score 0 --variant 2" "--loss 1 --score 0 --variant 0" "--loss 1 --score 0 --variant 2" "--loss 0 --score 1 --variant 0"; do
    echo "# $cfg" >> gpurun_out/dev_bench.log
    timeout 300 python tools/dev_bench.py $cfg >> gpurun_out/dev_bench.log 2>&1
  done
  cat gpurun_out/dev_bench.log ;;
bench)
  timeout 900 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
cpu)
  for t in 0 32 10; do
    timeout 900 python bench.py --impl reference --steps 2 --warmup 1 --cpu-threads $t > gpurun_out/bench_ref_$t.json 2> gpurun_out/bench_ref.err
    python -c "import json;d=json.load(open('gpurun_out/bench_ref_$t.json'));print('cpu threads',d['cpu_baseline']['cores'],'value',d['value'],'ms',d['ms_per_step'])"
  done ;;
esac
done
