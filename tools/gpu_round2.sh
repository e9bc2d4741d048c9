#!/bin/bash
# 2-GPU call: full GPU test suite (incl. the NCCL sharded test), config-2/4 timings, 2-rank bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|Error" gpurun_out/pytest_gpu.log | head -40
rm -f gpurun_out/dev_bench.log
for cfg in "--loss 0 --score 0" "--loss 1 --score 0" "--loss 0 --score 1" "--loss 1 --score 1"; do
  echo "# $cfg" >> gpurun_out/dev_bench.log
  timeout 300 python tools/dev_bench.py $cfg >> gpurun_out/dev_bench.log 2>&1
done
cat gpurun_out/dev_bench.log | cut -c1-330
timeout 600 python tools/dev_bench_c4.py > gpurun_out/dev_c4.json 2> gpurun_out/dev_c4.err; cat gpurun_out/dev_c4.json; tail -3 gpurun_out/dev_c4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench2 exit $?"; cat gpurun_out/bench_2gpu.json | cut -c1-600; tail -3 gpurun_out/bench_2gpu.err
