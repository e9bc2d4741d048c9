#!/bin/bash
# 2+ GPUs: NCCL parity of both sharded formulations + owner-compute timings (config 2 rows, config 5 rows)
NG=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -x --timeout 600 > gpurun_out/pytest_sharded.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sharded.log
grep -E "passed|failed|FAILED|ERROR|Error|assert|^E " gpurun_out/pytest_sharded.log | head -30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
rm -f gpurun_out/dev_bench_owner_g$NG.log
for cfg in "RSB_N=10000001 RSB_LOSS=0" "RSB_N=10000001 RSB_LOSS=1" "RSB_N=100000001 RSB_LOSS=0"; do
  echo "# $cfg" >> gpurun_out/dev_bench_owner_g$NG.log
  env $cfg timeout 600 $TR tools/dev_bench_owner.py 2>&1 | grep -v "^W\|^\*\*\*\|NCCL version" >> gpurun_out/dev_bench_owner_g$NG.log
done
cut -c1-1200 gpurun_out/dev_bench_owner_g$NG.log
