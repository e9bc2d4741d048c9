#!/bin/bash
# N GPUs: owner-compute timings at config-2 rows and config-5 rows (no pytest: parity is covered at 2 GPUs)
NG=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
rm -f gpurun_out/dev_bench_owner_g$NG.log
for cfg in "RSB_N=10000001 RSB_LOSS=0" "RSB_N=100000001 RSB_LOSS=0" "RSB_N=100000001 RSB_LOSS=1"; do
  echo "# $cfg" >> gpurun_out/dev_bench_owner_g$NG.log
  env $cfg timeout 400 $TR tools/dev_bench_owner.py 2>&1 | grep -v "^W\|^\*\*\*\|NCCL version\|OMP_NUM_THREADS\|^$" >> gpurun_out/dev_bench_owner_g$NG.log
done
cut -c1-1300 gpurun_out/dev_bench_owner_g$NG.log
