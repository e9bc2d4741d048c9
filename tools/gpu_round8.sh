#!/bin/bash
# bench.py --workload c5-sharded at NG GPUs (torchrun for NG > 1)
NG=${1:-2}
mkdir -p gpurun_out
if [ "$NG" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512"; fi
timeout 800 $RUN bench.py --gpus $NG --workload c5-sharded > gpurun_out/bench_c5_g$NG.log 2>&1
tail -1 gpurun_out/bench_c5_g$NG.log | cut -c1-2500
