#!/bin/bash
# 2-GPU call: attention/SASRec timing after the staging change, sharded NCCL tests + sharded step timing, 2-rank bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_attention.py -q -m gpu --timeout 900 > gpurun_out/pytest_gpu2.log 2>&1
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_gpu2.log | head -20
timeout 600 python tools/dev_bench_c3.py > gpurun_out/dev_c3.json 2> gpurun_out/dev_c3.err; cat gpurun_out/dev_c3.json; tail -3 gpurun_out/dev_c3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dev_bench_sharded.py > gpurun_out/dev_sharded.json 2> gpurun_out/dev_sharded.err
cat gpurun_out/dev_sharded.json; tail -4 gpurun_out/dev_sharded.err
