#!/bin/bash
# 2 GPUs: NCCL parity of the owner-compute step incl. owner-side regeneration, then the dev bench (10M x 128 table)
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 500 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | cut -c1-300 | head -12
RSB_STEPS=30 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dev_bench_owner.py 2>&1 | tail -1 | cut -c1-1300
