#!/bin/bash
# owner-compute (query shipping) sharded step: single-GPU emulated-owner parity + world-1 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard_step.py -q -x --timeout 600 > gpurun_out/pytest_shard.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_shard.log
grep -E "passed|failed|FAILED|ERROR|Error|assert|^E " gpurun_out/pytest_shard.log | head -30
timeout 600 python tools/dev_bench_owner.py > gpurun_out/dev_bench_owner1.log 2>&1
tail -2 gpurun_out/dev_bench_owner1.log | cut -c1-1200
