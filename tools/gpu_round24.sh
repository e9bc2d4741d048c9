#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_shard_step.py -q -m gpu --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | cut -c1-300 | head -12
RSB_STEPS=30 timeout 600 python tools/dev_bench_owner.py 2>&1 | tail -1 | cut -c1-1200
