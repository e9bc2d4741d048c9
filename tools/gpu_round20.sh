#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_rowopt.py -q -m gpu --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | cut -c1-200 | head -10; done
