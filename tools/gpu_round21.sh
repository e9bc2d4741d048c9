#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py -q -m gpu --timeout 600 -k "config2" 2>&1 | grep -E "^E  |FAILED|passed|failed|Error|error" | cut -c1-300 | head -20
