#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_attention.py -q -m gpu --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | cut -c1-300 | head -12
