#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowopt.py tests/test_gpu_pair.py -q -m gpu --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | cut -c1-200 | head -10
timeout 900 python tools/dev_bench_opt.py > gpurun_out/dev_bench_opt.json 2> gpurun_out/dev_bench_opt.err
cat gpurun_out/dev_bench_opt.json; tail -3 gpurun_out/dev_bench_opt.err
