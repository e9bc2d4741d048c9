#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_midx.py -q -m gpu --timeout 600 2>&1 | tail -2
timeout 900 python tools/dev_bench_midx.py > gpurun_out/dev_bench_midx.json 2> gpurun_out/dev_bench_midx.err
cat gpurun_out/dev_bench_midx.json; tail -5 gpurun_out/dev_bench_midx.err
