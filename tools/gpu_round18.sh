#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py -q -m gpu --timeout 600 2>&1 | tail -2
timeout 600 python tools/dev_bench.py --steps 10 --loss 1 --sampler popular --mode 0 2>&1 | tail -1 | cut -c1-200
timeout 900 python tools/dev_bench.py --steps 5 --N 100000001 --loss 1 --sampler popular --mode 0 2>&1 | tail -1 | cut -c1-200
