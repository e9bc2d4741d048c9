#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -q -m gpu -k "variants or golden_steps or random_vs" --timeout 600 2>&1 | tail -2
timeout 600 python tools/dev_bench.py --steps 20 --variants 0,5,32,0,5 > gpurun_out/dev_bench_occ.jsonl 2> gpurun_out/dev_bench_occ.err
cut -c1-330 gpurun_out/dev_bench_occ.jsonl; tail -3 gpurun_out/dev_bench_occ.err
timeout 600 python tools/dev_bench.py --steps 20 --loss 1 --variants 0,5 >> gpurun_out/dev_bench_occ.jsonl 2>> gpurun_out/dev_bench_occ.err
tail -2 gpurun_out/dev_bench_occ.jsonl | cut -c1-330
