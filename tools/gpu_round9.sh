#!/bin/bash
# L2 eviction-hint variants (16..31) of the fused step at config 2: parity + per-phase timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py -q -m gpu -k "variants or golden_steps" --timeout 600 > gpurun_out/pytest_variants.log 2>&1
tail -3 gpurun_out/pytest_variants.log
timeout 600 python tools/dev_bench.py --steps 20 --variants 0,17,19,22,23,31,24,0 > gpurun_out/dev_bench_hints.jsonl 2> gpurun_out/dev_bench_hints.err
cut -c1-420 gpurun_out/dev_bench_hints.jsonl; tail -3 gpurun_out/dev_bench_hints.err
timeout 600 python tools/dev_bench.py --steps 20 --loss 1 --variants 0,23,31 >> gpurun_out/dev_bench_hints.jsonl 2>> gpurun_out/dev_bench_hints.err
tail -3 gpurun_out/dev_bench_hints.jsonl | cut -c1-420
