#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 3 -o gpurun_out/prof_attn -f \
    python tools/dev_bench_c3.py > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/prof_attn.ncu-rep; tail -2 gpurun_out/ncu_attn.log
