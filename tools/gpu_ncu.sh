#!/bin/bash
# ncu evidence for the dominant kernels at config 2 (one GPU).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 42 -c 14 --csv --log-file gpurun_out/launches.csv \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_fwd -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 6 -c 1 -o gpurun_out/prof_scatter -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_scatter.log 2>&1
ncu --set full --clock-control none -k regex:score_gmax_kernel -s 1 -c 1 -o gpurun_out/prof_topk -f \
    python tools/dev_bench_c4.py > gpurun_out/ncu_topk.log 2>&1
ls -la gpurun_out | head -30
