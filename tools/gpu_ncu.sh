#!/bin/bash
# ncu evidence for the two dominant kernels at config 2 (one GPU).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 28 -c 28 --csv --log-file gpurun_out/launches.csv \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_fwd -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 4 -c 1 -o gpurun_out/prof_scatter -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_scatter.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/ncu_fwd.log
