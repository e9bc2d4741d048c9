#!/bin/bash
# final-state ncu evidence (one GPU): per-step launch list + full captures of the two dominant kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 15 --csv --log-file gpurun_out/launches_final.csv \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_fwd_final -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 6 -c 1 -o gpurun_out/prof_scatter_final -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_scatter.log 2>&1
ls -la gpurun_out/*final*
