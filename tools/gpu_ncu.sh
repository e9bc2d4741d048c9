#!/bin/bash
# ncu evidence for the dominant kernels at config 2 (one GPU).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
# launch list of the contract bench command (kernel SHARES of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench_launches.log 2>&1
# per-phase launch list of one step (sample 1, count 3, scan 6, resolve 1, fwd 2, scatter 2 = 15 kernels per step)
ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 15 --csv --log-file gpurun_out/launches.csv \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_fwd -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 6 -c 1 -o gpurun_out/prof_scatter -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_scatter.log 2>&1
ncu --set full --clock-control none -k regex:count_kernel -s 9 -c 1 -o gpurun_out/prof_count -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_count.log 2>&1
ncu --set full --clock-control none -k regex:resolve_kernel -s 3 -c 1 -o gpurun_out/prof_resolve -f \
    python tools/dev_bench.py --steps 1 > gpurun_out/ncu_resolve.log 2>&1
ncu --set full --clock-control none -k regex:"kmeans_assign|kmeans_update|score_ids_stream" -c 3 -o gpurun_out/prof_midx -f \
    python tools/dev_bench_midx.py --N 2000000 > gpurun_out/ncu_midx.log 2>&1
ls -la gpurun_out | head -40
