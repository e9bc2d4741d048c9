#!/bin/bash
# attention re-test + timing after the 16-warp restructure, then smoke() and the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py -q -x --timeout 600 > gpurun_out/pytest_attn.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_attn.log
grep -E "passed|failed|FAILED|ERROR|Error|assert|^E " gpurun_out/pytest_attn.log | head -30
timeout 600 python tools/dev_bench_c3.py > gpurun_out/dev_bench_c3.log 2>&1
cut -c1-600 gpurun_out/dev_bench_c3.log | tail -8
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-1500
