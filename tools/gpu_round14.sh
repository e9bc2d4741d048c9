#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
