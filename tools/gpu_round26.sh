#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -q -m gpu --timeout 500 -k "graphed or golden_steps" 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | cut -c1-300 | head -12
timeout 600 python bench.py --no-cpu > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -2 gpurun_out/bench_graph.err; python -c "
import json; d=json.load(open('gpurun_out/bench_graph.json')); print('graph   ', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --no-cpu --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; tail -2 gpurun_out/bench_nograph.err; python -c "
import json; d=json.load(open('gpurun_out/bench_nograph.json')); print('no-graph', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])"
