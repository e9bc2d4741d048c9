#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_plugins.py tests/test_gpu_rowopt.py tests/test_gpu_sampling_methods.py -q -m gpu --timeout 600 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | cut -c1-250 | head -12
timeout 600 python tools/dev_bench.py --steps 60 --interleave --variants 0,7,6 > gpurun_out/dev_bench_stage.jsonl 2> gpurun_out/dev_bench_stage.err
timeout 600 python tools/dev_bench.py --steps 60 --interleave --loss 1 --variants 0,7 >> gpurun_out/dev_bench_stage.jsonl 2>> gpurun_out/dev_bench_stage.err
python - <<'PY'
import json
for l in open("gpurun_out/dev_bench_stage.jsonl"):
    d = json.loads(l); print(d["variant"], {k: round(v, 4) for k, v in d["ms"].items()}, round(d["step_ms"], 4))
PY
tail -3 gpurun_out/dev_bench_stage.err
