#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_shard_step.py tests/test_gpu_plugins.py -q -m gpu --timeout 600 2>&1 | tail -2
timeout 600 python tools/dev_bench.py --steps 60 --interleave --variants 0,6,40 > gpurun_out/dev_bench_hints2.jsonl 2> gpurun_out/dev_bench_hints2.err
python - <<'PY'
import json
for l in open("gpurun_out/dev_bench_hints2.jsonl"):
    d = json.loads(l); print(d["variant"], {k: round(v, 4) for k, v in d["ms"].items()}, round(d["step_ms"], 4))
PY
tail -3 gpurun_out/dev_bench_hints2.err
