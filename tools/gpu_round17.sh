#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_plugins.py tests/test_gpu_shard_step.py -q -m gpu --timeout 600 2>&1 | tail -2
timeout 600 python tools/dev_bench.py --steps 20 --variants 0,40,0,40 > gpurun_out/dev_bench_scat.jsonl 2> gpurun_out/dev_bench_scat.err
python - <<'PY'
import json
for l in open("gpurun_out/dev_bench_scat.jsonl"):
    d = json.loads(l); print(d["variant"], {k: round(v, 4) for k, v in d["ms"].items()}, round(d["step_ms"], 4))
PY
tail -3 gpurun_out/dev_bench_scat.err
