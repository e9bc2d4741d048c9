#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_midx.py tests/test_gpu_shard_step.py tests/test_gpu_sharded.py -q -m gpu --timeout 600 2>&1 | tail -2
timeout 900 python tools/dev_bench_midx.py > gpurun_out/dev_bench_midx.json 2> gpurun_out/dev_bench_midx.err
cat gpurun_out/dev_bench_midx.json; tail -3 gpurun_out/dev_bench_midx.err
python - <<'PY'
import torch, json
x = torch.empty(2_900_000_000 // 4, device="cuda")
res = {}
for name, fn in (("memset_2.9GB", lambda: x.zero_()), ("fill_2.9GB", lambda: x.fill_(1.5))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[name] = {"ms": ms, "GBps": x.numel() * 4 / ms / 1e6}
y = torch.empty_like(x[: x.numel() // 2]); 
def cp(): y.copy_(x[: y.numel()])
cp(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): cp()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
res["copy_1.45GB"] = {"ms": ms, "GBps": 2 * y.numel() * 4 / ms / 1e6}
s = 0
def rd(): return x.sum()
rd(); torch.cuda.synchronize()
e0.record()
for _ in range(10): rd()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
res["read_sum_2.9GB"] = {"ms": ms, "GBps": x.numel() * 4 / ms / 1e6}
print(json.dumps(res))
open("gpurun_out/bw_probe.json", "w").write(json.dumps(res))
PY
bash tools/gpu_ncu.sh > gpurun_out/ncu_all.log 2>&1; tail -5 gpurun_out/ncu_all.log
