"""Developer timing of the config-4 kernels (1M items x 128): full-catalog top-k (eval) and
full-softmax forward+backward (training).  Reports achieved fp32 FLOP/s next to the HBM floor."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recstudio_b200 import _lib, plugins, topk  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device("cuda:0")
    N, d = 1_000_001, 128
    torch.manual_seed(0)
    w = torch.randn(N, d, device=dev) * 0.1; w[0] = 0
    out = {}
    for Be, k, H in ((128, 10, 64), (128, 100, 64), (1024, 100, 64)):
        q = torch.randn(Be, d, device=dev) * 0.1
        hist = torch.randint(1, N, (Be, H), device=dev)
        ms = timeit(lambda: topk.topk_full(q, w, k, hist))
        ref = timeit(lambda: torch.topk(q @ w[1:].T, k + H))
        out[f"topk_Be{Be}_k{k}"] = {"ms": ms, "tflops": 2 * Be * (N - 1) * d / ms / 1e9, "torch_matmul_topk_ms": ref,
                                    "hbm_floor_ms": (N - 1) * d * 4 / 6.49e9}
    for B in (1024, 4096):
        q = torch.randn(B, d, device=dev) * 0.1
        pos = torch.randint(1, N, (B,), device=dev)
        wi = w.clone().requires_grad_(True)

        def step():
            qq = q.clone().requires_grad_(True)
            plugins._FullSoftmaxFn.apply(qq, wi, pos).backward()
            wi.grad = None

        def ref_step():
            qq = q.clone().requires_grad_(True)
            s = qq @ wi[1:].T
            (torch.logsumexp(s, -1) - (qq * wi[pos]).sum(-1)).mean().backward()
            wi.grad = None
        ms = timeit(step, iters=3, warm=1)
        rms = timeit(ref_step, iters=3, warm=1)
        out[f"fullsoftmax_B{B}"] = {"ms": ms, "tflops_4gemm": 8 * B * (N - 1) * d / ms / 1e9, "torch_ms": rms}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
