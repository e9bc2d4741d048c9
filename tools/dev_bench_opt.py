"""Developer benchmark of the FULL training iteration at BASELINE config 2 (training_step + backward + optimizer.step())
for the gradient hand-off modes: 'rows' (gradient rows -> rsb200_rows_update) vs 'apply' (optimizer fused into the scatter
epilogue, RSB200_SINK_APPLY).  Not the contract bench: bench.py times the gather -> scatter path without the optimizer."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recstudio_b200 import retriever, rowopt  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    N, U, d, B, n = 10_000_001, 1_000_001, 128, 8192, 1024
    g = torch.Generator(device=dev).manual_seed(0)
    batches = [{"user_id": torch.randint(1, U, (B,), device=dev, generator=g), "item_id": torch.randint(1, N, (B,), device=dev, generator=g),
                "rating": torch.ones(B, device=dev)} for _ in range(8)]
    out = {}
    for learner in ("sgd", "adagrad", "sparse_adam"):
        for mode in ("rows", "apply"):
            m = retriever.build_synthetic(U, N, d, n, fused_grad=mode, device=dev)
            opt = rowopt.FusedRowOptimizer(m, learner, lr=0.01)

            def it(i):
                loss = m.training_step(batch=dict(batches[i % 8]))
                loss.backward()
                opt.step()
                return loss
            for i in range(5):
                it(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K = 40
            for i in range(K):
                loss = it(i)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            out["%s_%s" % (learner, mode)] = {"ms_per_iteration": ms, "interactions_per_s": B / ms * 1e3, "loss": float(loss.item())}
            del m, opt
            torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
