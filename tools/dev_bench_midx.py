"""Developer micro-benchmark of the 8(f)-4 index build (Sampler.update) and the 8(f)-2 sampling methods at
BASELINE config-2 table size (10M x 128).  Times the CUDA kernels and, next to them, the reference's own torch
formulation run on the same GPU (dense [N,K] distance / one-hot matrices).  Not the contract bench."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recstudio_b200 import midx, plugins, retriever  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--K", type=int, default=64)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    emb = torch.randn(a.N, a.d, device=dev) * 0.1
    X = emb[:, : a.d // 2]
    C = X[torch.randperm(a.N)[: a.K].to(dev)].clone()
    res = {"N": a.N, "d": a.d, "K": a.K}
    ms, (assign, loss) = timed(lambda: midx.kmeans_assign(X, C))
    res["kmeans_assign_ms"] = ms
    res["kmeans_assign_TFLOPs"] = 2.0 * a.N * a.K * (a.d // 2) / ms / 1e9
    res["kmeans_assign_GBps"] = a.N * (a.d // 2) * 4 / ms / 1e6
    ms, _ = timed(lambda: midx.kmeans_update(X, assign, a.K))
    res["kmeans_update_ms"] = ms
    res["kmeans_update_GBps"] = a.N * (a.d // 2) * 4 / ms / 1e6

    def ref_iter():        # sampler.py:19-31 as written, on the GPU
        dist = torch.sum(X * X, dim=-1, keepdim=True) - 2 * (X @ C.T) + torch.sum(C * C, dim=-1).unsqueeze(0)
        asg = dist.argmin(-1)
        am = X.new_zeros(a.N, a.K)
        am[(torch.arange(a.N, device=dev), asg)] = 1
        lossr = torch.sum(torch.square(X - C[asg, :]))
        return am.T @ X, am.sum(0), lossr
    ms, _ = timed(ref_iter, reps=3)
    res["reference_torch_iter_ms"] = ms
    codes = torch.randint(0, a.K * a.K, (a.N,), device=dev)
    ms, (ind, ptr_) = timed(lambda: midx.construct_index(codes, a.K * a.K))
    res["construct_index_ms"] = ms
    ms, _ = timed(lambda: torch.sort(codes, stable=True), reps=3)
    res["reference_torch_sort_ms"] = ms
    w = torch.rand(a.N, device=dev)
    ms, (cp, tot) = timed(lambda: midx.segment_cdf(w, ind, ptr_))
    res["segment_cdf_ms"] = ms
    k01 = torch.randint(0, a.K * a.K, (8192, 1024), device=dev)
    u = torch.rand(8192, 1024, device=dev)
    p = torch.cat([w.new_ones(1), w])
    ms, _ = timed(lambda: midx.segment_search(k01, u, cp, ind, ptr_, p))
    res["segment_search_8192x1024_ms"] = ms
    del emb, X, codes, ind, cp, w, k01, u, p
    torch.cuda.empty_cache()

    # 8(f)-2: dns step at config-2 shape: pool 1024 -> keep 256 hardest, BPR
    m = retriever.build_synthetic(1_000_001, a.N + 1, a.d, [1024, 256], loss="bpr", device=dev, sampling_method="dns", fused_grad="rows")
    g = torch.Generator(device=dev).manual_seed(1)
    batch = {"user_id": torch.randint(1, 1_000_001, (8192,), device=dev, generator=g),
             "item_id": torch.randint(1, a.N + 1, (8192,), device=dev, generator=g), "rating": torch.ones(8192, device=dev)}

    def dns_step():
        loss = m.training_step(batch)
        loss.backward()
        return loss
    ms, _ = timed(dns_step, reps=10)
    res["dns_1024to256_step_ms"] = ms
    res["dns_interactions_per_s"] = 8192 / ms * 1e3
    q = m.query_encoder(batch["user_id"])
    pool = torch.randint(1, a.N + 1, (8192, 1024), device=dev)
    ms, _ = timed(lambda: plugins.score_ids(0, q, m.item_encoder.weight, pool), reps=10)
    res["score_ids_8192x1024_ms"] = ms
    res["score_ids_GBps"] = 8192 * 1024 * a.d * 4 / ms / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
