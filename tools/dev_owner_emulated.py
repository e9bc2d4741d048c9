"""Developer tool: kernel times of ONE owner of a P-way row-sharded table on a single GPU, without NCCL: the global batch
(G = P x 8192 queries) is synthesised locally and the owner's phases of rsb200_shard_step are timed with CUDA events
(exchanges skipped: sp / stats / dq are whatever this owner computed -- timing only, not a parity run).
    python tools/dev_owner_emulated.py --world 8 [--rank 0] [--items 10000001] [--sampler uniform|popular]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from recstudio_b200 import _lib, sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--items", type=int, default=bench.N_ITEMS)
    ap.add_argument("--sampler", default="uniform")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--loss", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    P, r, N, d, B, n = a.world, a.rank, a.items, bench.DIM, bench.BATCH, bench.NEG
    per = sharded.rows_per_rank(N, P)
    row0 = r * per
    local = max(0, min(per, N - row0))
    w = torch.empty(local, d, device=dev).normal_(0, 0.05)
    G = P * B
    pop, touches = None, None
    if a.sampler == "popular":
        g = torch.Generator().manual_seed(0)
        u = torch.rand(N, generator=g, dtype=torch.float64).clamp_(min=1e-300)
        counts = torch.floor(u.pow_(-1.0 / 0.05).clamp_(max=1e9)).to(torch.float32); counts[0] = 0
        table, prob = sharded.PopularSlice.tables(counts, mode=0)
        pop = sharded.PopularSlice(table, prob, row0, local, dev)
        touches = int(G * (n + 1) * (min(pop.cdf_hi, 1.0) - max(pop.cdf_lo, 0.0))) + 1
    eng = sharded.OwnerComputeCuda(N, row0, local, w, P, r, G, n, with_logq=pop is not None, expected_touches=touches)
    q_all = torch.empty(G, d, device=dev).normal_(0, 0.05)
    names = ["prep_neg", "prep_pos", "fwd", "finish", "scatter"]
    tot = {k: 0.0 for k in names}
    for it in range(a.steps + 3):
        pos_all = torch.randint(1, N, (G,), device=dev)
        state = torch.stack([torch.arange(100, 100 + P, dtype=torch.int64), torch.full((P,), 4 * (1000 + 8192 * it), dtype=torch.int64)], 1).to(dev)
        kw = {"pop": pop} if pop is not None else {}
        eng.bind(q_all, pos_all, None, a.loss, _lib.SCORE_IP, regen_state=state, **kw)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        torch.cuda.synchronize()
        ev[0].record(); eng.prep_neg()
        ev[1].record(); eng.prep_pos()
        ev[2].record(); eng.fwd()
        ev[3].record(); eng.finish()
        ev[4].record(); eng.scatter()
        ev[5].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, nm in enumerate(names):
                tot[nm] += ev[i].elapsed_time(ev[i + 1])
    eng.check()
    ph = {k: round(v / a.steps, 4) for k, v in tot.items()}
    print(json.dumps({"emulated_owner": r, "world": P, "items": N, "sampler": a.sampler, "loss_kind": a.loss, "phase_ms": ph,
                      "sum_ms": round(sum(ph.values()), 4), "bin_shift": eng.bin_shift, "owned_touches": int(eng.totals[0].item()),
                      "owned_unique_rows": int(eng.totals[1].item())}), flush=True)


if __name__ == "__main__":
    main()
