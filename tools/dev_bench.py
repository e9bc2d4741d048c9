"""Developer micro-benchmark: per-phase CUDA-event timing of the fused step at BASELINE
config 2 (10M x 128, B = 8192, n = 1024).  Not the contract bench (see bench.py)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recstudio_b200 import _lib, fused, sampling  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=10_000_001)
    ap.add_argument("--U", type=int, default=1_000_001)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--B", type=int, default=8192)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--loss", type=int, default=0)
    ap.add_argument("--score", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--variants", type=str, default="", help="comma list: time each variant in this process")
    ap.add_argument("--interleave", action="store_true", help="alternate the variants step by step (cancels drift between runs)")
    ap.add_argument("--sampler", default="uniform", choices=["uniform", "popular"])
    ap.add_argument("--mode", type=int, default=0, help="PopularSamplerModel mode (0 log, 2 count^0.75)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(2022)
    wi = torch.empty(a.N, a.d, device=dev).normal_(0, 0.01); wi[0] = 0
    wu = torch.empty(a.U, a.d, device=dev).normal_(0, 0.1); wu[0] = 0
    user = torch.randint(1, a.U, (a.B,), device=dev)
    pos = torch.randint(1, a.N, (a.B,), device=dev)
    ws = fused.PairWorkspace(a.N, a.U, a.B, a.n, a.d, dev, stage_entries=True)
    pop = None
    if a.sampler == "popular":
        import numpy as np
        from recstudio_b200 import plugins
        rng = np.random.RandomState(0)
        counts = np.floor(rng.zipf(1.05, size=a.N)).clip(max=1e9)            # Zipf(1.05) interaction counts
        counts[0] = 0
        pop = plugins.FusedPopularSampler(counts, mode=a.mode).to(dev)
    names = ["sample", "count", "scan", "fwd", "scatter"]
    variants = [int(v) for v in a.variants.split(",")] if a.variants else [a.variant]
    if a.interleave:
        run_interleaved(a, dev, wi, wu, user, pos, ws, variants, names)
        return
    for variant in variants:
        a.variant = variant
        run_variant(a, dev, wi, wu, user, pos, ws, pop, names)


def run_interleaved(a, dev, wi, wu, user, pos, ws, variants, names):
    tot = {v: {k: 0.0 for k in names} for v in variants}
    for it in range(a.steps + 3):
        for v in variants:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            _, neg32 = sampling.uniform_draw(a.N, a.B, a.n, dev, want_i64=False, want_i32=True)
            ev[1].record()
            for i, ph in enumerate((_lib.PHASE_COUNT, _lib.PHASE_SCAN, _lib.PHASE_FWD, _lib.PHASE_SCATTER)):
                fused.pair_step(ws, wi, wu, user, pos, neg32, a.loss, a.score, phases=ph, variant=v)
                ev[i + 2].record()
            torch.cuda.synchronize()
            if it >= 3:
                for i, k in enumerate(names):
                    tot[v][k] += ev[i].elapsed_time(ev[i + 1])
    for v in variants:
        ms = {k: x / a.steps for k, x in tot[v].items()}
        print(json.dumps({"variant": v, "interleaved": True, "ms": ms, "step_ms": sum(ms.values())}))


def run_variant(a, dev, wi, wu, user, pos, ws, pop, names):
    tot = {k: 0.0 for k in names}
    for it in range(a.steps + 3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        if pop is None:
            _, neg32 = sampling.uniform_draw(a.N, a.B, a.n, dev, want_i64=False, want_i32=True)
        else:
            neg32, _lq = pop.fused_draw(a.B, a.n, dev)
        ev[1].record()
        for i, ph in enumerate((_lib.PHASE_COUNT, _lib.PHASE_SCAN, _lib.PHASE_FWD, _lib.PHASE_SCATTER)):
            loss = fused.pair_step(ws, wi, wu, user, pos, neg32, a.loss, a.score, phases=ph, variant=a.variant)
            ev[i + 2].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, k in enumerate(names):
                tot[k] += ev[i].elapsed_time(ev[i + 1])
    ms = {k: v / a.steps for k, v in tot.items()}
    step = sum(ms.values())
    rows = (a.n + 2) * a.B
    alg = 2 * rows * a.d * 4
    t = ws.totals.tolist()
    print(json.dumps({"variant": a.variant, "ms": ms, "step_ms": step, "loss": float(loss.item()), "interactions_per_s": a.B / step * 1e3,
                      "fwd_GBps": rows * a.d * 4 / ms["fwd"] / 1e6, "step_alg_GBps": alg / step / 1e6,
                      "unique_item_rows": t[1], "entries": t[0], "ws_GB": ws.nbytes() / 1e9}))


if __name__ == "__main__":
    main()
