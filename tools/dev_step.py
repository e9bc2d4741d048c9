"""Developer micro-benchmark: per-phase CUDA-event timing of the fused step (default: BASELINE config 2,
10M x 128, B = 8192, n = 1024) for one or several (grouping, bin_shift, variant) settings, interleaved step by
step so that box-to-box drift cancels.  One JSON line per setting.  Not the contract bench (see bench.py).

    python tools/dev_step.py --set 0:0:0 --set 1:11:0 --set 1:12:0       # grouping:bin_shift:variant
"""
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--set", action="append", default=[], help="grouping:bin_shift:variant (bin_shift 0 = library default)")
    ap.add_argument("--tag", default="")
    ap.add_argument("--fused-draw", action="store_true", help="grouping 1: rsb200_pair_draw_count instead of draw + PHASE_COUNT")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(2022)
    wi = torch.empty(a.N, a.d, device=dev).normal_(0, 0.05); wi[0] = 0
    wu = torch.empty(a.U, a.d, device=dev).normal_(0, 0.05); wu[0] = 0
    user = torch.randint(1, a.U, (a.B,), device=dev)
    pos = torch.randint(1, a.N, (a.B,), device=dev)
    settings = [tuple(int(x) for x in s.split(":")) for s in (a.set or ["0:0:0", "1:0:0"])]
    wss = [fused.PairWorkspace(a.N, a.U, a.B, a.n, a.d, dev, grouping=g, bin_shift=(sh or None)) for g, sh, _ in settings]
    nbuf = torch.zeros(a.B, a.n, dtype=torch.int32, device=dev)
    names = ["sample", "count", "scan", "fwd", "scatter"]
    tot = [{k: 0.0 for k in names} for _ in settings]
    losses = [0.0] * len(settings)
    P = _lib
    for it in range(a.steps + 3):
        for k, ((g, sh, variant), ws) in enumerate(zip(settings, wss)):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            if a.fused_draw and g:           # draw + bin histogram in one kernel: reported under "sample", "count" = 0
                gen = torch.cuda.default_generators[0]
                neg32 = nbuf
                fused.pair_step(ws, wi, wu, user, pos, neg32, a.loss, a.score, phases=0,
                                draw={"seed": gen.initial_seed(), "offset": gen.get_offset()})
                gen.set_offset(gen.get_offset() + sampling.counter_offset(a.B * a.n, dev))
                ev[1].record()
                phs = (0, P.PHASE_SCAN, P.PHASE_FWD, P.PHASE_SCATTER)
            else:
                _, neg32 = sampling.uniform_draw(a.N, a.B, a.n, dev, want_i64=False, want_i32=True)
                ev[1].record()
                phs = (P.PHASE_COUNT, P.PHASE_SCAN, P.PHASE_FWD, P.PHASE_SCATTER)
            for i, ph in enumerate(phs):
                if ph:
                    loss = fused.pair_step(ws, wi, wu, user, pos, neg32, a.loss, a.score, phases=ph, variant=variant)
                ev[2 + i].record()
            torch.cuda.synchronize()
            if it >= 3:
                for i, nm in enumerate(names):
                    tot[k][nm] += ev[i].elapsed_time(ev[i + 1])
                losses[k] = float(loss.item())
    for k, ((g, sh, variant), ws) in enumerate(zip(settings, wss)):
        ph = {nm: tot[k][nm] / a.steps for nm in names}
        step = sum(ph.values())
        t = ws.totals.tolist()
        print(json.dumps({"tag": a.tag, "fused_draw": bool(a.fused_draw and g), "grouping": g, "bin_shift": ws.bin_shift, "variant": variant, "N": a.N, "B": a.B, "n": a.n,
                          "d": a.d, "loss_kind": a.loss, "score_kind": a.score, "ms": {k2: round(v, 4) for k2, v in ph.items()},
                          "step_ms": round(step, 4), "interactions_per_s": a.B / (step / 1e3), "loss": losses[k],
                          "entries": t[0], "unique_rows": t[1]}), flush=True)


if __name__ == "__main__":
    main()
