"""Developer tool: per-phase CUDA-event timing of the row-sharded owner-compute step (bench.py's N > 1 step) on rank 0.
    python tools/dev_shard_phases.py [--items 10000001] [--sampler uniform|popular]            # one owner
    torchrun --nproc-per-node N ... tools/dev_shard_phases.py ...                               # N owners over NCCL
One JSON line (rank 0)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from recstudio_b200 import _lib, sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=bench.N_ITEMS)
    ap.add_argument("--sampler", default="uniform")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    world, rank, local = bench._dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if "MASTER_ADDR" not in os.environ:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", "29578"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(2022 + rank)
    sb = bench.ShardedBench(a.items, dev, world, rank, sampler=a.sampler)
    eng, sh = sb.eng, sharded
    gen = torch.Generator(device=dev).manual_seed(rank)
    names = ["gather_q+state", "prep_neg (all-gather of q, pos in flight)", "all-gather wait + prep_pos", "allreduce(sp)", "fwd",
             "allgather(stats)", "finish", "scatter", "allreduce(dq)"]
    tot = {k: 0.0 for k in names}
    for it in range(a.steps + 5):
        u = torch.randint(1, bench.N_USERS, (bench.BATCH,), device=dev, generator=gen)
        p = torch.randint(1, a.items, (bench.BATCH,), device=dev, generator=gen)
        dist.barrier(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record()
        q = sh.CudaOps.gather_rows(sb.wu, u)
        state = sb.states.next()
        ev[1].record()
        w1 = dist.all_gather_into_tensor(sb.q_all, q, async_op=True)
        w2 = dist.all_gather_into_tensor(sb.pos_all, p, async_op=True)
        kw = {"pop": sb.pop} if sb.pop is not None else {}
        eng.bind(sb.q_all, sb.pos_all, None, _lib.LOSS_BPR, _lib.SCORE_IP, regen_state=state, **kw)
        eng.prep_neg()
        ev[2].record()
        w1.wait(); w2.wait()
        sp = eng.prep_pos()
        ev[3].record()
        dist.all_reduce(sp)
        ev[4].record()
        mine = eng.fwd()
        ev[5].record()
        sh.exchange_stats(eng, mine)
        ev[6].record()
        loss, dq = eng.finish()
        ev[7].record()
        eng.scatter()
        ev[8].record()
        dist.all_reduce(dq)
        ev[9].record()
        torch.cuda.synchronize()
        if it >= 5:
            for i, nm in enumerate(names):
                tot[nm] += ev[i].elapsed_time(ev[i + 1])
    if rank == 0:
        ph = {k: round(v / a.steps, 4) for k, v in tot.items()}
        print(json.dumps({"world": world, "items": a.items, "sampler": a.sampler, "phase_ms_rank0": ph, "sum_ms": round(sum(ph.values()), 4),
                          "bin_shift": eng.bin_shift, "owned_touches": int(eng.totals[0].item()), "owned_unique_rows": int(eng.totals[1].item()),
                          "loss": float(loss.item())}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
