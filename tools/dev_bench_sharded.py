"""Developer timing of the row-sharded step (sharded.py: all-to-all row lookup, the north star's literal
design) -- run with torchrun, one process per GPU.  Reports interactions/s and the bytes that crossed
NVLink per step, next to what the replicated fused step does on the same shape."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recstudio_b200 import _lib, sampling, sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N = int(os.environ.get("RSB_N", 10_000_001)); U, d, B, n = 1_000_001, 128, 8192, 1024
    items = sharded.ShardedRows(N, d, dev, init_std=0.05, seed=1)
    wu = torch.empty(U, d, device=dev).normal_(0, 0.05); wu[0] = 0
    gen = torch.Generator(device=dev).manual_seed(rank)
    step = sharded.make_fused_step(B, n, d, U, dev)
    times, moved = [], 0
    for it in range(6):
        user = torch.randint(1, U, (B,), device=dev, generator=gen)
        pos = torch.randint(1, N, (B,), device=dev, generator=gen)
        neg, _ = sampling.uniform_draw(N, B, n, dev)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss, (orow, oval), (ur, uv) = sharded.sharded_training_step(items, wu, user, pos, neg, _lib.LOSS_BPR, _lib.SCORE_IP,
                                                                     fused_step=step)
        torch.cuda.synchronize(); dist.barrier()
        if it >= 2:
            times.append(time.perf_counter() - t0)
        moved = int(orow.numel())
    ms = 1e3 * sum(times) / len(times)
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "N": N, "ms_per_step": t.item(), "interactions_per_s": world * B / t.item() * 1e3,
                          "loss": float(loss), "owned_grad_rows_rank0": moved,
                          "note": "host-synchronous exchange plan (counts via .tolist()), torch.unique for the id set; "
                                  "NVLink bytes per rank per step ~ 2 x unique_rows x 512 B x (world-1)/world"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
