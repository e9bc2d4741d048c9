"""Developer timing of the OWNER-COMPUTE step of the row-sharded table (sharded.owner_compute_step:
ship queries, not rows) -- run with torchrun, one process per GPU (or plain python for world 1).
Reports whole-job interactions/s (max over ranks, CUDA events), the per-phase split of one rank and
the bytes each rank puts on NVLink per step.  RSB_N = global rows (default 10 000 001; 100 000 001 =
BASELINE config 5), RSB_LOSS = 0 BPR | 1 SSM."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recstudio_b200 import _lib, sampling, sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    N = int(os.environ.get("RSB_N", 10_000_001)); U, d, B, n = 1_000_001, 128, 8192, 1024
    loss_kind = int(os.environ.get("RSB_LOSS", 0))
    steps, warm = int(os.environ.get("RSB_STEPS", 20)), 3
    items = sharded.ShardedRows(N, d, dev, init_std=0.05, seed=1)
    wu = torch.empty(U, d, device=dev).normal_(0, 0.05); wu[0] = 0
    G = world * B
    eng = sharded.OwnerComputeCuda(N, items.row0, items.local_rows, items.weight, world, rank, G, n)
    gen = torch.Generator(device=dev).manual_seed(rank)
    users = torch.randint(1, U, (steps + warm, B), device=dev, generator=gen)
    poss = torch.randint(1, N, (steps + warm, B), device=dev, generator=gen)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(i, marks=None):
        _, neg = sampling.uniform_draw(N, B, n, dev, want_i64=False, want_i32=True)
        q = wu[users[i]]
        if marks: marks[0].record()
        q_all, pos_all, neg_all = (sharded._all_gather_cat(t) for t in (q, poss[i], neg))
        if marks: marks[1].record()
        eng.bind(q_all, pos_all, neg_all, loss_kind, _lib.SCORE_IP)
        sp = eng.prep()
        if marks: marks[2].record()
        dist.all_reduce(sp)
        mine = eng.fwd()
        if marks: marks[3].record()
        sharded.exchange_stats(eng, mine)
        loss, dq = eng.finish()
        if marks: marks[4].record()
        rows = eng.scatter()
        if marks: marks[5].record()
        dist.all_reduce(dq)
        if marks: marks[6].record()
        return loss

    # the API call itself, and the same with the negative ids of step i+1 drawn + all-gathered on a side stream
    # while step i computes (ids do not depend on the weights; queries / positives are gathered in-step)
    side = torch.cuda.Stream()

    def produce(i):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _, neg = sampling.uniform_draw(N, B, n, dev, want_i64=False, want_i32=True)
            neg_all = sharded._all_gather_cat(neg)
            e = ev(); e.record()
        return neg_all, e

    def api_step(i, pre=None):
        q = wu[users[i]]
        if pre is None:
            _, neg = sampling.uniform_draw(N, B, n, dev, want_i64=False, want_i32=True)
            return sharded.owner_compute_step(eng, q, poss[i], neg, loss_kind, _lib.SCORE_IP)[0]
        neg_all, e = pre
        torch.cuda.current_stream().wait_event(e)
        neg_all.record_stream(torch.cuda.current_stream())
        g = (sharded._all_gather_cat(q), sharded._all_gather_cat(poss[i]), neg_all, None, None)
        return sharded.owner_compute_step(eng, q, poss[i], None, loss_kind, _lib.SCORE_IP, gathered=g)[0]

    def regen_step(i):
        st = sharded.uniform_regen_state(dev, B, n)                       # 16 bytes per rank instead of B*n*4
        return sharded.owner_compute_step(eng, wu[users[i]], poss[i], None, loss_kind, _lib.SCORE_IP, regen_state=st)[0]

    results = {}
    for i in range(warm):
        regen_step(i)
    dist.barrier(); torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(steps):
        loss_regen = regen_step(warm + i)
    t1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([t0.elapsed_time(t1) / steps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    results["regen"] = t.item()
    for mode in ("plain",):
        for i in range(warm):
            api_step(i)
        dist.barrier(); torch.cuda.synchronize()
        t0, t1 = ev(), ev()
        pre = produce(warm) if mode == "prefetch" else None
        t0.record()
        for i in range(steps):
            if mode == "prefetch":
                nxt = produce(warm + i + 1) if i + 1 < steps else None
                loss = api_step(warm + i, pre)
                pre = nxt
            else:
                loss = api_step(warm + i)
        t1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([t0.elapsed_time(t1) / steps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[mode] = t.item()
    ms = results["plain"]
    marks = [ev() for _ in range(7)]
    s0 = ev(); s0.record()
    step(warm, marks)
    torch.cuda.synchronize()
    names = ["sample+q", "allgather(batch)", "prep(filter+count+scan)", "allreduce(sp)+fwd", "allgather(stats)+finish", "scatter", "allreduce(dq)"]
    split = {names[0]: s0.elapsed_time(marks[0])}
    for k in range(1, 7):
        split[names[k]] = marks[k - 1].elapsed_time(marks[k])
    eng.check()
    if rank == 0:
        wire = (world - 1) * B * (d * 4 + 8 + n * 4) + 2 * (world - 1) / world * G * (4 + d * 4) + (world - 1) * G * 8
        print(json.dumps({"path": "owner-compute (queries shipped)", "world": world, "N": N, "loss_kind": loss_kind,
                          "ms_per_step": ms, "interactions_per_s": G / ms * 1e3,
                          "ms_per_step_regen_ids": results["regen"], "interactions_per_s_regen_ids": G / results["regen"] * 1e3,
                          "loss": float(loss),
                          "owned_unique_rows_rank0": int(eng.totals[1].item()), "phase_ms_rank0": split,
                          "approx_nvlink_bytes_per_rank_per_step": int(wire)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
