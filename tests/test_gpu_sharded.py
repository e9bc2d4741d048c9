"""Row-sharded item table over NCCL (world size 2, one process per GPU): the sharded step must
reproduce the single-table result -- loss, owner-side item gradients, replicated user
gradients -- and every rank must draw exactly the negatives the reference sampler would draw
from its own generator state.  Skipped on a single-GPU box (the host logic is covered by the
world_size-2 gloo test on CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retriever as R

pytestmark = pytest.mark.gpu
WORLD = 2
N, U, D, B, NNEG = 50_001, 2_001, 64, 256, 300


def _data():
    g = torch.Generator().manual_seed(0)
    w_item = torch.randn(N, D, generator=g) * 0.3; w_item[0] = 0
    w_user = torch.randn(U, D, generator=g) * 0.3; w_user[0] = 0
    user = torch.randint(1, U, (WORLD, B), generator=g)
    pos = torch.randint(1, N, (WORLD, B), generator=g)
    return w_item, w_user, user, pos


def _worker(rank, port, loss_kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        from recstudio_b200 import sampling, sharded
        w_item, w_user, user, pos = _data()
        items = sharded.ShardedRows(N, D, dev)
        items.weight.copy_(w_item[items.row0:items.row0 + items.local_rows])
        torch.manual_seed(100 + rank)
        neg, _ = sampling.uniform_draw(N, B, NNEG, dev)                     # global ids, this rank's own stream
        torch.manual_seed(100 + rank)
        assert torch.equal(neg, torch.randint(1, N, (B, NNEG), device=dev))
        step = sharded.make_fused_step(B, NNEG, D, U, dev)
        step.loss_kind = loss_kind
        loss, (orow, oval), (urow, uval) = sharded.sharded_training_step(
            items, w_user.to(dev), user[rank].to(dev), pos[rank].to(dev), neg, loss_kind, R.IP, fused_step=step)
        torch.cuda.synchronize()
        q.put((rank, loss.item(), neg.cpu().numpy(), (orow + items.row0).cpu().numpy(), oval.cpu().numpy(),
               urow.cpu().numpy(), uval.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("loss_kind", [R.BPR, R.SSM])
def test_sharded_step_on_two_gpus(loss_kind):
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, port, loss_kind, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w_item, w_user, user, pos = _data()
    neg = torch.from_numpy(np.concatenate([r[2] for r in res]))
    ref = R.training_step_aten(w_item, w_user, user.reshape(-1), pos.reshape(-1), neg, loss=loss_kind, scorer=R.IP)
    d_item = np.zeros((N, D))
    for rank, loss, _, orow, oval, urow, uval in res:
        assert abs(loss - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
        assert np.all(np.diff(orow) > 0) and np.all(orow // ((N + WORLD - 1) // WORLD) == rank)
        d_item[orow] += oval
        du = np.zeros((U, D)); du[urow] = uval
        assert np.abs(du - ref["d_user"].numpy()).max() <= 1e-5 * np.abs(ref["d_user"].numpy()).max()
    assert np.abs(d_item - ref["d_item"].numpy()).max() <= 1e-5 * np.abs(ref["d_item"].numpy()).max()


def _oc_worker(rank, port, loss_kind, q, regen=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        from recstudio_b200 import sampling, sharded
        w_item, w_user, user, pos = _data()
        items = sharded.ShardedRows(N, D, dev)
        items.weight.copy_(w_item[items.row0:items.row0 + items.local_rows])
        torch.manual_seed(100 + rank)
        _, neg32 = sampling.uniform_draw(N, B, NNEG, dev, want_i64=False, want_i32=True)
        eng = sharded.OwnerComputeCuda(N, items.row0, items.local_rows, items.weight, WORLD, rank, WORLD * B, NNEG)
        if regen:       # the owners recompute every rank's draw from the all-gathered (seed, offset) pairs: no id exchange
            torch.manual_seed(100 + rank)
            torch.randint(1, N, (B, NNEG), device=dev)
            after = torch.rand(3, device=dev)                       # what follows the draw on this rank's stream
            torch.manual_seed(100 + rank)
            state = sharded.uniform_regen_state(dev, B, NNEG, group=items.group)
            assert torch.equal(torch.rand(3, device=dev), after)    # generator advanced exactly like the draw it replaces
            loss, (orow, oval), (urow, uval) = sharded.owner_compute_training_step(
                items, eng, w_user.to(dev), user[rank].to(dev), pos[rank].to(dev), None, loss_kind, R.IP, regen_state=state)
        else:
            loss, (orow, oval), (urow, uval) = sharded.owner_compute_training_step(
                items, eng, w_user.to(dev), user[rank].to(dev), pos[rank].to(dev), neg32, loss_kind, R.IP)
        eng.check()
        torch.cuda.synchronize()
        q.put((rank, loss.item(), neg32.cpu().numpy(), (orow + items.row0).cpu().numpy(), oval.cpu().numpy(),
               urow.cpu().numpy(), uval.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("regen", [False, True])
@pytest.mark.parametrize("loss_kind", [R.BPR, R.SSM])
def test_owner_compute_step_on_two_gpus(loss_kind, regen):
    """The query-shipping formulation (rsb200_shard_step + three small NCCL exchanges): same contract, same answer;
    regen=True: the negative ids are not exchanged at all, every owner regenerates both ranks' UniformSampler draws."""
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_oc_worker, args=(r, port, loss_kind, q, regen)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w_item, w_user, user, pos = _data()
    neg = torch.from_numpy(np.concatenate([r[2] for r in res])).long()
    ref = R.training_step_aten(w_item, w_user, user.reshape(-1), pos.reshape(-1), neg, loss=loss_kind, scorer=R.IP)
    d_item = np.zeros((N, D))
    for rank, loss, _, orow, oval, urow, uval in res:
        assert abs(loss - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
        assert np.all(np.diff(orow) > 0) and np.all(orow // ((N + WORLD - 1) // WORLD) == rank)
        d_item[orow] += oval
        du = np.zeros((U, D)); du[urow] = uval
        assert np.abs(du - ref["d_user"].numpy()).max() <= 1e-5 * np.abs(ref["d_user"].numpy()).max()
    assert np.abs(d_item - ref["d_item"].numpy()).max() <= 1e-5 * np.abs(ref["d_item"].numpy()).max()
