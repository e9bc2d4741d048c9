"""world_size-2 `gloo` test (CPU) of the row-sharded step's HOST logic: ownership plan,
variable-size all-to-all exchanges, compact-table id remapping, owner-side accumulation and
the replicated-user all-gather.  The arithmetic is injected from the CPU oracle (test
infrastructure), so what is verified is exactly the exchange code that runs under NCCL on the
GPUs: the sharded result must equal the single-process result on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retriever as R

WORLD = 2
N, U, D, B, NNEG = 103, 17, 8, 12, 9


class TorchOps:
    """CPU stand-ins for the CUDA ops (test only)."""

    @staticmethod
    def gather_rows(weight, ids):
        return weight[ids]

    @staticmethod
    def coalesce_rows(ids, vals, num_rows, skip_row0=False):
        if skip_row0:
            keep = ids != 0
            ids, vals = ids[keep], vals[keep]
        rows, inv = torch.unique(ids, sorted=True, return_inverse=True)
        out = torch.zeros(rows.numel(), vals.shape[1], dtype=vals.dtype)
        out.index_add_(0, inv, vals)
        return rows, out


def oracle_fused_step(loss_kind, score_kind):
    def step(w_local, w_user, user, pos_c, neg32, lqp, lqn, grad_scale):
        out = R.training_step_aten(w_local, w_user, user, pos_c, neg32.long(), loss=loss_kind, scorer=score_kind,
                                   log_pos_prob=lqp, log_neg_prob=lqn)
        gi, gu = out["d_item"] * grad_scale, out["d_user"] * grad_scale
        ri = torch.nonzero(gi.abs().sum(-1) > 0).flatten()
        ru = torch.nonzero(gu.abs().sum(-1) > 0).flatten()
        return out["loss"], ri, gi[ri], ru, gu[ru]
    return step


def _data():
    g = torch.Generator().manual_seed(0)
    w_item = torch.randn(N, D, generator=g) * 0.4; w_item[0] = 0
    w_user = torch.randn(U, D, generator=g) * 0.4; w_user[0] = 0
    user = torch.randint(1, U, (WORLD, B), generator=g)
    pos = torch.randint(1, N, (WORLD, B), generator=g)
    neg = torch.randint(0, N, (WORLD, B, NNEG), generator=g)       # includes the padding id 0
    pos[0, 0] = 0
    return w_item, w_user, user, pos, neg


def _worker(rank, port, loss_kind, score_kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from recstudio_b200 import sharded
        w_item, w_user, user, pos, neg = _data()
        items = sharded.ShardedRows(N, D, "cpu", ops=TorchOps)
        assert items.per_rank == (N + WORLD - 1) // WORLD
        items.weight.copy_(w_item[items.row0:items.row0 + items.local_rows])
        loss, (orow, oval), (urow, uval) = sharded.sharded_training_step(
            items, w_user, user[rank], pos[rank], neg[rank], loss_kind, score_kind,
            fused_step=oracle_fused_step(loss_kind, score_kind))
        q.put((rank, loss.item(), (orow + items.row0).numpy(), oval.numpy(), urow.numpy(), uval.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("loss_kind,score_kind", [(R.BPR, R.IP), (R.SSM, R.EUCLID)])
def test_sharded_step_equals_single_process(loss_kind, score_kind):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, port, loss_kind, score_kind, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w_item, w_user, user, pos, neg = _data()
    # reference: mean over ranks of the per-rank mean losses == one big batch of WORLD*B interactions
    ref = R.training_step_aten(w_item, w_user, user.reshape(-1), pos.reshape(-1), neg.reshape(-1, NNEG),
                               loss=loss_kind, scorer=score_kind)
    d_item = np.zeros((N, D)); d_user = None
    for rank, loss, orow, oval, urow, uval in res:
        assert abs(loss - ref["loss"].item()) < 1e-6 * max(1.0, abs(ref["loss"].item()))
        assert np.all(np.diff(orow) > 0) and np.all(orow // ((N + WORLD - 1) // WORLD) == rank)   # owners only get their rows
        d_item[orow] += oval
        du = np.zeros((U, D)); du[urow] = uval
        if d_user is None:
            d_user = du
        else:
            np.testing.assert_allclose(du, d_user, rtol=0, atol=1e-7)       # every replica sees the same user gradient
    np.testing.assert_allclose(d_item, ref["d_item"].numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(d_user, ref["d_user"].numpy(), rtol=1e-5, atol=1e-7)
    assert np.all(d_item[0] == 0)


def _oc_worker(rank, port, loss_kind, score_kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from oracle import shard_step as S
        from recstudio_b200 import sharded
        w_item, w_user, user, pos, neg = _data()
        items = sharded.ShardedRows(N, D, "cpu", ops=TorchOps)
        items.weight.copy_(w_item[items.row0:items.row0 + items.local_rows])
        g = torch.Generator().manual_seed(5)
        lqp, lqn = torch.randn(WORLD, B, generator=g), torch.randn(WORLD, B, NNEG, generator=g)
        eng = S.OwnerComputeOracle(N, items.row0, items.local_rows, items.weight, WORLD, rank, WORLD * B, NNEG, with_logq=True)
        loss, (orow, oval), (urow, uval) = sharded.owner_compute_training_step(
            items, eng, w_user, user[rank], pos[rank], neg[rank], loss_kind, score_kind,
            logq_pos=lqp[rank] if loss_kind == R.SSM else None, logq_neg=lqn[rank] if loss_kind == R.SSM else None)
        q.put((rank, loss.item(), (orow + items.row0).numpy(), oval.numpy(), urow.numpy(), uval.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("loss_kind,score_kind", [(R.BPR, R.IP), (R.BPR, R.EUCLID), (R.SSM, R.IP), (R.SSM, R.EUCLID)])
def test_owner_compute_step_equals_single_process(loss_kind, score_kind):
    """The owner-compute ("ship queries, not rows") choreography of sharded.owner_compute_step: all-gather of
    the batch, all-reduce of the positive scores, all-gather of the per-owner statistics, all-reduce of dq --
    with the per-owner arithmetic injected from oracle/shard_step.py.  Sum over owners == the reference step
    on the whole table."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_oc_worker, args=(r, port, loss_kind, score_kind, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w_item, w_user, user, pos, neg = _data()
    g = torch.Generator().manual_seed(5)
    lqp, lqn = torch.randn(WORLD, B, generator=g), torch.randn(WORLD, B, NNEG, generator=g)
    kw = dict(log_pos_prob=lqp.reshape(-1), log_neg_prob=lqn.reshape(-1, NNEG)) if loss_kind == R.SSM else {}
    ref = R.training_step_aten(w_item, w_user, user.reshape(-1), pos.reshape(-1), neg.reshape(-1, NNEG),
                               loss=loss_kind, scorer=score_kind, **kw)
    d_item = np.zeros((N, D))
    for rank, loss, orow, oval, urow, uval in res:
        assert abs(loss - ref["loss"].item()) < 2e-6 * max(1.0, abs(ref["loss"].item()))
        assert np.all(np.diff(orow) > 0) and np.all(orow // ((N + WORLD - 1) // WORLD) == rank)
        d_item[orow] += oval
        du = np.zeros((U, D)); du[urow] = uval
        np.testing.assert_allclose(du, ref["d_user"].numpy(), rtol=1e-4, atol=2e-7)
    np.testing.assert_allclose(d_item, ref["d_item"].numpy(), rtol=1e-4, atol=2e-7)
    assert np.all(d_item[0] == 0)


def test_owner_counts_plan():
    from recstudio_b200 import sharded
    ids = torch.tensor([0, 3, 4, 5, 9, 10, 11, 19])
    assert sharded.rows_per_rank(20, 4) == 5
    assert sharded.owner_counts(ids, 5, 4).tolist() == [3, 2, 2, 1]
    assert sharded.owner_counts(torch.tensor([], dtype=torch.int64), 5, 4).tolist() == [0, 0, 0, 0]
    assert sharded.owner_counts(torch.tensor([7, 8]), 5, 4).tolist() == [0, 2, 0, 0]
