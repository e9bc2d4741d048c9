"""rsb200_shard_step (owner-compute step of the row-sharded table, csrc/shard.cu) on ONE GPU: the
owners are emulated one after the other on the same device and the three exchanges (sum of the
positive scores, gather of the statistics, sum of dq) are done by hand, so the kernels are checked
without NCCL.  Sum over owners must equal the reference step on the whole table
(oracle.retriever.training_step_aten = baseretriever.py:142-176,399-404 op for op) and, at a
larger size, the single-table fused step (rsb200_pair_step)."""
import numpy as np
import pytest
import torch

from oracle import retriever as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_owners(w_item, q_all, pos, neg, world, loss_kind, score_kind, lqp=None, lqn=None, grouping=None):
    """-> (loss, dense d_item [N, d], dq [G, d]) assembled from `world` owners run back to back."""
    from recstudio_b200 import sharded
    N, d = w_item.shape
    G, n = neg.shape
    per = sharded.rows_per_rank(N, world)
    engines = []
    for r in range(world):
        row0 = r * per
        local = max(0, min(per, N - row0))
        e = sharded.OwnerComputeCuda(N, row0, local, w_item[row0:row0 + local].contiguous(), world, r, G, n,
                                     with_logq=lqn is not None, grouping=grouping)
        assert grouping is None or e.grouping == grouping
        e.bind(q_all, pos, neg, loss_kind, score_kind, lqp, lqn)
        engines.append(e)
    sp = torch.stack([e.prep().clone() for e in engines]).sum(0)          # all-reduce SUM
    for e in engines:
        e.sp[:G] = sp
    stats = torch.stack([e.fwd().clone() for e in engines])               # all-gather
    for e in engines:
        e.stats_all[:, :G] = stats
    losses, dq = [], torch.zeros(G, d, device=DEV, dtype=torch.float64)
    d_item = torch.zeros(N, d, device=DEV)
    for e in engines:
        loss, dqe = e.finish()
        losses.append(loss.item()); dq += dqe.double()                    # all-reduce SUM
        rows, vals, totals = e.scatter()
        R_ = int(totals[1].item())
        rr = rows[:R_]
        assert torch.all(rr[1:] > rr[:-1]) and (R_ == 0 or (rr[0] >= 0 and rr[-1] < e.local_rows))
        d_item[rr + e.row0] += vals[:R_]
        e.check()
    assert max(losses) - min(losses) == 0.0, "every owner must compute the identical global loss"
    return losses[0], d_item, dq.float()


def _case(N, U, d, G, n, seed, pad=True):
    g = torch.Generator().manual_seed(seed)
    w_item = torch.randn(N, d, generator=g) * 0.4; w_item[0] = 0
    w_user = torch.randn(U, d, generator=g) * 0.4; w_user[0] = 0
    user = torch.randint(1, U, (G,), generator=g)
    pos = torch.randint(1, N, (G,), generator=g)
    neg = torch.randint(0 if pad else 1, N, (G, n), generator=g)
    if pad:
        pos[0] = 0
    lqp, lqn = torch.randn(G, generator=g), torch.randn(G, n, generator=g)
    return w_item, w_user, user, pos, neg, lqp, lqn


@pytest.mark.parametrize("grouping", [0, 1])      # 0: counting sort over the owner's rows, 1: bins (default)
@pytest.mark.parametrize("world", [1, 2, 3, 7])
@pytest.mark.parametrize("loss_kind,score_kind", [(R.BPR, R.IP), (R.BPR, R.EUCLID), (R.SSM, R.IP), (R.SSM, R.EUCLID)])
def test_owners_sum_to_the_reference_step(world, loss_kind, score_kind, grouping):
    N, U, d, G, n = 211, 40, 24, 37, 45          # n not a multiple of 32, d < 128, ids include the padding row
    w_item, w_user, user, pos, neg, lqp, lqn = _case(N, U, d, G, n, seed=world)
    ssm = loss_kind == R.SSM
    kw = dict(log_pos_prob=lqp, log_neg_prob=lqn) if ssm else {}
    ref = R.training_step_aten(w_item, w_user, user, pos, neg, loss=loss_kind, scorer=score_kind, **kw)
    q_all = w_user[user].to(DEV)
    loss, d_item, dq = run_owners(w_item.to(DEV), q_all, pos.to(DEV), neg.to(DEV).int(), world, loss_kind, score_kind,
                                  lqp.to(DEV) if ssm else None, lqn.to(DEV) if ssm else None, grouping=grouping)
    assert abs(loss - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
    gi = ref["d_item"].numpy()
    assert np.abs(d_item.cpu().numpy() - gi).max() <= 1e-5 * np.abs(gi).max()
    assert torch.all(d_item[0] == 0)
    # d loss / d query  ->  user-table gradient
    du = torch.zeros(U, d, dtype=torch.float64).index_add_(0, user, dq.cpu().double()); du[0] = 0
    gu = ref["d_user"].numpy()
    assert np.abs(du.numpy() - gu).max() <= 1e-5 * np.abs(gu).max()


def test_owner_without_any_owned_negative_and_wide_rows():
    """d = 256 (two 16-byte loads per lane), more owners than most queries have negatives: several owners hold
    nothing of a query (empty lists, m = -inf in the SampledSoftmax merge)."""
    N, U, d, G, n = 64, 9, 256, 10, 3
    w_item, w_user, user, pos, neg, lqp, lqn = _case(N, U, d, G, n, seed=3, pad=False)
    for loss_kind in (R.BPR, R.SSM):
        ref = R.training_step_aten(w_item, w_user, user, pos, neg, loss=loss_kind, scorer=R.IP)
        loss, d_item, dq = run_owners(w_item.to(DEV), w_user[user].to(DEV), pos.to(DEV), neg.to(DEV).int(), 8, loss_kind, R.IP)
        assert abs(loss - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
        gi = ref["d_item"].numpy()
        assert np.abs(d_item.cpu().numpy() - gi).max() <= 1e-5 * np.abs(gi).max()


def test_bad_ids_are_reported():
    from recstudio_b200 import _lib, sharded
    N, d, G, n = 50, 8, 4, 5
    w = torch.zeros(N, d, device=DEV)
    e = sharded.OwnerComputeCuda(N, 0, N, w, 1, 0, G, n)
    neg = torch.full((G, n), 3, dtype=torch.int32, device=DEV); neg[1, 2] = N + 4
    e.bind(torch.zeros(G, d, device=DEV), torch.ones(G, dtype=torch.int64, device=DEV), neg, R.BPR, R.IP)
    e.prep()
    with pytest.raises(_lib.Rsb200Error):
        e.check()


@pytest.mark.parametrize("loss_kind", [R.BPR, R.SSM])
def test_matches_the_single_table_fused_step(loss_kind):
    """Config-2-like shape scaled down (n = 1024, d = 128): 4 owners vs rsb200_pair_step on the whole table."""
    from recstudio_b200 import _lib, fused
    N, U, d, G, n = 200_001, 5_001, 128, 512, 1024
    gen = torch.Generator(device=DEV).manual_seed(1)
    w_item = torch.randn(N, d, device=DEV, generator=gen) * 0.1; w_item[0] = 0
    w_user = torch.randn(U, d, device=DEV, generator=gen) * 0.1; w_user[0] = 0
    user = torch.randint(1, U, (G,), device=DEV, generator=gen)
    pos = torch.randint(1, N, (G,), device=DEV, generator=gen)
    neg = torch.randint(1, N, (G, n), device=DEV, generator=gen).int()
    ws = fused.PairWorkspace(N, U, G, n, d, torch.device(DEV))
    loss_ref = fused.pair_step(ws, w_item, w_user, user, pos, neg, loss_kind, _lib.SCORE_IP).item()
    (ri, vi), (ru, vu) = fused.sparse_grads(ws)
    ref_item = torch.zeros(N, d, device=DEV); ref_item[ri] = vi
    ref_dq = ws.dq_buf.view(G, d).clone()
    loss, d_item, dq = run_owners(w_item, w_user[user], pos, neg, 4, loss_kind, _lib.SCORE_IP)
    assert abs(loss - loss_ref) <= 2e-6 * abs(loss_ref)
    assert (d_item - ref_item).abs().max().item() <= 1e-5 * ref_item.abs().max().item()
    assert (dq - ref_dq).abs().max().item() <= 1e-5 * ref_dq.abs().max().item()


@pytest.mark.parametrize("shape", [(24, 300), (48, 256), (1300, 256)])   # T % n != 0 (per-id Philox) | T = t n: shared Philox blocks, 1 and 2 rounds
@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("loss_kind", [R.BPR, R.SSM])
def test_owner_side_regeneration_of_the_uniform_draw(world, loss_kind, shape):
    """Instead of receiving every rank's negative ids, the owners recompute them from the ranks' generator states
    (rsb200_shard_args.regen_state): rank r's ids are exactly torch.randint(1, N, (B, n), device=cuda) for its
    (seed, offset).  The step must be bit-identical to the one fed with the explicitly drawn, gathered ids."""
    from recstudio_b200 import sampling, sharded
    N, U, d = 40_001, 301, 64
    B, n = shape
    G = world * B
    g = torch.Generator().manual_seed(world)
    w_item = (torch.randn(N, d, generator=g) * 0.3).to(DEV); w_item[0] = 0
    q_all = (torch.randn(G, d, generator=g) * 0.3).to(DEV)
    pos = torch.randint(1, N, (G,), generator=g).to(DEV)
    states, negs = [], []
    for r in range(world):                                 # what every rank would draw on its own generator
        torch.manual_seed(1000 + r)
        torch.rand(17 * (r + 1), device=DEV)                # ranks sit at different offsets
        gen = torch.cuda.default_generators[0]
        states.append([gen.initial_seed(), gen.get_offset()])
        want = torch.randint(1, N, (B, n), device=DEV)
        negs.append(want.int())
    neg_all = torch.cat(negs, 0)
    state = torch.tensor(states, dtype=torch.int64, device=DEV)
    per = sharded.rows_per_rank(N, world)
    outs = []
    for regen in (False, True):
        engines = []
        for r in range(world):
            row0 = r * per; local = max(0, min(per, N - row0))
            e = sharded.OwnerComputeCuda(N, row0, local, w_item[row0:row0 + local].contiguous(), world, r, G, n)
            if regen:
                e.bind(q_all, pos, None, loss_kind, R.IP, regen_state=state)
            else:
                e.bind(q_all, pos, neg_all, loss_kind, R.IP)
            engines.append(e)
        sp = torch.stack([e.prep().clone() for e in engines]).sum(0)
        for e in engines:
            e.sp[:G] = sp
        stats = torch.stack([e.fwd().clone() for e in engines])
        res = []
        for e in engines:
            e.stats_all[:, :G] = stats
        for e in engines:
            loss, dq = e.finish()
            rows, vals, totals = e.scatter()
            R_ = int(totals[1].item())
            res.append((loss.clone(), dq.clone(), rows[:R_].clone(), vals[:R_].clone(), e.ncount[:G].clone(), int(totals[0].item())))
            e.check()
        outs.append(res)
    for a, b in zip(*outs):
        assert torch.equal(a[4], b[4]) and a[5] == b[5]              # same owned negatives per query, same touch count
        assert torch.equal(a[2], b[2])                               # same gradient rows
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])   # loss and dq bit-identical (same compaction order)
        assert (a[3] - b[3]).abs().max().item() <= 1e-6 * max(a[3].abs().max().item(), 1e-30)   # entry order inside a row may differ
    # uniform_regen_state advances the generator exactly like the draw it replaces
    torch.manual_seed(5)
    ref_after = (torch.randint(1, N, (B, n), device=DEV), torch.rand(2, device=DEV))[1]
    torch.manual_seed(5)
    import torch.distributed as dist
    if not dist.is_initialized():
        import os
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29571")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    st = sharded.uniform_regen_state(DEV, B, n)
    assert st.shape == (1, 2) and torch.equal(torch.rand(2, device=DEV), ref_after)


def _popular_case(N, world, B, n, mode, seed):
    """every rank's PopularSamplerModel draw on its own generator state, with the single-GPU sampler (bit-equal to
    torch.rand + searchsorted, tests/test_gpu_sampler.py) -> (tables, states, gathered ids, gathered log Q)"""
    from recstudio_b200 import sampling, sharded
    rng = np.random.RandomState(seed)
    counts = np.floor(rng.zipf(1.3, size=N)).clip(max=1e6); counts[0] = 0
    table, prob = sharded.PopularSlice.tables(counts, mode)
    tab_d, prob_d = table.to(DEV), prob.to(DEV)
    guide, bits = sampling.build_guide(tab_d)
    states, negs, lqs = [], [], []
    for r in range(world):
        torch.manual_seed(500 + r)
        torch.rand(13 * (r + 1), device=DEV)
        gen = torch.cuda.default_generators[0]
        states.append([gen.initial_seed(), gen.get_offset()])
        _, neg32, lq = sampling.popular_draw(tab_d, prob_d, B, n, guide, bits, want_i64=False, want_i32=True)
        negs.append(neg32); lqs.append(lq)
    return table, prob, torch.tensor(states, dtype=torch.int64, device=DEV), torch.cat(negs, 0), torch.cat(lqs, 0)


@pytest.mark.parametrize("shape", [(24, 300), (48, 256), (1300, 256)])
@pytest.mark.parametrize("world", [1, 2, 5])
@pytest.mark.parametrize("loss_kind,mode", [(R.BPR, 0), (R.SSM, 2)])
def test_owner_side_regeneration_of_the_popularity_draw(world, loss_kind, mode, shape):
    """regen_kind 1: every owner regenerates every rank's PopularSamplerModel draw from the generator states and ITS slice of
    the cumulative table (PopularSlice): same owned negatives, same log Q, bit-identical loss / dq as the step fed with
    the gathered ids and log-probabilities; SampledSoftmax takes log Q(pos) from the positive's owner."""
    from recstudio_b200 import sampling, sharded
    N, d = 30_011, 64
    B, n = shape
    G = world * B
    table, prob, state, neg_all, lq_all = _popular_case(N, world, B, n, mode, seed=world + n)
    assert int(neg_all.min()) >= 0 and int(neg_all.max()) < N
    g = torch.Generator().manual_seed(7)
    w_item = (torch.randn(N, d, generator=g) * 0.3).to(DEV); w_item[0] = 0
    q_all = (torch.randn(G, d, generator=g) * 0.3).to(DEV)
    pos = torch.randint(1, N, (G,), generator=g).to(DEV)
    lqp = sampling.popular_logq(prob.to(DEV), pos)
    ssm = loss_kind == R.SSM
    per = sharded.rows_per_rank(N, world)
    outs = []
    for regen in (False, True):
        engines = []
        for r in range(world):
            row0 = r * per; local = max(0, min(per, N - row0))
            e = sharded.OwnerComputeCuda(N, row0, local, w_item[row0:row0 + local].contiguous(), world, r, G, n, with_logq=True)
            if regen:
                e.bind(q_all, pos, None, loss_kind, R.IP, regen_state=state, pop=sharded.PopularSlice(table, prob, row0, local, DEV))
            else:
                e.bind(q_all, pos, neg_all, loss_kind, R.IP, lqp if ssm else None, lq_all if ssm else None)
            engines.append(e)
        sp = torch.stack([e.prep().clone() for e in engines]).sum(0)
        for e in engines:
            if regen and ssm:
                e.sp2[:, :G] = sp
            else:
                e.sp[:G] = sp
        if regen and ssm:                                    # log Q(pos) assembled from the owners == the global lookup
            assert torch.equal(sp[1], lqp)
        stats = torch.stack([e.fwd().clone() for e in engines])
        for e in engines:
            e.stats_all[:, :G] = stats
        res = []
        for e in engines:
            loss, dq = e.finish()
            rows, vals, totals = e.scatter()
            R_ = int(totals[1].item())
            res.append((loss.clone(), dq.clone(), rows[:R_].clone(), vals[:R_].clone(), e.ncount[:G].clone(), int(totals[0].item()),
                        e.neg_c.clone(), e.lq_c.clone()))
            e.check()
        outs.append(res)
    for a, b in zip(*outs):
        assert torch.equal(a[4], b[4]) and a[5] == b[5]
        cnt = a[4].cpu().numpy()
        for q in (0, G // 2, G - 1):                         # compacted local ids / log Q of a few queries, element by element
            sl = slice(q * n, q * n + int(cnt[q]))
            assert torch.equal(a[6][sl], b[6][sl])
            if ssm:
                assert torch.equal(a[7][sl], b[7][sl])
        assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])       # gradient rows: bit-identical (deterministic bins)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert sum(r[5] for r in outs[1]) == int((neg_all > 0).sum()) + G   # every non-padding touch has exactly one owner


def test_draw_states_track_the_generator_on_the_device():
    """sharded.DrawStates: the [world, 2] (seed, offset) tensor advanced on the device step by step equals what
    uniform_regen_state would gather each step, and the local torch generator ends where the replaced draws leave it."""
    import torch.distributed as dist
    from recstudio_b200 import _lib, sharded
    if not dist.is_initialized():
        import os
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29571")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    B, n, N = 48, 256, 5001
    torch.manual_seed(77)
    want = []
    for _ in range(3):
        gen = torch.cuda.default_generators[0]
        want.append((gen.initial_seed(), gen.get_offset()))
        torch.randint(1, N, (B, n), device=DEV)
    after = torch.rand(3, device=DEV)
    torch.manual_seed(77)
    ds = sharded.DrawStates(DEV, B, n)
    got = [ds.next().clone().cpu().tolist()[0] for _ in range(3)]
    assert [tuple(g) for g in got] == want
    assert torch.equal(torch.rand(3, device=DEV), after)
    with pytest.raises(_lib.Rsb200Error):                # the rand above consumed the generator behind its back
        ds.next()
    ds.resync()
    assert tuple(ds.next().cpu().tolist()[0]) == (torch.cuda.default_generators[0].initial_seed(),
                                                  torch.cuda.default_generators[0].get_offset() - ds.inc)
