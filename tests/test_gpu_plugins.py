"""GPU parity of the standalone plugins and of FusedRetriever.training_step (all three
gradient hand-off modes) against the CPU oracle / golden vectors of the reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import retriever as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


def _close(got, want, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got - want).max() <= rtol * scale, (np.abs(got - want).max(), scale)


def test_embedding_forward_backward():
    from recstudio_b200 import plugins
    torch.manual_seed(0)
    emb = plugins.FusedEmbedding(500, 64).to(DEV)
    with torch.no_grad():
        emb.weight[0] = 0
    for shape in ((7,), (5, 9), (3, 4, 6)):
        ids = torch.randint(0, 500, shape, device=DEV)
        out = emb(ids)
        assert out.shape == shape + (64,)
        assert torch.equal(out, F.embedding(ids, emb.weight))            # a gather is exact
        g = torch.randn_like(out)
        emb.weight.grad = None
        out.backward(g)
        w2 = emb.weight.detach().clone().requires_grad_(True)
        F.embedding(ids, w2, padding_idx=0).backward(g)
        _close(emb.weight.grad.cpu(), w2.grad.cpu())
        assert float(emb.weight.grad[0].abs().sum()) == 0.0                 # padding row frozen (E2)


@pytest.mark.parametrize("kind", ["ip", "eu"])
def test_scorers_all_shapes(kind):
    from recstudio_b200 import plugins
    sc = plugins.FusedInnerProductScorer() if kind == "ip" else plugins.FusedEuclideanScorer()
    ok = R.IP if kind == "ip" else R.EUCLID
    g = torch.Generator().manual_seed(3)
    B, n, N, d, Lq = 6, 5, 11, 32, 4
    shapes = [((B, d), (B, d)), ((B, d), (B, n, d)), ((B, d), (N, d)), ((B, Lq, d), (B, Lq, d)), ((B, Lq, d), (B, Lq, n, d))]
    for qs, its in shapes:
        q = torch.randn(qs, generator=g); it = torch.randn(its, generator=g)
        want = R.score(ok, q, it)
        qd = q.to(DEV).requires_grad_(True); itd = it.to(DEV).requires_grad_(True)
        got = sc(qd, itd)
        assert got.shape == want.shape
        _close(got.detach().cpu(), want)
        go = torch.randn(want.shape, generator=g)
        qc = q.clone().requires_grad_(True); ic = it.clone().requires_grad_(True)
        R.score(ok, qc, ic).backward(go)
        got.backward(go.to(DEV))
        _close(qd.grad.cpu(), qc.grad); _close(itd.grad.cpu(), ic.grad)
    # Appendix A known answers
    ga = load_golden("appendix_a")
    got = sc(torch.from_numpy(ga["q"]).to(DEV), torch.from_numpy(ga["vn"]).to(DEV))
    _close(got.cpu(), ga[f"{kind}_neg"])


def test_losses_standalone():
    from recstudio_b200 import plugins
    ga = load_golden("appendix_a")
    for kind in ("ip", "eu"):
        ps = torch.from_numpy(ga[f"{kind}_pos"]).to(DEV); ns = torch.from_numpy(ga[f"{kind}_neg"]).to(DEV)
        lqp = torch.from_numpy(ga["lqp"]).to(DEV); lqn = torch.from_numpy(ga["lqn"]).to(DEV)
        assert abs(plugins.FusedBPRLoss()(None, ps, None, ns, None).item() - ga[f"{kind}_bpr"].item()) < 2e-6
        z = torch.zeros_like(ps, dtype=torch.int64); zn = torch.zeros_like(ns, dtype=torch.int64)   # UniformSampler's int64 zeros
        assert abs(plugins.FusedSampledSoftmaxLoss()(None, ps, z, ns, zn).item() - ga[f"{kind}_ssm0"].item()) < 5e-6
        assert abs(plugins.FusedSampledSoftmaxLoss()(None, ps, lqp, ns, lqn).item() - ga[f"{kind}_ssmq"].item()) < 5e-6
    g = torch.Generator().manual_seed(5)
    ps = torch.randn(37, generator=g) * 2; ns = torch.randn(37, 129, generator=g) * 2
    lqp = torch.randn(37, generator=g); lqn = torch.randn(37, 129, generator=g)
    for cls, ref in ((plugins.FusedBPRLoss, lambda a, b: R.bpr_loss(a, b)),
                     (plugins.FusedSampledSoftmaxLoss, lambda a, b: R.sampled_softmax_loss(a, lqp, b, lqn))):
        a = ps.clone().requires_grad_(True); b = ns.clone().requires_grad_(True)
        want = ref(a, b); want.backward()
        ad = ps.to(DEV).requires_grad_(True); bd = ns.to(DEV).requires_grad_(True)
        got = cls()(None, ad, lqp.to(DEV), bd, lqn.to(DEV)); (got * 1.0).backward()
        assert abs(got.item() - want.item()) <= RTOL * abs(want.item())
        _close(ad.grad.cpu(), a.grad); _close(bd.grad.cpu(), b.grad)


def test_sampler_plugins_contract():
    from recstudio_b200 import plugins
    s = plugins.FusedUniformSampler(1000).to(DEV)
    q = torch.zeros(5, 7, 8, device=DEV)
    torch.manual_seed(4)
    want = torch.randint(1, 1000, (35, 9), device=DEV).reshape(5, 7, 9)
    torch.manual_seed(4)
    neg, lp = s(q, 9)
    assert torch.equal(neg, want) and lp.dtype == torch.int64 and not lp.any()
    torch.manual_seed(4)
    pos = torch.tensor([3, 4], device=DEV)
    lpp, neg2, lnn = s(2, 4, pos_items=pos)
    assert lpp.dtype == torch.int64 and neg2.shape == (2, 4) and lnn.shape == (2, 4)
    g = load_golden("popular")
    ps = plugins.FusedPopularSampler(g["big_count"], mode=0).to(DEV)
    torch.manual_seed(8)
    seeds = torch.rand(6 * 3, 11, device=DEV)
    want = torch.searchsorted(ps.table, seeds).reshape(6, 3, 11)
    pos_items = torch.randint(0, 5000, (6, 3), device=DEV)
    torch.manual_seed(8)
    lp, neg, ln = ps(torch.zeros(6, 3, 4, device=DEV), 11, pos_items=pos_items)
    assert torch.equal(neg, want) and torch.equal(ln, torch.log(ps.pop_prob[want])) and lp.shape == (6, 3)


CASES = [("bpr", "ip", "uniform"), ("ssm", "ip", "uniform"), ("bpr", "eu", "uniform"), ("ssm", "eu", "popular"),
         ("ssm", "ip", "popular")]


@pytest.mark.parametrize("loss,scorer,sampler", CASES)
@pytest.mark.parametrize("mode", ["dense", "sparse", "rows"])
def test_fused_retriever_training_step(loss, scorer, sampler, mode):
    from recstudio_b200 import retriever
    U, N, d, B, n = 300, 5000, 64, 48, 300
    g = load_golden("popular")
    m = retriever.build_synthetic(U, N, d, n, loss=loss, scorer=scorer, sampler=sampler, pop_count=g["big_count"],
                                  fused_grad=mode, device=DEV, init_std=0.3)
    gen = torch.Generator().manual_seed(1)
    batch = {"user_id": torch.randint(1, U, (B,), generator=gen), "item_id": torch.randint(1, N, (B,), generator=gen),
             "rating": torch.ones(B)}                                 # host batch, as the trainer hands it over
    batch = m._to_device(batch, DEV)
    torch.manual_seed(77)
    out = m.training_step(batch)
    assert out.dim() == 0 and out.requires_grad
    out.backward()
    neg = m.fused_last_neg_id().long().cpu()
    # the negatives are what the reference sampler would have drawn from the same generator state
    torch.manual_seed(77)
    if sampler == "uniform":
        assert torch.equal(neg, torch.randint(1, N, (B, n), device=DEV).cpu())
        lqp = lqn = None
    else:
        want = torch.searchsorted(m.sampler.table, torch.rand(B, n, device=DEV))
        assert torch.equal(neg, want.cpu())
        lqn = torch.log(m.sampler.pop_prob[want]).cpu(); lqp = torch.log(m.sampler.pop_prob[batch["item_id"]]).cpu()
    wi, wu = m.item_encoder.weight.detach().cpu(), m.query_encoder.weight.detach().cpu()
    ref = R.training_step_aten(wi, wu, batch["user_id"].cpu(), batch["item_id"].cpu(), neg,
                               loss=R.SSM if loss == "ssm" else R.BPR, scorer=R.EUCLID if scorer == "eu" else R.IP,
                               log_pos_prob=lqp, log_neg_prob=lqn)
    assert abs(out.item() - ref["loss"].item()) <= RTOL * abs(ref["loss"].item())
    if mode == "dense":
        gi, gu = m.item_encoder.weight.grad, m.query_encoder.weight.grad
        assert gi.layout == torch.strided
    elif mode == "sparse":
        gi, gu = m.item_encoder.weight.grad, m.query_encoder.weight.grad
        assert gi.is_sparse and gi.is_coalesced()
        gi, gu = gi.to_dense(), gu.to_dense()
    else:
        assert m.item_encoder.weight.grad is None
        ws = next(iter(m._fused_ws_cache.values()))
        ri, vi, ru, vu, tot = ws.row_grads
        t = tot.tolist()
        gi = torch.zeros_like(m.item_encoder.weight); gi[ri[:t[1]]] = vi[:t[1]]
        gu = torch.zeros_like(m.query_encoder.weight); gu[ru[:t[3]]] = vu[:t[3]]
    _close(gi.cpu(), ref["d_item"]); _close(gu.cpu(), ref["d_user"])


def test_fused_retriever_falls_through_for_unknown_combinations():
    """A loss the kernels do not implement => the reference-shaped forward on standalone plugins."""
    from recstudio_b200 import iface, plugins, retriever

    class HingeLike(iface.PairwiseLoss):
        def forward(self, label, pos_score, log_pos_prob, neg_score, log_neg_prob):
            return torch.relu(1.0 - pos_score.unsqueeze(-1) + neg_score).mean()

    U, N, d, B, n = 50, 400, 32, 16, 10
    m = retriever.build_synthetic(U, N, d, n, device=DEV, init_std=0.4)
    m.loss_fn = HingeLike()
    gen = torch.Generator().manual_seed(2)
    batch = {"user_id": torch.randint(1, U, (B,), generator=gen).to(DEV), "item_id": torch.randint(1, N, (B,), generator=gen).to(DEV),
             "rating": torch.ones(B, device=DEV)}
    torch.manual_seed(5)
    loss = m.training_step(batch); loss.backward()
    torch.manual_seed(5)
    neg = torch.randint(1, N, (B, n), device=DEV).cpu()
    wi = m.item_encoder.weight.detach().cpu().requires_grad_(True); wu = m.query_encoder.weight.detach().cpu().requires_grad_(True)
    q = F.embedding(batch["user_id"].cpu(), wu, padding_idx=0)
    ps = R.inner_product_score(q, F.embedding(batch["item_id"].cpu(), wi, padding_idx=0))
    ns = R.inner_product_score(q, F.embedding(neg, wi, padding_idx=0))
    want = torch.relu(1.0 - ps.unsqueeze(-1) + ns).mean(); want.backward()
    assert abs(loss.item() - want.item()) <= RTOL * abs(want.item())
    _close(m.item_encoder.weight.grad.cpu(), wi.grad); _close(m.query_encoder.weight.grad.cpu(), wu.grad)


def test_optimizer_steps_on_both_gradient_layouts():
    """dense grads feed the reference's default dense Adam; sparse grads feed SGD/SparseAdam."""
    from recstudio_b200 import retriever
    for mode, opt_cls in (("dense", torch.optim.Adam), ("sparse", torch.optim.SparseAdam), ("sparse", torch.optim.SGD)):
        m = retriever.build_synthetic(100, 1000, 32, 20, fused_grad=mode, device=DEV, init_std=0.1)
        opt = opt_cls(m.parameters(), lr=0.05)
        batch = {"user_id": torch.randint(1, 100, (64,), device=DEV), "item_id": torch.randint(1, 1000, (64,), device=DEV),
                 "rating": torch.ones(64, device=DEV)}
        losses = []
        for _ in range(30):
            opt.zero_grad()
            loss = m.training_step(batch); loss.backward(); opt.step()
            losses.append(loss.item())
        margin = 0.05 if opt_cls is not torch.optim.SGD else 1e-3          # plain SGD at lr 0.05 moves slowly
        assert losses[-1] < losses[0] - margin, (mode, opt_cls.__name__, losses[0], losses[-1])
