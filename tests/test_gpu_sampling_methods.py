"""GPU parity of SURVEY 8(f)-2: MaskedUniformSampler (bit-exact ids on the CUDA generator's stream) and the
sampling methods dns / sir / toprand / top&rand / brute of BaseRetriever.sampling running on the CUDA ops
and feeding the fused step -- against the CPU oracle and the golden vectors of the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import retriever as R, samplers as S, topk_eval as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


def _hist(B, H, N, g, dup=False, left=False):
    h = torch.zeros(B, H, dtype=torch.int64)
    for b in range(B):
        c = int(torch.randint(0, H + 1, (1,), generator=g))
        h[b, :c] = torch.randperm(N - 1, generator=g)[:c] + 1
    if dup and H > 1:
        h[:, 1] = h[:, 0]
    if left:
        h = torch.flip(h, dims=[1])
    return h


@pytest.mark.parametrize("case", [
    # B, H, N (rows incl. padding), per_user, dup, left
    (12, 9, 401, 7, False, False),
    (12, 9, 401, 7, True, False),          # duplicate history items: non-monotone adjusted row
    (33, 1, 50, 5, False, False),          # H = 1
    (7, 100, 102, 300, False, True),       # almost every item excluded, left padding, H not a power of two
    (5, 4096, 100001, 64, False, False),   # maximum supported history width
    (2048, 50, 1_000_001, 128, False, False),
])
def test_masked_uniform_bit_exact(case):
    from recstudio_b200 import sampling
    B, H, N, per_user, dup, left = case
    g = torch.Generator().manual_seed(B + H)
    hist = _hist(B, H, N, g, dup, left)
    torch.manual_seed(77)
    seeds = torch.rand(B, per_user, device=DEV)                    # what the reference would draw (sampler.py:130)
    after = torch.rand(3, device=DEV)
    torch.manual_seed(77)
    neg64, neg32 = sampling.masked_uniform_draw(N, hist.to(DEV), per_user, want_i64=True, want_i32=True)
    assert torch.equal(torch.rand(3, device=DEV), after)           # generator advanced exactly like torch.rand
    want = S.masked_uniform_from_seeds(N - 1, hist.numpy(), seeds.cpu().numpy())
    np.testing.assert_array_equal(neg64.cpu().numpy(), want)
    np.testing.assert_array_equal(neg32.cpu().numpy(), want.astype(np.int32))
    if not dup:
        hs = [set(r[r > 0].tolist()) for r in hist.numpy()]
        assert all(not (set(want[b].tolist()) & hs[b]) for b in range(B))
        assert want.min() >= 1 and want.max() <= N - 1


def test_masked_sampler_plugin_shapes_and_errors():
    from recstudio_b200 import _lib, plugins
    g = torch.Generator().manual_seed(1)
    hist = _hist(6, 5, 300, g).to(DEV)
    smp = plugins.FusedMaskedUniformSampler(300)
    lp, neg, ln = smp(torch.zeros(6, 8, device=DEV), 4, pos_items=torch.arange(6, device=DEV), user_hist=hist)
    assert neg.shape == (6, 4) and neg.dtype == torch.int64 and ln.dtype == torch.float32 and lp.shape == (6,)
    assert float(ln.abs().sum()) == 0 and float(lp.abs().sum()) == 0
    torch.manual_seed(3)
    neg3, _ = smp(torch.zeros(6, 3, 8, device=DEV), 4, user_hist=hist)
    torch.manual_seed(3)
    flat, _ = smp(torch.zeros(6, 8, device=DEV), 12, user_hist=hist)
    assert neg3.shape == (6, 3, 4) and torch.equal(neg3.reshape(6, 12), flat)
    with pytest.raises(ValueError):
        smp(torch.zeros(6, 8, device=DEV), 4)
    with pytest.raises(_lib.Rsb200Error):
        smp(torch.zeros(6, 8, device=DEV), 4, user_hist=hist.cpu())


def test_score_ids_streaming_kernel():
    from recstudio_b200 import plugins
    g = torch.Generator().manual_seed(5)
    for N, d, B, n in ((5000, 128, 37, 1000), (300, 64, 9, 33), (300, 100, 4, 8), (300, 256, 5, 40), (300, 32, 6, 3)):
        w = torch.randn(N, d, generator=g); q = torch.randn(B, d, generator=g)
        ids = torch.randint(0, N, (B, n), generator=g)
        for kind in (R.IP, R.EUCLID):
            got = plugins.score_ids(kind, q.to(DEV), w.to(DEV), ids.to(DEV)).cpu()
            want = R.score(kind, q, w[ids])
            assert (got - want).abs().max().item() <= RTOL * max(1.0, want.abs().max().item())


def _model(g, method, nc, loss, fused_grad="dense", sampler="uniform", excluding_hist=False):
    from recstudio_b200 import retriever
    U, N, d = g["w_user"].shape[0], g["w_item"].shape[0], g["w_item"].shape[1]
    m = retriever.build_synthetic(U, N, d, nc, loss=loss, scorer="ip", sampler=sampler, fused_grad=fused_grad, device=DEV,
                                  sampling_method=method, excluding_hist=excluding_hist)
    with torch.no_grad():
        m.item_encoder.weight.copy_(torch.from_numpy(g["w_item"]))
        m.query_encoder.weight.copy_(torch.from_numpy(g["w_user"]))
    m._update_item_vector()
    return m


def _close(got, want):
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    assert np.abs(got - want).max() <= RTOL * max(np.abs(want).max(), 1e-30), (np.abs(got - want).max(), np.abs(want).max())


@pytest.mark.parametrize("method", ["dns", "sir", "toprand", "top&rand", "brute"])
def test_sampling_methods_replay_reference_golden(method, monkeypatch):
    """The reference's recorded draws (pool ids, multinomial / randint outcomes) are injected in place of the
    device generator; everything else -- pool scoring, top-k candidates, selection, the fused step -- runs on
    the CUDA kernels and must reproduce the reference's negatives exactly and its loss / gradients to 1e-5."""
    from recstudio_b200 import sampling
    g = load_golden("sampling_" + method.replace("&", "and"))
    nc = g["negative_count"].tolist()
    loss = "ssm" if method in ("sir", "brute") else "bpr"
    m = _model(g, method, nc, loss)
    batch = {"user_id": torch.from_numpy(g["user"]), "item_id": torch.from_numpy(g["pos"]),
             "rating": torch.ones(len(g["user"])), "user_hist": torch.from_numpy(g["hist"])}
    if "rec_pool" in g:
        pool = torch.from_numpy(g["rec_pool"]).to(DEV)

        def forward(query, num_neg, pos_items=None, device=None):
            return torch.zeros_like(pos_items), pool, torch.zeros_like(pool)
        m.sampler.forward = forward
    if "rec_multinomial" in g:
        monkeypatch.setattr(torch, "multinomial", lambda *a, **k: torch.from_numpy(g["rec_multinomial"]).to(DEV))
    if method == "toprand":
        monkeypatch.setattr(torch, "randint", lambda *a, **k: torch.from_numpy(g["rec_randint"]).to(DEV))
    if method == "top&rand":
        monkeypatch.setattr(sampling, "uniform_draw", lambda *a, **k: (torch.from_numpy(g["rec_randint"]).to(DEV), None))
    loss_t = m.training_step(batch)
    assert loss_t.grad_fn is not None and type(loss_t.grad_fn).__name__.startswith("_FusedStepFn")
    loss_t.backward()
    np.testing.assert_array_equal(m.fused_last_neg_id().cpu().numpy(), g["neg"])
    assert abs(loss_t.item() - g["loss"].item()) <= RTOL * abs(g["loss"].item())
    _close(m.item_encoder.weight.grad.cpu(), g["d_item"])
    _close(m.query_encoder.weight.grad.cpu(), g["d_user"])


@pytest.mark.parametrize("method,loss,nc", [("dns", "bpr", [64, 8]), ("sir", "ssm", [64, 16]), ("toprand", "bpr", [32, 8]),
                                            ("top&rand", "ssm", 10), ("brute", "ssm", 7)])
def test_sampling_methods_live_vs_oracle(method, loss, nc):
    """Un-patched run on the CUDA generator: the oracle replays the same draws (torch's own CUDA randint /
    multinomial for the same seed) and must agree on the selected ids, the loss and the gradients."""
    g = torch.Generator().manual_seed(17)
    U, N, d, B, H = 40, 2001, 64, 48, 5
    w = {"w_item": (torch.randn(N, d, generator=g) * 0.3).numpy(), "w_user": (torch.randn(U, d, generator=g) * 0.3).numpy()}
    w["w_item"][0] = 0; w["w_user"][0] = 0
    m = _model(w, method, nc, loss)
    wi, wu = torch.from_numpy(w["w_item"]), torch.from_numpy(w["w_user"])
    hist = _hist(B, H, N, g)
    batch = {"user_id": torch.randint(1, U, (B,), generator=g), "item_id": torch.randint(1, N, (B,), generator=g),
             "rating": torch.ones(B), "user_hist": hist}
    n0, n1 = (nc, nc) if isinstance(nc, int) else nc
    query = wu[batch["user_id"]]
    lqp = lqn = None
    torch.manual_seed(99)
    loss_t = m.training_step(batch)
    loss_t.backward()
    got_neg = m.fused_last_neg_id().cpu()
    torch.manual_seed(99)                           # replay the reference's RNG call sequence with torch's own CUDA ops
    if method in ("dns", "sir"):
        pool = torch.randint(1, N, (B, n0), device=DEV).cpu()
        if method == "dns":
            sel = R.select_from_pool("dns", query, wi, pool, n1)
        else:
            dev_scores = R.score(R.IP, query, wi[pool]).to(DEV)
            res = torch.multinomial(torch.softmax(dev_scores + torch.finfo(torch.float32).eps, -1), n1, replacement=True).cpu()
            sel = R.select_from_pool("sir", query, wi, pool, n1, resampled_id=res)
            # multinomial consumes probabilities computed on the device: identical pools but fp32 softmax noise can move
            # a draw that sits exactly on a CDF boundary -- compare through the pool membership instead of equality
            assert all(set(got_neg[b].tolist()) <= set(pool[b].tolist()) for b in range(B))
            lqn = R.score(R.IP, query, wi[got_neg])
            lqp = R.score(R.IP, query, wi[batch["item_id"]])
        if method == "dns":
            np.testing.assert_array_equal(got_neg.numpy(), sel["neg_id"].numpy())
    elif method == "toprand":
        _, cand = T.topk_aten(query, wi[1:], n0, hist)
        ridx = torch.randint(0, n0, (B, n1), device=DEV).cpu()
        np.testing.assert_array_equal(got_neg.numpy(), torch.gather(cand, -1, ridx).numpy())
    elif method == "top&rand":
        _, cand = T.topk_aten(query, wi[1:], n1 // 2, hist)
        rnd = torch.randint(1, N, (B, n1 - n1 // 2), device=DEV).cpu()
        np.testing.assert_array_equal(got_neg.numpy(), torch.cat([cand, rnd], -1).numpy())
    else:
        prob = torch.nn.functional.pad(torch.softmax(R.score(R.IP, query, wi[1:]), -1), (1, 0))
        assert got_neg.min() >= 1                                      # the padding column has probability 0
        lqn = torch.log(torch.gather(prob, -1, got_neg))
        lqp = torch.log(torch.gather(prob, -1, batch["item_id"].view(-1, 1))).view(-1)
    ref = R.training_step_aten(wi, wu, batch["user_id"], batch["item_id"], got_neg, loss=R.SSM if loss == "ssm" else R.BPR,
                               scorer=R.IP, log_pos_prob=lqp, log_neg_prob=lqn)
    assert abs(loss_t.item() - ref["loss"].item()) <= 2 * RTOL * abs(ref["loss"].item())
    _close(m.item_encoder.weight.grad.cpu(), ref["d_item"].numpy())
    _close(m.query_encoder.weight.grad.cpu(), ref["d_user"].numpy())


@pytest.mark.parametrize("method,nc", [("none", 12), ("dns", [24, 6])])
def test_excluding_hist_masked_sampler_fused_step(method, nc):
    """excluding_hist=True + MaskedUniformSampler: the fused step draws history-free negatives on the
    generator's stream (ids equal the oracle's) and matches the oracle step on them."""
    g = torch.Generator().manual_seed(23)
    U, N, d, B, H = 30, 501, 32, 40, 7
    w = {"w_item": (torch.randn(N, d, generator=g) * 0.3).numpy(), "w_user": (torch.randn(U, d, generator=g) * 0.3).numpy()}
    w["w_item"][0] = 0; w["w_user"][0] = 0
    m = _model(w, method, nc, "bpr", sampler="masked", excluding_hist=True)
    hist = _hist(B, H, N, g)
    batch = {"user_id": torch.randint(1, U, (B,), generator=g), "item_id": torch.randint(1, N, (B,), generator=g),
             "rating": torch.ones(B), "user_hist": hist.to(DEV)}
    n0, n1 = (nc, nc) if isinstance(nc, int) else nc
    torch.manual_seed(5)
    seeds = torch.rand(B, n0, device=DEV).cpu().numpy()
    torch.manual_seed(5)
    loss_t = m.training_step(batch)
    loss_t.backward()
    got = m.fused_last_neg_id().cpu()
    pool = torch.from_numpy(S.masked_uniform_from_seeds(N - 1, hist.numpy(), seeds))
    wi, wu = torch.from_numpy(w["w_item"]), torch.from_numpy(w["w_user"])
    want = pool if method == "none" else R.select_from_pool("dns", wu[batch["user_id"]], wi, pool, n1)["neg_id"]
    np.testing.assert_array_equal(got.numpy(), want.numpy())
    hs = [set(r[r > 0].tolist()) for r in hist.numpy()]
    assert all(not (set(got[b].tolist()) & hs[b]) for b in range(B))
    ref = R.training_step_aten(wi, wu, batch["user_id"], batch["item_id"], got.long(), loss=R.BPR, scorer=R.IP)
    assert abs(loss_t.item() - ref["loss"].item()) <= RTOL * abs(ref["loss"].item())
    _close(m.item_encoder.weight.grad.cpu(), ref["d_item"].numpy())
