"""GPU parity of SURVEY 8(f)-4: k-means, construct_index and the MIDX / Cluster samplers' index build and item draw
(csrc/midx.cu behind recstudio_b200/midx.py) against the golden vectors of the unmodified reference
(tests/golden/midx.npz) and the CPU oracle (oracle/midx.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import midx as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _blobs(N, d, k, g, spread=0.4):
    cent = torch.randn(k, d, generator=g) * 3
    return cent[torch.randint(0, k, (N,), generator=g)] + torch.randn(N, d, generator=g) * spread


@pytest.mark.parametrize("N,d,K,chunked", [(5000, 64, 13, True), (3000, 32, 200, False), (1, 8, 1, False), (777, 128, 129, False),
                                           (129, 4, 3, True)])
def test_kmeans_assign_and_update_vs_oracle(N, d, K, chunked):
    from recstudio_b200 import midx
    g = torch.Generator().manual_seed(N + K)
    full = _blobs(N, 2 * d if chunked else d, max(K, 2), g)
    X = full[:, d:] if chunked else full                      # a column chunk: row stride 2d (torch.chunk view, sampler.py:275)
    C = X[torch.randperm(N, generator=g)[:K]].clone()
    Xd = full.to(DEV)[:, d:] if chunked else full.to(DEV)
    assign, loss = midx.kmeans_assign(Xd, C.to(DEV))
    dist = torch.sum(X * X, -1, keepdim=True) - 2 * (X @ C.T) + torch.sum(C * C, -1).unsqueeze(0)
    want = dist.argmin(-1)
    got = assign.cpu()
    # exact wherever the two best centres differ by more than fp32 evaluation noise
    top2 = torch.topk(dist, min(2, K), largest=False).values
    clear = torch.ones(N, dtype=torch.bool) if K == 1 else (top2[:, 1] - top2[:, 0]) > 1e-4 * dist.abs().max()
    assert torch.equal(got[clear], want[clear]) and clear.float().mean() > 0.95
    sse = torch.sum(torch.square(X - C[got])).item()
    assert abs(loss.item() - sse) <= 1e-5 * max(sse, 1e-12)
    sums, counts = midx.kmeans_update(Xd, assign, K)
    np.testing.assert_array_equal(counts.cpu().numpy(), torch.bincount(got, minlength=K).float().numpy())
    want_s = torch.zeros(K, d, dtype=torch.float64).index_add_(0, got, X.double())
    assert (sums.cpu().double() - want_s).abs().max().item() <= 1e-5 * max(want_s.abs().max().item(), 1e-12)


def test_kmeans_matches_reference_golden():
    from recstudio_b200 import midx
    g = load_golden("midx")
    X = torch.from_numpy(g["km_X"]).to(DEV)
    for it in (1, 4, 50):
        C, assign, _, loss = midx.kmeans(X, X[:7].clone(), max_iter=it)
        np.testing.assert_array_equal(assign.cpu().numpy(), g[f"km_assign_{it}"])
        np.testing.assert_allclose(C.cpu().numpy(), g[f"km_C_{it}"], rtol=1e-5, atol=1e-6)
        assert abs(loss - g[f"km_loss_{it}"].item()) <= 1e-5 * g[f"km_loss_{it}"].item()
    torch.manual_seed(3)                      # int K: the initial centres come from the CPU generator like the reference's
    C, assign, _, loss = midx.kmeans(X, 5, max_iter=6)
    np.testing.assert_array_equal(assign.cpu().numpy(), g["km_assignr"])
    np.testing.assert_allclose(C.cpu().numpy(), g["km_Cr"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(torch.rand(2).numpy(), g["km_next_rand"])     # and leave it in the same state


@pytest.mark.parametrize("N,nb", [(1000, 37), (1, 1), (2049, 256), (300000, 257), (1_000_003, 4096), (500000, 65536), (4096, 5)])
def test_construct_index_is_a_stable_sort(N, nb):
    from recstudio_b200 import midx
    g = torch.Generator().manual_seed(N)
    codes = torch.randint(0, nb, (N,), generator=g)
    if N > 10:
        codes[: N // 3] = codes[0]                                           # one heavy bucket
    ind, ptr_ = midx.construct_index(codes.to(DEV), nb)
    w_ind, w_ptr = M.construct_index(codes.numpy(), nb)
    np.testing.assert_array_equal(ind.cpu().numpy(), w_ind)
    np.testing.assert_array_equal(ptr_.cpu().numpy(), w_ptr)
    if N == 1000:
        gg = load_golden("midx")
        ind, ptr_ = midx.construct_index(torch.from_numpy(gg["ci_codes"]).to(DEV), 37)
        np.testing.assert_array_equal(ind.cpu().numpy(), gg["ci_indices"]); np.testing.assert_array_equal(ptr_.cpu().numpy(), gg["ci_indptr"])


def test_segment_cdf_and_search_vs_oracle():
    from recstudio_b200 import midx
    g = torch.Generator().manual_seed(8)
    N, nb = 20000, 50
    codes = torch.randint(0, nb, (N,), generator=g); codes[codes == 7] = 8           # bucket 7 empty
    w = torch.rand(N, generator=g) + 0.01
    ind, ptr_ = M.construct_index(codes.numpy(), nb)
    cp, total = midx.segment_cdf(w.to(DEV), torch.from_numpy(ind).to(DEV), torch.from_numpy(ptr_).to(DEV))
    w_cp, w_tot = M.bucket_cdf(w.numpy(), ind, ptr_)
    np.testing.assert_allclose(cp.cpu().numpy(), w_cp, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(total.cpu().numpy(), w_tot, rtol=2e-5)
    k01 = torch.randint(0, nb, (64, 33), generator=g); k01[k01 == 7] = 9
    u = torch.rand(64, 33, generator=g)
    p01 = torch.randn(64, 33, generator=g)
    p = torch.cat([torch.ones(1), w])
    neg, logp = midx.segment_search(k01.to(DEV), u.to(DEV), cp, torch.from_numpy(ind).to(DEV), torch.from_numpy(ptr_).to(DEV), p.to(DEV))
    w_neg, w_prob = M.sample_item_with_pop(k01.numpy(), p01.numpy(), u.numpy(), cp.cpu().numpy(), ind, ptr_, p.numpy())
    np.testing.assert_array_equal(neg.cpu().numpy(), w_neg)
    np.testing.assert_allclose((p01 + logp.cpu()).numpy(), w_prob, rtol=1e-5, atol=1e-6)


def _make(tag, g):
    from recstudio_b200 import midx, plugins
    N, K = g["mx_emb"].shape[0], 4
    pop = torch.from_numpy(g["mx_pop"])
    return {"mu_ip": lambda: midx.FusedMIDXSamplerUniform(N + 1, K, plugins.FusedInnerProductScorer()),
            "mu_eu": lambda: midx.FusedMIDXSamplerUniform(N + 1, K, plugins.FusedEuclideanScorer()),
            "mp_ip": lambda: midx.FusedMIDXSamplerPop(pop, K, plugins.FusedInnerProductScorer(), mode=1),
            "cl_ip": lambda: midx.FusedClusterSamplerUniform(N + 1, 2 * K, plugins.FusedInnerProductScorer()),
            "cp_ip": lambda: midx.FusedClusterSamplerPop(pop, 2 * K, plugins.FusedInnerProductScorer(), mode=2)}[tag]().to(DEV)


@pytest.mark.parametrize("tag", ["mu_ip", "mu_eu", "mp_ip", "cl_ip", "cp_ip"])
def test_sampler_update_and_draw_match_reference_golden(tag, monkeypatch):
    """Sampler.update() (k-means on the CPU generator's seeds, index build, wkk / p / cp), sample_item for the
    reference's recorded seeds, compute_item_p -- all five model-based samplers."""
    g = load_golden("midx")
    smp = _make(tag, g)
    emb = torch.from_numpy(g["mx_emb"]).to(DEV)
    torch.manual_seed(17)
    smp.update(emb, max_iter=8)
    for name in ("cd0", "cd1", "cd", "indices", "indptr"):
        if f"{tag}_{name}" in g:
            np.testing.assert_array_equal(getattr(smp, name).cpu().numpy(), g[f"{tag}_{name}"])
    for name in ("c0", "c1", "c", "wkk", "p", "cp"):
        if f"{tag}_{name}" in g:
            np.testing.assert_allclose(getattr(smp, name).detach().cpu().numpy(), g[f"{tag}_{name}"], rtol=2e-5, atol=1e-6)
    assert hasattr(smp, "cp") == (f"{tag}_cp" in g)
    monkeypatch.setattr(torch, "rand_like", lambda *a, **k: torch.from_numpy(g[f"{tag}_u"]).to(DEV))
    neg, prob = smp.sample_item(torch.from_numpy(g[f"{tag}_k01"]).to(DEV), torch.from_numpy(g[f"{tag}_p01"]).to(DEV))
    np.testing.assert_array_equal(neg.cpu().numpy(), g[f"{tag}_neg"])
    np.testing.assert_allclose(prob.cpu().numpy(), g[f"{tag}_negprob"], rtol=1e-5, atol=1e-6)
    query = torch.from_numpy(g["mx_query"]).to(DEV)
    for key in ("pos1", "pos2"):
        if tag == "cp_ip" and key == "pos1":
            continue        # the reference broadcasts [B] + [B,1] -> [B,B] here (sampler.py:486-490) and then fails in view_as
        got = smp.compute_item_p(query, torch.from_numpy(g[f"mx_{key}"]).to(DEV))
        # atol 1e-5: the centres come out of k-means with shared-memory fp32 atomics (summation order varies run to run by
        # ~1e-6 relative) and these log-probabilities are differences of O(10) scores
        np.testing.assert_allclose(got.cpu().numpy(), g[f"{tag}_{key}_p"], rtol=1e-5, atol=1e-5)
    if tag == "mu_ip":      # the next epoch's update warm-starts from the stored centres
        smp.update(torch.from_numpy(g["mu_ip2_emb"]).to(DEV), max_iter=3)
        np.testing.assert_array_equal(smp.indices.cpu().numpy(), g["mu_ip2_indices"])
        np.testing.assert_allclose(smp.wkk.cpu().numpy(), g["mu_ip2_wkk"])


@pytest.mark.parametrize("tag", ["mu_ip", "mp_ip", "cl_ip"])
def test_sampler_forward_and_fused_step(tag):
    """forward(): ids are drawn from the chosen buckets (valid ids, finite proposal log-probs, the same shapes as the
    reference), and a FusedRetriever with this sampler takes the fused step on the drawn ids."""
    from recstudio_b200 import retriever
    from oracle import retriever as R
    g = load_golden("midx")
    N, d = g["mx_emb"].shape
    smp = _make(tag, g)
    m = retriever.build_synthetic(30, N + 1, d, 12, loss="ssm", scorer="ip", device=DEV)
    with torch.no_grad():
        m.item_encoder.weight[1:] = torch.from_numpy(g["mx_emb"]).to(DEV)
    m.sampler = smp
    m._update_item_vector()
    torch.manual_seed(1)
    smp.update(m.item_vector, max_iter=5)
    q = torch.randn(9, d, device=DEV)
    lp, neg, ln = smp(q, 12, pos_items=torch.arange(1, 10, device=DEV))
    assert neg.shape == (9, 12) and ln.shape == (9, 12) and lp.shape == (9,)
    lo = 0 if tag == "mp_ip" else 1           # the popularity variant returns 0-based positions (sampler.py:362)
    assert int(neg.min()) >= lo and int(neg.max()) <= N and torch.isfinite(ln).all()
    B = 16
    gen = torch.Generator().manual_seed(2)
    batch = {"user_id": torch.randint(1, 30, (B,), generator=gen), "item_id": torch.randint(1, N + 1, (B,), generator=gen),
             "rating": torch.ones(B)}
    torch.manual_seed(5)
    loss = m.training_step(batch)
    assert type(loss.grad_fn).__name__.startswith("_FusedStepFn")
    loss.backward()
    neg_used = m.fused_last_neg_id().cpu().long()
    torch.manual_seed(5)
    qv = m.query_encoder(batch["user_id"].to(DEV))
    lpp, neg2, lnp = smp(qv, 12, pos_items=batch["item_id"].to(DEV))
    assert torch.equal(neg2.cpu(), neg_used)
    ref = R.training_step_aten(m.item_encoder.weight.detach().cpu(), m.query_encoder.weight.detach().cpu(), batch["user_id"],
                               batch["item_id"], neg_used, loss=R.SSM, scorer=R.IP, log_pos_prob=lpp.cpu(), log_neg_prob=lnp.cpu())
    assert abs(loss.item() - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
    gi = m.item_encoder.weight.grad.cpu().numpy()
    assert np.abs(gi - ref["d_item"].numpy()).max() <= 1e-5 * np.abs(ref["d_item"].numpy()).max()
