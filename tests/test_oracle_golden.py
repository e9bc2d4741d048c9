"""The CPU oracle against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and SURVEY.md Appendix A."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import philox, retriever as R, samplers as S, topk_eval as T
from conftest import GOLDEN, load_golden

STEP_FILES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "step_*_*_*.npz"))
                    if "full_softmax" not in p)


def _cfg(name):
    return (R.SSM if "_ssm_" in name else R.BPR), (R.EUCLID if name.endswith("_eu") else R.IP)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_philox_offset_and_mapping_properties():
    sm, mt = 148, 2048
    # counter_offset formula, DistributionTemplates.h:60
    assert philox.torch_cuda_counter_offset(1, sm, mt) == 4
    assert philox.torch_cuda_counter_offset(8192 * 1024, sm, mt) == ((8192 * 1024 - 1) // (256 * 1184 * 4) + 1) * 4
    # element li in round r of thread idx uses counter offset/4 + r: consecutive calls never overlap
    a = philox.torch_cuda_raw_u32(2022, 0, 5000, sm, mt)
    off = philox.torch_cuda_counter_offset(5000, sm, mt)
    b = philox.torch_cuda_raw_u32(2022, off, 5000, sm, mt)
    assert not np.array_equal(a, b)
    # a longer call shares its first-round words with a shorter one only if the grid is equal
    c = philox.torch_cuda_raw_u32(2022, 0, 256 * 1184 * 4 + 10, sm, mt)
    d = philox.torch_cuda_raw_u32(2022, 0, 256 * 1184 * 4, sm, mt)
    assert np.array_equal(c[:d.size], d)
    ids = philox.torch_cuda_randint(2022, 0, 1, 1000, 100000, sm, mt)
    assert ids.min() >= 1 and ids.max() <= 999 and ids.dtype == np.int64
    u = philox.torch_cuda_rand(2022, 0, 100000, sm, mt)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0


def test_appendix_a_scores_and_losses():
    g = load_golden("appendix_a")
    q, vp, vn = (torch.from_numpy(g[k]) for k in ("q", "vp", "vn"))
    lqp, lqn = torch.from_numpy(g["lqp"]), torch.from_numpy(g["lqn"])
    survey = {"ip": (0.9947767, 2.2288237, 2.8037701, 2.0457540), "eu": (1.9598691, 4.0924187, 4.6389341, 4.0754628)}
    for name, sc in (("ip", R.IP), ("eu", R.EUCLID)):
        ps, ns = R.score(sc, q, vp), R.score(sc, q, vn)
        np.testing.assert_array_equal(ps.numpy(), g[f"{name}_pos"])
        np.testing.assert_array_equal(ns.numpy(), g[f"{name}_neg"])
        vals = (R.bpr_loss(ps, ns), R.sampled_softmax_loss(ps, torch.zeros_like(ps), ns, torch.zeros_like(ns)),
                R.sampled_softmax_loss(ps, lqp, ns, lqn), R.softmax_loss(ps, ns))
        for v, key, s in zip(vals, ("bpr", "ssm0", "ssmq", "softmax"), survey[name]):
            assert v.item() == g[f"{name}_{key}"].item()
            assert abs(v.item() - s) < 2e-6


@pytest.mark.parametrize("name", STEP_FILES)
def test_training_step_aten_matches_reference_bitwise(name):
    g = load_golden(name)
    loss, scorer = _cfg(name)
    out = R.training_step_aten(torch.from_numpy(g["w_item"]), torch.from_numpy(g["w_user"]),
                               torch.from_numpy(g["user"]), torch.from_numpy(g["pos"]), torch.from_numpy(g["neg"]),
                               loss=loss, scorer=scorer,
                               log_pos_prob=torch.from_numpy(g["log_pos_prob"]),
                               log_neg_prob=torch.from_numpy(g["log_neg_prob"]))
    # same ATen ops in the same order on the same CPU => identical bits
    assert out["loss"].item() == g["loss"].item()
    np.testing.assert_array_equal(out["pos_score"].numpy(), g["pos_score"])
    np.testing.assert_array_equal(out["neg_score"].numpy(), g["neg_score"])
    np.testing.assert_array_equal(out["d_item"].numpy(), g["d_item"])
    np.testing.assert_array_equal(out["d_user"].numpy(), g["d_user"])
    assert np.all(g["d_item"][0] == 0)          # padding row never receives gradient (E2)


@pytest.mark.parametrize("name", STEP_FILES)
def test_closed_form_matches_reference(name):
    g = load_golden(name)
    loss, scorer = _cfg(name)
    cf = R.closed_form_step(g["w_item"], g["w_user"], g["user"], g["pos"], g["neg"], loss=loss, scorer=scorer,
                            log_pos_prob=g["log_pos_prob"], log_neg_prob=g["log_neg_prob"])
    assert abs(cf["loss"] - g["loss"].item()) <= 2e-6 * max(1.0, abs(g["loss"].item()))
    np.testing.assert_allclose(cf["pos_score"], g["pos_score"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cf["neg_score"], g["neg_score"], rtol=1e-5, atol=1e-5)
    di = R.dict_to_dense(cf["d_item"], g["d_item"].shape)
    du = R.dict_to_dense(cf["d_user"], g["d_user"].shape)
    scale_i, scale_u = np.abs(g["d_item"]).max(), np.abs(g["d_user"]).max()
    assert np.abs(di - g["d_item"]).max() <= 1e-5 * scale_i
    assert np.abs(du - g["d_user"]).max() <= 1e-5 * scale_u


def test_popular_tables_and_draw():
    g = load_golden("popular")
    for tag in ("small", "big"):
        for mode in (0, 1, 2):
            prob, table = S.popular_tables(g[f"{tag}_count"], mode)
            np.testing.assert_array_equal(prob, g[f"{tag}_m{mode}_prob"])
            np.testing.assert_array_equal(table, g[f"{tag}_m{mode}_table"])
    prob, table = S.popular_tables(g["small_count"], 0)
    # SURVEY Appendix A
    np.testing.assert_allclose(prob, [0.13756868, 0.24648999, 0.09535535, 0, 0.32987529, 1.0], atol=1e-7)
    np.testing.assert_allclose(table, [0.13756868, 0.38405865, 0.47941402, 0.47941402, 0.80928934, 1.0], atol=1e-7)
    idx, _ = S.popular_draw(table, prob, g["probe_seeds"])
    np.testing.assert_array_equal(idx, g["probe_idx"])
    np.testing.assert_array_equal(idx, [0, 1, 1, 4, 4, 5])
    prob, table = S.popular_tables(g["big_count"], 0)
    idx, logq = S.popular_draw(table, prob, g["draw_seeds"].reshape(-1))
    np.testing.assert_array_equal(idx.reshape(g["draw_idx"].shape), g["draw_idx"])
    np.testing.assert_array_equal(logq.reshape(g["draw_logq"].shape), g["draw_logq"])


def test_uniform_sampler_contract():
    g = load_golden("uniform_cpu")
    # reference returns int64 zeros for both log-prob tensors
    assert g["log_pos"].dtype == np.int64 and g["log_neg"].dtype == np.int64 and not g["log_neg"].any()
    assert g["neg"].min() >= 1 and g["neg"].max() <= 9
    assert g["neg2"].shape == (5, 7, 9)
    lp, neg, ln, off = S.uniform_sampler(2022, 0, 10, 2, 4, np.array([3, 4]), 148, 2048)
    assert lp.dtype == np.int64 and ln.dtype == np.int64 and neg.dtype == np.int64
    assert neg.shape == (2, 4) and neg.min() >= 1 and neg.max() <= 9 and off == 4


def test_topk_history_mask():
    g = load_golden("topk_eval")
    q = torch.from_numpy(g["a_w_user"])[torch.from_numpy(g["a_user"])]
    sc, ids = T.topk_aten(q, torch.from_numpy(g["a_w_item"])[1:], 3, torch.from_numpy(g["a_hist"]))
    np.testing.assert_array_equal(ids.numpy(), [[3, 4, 5], [5, 4, 3]])
    np.testing.assert_array_equal(ids.numpy(), g["a_ids"]); np.testing.assert_array_equal(sc.numpy(), g["a_score"])
    q = torch.from_numpy(g["r_w_user"])[torch.from_numpy(g["r_user"])]
    iv = torch.from_numpy(g["r_w_item"])[1:]
    hist = torch.from_numpy(g["r_hist"])
    for k, ks, ki in ((10, "r_score", "r_ids"), (100, "r_score100", "r_ids100")):
        sc, ids = T.topk_aten(q, iv, k, hist)
        np.testing.assert_array_equal(ids.numpy(), g[ki]); np.testing.assert_array_equal(sc.numpy(), g[ks])
        es, ei = T.topk_exact(q.numpy(), iv.numpy(), k, g["r_hist"])
        np.testing.assert_array_equal(ei, g[ki])
        np.testing.assert_allclose(es, g[ks], rtol=1e-5, atol=1e-5)
    sc, ids = T.topk_aten(q, iv, 10, None)
    np.testing.assert_array_equal(ids.numpy(), g["r_ids_nohist"])
    es, ei = T.topk_exact(q.numpy(), iv.numpy(), 10, None)
    np.testing.assert_array_equal(ei, g["r_ids_nohist"])
    # the masked history ids never appear
    for b in range(ids.shape[0]):
        assert not np.isin(g["r_ids"][b], g["r_hist"][b][g["r_hist"][b] > 0]).any()


def test_rank_metrics():
    g = load_golden("topk_eval")
    survey = {"recall": (0.1111111, 0.2222222, 0.5555556), "precision": (0.3333333, 0.2222222, 0.2),
              "ndcg": (0.3333333, 0.2346394, 0.3781982), "map": (0.3333333, 0.1851852, 0.2685185),
              "mrr": (0.3333333, 0.3333333, 0.4166667), "hit": (0.3333333, 0.3333333, 0.6666667)}
    for name, fn in T.METRICS.items():
        for i, k in enumerate((1, 3, 5)):
            v = fn(g["m_label"], g["m_target"], k)
            assert abs(v - g[f"m_{name}_{k}"].item()) < 1e-6, (name, k)
            assert abs(v - survey[name][i]) < 1e-6
    label = T.hit_matrix(g["r_ids100"], g["e_target"])
    for name, fn in T.METRICS.items():
        for k in (5, 10, 20):
            assert abs(fn(label, g["e_rating"], k) - g[f"e_{name}_at_{k}"].item()) < 1e-6, (name, k)


def test_full_softmax_step():
    g = load_golden("step_full_softmax")
    out = R.full_softmax_step_aten(torch.from_numpy(g["w_item"]), torch.from_numpy(g["w_user"]),
                                   torch.from_numpy(g["user"]), torch.from_numpy(g["pos"]))
    assert out["loss"].item() == g["loss"].item()
    np.testing.assert_array_equal(out["d_item"].numpy(), g["d_item"])
    np.testing.assert_array_equal(out["d_user"].numpy(), g["d_user"])


# ---------------------------------------------------------------- 8(f)-2: masked sampler + sampling methods
def test_masked_uniform_matches_reference():
    """uniform_sample_masked_hist / MaskedUniformSampler (sampler.py:117-147,187-214): the oracle reproduces
    the reference's ids bit for bit from the recorded torch.rand seeds -- distinct, duplicate (non-monotone
    adjusted row), left-padded and full histories, and the [B, n_q, n] multi-query shape."""
    g = load_golden("masked_uniform")
    for tag in "abcd":
        want = g[f"{tag}_neg"]
        got = S.masked_uniform_from_seeds(int(g[f"{tag}_num_items"]), g[f"{tag}_hist"], g[f"{tag}_seeds"])
        np.testing.assert_array_equal(got.reshape(want.shape), want)
        if tag in "acd":        # distinct histories: never a history item, always a real item
            for b in range(want.shape[0]):
                assert not set(want[b].ravel().tolist()) & set(g[f"{tag}_hist"][b][g[f"{tag}_hist"][b] > 0].tolist())
            assert want.min() >= 1 and want.max() <= int(g[f"{tag}_num_items"])
    got = S.masked_uniform_from_seeds(400, g["a_hist"], g["s_seeds"])
    np.testing.assert_array_equal(got, g["s_neg"])
    assert g["s_log_neg"].dtype == np.float32 and np.all(g["s_log_neg"] == 0) and g["s_log_pos"].dtype == np.float32


@pytest.mark.parametrize("method", ["dns", "sir", "toprand", "topandrand", "brute"])
def test_sampling_methods_match_reference(method):
    """BaseRetriever.sampling methods (baseretriever.py:280-355) inside training_step: the oracle re-derives
    the selected negatives / proposal log-probabilities from the recorded draws, and the step on those
    negatives reproduces the reference's loss and dense gradients."""
    g = load_golden("sampling_" + method)
    wi, wu = torch.from_numpy(g["w_item"]), torch.from_numpy(g["w_user"])
    user, pos, hist = (torch.from_numpy(g[k]) for k in ("user", "pos", "hist"))
    query = wu[user]
    nc = g["negative_count"].tolist()
    n0, n1 = (nc, nc) if isinstance(nc, int) else nc
    loss_kind = R.SSM if method in ("sir", "brute") else R.BPR
    if method in ("dns", "sir"):
        np.testing.assert_array_equal(g["rec_pool"], g["rec_randint"])       # the pool IS the sampler's randint draw
        sel = R.select_from_pool(method, query, wi, g["rec_pool"], n1, R.IP, resampled_id=g.get("rec_multinomial"))
        np.testing.assert_array_equal(sel["neg_id"].numpy(), g["neg"])
        np.testing.assert_array_equal(sel["log_neg_prob"].numpy(), g["log_neg_prob"])
        if method == "sir":
            np.testing.assert_array_equal(g["log_pos_prob"], g["pos_score"])  # :335 proposal log-prob of the positive = its score
    elif method == "toprand":
        _, cand = T.topk_aten(query, wi[1:], n0, hist)
        np.testing.assert_array_equal(torch.gather(cand, -1, torch.from_numpy(g["rec_randint"])).numpy(), g["neg"])
    elif method == "topandrand":
        _, cand = T.topk_aten(query, wi[1:], n1 // 2, hist)
        np.testing.assert_array_equal(torch.cat([cand, torch.from_numpy(g["rec_randint"])], -1).numpy(), g["neg"])
    else:                                                                     # brute: softmax over the catalog
        prob = torch.nn.functional.pad(torch.softmax(R.score(R.IP, query, wi[1:]), -1), (1, 0))
        np.testing.assert_array_equal(g["rec_multinomial"], g["neg"])
        np.testing.assert_allclose(torch.log(torch.gather(prob, -1, torch.from_numpy(g["neg"]))).numpy(), g["log_neg_prob"], rtol=1e-6)
        np.testing.assert_allclose(torch.log(torch.gather(prob, -1, pos.view(-1, 1))).view(-1).numpy(), g["log_pos_prob"], rtol=1e-6)
    ref = R.training_step_aten(wi, wu, user, pos, torch.from_numpy(g["neg"]), loss=loss_kind, scorer=R.IP,
                               log_pos_prob=torch.from_numpy(g["log_pos_prob"]), log_neg_prob=torch.from_numpy(g["log_neg_prob"]))
    np.testing.assert_allclose(ref["loss"].item(), g["loss"].item(), rtol=1e-6)
    np.testing.assert_allclose(ref["d_item"].numpy(), g["d_item"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(ref["d_user"].numpy(), g["d_user"], rtol=1e-5, atol=1e-9)


# ---------------------------------------------------------------- 8(f)-4: kmeans / construct_index / MIDX samplers
def test_midx_oracle_matches_reference():
    from oracle import midx as M
    g = load_golden("midx")
    X = torch.from_numpy(g["km_X"])
    for it in (1, 4, 50):
        C, assign, loss, _ = M.kmeans(X, X[:7].clone(), max_iter=it)
        np.testing.assert_array_equal(assign.numpy(), g[f"km_assign_{it}"])
        np.testing.assert_allclose(C.numpy(), g[f"km_C_{it}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(loss, g[f"km_loss_{it}"].item(), rtol=1e-6)
    torch.manual_seed(3)                                     # int K: randperm on the CPU generator, same stream afterwards
    C, assign, loss, _ = M.kmeans(X, 5, max_iter=6)
    np.testing.assert_array_equal(assign.numpy(), g["km_assignr"])
    np.testing.assert_allclose(C.numpy(), g["km_Cr"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(torch.rand(2).numpy(), g["km_next_rand"])
    ind, ptr_ = M.construct_index(g["ci_codes"], 37)
    np.testing.assert_array_equal(ind, g["ci_indices"]); np.testing.assert_array_equal(ptr_, g["ci_indptr"])

    emb, query = torch.from_numpy(g["mx_emb"]), torch.from_numpy(g["mx_query"])
    pos1, pos2 = torch.from_numpy(g["mx_pos1"]), torch.from_numpy(g["mx_pos2"])
    K = 4
    norms = {"mu_ip": None, "mu_eu": torch.exp(-0.5 * torch.sum(emb ** 2, -1)).numpy(),
             "mp_ip": (torch.log(torch.from_numpy(g["mx_pop"]) + 1) + 1e-6).numpy()}
    for tag, norm in norms.items():
        torch.manual_seed(17)
        b = M.midx_build(emb, K, None, None, 8, norm)
        for name in ("cd0", "cd1"):
            np.testing.assert_array_equal(b[name].numpy(), g[f"{tag}_{name}"])
        np.testing.assert_array_equal(b["indices"], g[f"{tag}_indices"]); np.testing.assert_array_equal(b["indptr"], g[f"{tag}_indptr"])
        np.testing.assert_allclose(b["c0"].numpy(), g[f"{tag}_c0"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(b["wkk"], g[f"{tag}_wkk"], rtol=1e-5)
        if norm is None:
            neg, prob = M.sample_item_uniform(g[f"{tag}_k01"], g[f"{tag}_p01"], g[f"{tag}_u"], b["indices"], b["indptr"])
        else:
            np.testing.assert_allclose(b["cp"], g[f"{tag}_cp"], rtol=1e-5)
            np.testing.assert_allclose(b["p"], g[f"{tag}_p"], rtol=1e-6)
            neg, prob = M.sample_item_with_pop(g[f"{tag}_k01"], g[f"{tag}_p01"], g[f"{tag}_u"], g[f"{tag}_cp"], b["indices"],
                                               b["indptr"], g[f"{tag}_p"])
        np.testing.assert_array_equal(neg, g[f"{tag}_neg"])
        np.testing.assert_allclose(prob, g[f"{tag}_negprob"], rtol=1e-5, atol=1e-6)
        for pos, key in ((pos1, "pos1_p"), (pos2, "pos2_p")):
            got = M.midx_item_p(query, pos, b["c0"], b["c1"], b["cd0"], b["cd1"], b.get("p"))
            np.testing.assert_allclose(got.numpy(), g[f"{tag}_{key}"], rtol=1e-5, atol=1e-6)
    # second update warm-starts from the previous centers
    torch.manual_seed(17)
    b = M.midx_build(emb, K, None, None, 8, None)
    b2 = M.midx_build(torch.from_numpy(g["mu_ip2_emb"]), K, b["c0"], b["c1"], 3, None)
    np.testing.assert_array_equal(b2["indices"], g["mu_ip2_indices"])
    np.testing.assert_allclose(b2["wkk"], g["mu_ip2_wkk"])


# ---------------------------------------------------------------- size-independent properties of the new oracle pieces
def test_masked_uniform_oracle_is_the_order_preserving_map_onto_the_complement():
    """For a duplicate-free history the reference's arithmetic maps the k-th raw draw value (k = floor(u (N - c)) + 1) to
    the k-th smallest item that is NOT in the history: checked exhaustively over every k for random histories."""
    rng = np.random.RandomState(0)
    for _ in range(25):
        N = int(rng.randint(5, 60)); H = int(rng.randint(1, 12))
        c = int(rng.randint(0, min(H, N - 1) + 1))
        hist = np.zeros((1, H), dtype=np.int64)
        hist[0, rng.permutation(H)[:c]] = rng.permutation(N)[:c] + 1          # padding zeros anywhere in the row
        complement = np.array([i for i in range(1, N + 1) if i not in set(hist[0].tolist())])
        m = N - c
        seeds = ((np.arange(m) + 0.5) / m).astype(np.float32)[None, :]        # hits every k = 1 .. N - c exactly once
        got = S.masked_uniform_from_seeds(N, hist, seeds)[0]
        np.testing.assert_array_equal(got, complement)


def test_construct_index_oracle_properties():
    from oracle import midx as M
    rng = np.random.RandomState(1)
    for nb in (1, 7, 300):
        codes = rng.randint(0, nb, size=2000)
        ind, ptr_ = M.construct_index(codes, nb)
        assert sorted(ind.tolist()) == list(range(2000)) and ptr_[0] == 0 and ptr_[-1] == 2000
        for c in range(nb):
            seg = ind[ptr_[c]:ptr_[c + 1]]
            assert np.all(codes[seg] == c) and np.all(np.diff(seg) > 0)        # grouped by bucket, stable (ascending ids)
        cp, tot = M.bucket_cdf(np.ones(2000, np.float32), ind, ptr_)
        np.testing.assert_allclose(tot, np.diff(ptr_))
        assert all(abs(cp[ptr_[c + 1] - 1] - 1.0) < 1e-6 for c in range(nb) if ptr_[c + 1] > ptr_[c])
