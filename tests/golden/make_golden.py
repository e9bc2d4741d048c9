#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (ustcml/RecStudio @ 6f628ddd) is pure Python and imports with two
stub packages (`nni`, `torchmetrics`, see baseline/shim/).  It cannot travel to
the GPU box, so the vectors it produces are committed as small fixtures next
to this script.  Everything is computed on the CPU with torch
{torch.__version__ recorded in each file}.  Nothing here is imported by the
product or by the GPU tests; the tests only read the .npz files.
"""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "baseline", "shim"))
sys.path.insert(0, "/root/reference")
os.chdir(tempfile.mkdtemp(prefix="rs_golden_"))     # recstudio.utils creates ./log, ./.recstudio

import logging  # noqa: E402
import warnings  # noqa: E402

warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402

# NB: recstudio.model must be imported BEFORE recstudio.ann.sampler: sampler.py:6 imports
# recstudio.model.scorer, whose package __init__ pulls baseretriever.py:9
# (`from recstudio.ann.sampler import *`) while sampler.py is still half-initialised.
from recstudio.model.basemodel import BaseRetriever  # noqa: E402
import recstudio.eval as rs_eval  # noqa: E402
from recstudio.ann import sampler as rs_sampler  # noqa: E402
from recstudio.model import loss_func as rs_loss  # noqa: E402
from recstudio.model import scorer as rs_scorer  # noqa: E402
from recstudio.utils import get_model  # noqa: E402

META = dict(torch=torch.__version__, reference="ustcml/RecStudio@6f628ddd")


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    out["_meta"] = np.array(repr(META))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in out.items() if k != "_meta"})


# --------------------------------------------------------------------------
class _Feat:
    def __init__(self, fields):
        self.fields = fields


class FakeData:
    """Duck-typed dataset: just what Recommender._init_model / BaseRetriever._init_model
    read (recommender.py:66-77, baseretriever.py:54-68)."""
    name = "fake"
    fuid, fiid, frating = "user_id", "item_id", "rating"

    def __init__(self, num_users, num_items):
        self.num_users, self.num_items = num_users, num_items
        self.use_field = {self.fuid, self.fiid, self.frating}
        self.user_feat = _Feat([self.fuid])
        self.item_feat = _Feat([self.fiid])

    def drop_feat(self, fields):
        pass


def build_retriever(num_users, num_items, d, n, loss, scorer, sampler=None, seed=2022):
    conf = get_model("BPR")[1]
    conf["train"]["gpu"] = None
    conf["train"]["negative_count"] = n
    conf["train"]["seed"] = seed
    conf["model"]["embed_dim"] = d
    kwargs = dict(loss=loss, scorer=scorer)
    if sampler is not None:
        kwargs["sampler"] = sampler
    m = BaseRetriever(conf, **kwargs)
    m.logger = logging.getLogger("golden")
    m._init_model(FakeData(num_users, num_items))
    m._init_parameter()
    return m


def golden_appendix_a():
    """SURVEY.md Appendix A inputs through the reference scorer/loss classes."""
    q = torch.tensor([[1, 2, -1, 0.5], [0, -1, 1, 2]])
    vp = torch.tensor([[0.5, 0.5, 1, -1], [1, 1, 1, 1.0]])
    vn = torch.tensor([[[1, 0, 0, 0], [0, 1, 0, 2], [-1, -1, 0.5, 0]],
                       [[2, 0, 1, 0], [0, 0, 0, 0], [0.5, -0.5, 0.25, 1.0]]])
    lqp = torch.tensor([-1.0, -2.0]); lqn = torch.tensor([[-1, -2, -3], [-0.5, -1.5, -2.5]])
    out = dict(q=q, vp=vp, vn=vn, lqp=lqp, lqn=lqn)
    for name, sc in (("ip", rs_scorer.InnerProductScorer()), ("eu", rs_scorer.EuclideanScorer())):
        ps, ns = sc(q, vp), sc(q, vn)
        out[f"{name}_pos"] = ps; out[f"{name}_neg"] = ns
        out[f"{name}_bpr"] = rs_loss.BPRLoss()(None, ps, None, ns, None)
        z = torch.zeros_like(ps); zn = torch.zeros_like(ns)
        out[f"{name}_ssm0"] = rs_loss.SampledSoftmaxLoss()(None, ps.clone(), z, ns.clone(), zn)
        out[f"{name}_ssmq"] = rs_loss.SampledSoftmaxLoss()(None, ps.clone(), lqp, ns.clone(), lqn)
        out[f"{name}_softmax"] = rs_loss.SoftmaxLoss()(None, ps, ns)
    save("appendix_a", **out)


def golden_training_steps():
    """BaseRetriever.training_step + backward on seeded synthetic batches, all
    loss x scorer combinations, uniform (CPU mt19937) negatives recorded so the
    CUDA path can be fed the identical (user, pos, neg[]) batch."""
    cases = [("small", 57, 301, 32, 16, 12), ("d128", 40, 997, 128, 24, 40), ("d64dup", 9, 23, 64, 32, 50)]
    for tag, U, N, d, B, n in cases:
        for lname, loss_cls in (("bpr", rs_loss.BPRLoss), ("ssm", rs_loss.SampledSoftmaxLoss)):
            for sname, sc_cls in (("ip", rs_scorer.InnerProductScorer), ("eu", rs_scorer.EuclideanScorer)):
                m = build_retriever(U, N, d, n, loss_cls(), sc_cls())
                g = torch.Generator().manual_seed(7)
                with torch.no_grad():    # non-trivial scores: re-draw weights N(0, 0.3), keep row 0 zero
                    m.item_encoder.weight.copy_(torch.randn(N, d, generator=g) * 0.3)
                    m.query_encoder.weight.copy_(torch.randn(U, d, generator=g) * 0.3)
                    m.item_encoder.weight[0] = 0; m.query_encoder.weight[0] = 0
                batch = {"user_id": torch.randint(1, U, (B,), generator=g),
                         "item_id": torch.randint(1, N, (B,), generator=g),
                         "rating": torch.ones(B)}
                torch.manual_seed(123)
                loss = m.training_step(batch)
                loss.backward()
                torch.manual_seed(123)
                out = m.forward(batch, return_neg_id=True, return_query=True)
                save(f"step_{tag}_{lname}_{sname}",
                     w_item=m.item_encoder.weight, w_user=m.query_encoder.weight,
                     user=batch["user_id"], pos=batch["item_id"], neg=out["neg_id"],
                     loss=loss, pos_score=out["score"]["pos_score"], neg_score=out["score"]["neg_score"],
                     log_pos_prob=out["score"]["log_pos_prob"], log_neg_prob=out["score"]["log_neg_prob"],
                     d_item=m.item_encoder.weight.grad, d_user=m.query_encoder.weight.grad)


def golden_popular():
    """PopularSamplerModel tables (all modes), draws for recorded seeds, a
    training step with popularity negatives and non-zero logQ."""
    out = {}
    pc_small = np.array([0, 5, 1, 0, 10, 3])
    rng = np.random.RandomState(0)
    pc_big = np.floor(rng.zipf(1.3, size=5000)).astype(np.int64); pc_big[0] = 0
    for tag, pc in (("small", pc_small), ("big", pc_big)):
        out[f"{tag}_count"] = pc
        for mode in (0, 1, 2):
            s = rs_sampler.PopularSamplerModel(pc, mode=mode)
            out[f"{tag}_m{mode}_prob"] = s.pop_prob; out[f"{tag}_m{mode}_table"] = s.table
    s = rs_sampler.PopularSamplerModel(pc_big, mode=0)
    torch.manual_seed(5)
    seeds = torch.rand(64, 33)
    out["draw_seeds"] = seeds
    out["draw_idx"] = torch.searchsorted(s.table, seeds)
    out["draw_logq"] = s.compute_item_p(None, out["draw_idx"])
    probe = torch.tensor([0, 0.13757, 0.1377, 0.47944, 0.48, 0.9999])
    s0 = rs_sampler.PopularSamplerModel(pc_small, mode=0)
    out["probe_seeds"] = probe; out["probe_idx"] = torch.searchsorted(s0.table, probe)
    save("popular", **out)

    U, N, d, B, n = 31, 5000, 64, 20, 64
    smp = rs_sampler.PopularSamplerModel(pc_big, mode=0)
    m = build_retriever(U, N, d, n, rs_loss.SampledSoftmaxLoss(), rs_scorer.InnerProductScorer(), sampler=smp)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        m.item_encoder.weight.copy_(torch.randn(N, d, generator=g) * 0.2)
        m.query_encoder.weight.copy_(torch.randn(U, d, generator=g) * 0.2)
        m.item_encoder.weight[0] = 0; m.query_encoder.weight[0] = 0
    batch = {"user_id": torch.randint(1, U, (B,), generator=g),
             "item_id": torch.randint(1, N, (B,), generator=g), "rating": torch.ones(B)}
    torch.manual_seed(77)
    loss = m.training_step(batch); loss.backward()
    torch.manual_seed(77)
    o = m.forward(batch, return_neg_id=True)
    save("step_popular_ssm_ip", w_item=m.item_encoder.weight, w_user=m.query_encoder.weight,
         user=batch["user_id"], pos=batch["item_id"], neg=o["neg_id"], loss=loss,
         pos_score=o["score"]["pos_score"], neg_score=o["score"]["neg_score"],
         log_pos_prob=o["score"]["log_pos_prob"], log_neg_prob=o["score"]["log_neg_prob"],
         d_item=m.item_encoder.weight.grad, d_user=m.query_encoder.weight.grad,
         pop_count=pc_big)


def golden_uniform_cpu():
    """UniformSampler contract on CPU: dtypes/shapes/range (the stream itself is
    mt19937 on CPU, Philox on CUDA -- only the CUDA stream is a parity target)."""
    torch.manual_seed(2022)
    lp, neg, ln = rs_sampler.UniformSampler(10)(2, 4, pos_items=torch.tensor([3, 4]))
    torch.manual_seed(1)
    neg2, ln2 = rs_sampler.UniformSampler(1000)(torch.zeros(5, 7, 8), 9)
    save("uniform_cpu", log_pos=lp, neg=neg, log_neg=ln, neg2=neg2, log_neg2=ln2)


def golden_topk_eval():
    """BaseRetriever.topk with history mask (Appendix A example + random) and
    every rank metric of recstudio.eval."""
    out = {}
    m = build_retriever(3, 7, 2, 1, rs_loss.BPRLoss(), rs_scorer.InnerProductScorer())
    with torch.no_grad():
        m.item_encoder.weight.copy_(torch.tensor([[0, 0], [1, 0], [.9, 0], [.8, 0], [.7, 0], [.6, 0], [.5, 0]]))
        m.query_encoder.weight.copy_(torch.tensor([[0, 0], [1, 0], [-1.0, 0]]))
    m._update_item_vector()
    sc, ids = m.topk({"user_id": torch.tensor([1, 2])}, 3, torch.tensor([[2, 1], [6, 0]]))
    out.update(a_w_item=m.item_encoder.weight, a_w_user=m.query_encoder.weight, a_user=torch.tensor([1, 2]),
               a_hist=torch.tensor([[2, 1], [6, 0]]), a_score=sc, a_ids=ids)

    U, N, d, Be, k, H = 50, 2001, 48, 16, 10, 12
    m = build_retriever(U, N, d, 1, rs_loss.BPRLoss(), rs_scorer.InnerProductScorer())
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        m.item_encoder.weight.copy_(torch.randn(N, d, generator=g))
        m.query_encoder.weight.copy_(torch.randn(U, d, generator=g))
        m.item_encoder.weight[0] = 0; m.query_encoder.weight[0] = 0
    m._update_item_vector()
    users = torch.randint(1, U, (Be,), generator=g)
    hist = torch.stack([torch.randperm(N - 1, generator=g)[:H] + 1 for _ in range(Be)])
    hist[:, -3:] = torch.where(torch.rand(Be, 3, generator=g) < 0.5, torch.zeros(Be, 3, dtype=torch.long), hist[:, -3:])
    hist, _ = torch.sort(hist, dim=1, descending=True)       # right-padded with zeros
    with torch.no_grad():   # make sure some history items would otherwise rank in the top-k
        best = torch.topk(m.query_encoder(users) @ m.item_vector.T, 4).indices + 1
        hist[:, :2] = best[:, 1:3]
    sc, ids = m.topk({"user_id": users}, k, hist)
    sc100, ids100 = m.topk({"user_id": users}, 100, hist)
    sc_nh, ids_nh = m.topk({"user_id": users}, k, None)
    out.update(r_w_item=m.item_encoder.weight, r_w_user=m.query_encoder.weight, r_user=users, r_hist=hist,
               r_score=sc, r_ids=ids, r_score100=sc100, r_ids100=ids100, r_score_nohist=sc_nh, r_ids_nohist=ids_nh)

    # _test_step label + metrics (baseretriever.py:416-431)
    target = torch.zeros(Be, 5, dtype=torch.long)
    for b in range(Be):
        t = int(torch.randint(1, 6, (1,), generator=g))
        pick = torch.cat([ids100[b, torch.randperm(30, generator=g)[:max(t - 2, 0)]],
                          torch.randint(1, N, (t - max(t - 2, 0),), generator=g)])
        target[b, :t] = pick[:t]
    rating = (target > 0).float()
    batch = {"user_id": users, "item_id": target, "rating": rating, "user_hist": hist}
    m.config["eval"]["topk"] = 100
    res, bs = m._test_step(batch, ["ndcg", "recall", "precision", "map", "mrr", "hit"], [5, 10, 20])
    out.update(e_target=target, e_rating=rating)
    for kname, v in res.items():
        out["e_" + kname.replace("@", "_at_")] = v

    label = torch.tensor([[1, 0, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 0]], dtype=torch.bool)
    tgt = torch.tensor([[1, 1, 1], [1, 0, 0], [1, 1, 0.0]])
    out.update(m_label=label, m_target=tgt)
    for name, fn in rs_eval.get_rank_metrics(["ndcg", "recall", "precision", "map", "mrr", "hit"]):
        for kk in (1, 3, 5):
            out[f"m_{name}_{kk}"] = fn(label, tgt, kk)
    save("topk_eval", **out)


def golden_full_softmax():
    """Full-score branch + SoftmaxLoss (baseretriever.py:177-186, loss_func.py:41-42)."""
    U, N, d, B = 20, 777, 64, 12
    m = build_retriever(U, N, d, 0, rs_loss.SoftmaxLoss(), rs_scorer.InnerProductScorer())
    m.sampler = None
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        m.item_encoder.weight.copy_(torch.randn(N, d, generator=g) * 0.3)
        m.query_encoder.weight.copy_(torch.randn(U, d, generator=g) * 0.3)
        m.item_encoder.weight[0] = 0; m.query_encoder.weight[0] = 0
    batch = {"user_id": torch.randint(1, U, (B,), generator=g),
             "item_id": torch.randint(1, N, (B,), generator=g), "rating": torch.ones(B)}
    loss = m.training_step(batch); loss.backward()
    o = m.forward(batch, full_score=True)
    save("step_full_softmax", w_item=m.item_encoder.weight, w_user=m.query_encoder.weight,
         user=batch["user_id"], pos=batch["item_id"], loss=loss,
         pos_score=o["score"]["pos_score"], all_score=o["score"]["all_score"],
         d_item=m.item_encoder.weight.grad, d_user=m.query_encoder.weight.grad)


def golden_masked_uniform():
    """uniform_sample_masked_hist / MaskedUniformSampler (sampler.py:117-147,187-214): the torch.rand
    seeds are recorded (CPU mt19937 here, Philox on CUDA) so the integer arithmetic after the
    draw is pinned independently of the generator."""
    out = {}
    real_rand = torch.rand
    rec = {}

    def rand_spy(*a, **k):
        rec["seeds"] = real_rand(*a, **k)
        return rec["seeds"]

    g = torch.Generator().manual_seed(21)
    N = 400                                                     # Sampler.num_items (real items)
    hist_a = torch.zeros(12, 9, dtype=torch.long)               # distinct items, right-padded
    for b in range(12):
        c = int(torch.randint(0, 10, (1,), generator=g))
        hist_a[b, :c] = torch.randperm(N, generator=g)[:c] + 1
    hist_b = hist_a.clone(); hist_b[:, 1] = hist_b[:, 0]        # duplicate items -> non-monotone adjusted row
    hist_c = torch.flip(hist_a, dims=[1])                       # left-padded
    hist_d = torch.stack([torch.randperm(30, generator=g)[:16] + 1 for _ in range(5)])   # full rows, tiny catalog
    cases = (("a", N, hist_a, 7, None), ("b", N, hist_b, 7, None), ("c", N, hist_c, 4, 3), ("d", 30, hist_d, 40, None))
    torch.rand = rand_spy
    try:
        for tag, n_items, hist, n, nq in cases:
            torch.manual_seed(100 + len(tag) + n)
            neg = rs_sampler.uniform_sample_masked_hist(n_items, n, hist, nq)
            out[f"{tag}_num_items"] = n_items; out[f"{tag}_hist"] = hist; out[f"{tag}_seeds"] = rec["seeds"]
            out[f"{tag}_neg"] = neg
        smp = rs_sampler.MaskedUniformSampler(N + 1)
        torch.manual_seed(9)
        lp, neg, ln = smp(torch.zeros(12, 8), 5, pos_items=torch.arange(12), user_hist=hist_a)
        out.update(s_seeds=rec["seeds"], s_neg=neg, s_log_pos=lp, s_log_neg=ln)
    finally:
        torch.rand = real_rand
    save("masked_uniform", **out)


def golden_sampling_methods():
    """BaseRetriever.sampling methods dns / sir / toprand / top&rand / brute (baseretriever.py:248-369)
    inside a full training_step + backward.  Every random draw is recorded (pool ids, multinomial /
    randint outcomes) so the same step can be replayed on another device's generator."""
    U, N, d, B = 23, 301, 32, 14
    H = 6
    for method, loss_cls, lname, nc in (("dns", rs_loss.BPRLoss, "bpr", [12, 4]), ("sir", rs_loss.SampledSoftmaxLoss, "ssm", [12, 5]),
                                        ("toprand", rs_loss.BPRLoss, "bpr", [10, 4]), ("top&rand", rs_loss.BPRLoss, "bpr", 6),
                                        ("brute", rs_loss.SampledSoftmaxLoss, "ssm", 5)):
        m = build_retriever(U, N, d, nc, loss_cls(), rs_scorer.InnerProductScorer())
        m.config["train"]["sampling_method"] = method
        g = torch.Generator().manual_seed(31)
        with torch.no_grad():
            m.item_encoder.weight.copy_(torch.randn(N, d, generator=g) * 0.4)
            m.query_encoder.weight.copy_(torch.randn(U, d, generator=g) * 0.4)
            m.item_encoder.weight[0] = 0; m.query_encoder.weight[0] = 0
        m._update_item_vector()
        hist = torch.stack([torch.randperm(N - 1, generator=g)[:H] + 1 for _ in range(B)])
        hist[:, -2:] = 0
        batch = {"user_id": torch.randint(1, U, (B,), generator=g), "item_id": torch.randint(1, N, (B,), generator=g),
                 "rating": torch.ones(B), "user_hist": hist}
        rec = {}
        real_fwd, real_mult, real_randint = m.sampler.forward, torch.multinomial, torch.randint

        def fwd_spy(*a, **k):
            r = real_fwd(*a, **k)
            rec["pool"] = r[-2]
            return r

        def mult_spy(*a, **k):
            rec["multinomial"] = real_mult(*a, **k)
            return rec["multinomial"]

        def randint_spy(*a, **k):
            rec["randint"] = real_randint(*a, **k)
            return rec["randint"]

        m.sampler.forward = fwd_spy
        torch.multinomial, torch.randint = mult_spy, randint_spy
        try:
            torch.manual_seed(55)
            loss = m.training_step(batch)
            loss.backward()
            pool_train = {k: v.clone() for k, v in rec.items()}
            torch.manual_seed(55)
            o = m.forward(batch, return_neg_id=True, return_query=True)
        finally:
            m.sampler.forward = real_fwd
            torch.multinomial, torch.randint = real_mult, real_randint
        tag = method.replace("&", "and")
        extra = {f"rec_{k}": v for k, v in pool_train.items()}
        save(f"sampling_{tag}", w_item=m.item_encoder.weight, w_user=m.query_encoder.weight, user=batch["user_id"],
             pos=batch["item_id"], hist=hist, neg=o["neg_id"], loss=loss, pos_score=o["score"]["pos_score"],
             neg_score=o["score"]["neg_score"], log_pos_prob=o["score"]["log_pos_prob"],
             log_neg_prob=o["score"]["log_neg_prob"], d_item=m.item_encoder.weight.grad,
             d_user=m.query_encoder.weight.grad, negative_count=np.asarray(nc), **extra)


def golden_midx():
    """kmeans / construct_index / MIDX + Cluster samplers (sampler.py:9-45,261-559): index build (update) and the
    deterministic parts of the draw (sample_item for recorded seeds, compute_item_p)."""
    out = {}
    g = torch.Generator().manual_seed(41)
    # well separated blobs so that assignments do not sit on fp32 ties
    cent = torch.randn(7, 16, generator=g) * 3
    X = cent[torch.randint(0, 7, (600,), generator=g)] + torch.randn(600, 16, generator=g) * 0.5
    for it in (1, 4, 50):
        C, assign, assign_m, loss = rs_sampler.kmeans(X, X[:7].clone(), max_iter=it)
        out[f"km_C_{it}"] = C; out[f"km_assign_{it}"] = assign; out[f"km_loss_{it}"] = loss
    torch.manual_seed(3)
    C, assign, _, loss = rs_sampler.kmeans(X, 5, max_iter=6)          # int K: centers = X[randperm(N)[:K]] on the CPU generator
    out.update(km_X=X, km_Cr=C, km_assignr=assign, km_lossr=loss, km_next_rand=torch.rand(2))
    codes = torch.randint(0, 37, (1000,), generator=g)
    ind, ptr_ = rs_sampler.construct_index(codes, 37)
    out.update(ci_codes=codes, ci_indices=ind, ci_indptr=ptr_)

    N, d, K = 500, 16, 4
    emb = cent[torch.randint(0, 7, (N,), generator=g)][:, :d] + torch.randn(N, d, generator=g) * 0.6
    query = torch.randn(6, d, generator=g)
    pos1 = torch.randint(1, N + 1, (6,), generator=g)
    pos2 = torch.randint(0, N + 1, (6, 3), generator=g)
    popc = torch.floor(torch.rand(N, generator=g) * 50)
    real_rand_like = torch.rand_like
    rec = {}

    def spy(*a, **k):
        rec["u"] = real_rand_like(*a, **k)
        return rec["u"]

    def dump(tag, smp, has_cp):
        for name in ("c0", "c1", "cd0", "cd1", "c", "cd", "indices", "indptr", "wkk", "p", "cp"):
            if hasattr(smp, name):
                out[f"{tag}_{name}"] = getattr(smp, name)
        k01 = torch.randint(0, smp.indptr.numel() - 1, (6, 9), generator=g)
        sizes = smp.indptr[1:] - smp.indptr[:-1]
        k01 = torch.where(sizes[k01] > 0, k01, torch.argmax(sizes).expand_as(k01))     # only non-empty buckets are ever drawn
        p01 = torch.randn(6, 9, generator=g)
        torch.rand_like = spy
        try:
            neg, prob = smp.sample_item(k01, p01)
        finally:
            torch.rand_like = real_rand_like
        out.update({f"{tag}_k01": k01, f"{tag}_p01": p01, f"{tag}_u": rec["u"], f"{tag}_neg": neg, f"{tag}_negprob": prob})
        out[f"{tag}_pos1_p"] = smp.compute_item_p(query, pos1) if not (has_cp and tag.startswith("cl")) else torch.zeros(1)
        out[f"{tag}_pos2_p"] = smp.compute_item_p(query, pos2)

    out.update(mx_emb=emb, mx_query=query, mx_pos1=pos1, mx_pos2=pos2, mx_pop=popc)
    for tag, make, has_cp in (
            ("mu_ip", lambda: rs_sampler.MIDXSamplerUniform(N + 1, K, rs_scorer.InnerProductScorer()), False),
            ("mu_eu", lambda: rs_sampler.MIDXSamplerUniform(N + 1, K, rs_scorer.EuclideanScorer()), True),
            ("mp_ip", lambda: rs_sampler.MIDXSamplerPop(popc, K, rs_scorer.InnerProductScorer(), mode=1), True),
            ("cl_ip", lambda: rs_sampler.ClusterSamplerUniform(N + 1, K * 2, rs_scorer.InnerProductScorer()), False),
            ("cp_ip", lambda: rs_sampler.ClusterSamplerPop(popc, K * 2, rs_scorer.InnerProductScorer(), mode=2), True)):
        smp = make()
        torch.manual_seed(17)
        smp.update(emb, max_iter=8)
        dump(tag, smp, has_cp)
        if tag == "mu_ip":                    # a second update starts k-means from the previous centers (sampler.py:276-279)
            emb2 = emb + 0.05 * torch.randn(N, d, generator=g)
            smp.update(emb2, max_iter=3)
            out.update(mu_ip2_emb=emb2, mu_ip2_c0=smp.c0, mu_ip2_c1=smp.c1, mu_ip2_indices=smp.indices, mu_ip2_wkk=smp.wkk)
    save("midx", **out)


def golden_pooling():
    """SeqPoolingLayer (recstudio/model/module/layers.py:247-314) for every pooling type the SASRec / BERT4Rec query
    encoder can be configured with."""
    from recstudio.model import module as rs_module
    g = torch.Generator().manual_seed(77)
    B, L, D = 7, 9, 12
    x = torch.randn(B, L, D, generator=g)
    seqlen = torch.randint(1, L + 1, (B,), generator=g)
    mask_token = torch.rand(B, L, generator=g) < 0.3
    out = dict(x=x, seqlen=seqlen, mask_token=mask_token)
    for ptype in ("origin", "mask", "concat", "sum", "mean", "max", "last"):
        layer = rs_module.SeqPoolingLayer(pooling_type=ptype)
        r = layer(x, seqlen, mask_token=mask_token if ptype == "mask" else None)
        if ptype == "max":
            out["max_values"], out["max_indices"] = r.values, r.indices
        else:
            out[ptype] = r
    save("pooling", **out)


def golden_sasrec_encoder():
    """The reference's own SASRecQueryEncoder (recstudio/model/seq/sasrec.py:8-67; BERT4Rec uses it with
    bidirectional=True, bert4rec.py) at the config-3 shape class: d = 128, 2 heads, FFN 128, GELU, LayerNorm eps 1e-12, 2 layers,
    L = 200 (seq/config/sasrec.yaml:1-7), dropout 0 so the result is a function of the weights.  Weights, a right-padded id batch,
    the pooled output ('last' pooling) for the causal and the bidirectional mask, and gradients of sum(out * g) with respect
    to a spread of parameters (table, positions, first-layer attention, last-layer FFN and norms)."""
    from recstudio.model.seq.sasrec import SASRecQueryEncoder
    torch.manual_seed(31)
    N, d, L, B = 300, 128, 200, 6
    item = torch.nn.Embedding(N, d, padding_idx=0)
    enc = SASRecQueryEncoder("item_id", d, L, 2, 128, 0.0, "gelu", 1e-12, 2, item, bidirectional=False)
    with torch.no_grad():
        item.weight.normal_(0, 0.5); item.weight[0] = 0
        enc.position_emb.weight.normal_(0, 0.5)
    seqlen = torch.tensor([1, 7, 64, 129, 199, 200])
    ids = torch.randint(1, N, (B, L)) * (torch.arange(L)[None, :] < seqlen[:, None])
    batch = {"in_item_id": ids, "seqlen": seqlen}
    g = torch.randn(B, d)
    watch = ["item_encoder.weight", "position_emb.weight", "transformer_layer.layers.0.self_attn.in_proj_weight",
             "transformer_layer.layers.0.self_attn.out_proj.bias", "transformer_layer.layers.0.norm1.weight",
             "transformer_layer.layers.1.linear2.weight", "transformer_layer.layers.1.norm2.bias"]
    out = {"ids": ids, "seqlen": seqlen, "g": g, "watch": np.array(watch)}
    for k, v in enc.state_dict().items():
        out["w:" + k] = v
    enc.train()
    for tag, bidir in (("causal", False), ("bidir", True)):
        enc.bidirectional = bidir
        enc.zero_grad()
        o = enc(batch)
        (o * g).sum().backward()
        out["out_" + tag] = o
        params = dict(enc.named_parameters())
        for k in watch:
            out["grad_%s:%s" % (tag, k)] = params[k].grad.clone()
    enc.eval()
    enc.bidirectional = False
    out["out_eval_causal"] = enc(batch)
    save("sasrec_encoder", **out)


if __name__ == "__main__":
    ALL = [golden_appendix_a, golden_training_steps, golden_popular, golden_uniform_cpu, golden_topk_eval,
           golden_full_softmax, golden_masked_uniform, golden_sampling_methods,
           golden_midx, golden_pooling, golden_sasrec_encoder]
    want = sys.argv[1:]                       # optional: names of the generators to (re)run
    for fn in ALL:
        if not want or fn.__name__ in want:
            fn()
