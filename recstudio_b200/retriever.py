"""Retriever hosts for the fused plugins.

* ``FusedRetriever`` / ``FusedBPR``: when ``recstudio`` is importable these are real
  ``BaseRetriever`` subclasses (``FusedRetrieverMixin`` in front of it), so the reference's
  trainer (``fit`` / ``training_epoch`` / ``evaluate``) drives them unchanged.
* ``MiniRetriever``: where the Python reference cannot be imported (the GPU box), a mirror
  of the ``BaseRetriever`` surface FOR THIS PATH ONLY -- kwargs constructor, ``forward``,
  ``sampling('none')``, ``training_step``, ``topk``, ``_update_item_vector`` -- with the same
  names, argument meaning and error behaviour (recstudio/model/basemodel/baseretriever.py:
  14-46,117-192,204-278,360-404).  It holds no arithmetic of its own: everything goes through
  the plugins.
"""
from __future__ import annotations

import copy
import inspect
from typing import Dict, Optional

import torch

from . import iface, plugins

DEFAULT_CONFIG = {
    # the hot-path knobs of recstudio/model/basemodel/basemodel.yaml (:27,36,52,57,58,69-74)
    "model": {"embed_dim": 64, "item_bias": False},
    "train": {"negative_count": 1, "sampling_method": "none", "excluding_hist": False, "batch_size": 512,
              "ann": None, "gpu": None, "seed": 2022, "learner": "adam", "learning_rate": 0.001},
    "eval": {"topk": 100, "cutoff": [5, 10, 20], "batch_size": 128, "val_metrics": ["ndcg", "recall"],
             "test_metrics": ["ndcg", "recall", "precision", "map", "mrr", "hit"]},
}


class MiniRetriever(torch.nn.Module):
    def __init__(self, config: Dict = None, **kwargs):
        super().__init__()
        self.config = copy.deepcopy(DEFAULT_CONFIG)
        for grp, vals in (config or {}).items():
            self.config.setdefault(grp, {}).update(vals)
        self.embed_dim = self.config["model"]["embed_dim"]
        for key, attr in (("item_encoder", "item_encoder"), ("query_encoder", "query_encoder"), ("scorer", "score_func")):
            if key in kwargs:
                assert isinstance(kwargs[key], torch.nn.Module), "%s must be torch.nn.Module" % key   # baseretriever.py:17-36
                setattr(self, attr, kwargs[key])
            else:
                setattr(self, attr, None)
        if self.score_func is None:
            self.score_func = plugins.FusedInnerProductScorer()
        if "sampler" in kwargs:
            assert isinstance(kwargs["sampler"], iface.Sampler), "sampler must be recstudio.ann.sampler.Sampler"   # :38-41
            self.sampler = kwargs["sampler"]
        else:
            self.sampler = None
        if "loss" in kwargs:
            assert isinstance(kwargs["loss"], (iface.FullScoreLoss, iface.PairwiseLoss, iface.PointwiseLoss)), \
                "loss should be one of FullScoreLoss / PairwiseLoss / PointwiseLoss"                              # recommender.py:48-54
            self.loss_fn = kwargs["loss"]
        else:
            self.loss_fn = None
        self.fuid, self.fiid, self.frating = "user_id", "item_id", "rating"
        self.item_fields, self.query_fields = {self.fiid}, {self.fuid}
        self.neg_count = self.config["train"]["negative_count"]
        self.use_index = False

    # --- construction helpers (the reference builds these in _init_model from the dataset) ---------
    def init_tables(self, num_users: int, num_items: int):
        d = self.embed_dim
        if self.item_encoder is None:
            self.item_encoder = plugins.FusedEmbedding(num_items, d, padding_idx=0)
        if self.query_encoder is None:
            self.query_encoder = plugins.FusedEmbedding(num_users, d, padding_idx=0)
        if self.sampler is None:
            self.sampler = plugins.FusedUniformSampler(num_items)
        if self.loss_fn is None:
            self.loss_fn = plugins.FusedBPRLoss()
        return self

    @staticmethod
    def _to_device(batch, device):
        """recommender.py:699-714"""
        if isinstance(batch, (torch.Tensor, torch.nn.Module)):
            return batch.to(device)
        if isinstance(batch, dict):
            for k in batch:
                batch[k] = MiniRetriever._to_device(batch[k], device)
            return batch
        if isinstance(batch, (list, tuple)):
            out = [MiniRetriever._to_device(b, device) for b in batch]
            return out if isinstance(batch, list) else tuple(out)
        raise TypeError("`batch` is expected to be torch.Tensor, Dict, List or Tuple, but %s given." % type(batch))

    def _get_item_feat(self, data):
        return data[self.fiid] if isinstance(data, dict) else data

    def _get_query_feat(self, data):                                            # baseretriever.py:86-97
        if not isinstance(data, dict):
            return data
        if len(self.query_fields) == 1:
            return data[list(self.query_fields)[0]]
        return dict((field, value) for field, value in data.items() if field in self.query_fields)

    def _get_item_vector(self):
        return self.item_encoder.weight[1:]                                    # baseretriever.py:122-123

    def _update_item_vector(self):                                             # :131-140
        item_vector = self._get_item_vector()
        if not hasattr(self, "item_vector"):
            self.register_buffer("item_vector", item_vector.detach().clone())
        else:
            self.item_vector = item_vector

    # --- baseretriever.py:204-246 -------------------------------------------------------------------
    def _sample(self, batch, neg: int = 1, excluding_hist: bool = False, return_query: bool = True):
        query = self.query_encoder(self._get_query_feat(batch))
        pos_items = batch.get(self.fiid, None)
        if not isinstance(self.sampler, iface.Sampler):
            raise TypeError("`sampler` only support Sampler type.")
        kwargs = {"num_neg": neg, "pos_items": pos_items}
        params = inspect.signature(self.sampler.forward).parameters
        if "excluding_hist" in params:
            kwargs["excluding_hist"] = excluding_hist
        if "user_hist" in params:
            kwargs["user_hist"] = batch.get("user_hist", None) if excluding_hist else None
        kwargs["query"] = query
        pos_prob, neg_id, neg_prob = self.sampler(**kwargs)
        return (pos_prob, neg_id, neg_prob, query) if return_query else (pos_prob, neg_id, neg_prob)

    # --- baseretriever.py:248-278,360-369: method 'none'; dns / sir / toprand / top&rand / brute are implemented
    # on the CUDA ops by FusedRetrieverMixin.sampling, which sits in front of this class ------------
    def sampling(self, batch, num_neg, method="none", excluding_hist=False, t=1, return_query=False, query=None):
        if method != "none":
            raise NotImplementedError("MiniRetriever.sampling mirrors sampling_method='none' only; the other "
                                      "methods are provided by FusedRetrieverMixin.sampling")
        assert self.sampler is not None, "excepted sampler of retriever to be Sampler, but get None."
        if isinstance(num_neg, int):
            num_neg = [num_neg, num_neg]
        log_pos_prob, neg_id, log_neg_prob, query = self._sample(batch, num_neg[1], excluding_hist, True)
        pos_items = batch.get(self.fiid, None)
        if pos_items is not None:
            log_pos_prob = log_pos_prob.view_as(pos_items)
            result = (log_pos_prob.detach(), neg_id, log_neg_prob.detach())
        else:
            result = (None, neg_id, log_neg_prob.detach())
        return (result, query) if return_query else (result, None)

    # --- baseretriever.py:142-192 -------------------------------------------------------------------
    def forward(self, batch, full_score=False, return_query=False, return_item=False, return_neg_item=False,
                return_neg_id=False):
        output = {}
        pos_items = self._get_item_feat(batch)
        pos_item_vec = self.item_encoder(pos_items)
        if self.sampler is not None:
            if self.neg_count is None:
                raise ValueError("`negative_count` value is required when `sampler` is not none.")
            (log_pos_prob, neg_item_idx, log_neg_prob), query = self.sampling(
                batch=batch, num_neg=self.neg_count, excluding_hist=self.config["train"].get("excluding_hist", False),
                method=self.config["train"].get("sampling_method", "none"), return_query=True)
            pos_score = self.score_func(query, pos_item_vec)
            if batch[self.fiid].dim() > 1:
                pos_score[batch[self.fiid] == 0] = -float("inf")
            neg_item_vec = self.item_encoder(self._get_item_feat(neg_item_idx))
            neg_score = self.score_func(query, neg_item_vec)
            output["score"] = {"pos_score": pos_score, "log_pos_prob": log_pos_prob,
                               "neg_score": neg_score, "log_neg_prob": log_neg_prob}
            if return_neg_item:
                output["neg_item"] = neg_item_vec
            if return_neg_id:
                output["neg_id"] = neg_item_idx
        else:
            query = self.query_encoder(self._get_query_feat(batch))
            pos_score = self.score_func(query, pos_item_vec)
            if batch[self.fiid].dim() > 1:
                pos_score[batch[self.fiid] == 0] = -float("inf")
            output["score"] = {"pos_score": pos_score}
            if full_score:
                output["score"]["all_score"] = self.score_func(query, self._get_item_vector())
        if return_query:
            output["query"] = query
        if return_item:
            output["item"] = pos_item_vec
        return output

    def training_step(self, batch):                                            # :399-404
        output = self.forward(batch, isinstance(self.loss_fn, iface.FullScoreLoss))
        score = output["score"]
        score["label"] = batch[self.frating]
        return self.loss_fn(**score)

    def topk(self, batch, k, user_h=None, return_query=False):                 # :374-397 (see FusedRetrieverMixin)
        raise NotImplementedError("topk is provided by FusedRetrieverMixin (rsb200_topk_full)")

    def _test_step(self, batch, metric, cutoffs):                              # :416-431
        from . import rank_metrics
        rank_m = rank_metrics.get_rank_metrics(metric)
        topk = self.config["eval"]["topk"]
        bs = batch[self.frating].size(0)
        assert len(rank_m) > 0
        score, topk_items = self.topk(batch, topk, batch["user_hist"])
        if batch[self.fiid].dim() > 1:
            target, _ = batch[self.fiid].sort()
            idx_ = torch.searchsorted(target, topk_items)
            idx_[idx_ == target.size(1)] = target.size(1) - 1
            label = torch.gather(target, 1, idx_) == topk_items
            pos_rating = batch[self.frating]
        else:
            label = batch[self.fiid].view(-1, 1) == topk_items
            pos_rating = batch[self.frating].view(-1, 1)
        return {f"{name}@{cutoff}": func(label, pos_rating, cutoff) for cutoff in cutoffs for name, func in rank_m}, bs

    def validation_step(self, batch):                                          # :406-409
        cutoff = self.config["eval"]["cutoff"]
        return self._test_step(batch, self.config["eval"]["val_metrics"], [cutoff[0] if isinstance(cutoff, list) else cutoff])

    def test_step(self, batch):                                                # :411-414
        cutoff = self.config["eval"]["cutoff"]
        return self._test_step(batch, self.config["eval"]["test_metrics"], cutoff if isinstance(cutoff, list) else [cutoff])


_Base = iface.BaseRetriever if iface.HAVE_RECSTUDIO else MiniRetriever


class FusedRetriever(plugins.FusedRetrieverMixin, _Base):
    """``BaseRetriever`` with the fused CUDA training step.  Construct it exactly like the
    reference (test/test_retriever.py:13-19): kwargs ``item_encoder=, query_encoder=, scorer=,
    sampler=, loss=`` -- any of them may be a reference plugin, the fused path engages when the
    combination is recognised."""

    def __init__(self, config: Dict = None, fused_grad: str = "dense", device_loader: bool = False, **kwargs):
        super().__init__(config, **kwargs)
        if fused_grad not in ("dense", "sparse", "rows", "apply"):
            raise ValueError("fused_grad must be 'dense', 'sparse', 'rows' or 'apply'")
        self.fused_grad = fused_grad
        self.device_loader = device_loader

    def _get_train_loaders(self, train_data, ddp=False):
        """recommender.py:384-388.  With ``device_loader=True`` the interaction columns live on the GPU and
        batches are sliced there (same epoch permutation as the reference's DataSampler)."""
        if not self.device_loader:
            return super()._get_train_loaders(train_data, ddp)
        from .loader import DeviceBatchLoader
        dev = next(self.parameters()).device
        return [DeviceBatchLoader.from_dataset(train_data, self.config["train"]["batch_size"], dev, shuffle=True, drop_last=False)]

    # subclass hooks of BaseRetriever (baseretriever.py:83-115, recommender.py:369-370)
    def _get_item_encoder(self, train_data):
        return plugins.FusedEmbedding(train_data.num_items, self.embed_dim, padding_idx=0)

    def _get_query_encoder(self, train_data):
        return plugins.FusedEmbedding(train_data.num_users, self.embed_dim, padding_idx=0)

    def _get_score_func(self):
        return plugins.FusedInnerProductScorer()

    def _get_sampler(self, train_data):
        return plugins.FusedUniformSampler(train_data.num_items)

    def _get_loss_func(self):
        return plugins.FusedBPRLoss()

    def _get_optimizers(self):
        """The reference's documented multi-learner hook (recommender.py:403-406).  With
        ``fused_grad='rows'`` the two embedding tables are stepped by ``FusedRowOptimizer`` straight from
        the fused step's workspace; every other parameter keeps the reference's optimizer."""
        if self.fused_grad not in ("rows", "apply"):
            return super()._get_optimizers()
        from .rowopt import FusedRowOptimizer
        tr = self.config["train"]
        name = str(tr.get("learner", "adam")).lower()
        learner = {"sgd": "sgd", "adagrad": "adagrad"}.get(name, "sparse_adam")      # adam / sparse_adam -> SparseAdam semantics
        lr = tr.get("learning_rate", 0.001)
        opts = [{"optimizer": FusedRowOptimizer(self, learner, lr=lr)}]
        tables = {id(self.item_encoder.weight)}
        if isinstance(self.query_encoder, torch.nn.Embedding):
            tables.add(id(self.query_encoder.weight))
        rest = [p for p in self.parameters() if id(p) not in tables]
        if rest:
            opts.append({"optimizer": self._get_optimizer(name, rest, lr, tr.get("weight_decay", 0))})
        return opts


class FusedBPR(FusedRetriever):
    """The reference's BPR (recstudio/model/mf/bpr.py:7-25) on the fused path."""


def build_synthetic(num_users: int, num_items: int, d: int, n, loss: str = "bpr", scorer: str = "ip",
                    sampler: str = "uniform", pop_count=None, fused_grad: str = "dense", device="cuda:0",
                    init_std: Optional[float] = None, seed: int = 2022, sampling_method: str = "none",
                    excluding_hist: bool = False) -> FusedRetriever:
    """A FusedRetriever over plain embedding towers without a dataset object (bench / tests):
    the kwargs construction of test/test_retriever.py with synthetic table sizes."""
    loss_m = plugins.FusedBPRLoss() if loss == "bpr" else plugins.FusedSampledSoftmaxLoss()
    score_m = plugins.FusedInnerProductScorer() if scorer == "ip" else plugins.FusedEuclideanScorer()
    samp_m = {"uniform": lambda: plugins.FusedUniformSampler(num_items),
              "masked": lambda: plugins.FusedMaskedUniformSampler(num_items),
              "popular": lambda: plugins.FusedPopularSampler(pop_count)}[sampler]()
    item = plugins.FusedEmbedding(num_items, d, padding_idx=0)
    user = plugins.FusedEmbedding(num_users, d, padding_idx=0)
    extra = {"sampling_method": sampling_method, "excluding_hist": excluding_hist}
    cfg = {"model": {"embed_dim": d}, "train": {"negative_count": n, "seed": seed, **extra}}
    if iface.HAVE_RECSTUDIO:
        from recstudio.utils import get_model
        conf = get_model("BPR")[1]
        conf["train"].update({"negative_count": n, "gpu": None, "seed": seed, **extra})
        conf["model"]["embed_dim"] = d
        m = FusedRetriever(conf, fused_grad=fused_grad, item_encoder=item, query_encoder=user, scorer=score_m,
                           sampler=samp_m, loss=loss_m)
        m.fuid, m.fiid, m.frating = "user_id", "item_id", "rating"
        m.item_fields, m.query_fields, m.neg_count = {"item_id"}, {"user_id"}, n
    else:
        m = FusedRetriever(cfg, fused_grad=fused_grad, item_encoder=item, query_encoder=user, scorer=score_m,
                           sampler=samp_m, loss=loss_m)
    m = m.to(device)
    std = init_std if init_std is not None else (2.0 / (num_items + d)) ** 0.5    # xavier_normal_ (init.py:5-9)
    with torch.no_grad():
        g = torch.Generator(device=device).manual_seed(seed)
        m.item_encoder.weight.normal_(0, std, generator=g)
        m.query_encoder.weight.normal_(0, std if init_std is None else init_std, generator=g)
        m.item_encoder.weight[0] = 0
        m.query_encoder.weight[0] = 0
    return m


def build_sasrec_synthetic(num_items: int, d: int, n: int, max_seq_len: int = 200, n_head: int = 2, hidden_size: int = 128,
                           n_layer: int = 2, dropout: float = 0.0, loss: str = "ssm", fused_grad: str = "dense",
                           device="cuda:0", init_std: float = 0.02, seed: int = 2022, bidirectional: bool = False,
                           training_pooling_type: str = "last") -> FusedRetriever:
    """SASRec (recstudio/model/seq/sasrec.py:70-123) on the fused path without a dataset object: the item
    tower is a FusedEmbedding shared with the FusedSASRecQueryEncoder (sasrec.py:107), the head is the fused
    sampled-softmax / BPR step (BASELINE config 3).  Batches: {'in_item_id' [B, L], 'seqlen' [B], 'item_id' [B]}."""
    from . import attention
    loss_m = {"bpr": plugins.FusedBPRLoss, "ssm": plugins.FusedSampledSoftmaxLoss, "softmax": plugins.FusedSoftmaxLoss}[loss]()
    item = plugins.FusedEmbedding(num_items, d, padding_idx=0)
    enc = attention.FusedSASRecQueryEncoder(fiid="item_id", embed_dim=d, max_seq_len=max_seq_len, n_head=n_head,
                                            hidden_size=hidden_size, dropout=dropout, activation="gelu", layer_norm_eps=1e-12,
                                            n_layer=n_layer, item_encoder=item, bidirectional=bidirectional,
                                            training_pooling_type=training_pooling_type)
    kwargs = dict(item_encoder=item, query_encoder=enc, scorer=plugins.FusedInnerProductScorer(), loss=loss_m)
    if loss != "softmax":                      # full-softmax models (BERT4Rec) have no sampler (bert4rec.py:43-44)
        kwargs["sampler"] = plugins.FusedUniformSampler(num_items)
    if iface.HAVE_RECSTUDIO:
        from recstudio.utils import get_model
        conf = get_model("SASRec")[1]
        conf["train"].update({"negative_count": n, "gpu": None, "seed": seed})
        conf["model"]["embed_dim"] = d
        m = FusedRetriever(conf, fused_grad=fused_grad, **kwargs)
        m.fuid, m.fiid, m.frating = "user_id", "item_id", "rating"
        m.item_fields, m.neg_count = {"item_id"}, n
    else:
        m = FusedRetriever({"model": {"embed_dim": d}, "train": {"negative_count": n, "seed": seed}}, fused_grad=fused_grad, **kwargs)
    m.query_fields = {"in_item_id", "seqlen"} | ({"mask_token"} if training_pooling_type == "mask" else set())
    if loss == "softmax":
        m.sampler = None
    m = m.to(device)
    with torch.no_grad():      # init_method: normal (seq/config/sasrec.yaml:11, init.py:18-27), padding row re-zeroed
        g = torch.Generator(device=device).manual_seed(seed)
        for p in m.parameters():
            if p.dim() > 1:
                p.normal_(0, init_std, generator=g)
        m.item_encoder.weight[0] = 0
    return m


def build_bert4rec_synthetic(num_items: int, d: int, **kw) -> FusedRetriever:
    """BERT4Rec (recstudio/model/seq/bert4rec.py:8-58) on the fused kernels: the SASRec encoder with bidirectional attention and
    'mask' pooling, the item table extended by the mask-token row (id ``num_items``), full-catalog SoftmaxLoss, no sampler.
    Batches: {'in_item_id' [B, L] (masked positions hold the mask token), 'seqlen' [B], 'mask_token' bool [B, L],
    'item_id' [number of masked positions]} -- what ``_reconstruct_train_data`` produces."""
    return build_sasrec_synthetic(num_items + 1, d, 0, loss="softmax", bidirectional=True, training_pooling_type="mask", **kw)
