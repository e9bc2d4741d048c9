"""Retriever hosts for the fused plugins.

``FusedRetriever`` / ``FusedBPR`` are real ``BaseRetriever`` subclasses (``FusedRetrieverMixin`` in front of the
reference class, recstudio/model/basemodel/baseretriever.py:14-431), so the reference's own trainer
(``fit`` / ``training_epoch`` / ``evaluate``, recommender.py:594-691) drives them unchanged; ``forward``, ``_sample``,
``_test_step``, the rank metrics (recstudio.eval) and ``_to_device`` are the reference's code, not copies.
The ``build_*_synthetic`` helpers construct them over synthetic table sizes without a dataset object, the way
test/test_retriever.py:13-19 does with kwargs.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import iface, plugins


class FusedRetriever(plugins.FusedRetrieverMixin, iface.BaseRetriever):
    """``BaseRetriever`` with the fused CUDA training step.  Construct it exactly like the
    reference (test/test_retriever.py:13-19): kwargs ``item_encoder=, query_encoder=, scorer=,
    sampler=, loss=`` -- any of them may be a reference plugin, the fused path engages when the
    combination is recognised."""

    def __init__(self, config: Dict = None, fused_grad: str = "dense", device_loader: bool = False, fused_graph: bool = False,
                 **kwargs):
        super().__init__(config, **kwargs)
        if fused_grad not in ("dense", "sparse", "rows", "apply"):
            raise ValueError("fused_grad must be 'dense', 'sparse', 'rows' or 'apply'")
        self.fused_grad = fused_grad
        self.device_loader = device_loader
        # replay the fused step from two captured CUDA graphs (forward | backward) when the batch shape is stable:
        # fused_grad='rows', two embedding towers, FusedUniformSampler (plugins._GraphedStepFn)
        self.fused_graph = bool(fused_graph)

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
        import warnings
        from .rowopt import FusedRowOptimizer
        tr = self.config["train"]
        name = str(tr.get("learner", "adam")).lower()
        learner = {"sgd": "sgd", "adagrad": "adagrad", "adam": "sparse_adam", "sparse_adam": "sparse_adam"}.get(name)
        if learner is None:
            raise ValueError("fused_grad=%r steps the embedding tables with a touched-row optimizer: learner must be one of "
                             "sgd / adagrad / adam / sparse_adam, got %r" % (self.fused_grad, name))
        if name == "adam":
            warnings.warn("fused_grad=%r: the embedding tables are stepped with SparseAdam semantics (moments of untouched rows do "
                          "not decay), not the reference's dense Adam; use fused_grad='dense' for the reference's optimizer"
                          % self.fused_grad, stacklevel=2)
        wd = tr.get("weight_decay", 0) or 0
        if wd:
            warnings.warn("fused_grad=%r: train.weight_decay=%g is NOT applied to the embedding tables (a touched-row optimizer "
                          "cannot decay untouched rows); it still applies to every other parameter" % (self.fused_grad, wd), stacklevel=2)
        lr = tr.get("learning_rate", 0.001)
        row_opt = FusedRowOptimizer(self, learner, lr=lr)
        opts = [{"optimizer": row_opt}]
        sched = self._get_scheduler(tr.get("scheduler", None), row_opt)     # torch schedulers only touch param_groups[i]['lr']
        if sched:
            m = self.val_metric if getattr(self, "val_check", False) else "train_loss"      # set by fit() (recommender.py:129-136)
            opts[0]["lr_scheduler"] = {"scheduler": sched, "monitor": m, "interval": "epoch", "frequency": 1, "strict": False}
        tables = {id(self.item_encoder.weight)}
        if isinstance(self.query_encoder, torch.nn.Embedding):
            tables.add(id(self.query_encoder.weight))
        rest = [p for p in self.parameters() if id(p) not in tables]
        if rest:
            opt = self._get_optimizer(name, rest, lr, tr.get("weight_decay", 0))
            ent = {"optimizer": opt}
            sched = self._get_scheduler(tr.get("scheduler", None), opt)
            if sched:
                m = self.val_metric if getattr(self, "val_check", False) else "train_loss"
                ent["lr_scheduler"] = {"scheduler": sched, "monitor": m, "interval": "epoch", "frequency": 1, "strict": False}
            opts.append(ent)
        return opts


class FusedBPR(FusedRetriever):
    """The reference's BPR (recstudio/model/mf/bpr.py:7-25) on the fused path."""


def build_synthetic(num_users: int, num_items: int, d: int, n, loss: str = "bpr", scorer: str = "ip",
                    sampler: str = "uniform", pop_count=None, fused_grad: str = "dense", device="cuda:0",
                    init_std: Optional[float] = None, seed: int = 2022, sampling_method: str = "none",
                    excluding_hist: bool = False, fused_graph: bool = False) -> FusedRetriever:
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
    from recstudio.utils import get_model
    conf = get_model("BPR")[1]
    conf["train"].update({"negative_count": n, "gpu": None, "seed": seed, **extra})
    conf["model"]["embed_dim"] = d
    m = FusedRetriever(conf, fused_grad=fused_grad, fused_graph=fused_graph, item_encoder=item, query_encoder=user, scorer=score_m,
                       sampler=samp_m, loss=loss_m)
    # what _init_model reads off the dataset object (recommender.py:66-77, baseretriever.py:54-68)
    m.fuid, m.fiid, m.frating = "user_id", "item_id", "rating"
    m.item_fields, m.query_fields, m.neg_count = {"item_id"}, {"user_id"}, n
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
    from recstudio.utils import get_model
    conf = get_model("SASRec")[1]
    conf["train"].update({"negative_count": n, "gpu": None, "seed": seed})
    conf["model"]["embed_dim"] = d
    m = FusedRetriever(conf, fused_grad=fused_grad, **kwargs)
    m.fuid, m.fiid, m.frating = "user_id", "item_id", "rating"
    m.item_fields, m.neg_count = {"item_id"}, n
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
