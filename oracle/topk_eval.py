"""CPU restatement of full-catalog top-k with history mask (T1) and the rank
metrics that consume it (T2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch


# --------------------------------------------------------------------------- T1
def topk_aten(query: torch.Tensor, item_vector: torch.Tensor, k: int, user_hist=None):
    """``BaseRetriever.topk`` without ANN index, InnerProduct scorer
    (recstudio/model/basemodel/baseretriever.py:374-397), op for op.

    ``item_vector`` is ``weight[1:]`` (baseretriever.py:122-123) so column j is
    item id j+1 (:385).  Returns (score [B,k] f32, ids [B,k] i64 1-based).
    """
    more = user_hist.size(1) if user_hist is not None else 0                     # :376
    score, topk_items = torch.topk(torch.matmul(query, item_vector.T), k + more)  # :384
    topk_items = topk_items + 1                                                  # :385
    if user_hist is not None:
        existing, _ = user_hist.sort()                                           # :387
        idx_ = torch.searchsorted(existing, topk_items)                          # :388
        idx_[idx_ == existing.size(1)] = existing.size(1) - 1                    # :389
        score[torch.gather(existing, 1, idx_) == topk_items] = -float("inf")     # :390
        score, idx = score.topk(k)                                               # :391
        topk_items = torch.gather(topk_items, 1, idx)                            # :392
    return score, topk_items


def topk_exact(query, item_vector, k: int, user_hist=None):
    """Order-defined restatement in float64 numpy: rank all items by
    (score descending, id ascending), drop ids present in the user's history
    (padding 0 ignored), keep k.  With tie-free scores this equals
    ``topk_aten`` as long as fewer than ``more`` of the top k+more are masked,
    which always holds because at most ``more`` ids can be masked.

    Returns (score f64 [B,k], ids i64 [B,k] 1-based).
    """
    q = np.asarray(query, dtype=np.float64)
    v = np.asarray(item_vector, dtype=np.float64)
    s = q @ v.T
    B = s.shape[0]
    out_s = np.empty((B, k)); out_i = np.empty((B, k), dtype=np.int64)
    for b in range(B):
        sb = s[b].copy()
        if user_hist is not None:
            h = np.asarray(user_hist[b]); h = h[h > 0]
            sb[h - 1] = -np.inf
        order = np.lexsort((np.arange(sb.shape[0]), -sb))[:k]
        out_s[b] = sb[order]; out_i[b] = order + 1
    return out_s, out_i


def hit_matrix(topk_items: np.ndarray, target: np.ndarray) -> np.ndarray:
    """``_test_step`` label construction (baseretriever.py:422-430): 2-D padded
    targets -> membership; 1-D target -> equality."""
    topk_items = np.asarray(topk_items); target = np.asarray(target)
    if target.ndim > 1:
        lab = np.zeros(topk_items.shape, dtype=bool)
        for b in range(topk_items.shape[0]):
            lab[b] = np.isin(topk_items[b], target[b])
        return lab
    return target.reshape(-1, 1) == topk_items


# --------------------------------------------------------------------------- T2
def _dcg(pred: np.ndarray, k: int) -> np.ndarray:
    k = min(k, pred.shape[1])                                       # eval/__init__.py:105
    denom = np.log2(np.arange(k, dtype=np.float32) + np.float32(2.0)).reshape(1, -1)
    return (pred[:, :k] / denom).sum(-1, dtype=np.float32)


def ndcg(pred, target, k):
    """eval/__init__.py:110-128."""
    pred = np.asarray(pred, dtype=np.float32); target = np.asarray(target)
    pred_dcg = _dcg(pred, k)
    ideal = -np.sort(-(target > 0).astype(np.float32), axis=-1)
    ideal_dcg = _dcg(ideal, k)
    irrelevant = np.all(target <= np.finfo(np.float64).eps, axis=-1)
    out = np.where(irrelevant, np.float32(0), pred_dcg / np.where(irrelevant, np.float32(1), ideal_dcg))
    return out.mean(dtype=np.float32)


def recall(pred, target, k):
    """eval/__init__.py:26-30."""
    pred = np.asarray(pred); target = np.asarray(target)
    count = (target > 0).sum(-1)
    return (pred[:, :k].sum(-1).astype(np.float32) / count).mean(dtype=np.float32)


def precision(pred, target, k):
    """eval/__init__.py:53-56."""
    pred = np.asarray(pred)
    return (pred[:, :k].sum(-1).astype(np.float32) / k).mean(dtype=np.float32)


def map_(pred, target, k):
    """eval/__init__.py:96-101."""
    pred = np.asarray(pred)[:, :k].astype(np.float32); target = np.asarray(target)
    count = (target > 0).sum(-1)
    out = np.cumsum(pred, axis=-1) / np.arange(1, k + 1, dtype=np.float32)
    out = (out * pred).sum(-1) / np.minimum(count, k)
    return out.mean(dtype=np.float32)


def mrr(pred, target, k):
    """eval/__init__.py:143-150: reciprocal rank of the first hit within k."""
    pred = np.asarray(pred)[:, :k]
    first = np.where(pred.any(-1), pred.argmax(-1) + 1, 0)
    out = np.where(first == 0, 0.0, 1.0 / np.maximum(first, 1)).astype(np.float32)
    return out.mean(dtype=np.float32)


def hits(pred, target, k):
    """eval/__init__.py:165."""
    return np.asarray(pred)[:, :k].any(-1).astype(np.float32).mean(dtype=np.float32)


METRICS = {"ndcg": ndcg, "recall": recall, "precision": precision, "map": map_, "mrr": mrr, "hit": hits}
