"""T2: rank metrics on the boolean hit matrix of the fused top-k.

SURVEY.md 8(a) T2 keeps these as the reference's PyTorch (negligible work on [Be, k]
tensors).  When ``recstudio`` is importable the reference functions are used directly;
this module mirrors their names and semantics (recstudio/eval/__init__.py:9-165,245-251)
for environments where it is not.
"""
from __future__ import annotations

import sys

import torch


def recall(pred, target, k):                      # eval/__init__.py:26-30
    count = (target > 0).sum(-1)
    return (pred[:, :k].sum(dim=-1).float() / count).mean()


def precision(pred, target, k):                   # :53-56
    return (pred[:, :k].sum(dim=-1).float() / k).mean()


def map(pred, target, k):                         # :96-101  (name shadows the builtin, as in the reference)
    count = (target > 0).sum(-1)
    pred = pred[:, :k].float()
    output = pred.cumsum(dim=-1) / torch.arange(1, k + 1).type_as(pred)
    output = (output * pred).sum(dim=-1) / torch.minimum(count, k * torch.ones_like(count))
    return output.mean()


def _dcg(pred, k):                                # :104-107
    k = min(k, pred.size(1))
    denom = torch.log2(torch.arange(k).type_as(pred) + 2.0).view(1, -1)
    return (pred[:, :k] / denom).sum(dim=-1)


def ndcg(pred, target, k):                        # :110-128
    pred_dcg = _dcg(pred.float(), k)
    ideal_dcg = _dcg(torch.sort((target > 0).float(), descending=True)[0], k)
    all_irrelevant = torch.all(target <= sys.float_info.epsilon, dim=-1)
    pred_dcg[all_irrelevant] = 0
    pred_dcg[~all_irrelevant] /= ideal_dcg[~all_irrelevant]
    return pred_dcg.mean()


def mrr(pred, target, k):                         # :131-150
    row, col = torch.nonzero(pred[:, :k], as_tuple=True)
    row_uniq, counts = torch.unique_consecutive(row, return_counts=True)
    idx = torch.zeros_like(counts)
    idx[1:] = counts.cumsum(dim=-1)[:-1]
    first = col.new_zeros(pred.size(0)).scatter_(0, row_uniq, col[idx] + 1)
    output = 1.0 / first
    output[first == 0] = 0
    return output.mean()


def hits(pred, target, k):                        # :153-165
    return torch.any(pred[:, :k] > 0, dim=-1).float().mean()


metric_dict = {"ndcg": ndcg, "precision": precision, "recall": recall, "map": map, "hit": hits, "mrr": mrr}


def get_rank_metrics(metric):                     # :245-251
    if not isinstance(metric, list):
        metric = [metric]
    return [(m, metric_dict[m]) for m in metric if m in metric_dict]
