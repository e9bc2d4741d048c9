"""CPU restatement of the retriever training step (E1, E2, Q1, Q2, L1, L2, L3, R1).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two independent restatements are kept on purpose:

* ``*_aten`` functions follow the reference op for op with the SAME ATen calls
  (``F.embedding``, ``matmul``, ``logsigmoid``, ``logsumexp`` ...) and let
  autograd produce the gradients -- this is what the reference executes on the
  host, so it also serves as bench.py's CPU arm (``cpu_baseline.kind="port"``).
* ``closed_form_*`` functions compute loss and row gradients in float64 numpy
  from the closed forms of SURVEY.md 8(a) -- no autograd, no torch -- so the two
  can be checked against each other and against the golden fixtures.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IP, EUCLID = 0, 1
BPR, SSM = 0, 1


# ------------------------------------------------------------------ Q1 / Q2
def inner_product_score(query: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
    """``InnerProductScorer.forward`` (recstudio/model/scorer.py:5-17): the
    dispatch is on shapes only -- equal leading size => per-row branch."""
    if query.size(0) == items.size(0):
        if query.dim() < items.dim():                                  # :10-12
            out = torch.matmul(items, query.view(*query.shape, 1))
            out = out.view(out.shape[:-1])
        else:                                                          # :13-14
            out = torch.sum(query * items, dim=-1)
    else:                                                              # :15-16
        out = torch.matmul(query, items.T)
    return out


def euclidean_score(query: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
    """``EuclideanScorer.forward`` (scorer.py:28-34) = -||q - v||^2."""
    out = -2 * inner_product_score(query, items)
    out = out + torch.sum(torch.square(items), dim=-1)
    out = out + torch.sum(torch.square(query), dim=-1,
                          keepdim=(query.dim() != items.dim() or query.size(0) != items.size(0)))
    return -out


def score(scorer: int, query, items):
    return inner_product_score(query, items) if scorer == IP else euclidean_score(query, items)


# ------------------------------------------------------------------ L1 / L2 / L3
def bpr_loss(pos_score, neg_score):
    """``BPRLoss.forward`` with dns=False (recstudio/model/loss_func.py:55-59)."""
    loss = F.logsigmoid(pos_score.view(*pos_score.shape, 1) - neg_score)
    weight = F.softmax(torch.ones_like(neg_score), -1)
    return -torch.mean((loss * weight).sum(-1))


def sampled_softmax_loss(pos_score, log_pos_prob, neg_score, log_neg_prob):
    """``SampledSoftmaxLoss.forward`` (loss_func.py:80-90)."""
    new_pos = pos_score - log_pos_prob
    new_neg = neg_score - log_neg_prob
    if new_pos.dim() < new_neg.dim():
        new_pos = new_pos.unsqueeze(-1)
    new_neg = torch.cat([new_pos, new_neg], dim=-1)
    output = torch.logsumexp(new_neg, dim=-1, keepdim=True) - new_pos
    notpadnum = torch.logical_not(torch.isinf(new_pos)).float().sum(-1)
    output = torch.nan_to_num(output, posinf=0).sum(-1) / notpadnum
    return torch.mean(output)


def softmax_loss(pos_score, all_score):
    """``SoftmaxLoss.forward`` 1-D positive branch (loss_func.py:41-42)."""
    assert all_score.dim() > pos_score.dim()
    return torch.mean(torch.logsumexp(all_score, dim=-1) - pos_score)


# ------------------------------------------------------------------ R1 (+E1, E2)
def training_step_aten(w_item: torch.Tensor, w_user: torch.Tensor, user, pos, neg,
                       loss: int = BPR, scorer: int = IP,
                       log_pos_prob=None, log_neg_prob=None, backward: bool = True):
    """``BaseRetriever.training_step`` for the sampler branch with
    ``sampling_method='none'`` and given negatives
    (recstudio/model/basemodel/baseretriever.py:142-176,399-404) followed by
    ``loss.backward()`` (recommender.py:638).

    ``w_item`` / ``w_user`` are leaf tensors; like ``nn.Embedding(padding_idx=0)``
    the gathers go through ``F.embedding(..., padding_idx=0)`` so row 0 gets a
    zero gradient (E2).  Returns dict(loss, pos_score, neg_score, d_item, d_user)
    with DENSE gradients, exactly what the reference leaves in ``.grad``.
    """
    w_item = w_item.detach().clone().requires_grad_(backward)
    w_user = w_user.detach().clone().requires_grad_(backward)
    pos_vec = F.embedding(pos, w_item, padding_idx=0)                    # :154
    query = F.embedding(user, w_user, padding_idx=0)                     # :211
    pos_score = score(scorer, query, pos_vec)                            # :163
    neg_vec = F.embedding(neg, w_item, padding_idx=0)                    # :167-168
    neg_score = score(scorer, query, neg_vec)                            # :169
    if log_pos_prob is None:
        log_pos_prob = torch.zeros_like(pos)                             # sampler.py:113-114
    if log_neg_prob is None:
        log_neg_prob = torch.zeros_like(neg)
    if loss == BPR:
        val = bpr_loss(pos_score, neg_score)
    else:
        val = sampled_softmax_loss(pos_score, log_pos_prob, neg_score, log_neg_prob)
    out = {"loss": val.detach(), "pos_score": pos_score.detach(), "neg_score": neg_score.detach()}
    if backward:
        val.backward()
        out["d_item"] = w_item.grad
        out["d_user"] = w_user.grad
    return out


def full_softmax_step_aten(w_item, w_user, user, pos, backward: bool = True):
    """Full-score branch (baseretriever.py:177-186) + ``SoftmaxLoss`` (L3):
    ``all_score = query @ weight[1:].T``; column j is item j+1 (fact 4)."""
    w_item = w_item.detach().clone().requires_grad_(backward)
    w_user = w_user.detach().clone().requires_grad_(backward)
    pos_vec = F.embedding(pos, w_item, padding_idx=0)
    query = F.embedding(user, w_user, padding_idx=0)
    pos_score = inner_product_score(query, pos_vec)
    all_score = inner_product_score(query, w_item[1:])
    val = softmax_loss(pos_score, all_score)
    out = {"loss": val.detach(), "pos_score": pos_score.detach(),
           "lse": torch.logsumexp(all_score.detach(), dim=-1)}
    if backward:
        val.backward()
        out["d_item"] = w_item.grad
        out["d_user"] = w_user.grad
    return out


# ------------------------------------------------------------------ closed forms
def _softplus64(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def closed_form_step(w_item, w_user, user, pos, neg, loss: int = BPR, scorer: int = IP,
                     log_pos_prob=None, log_neg_prob=None):
    """float64 numpy closed forms of SURVEY.md 8(a) L1/L2 with Q1/Q2:

    BPR:  L = mean_{b,j} softplus(s-_bj - s+_b);  dL/ds-_bj = sigmoid(s-_bj - s+_b)/(B n)
    SSM:  z = score - logQ; L = mean_b (lse_j z_bj - z_b0); dL/ds-_bj = p_bj / B,
          dL/ds+_b = (p_b0 - 1)/B
    IP:   ds/dq = v, ds/dv = q.     Euclid: s = -|q-v|^2, ds/dq = 2(v-q), ds/dv = 2(q-v).
    Row 0 (padding) receives no gradient (E2).

    Returns dict(loss, pos_score [B], neg_score [B,n], coef_neg [B,n], coef_pos [B],
    d_item {row: vec}, d_user {row: vec}) with dict-of-rows sparse gradients.
    """
    W = np.asarray(w_item, dtype=np.float64)
    U = np.asarray(w_user, dtype=np.float64)
    user = np.asarray(user); pos = np.asarray(pos); neg = np.asarray(neg)
    B, n = neg.shape
    q = U[user]                       # [B,d]
    vp = W[pos]                       # [B,d]
    vn = W[neg]                       # [B,n,d]
    if scorer == IP:
        sp = (q * vp).sum(-1)
        sn = np.einsum("bd,bnd->bn", q, vn)
    else:
        sp = -((q - vp) ** 2).sum(-1)
        sn = -((q[:, None, :] - vn) ** 2).sum(-1)
    if loss == BPR:
        x = sn - sp[:, None]
        val = _softplus64(x).mean()
        cn = 1.0 / (1.0 + np.exp(-x)) / (B * n)
        cp = -cn.sum(-1)
    else:
        lp = np.zeros(B) if log_pos_prob is None else np.asarray(log_pos_prob, dtype=np.float64)
        ln = np.zeros((B, n)) if log_neg_prob is None else np.asarray(log_neg_prob, dtype=np.float64)
        z = np.concatenate([(sp - lp)[:, None], sn - ln], axis=1)
        m = z.max(-1, keepdims=True)
        lse = m[:, 0] + np.log(np.exp(z - m).sum(-1))
        val = (lse - z[:, 0]).mean()
        p = np.exp(z - lse[:, None])
        cn = p[:, 1:] / B
        cp = (p[:, 0] - 1.0) / B
    if scorer == IP:
        dq = (cn[:, :, None] * vn).sum(1) + cp[:, None] * vp
        dvn = cn[:, :, None] * q[:, None, :]
        dvp = cp[:, None] * q
    else:
        dq = 2.0 * ((cn[:, :, None] * (vn - q[:, None, :])).sum(1) + cp[:, None] * (vp - q))
        dvn = 2.0 * cn[:, :, None] * (q[:, None, :] - vn)
        dvp = 2.0 * cp[:, None] * (q - vp)
    d_item, d_user = {}, {}
    for b in range(B):
        r = int(pos[b])
        if r != 0:
            d_item[r] = d_item.get(r, 0.0) + dvp[b]
        u = int(user[b])
        if u != 0:
            d_user[u] = d_user.get(u, 0.0) + dq[b]
        for j in range(n):
            r = int(neg[b, j])
            if r != 0:
                d_item[r] = d_item.get(r, 0.0) + dvn[b, j]
    return {"loss": val, "pos_score": sp, "neg_score": sn, "coef_neg": cn, "coef_pos": cp,
            "dq": dq, "d_item": d_item, "d_user": d_user}


def dense_from_rows(rows, vals, shape):
    """Scatter a (rows, vals) sparse-row gradient into a dense float64 array."""
    out = np.zeros(shape, dtype=np.float64)
    np.add.at(out, np.asarray(rows, dtype=np.int64), np.asarray(vals, dtype=np.float64))
    return out


def dict_to_dense(d, shape):
    out = np.zeros(shape, dtype=np.float64)
    for r, v in d.items():
        out[r] = v
    return out


# ------------------------------------------------------------------ sampling methods (8(f)-2)
def select_from_pool(method: str, query, w_item, pool, num_keep: int, scorer: int = IP, resampled_id=None):
    """The pool stage of ``BaseRetriever.sampling`` for ``method`` 'dns' / 'sir'
    (recstudio/model/basemodel/baseretriever.py:331-355) given the drawn pool:
      scores = score_func(query, item_encoder(pool))                       :323-324
      dns: neg = pool[topk(scores, num_keep)], log_neg_prob = int64 zeros  :326-330
      sir: p = softmax(scores + eps); idx = multinomial(p, num_keep); neg = pool[idx],
           log_neg_prob = scores[idx]                                      :336-342
    ``resampled_id`` injects the multinomial outcome (its stream is device specific).
    Returns dict(scores, neg_id, log_neg_prob, probs)."""
    pool = torch.as_tensor(pool, dtype=torch.int64)
    vec = F.embedding(pool, w_item, padding_idx=0)
    scores = score(scorer, query, vec)
    out = {"scores": scores}
    if method == "dns":
        _, topk_id = torch.topk(scores, num_keep)
        out["neg_id"] = torch.gather(pool, -1, topk_id)
        out["log_neg_prob"] = torch.zeros_like(out["neg_id"])
    elif method == "sir":
        probs = torch.softmax(scores + torch.finfo(torch.float32).eps, dim=-1)
        out["probs"] = probs
        if resampled_id is None:
            resampled_id = torch.multinomial(probs, num_keep, replacement=True)
        resampled_id = torch.as_tensor(resampled_id, dtype=torch.int64)
        out["neg_id"] = torch.gather(pool, -1, resampled_id)
        out["log_neg_prob"] = torch.gather(scores, -1, resampled_id)
    else:
        raise ValueError(method)
    return out
