"""CPU restatement of the model-based samplers' index build and item draw.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Reference: recstudio/ann/sampler.py
  kmeans :9-36, construct_index :39-45, MIDXSamplerUniform :261-393, MIDXSamplerPop :396-423,
  ClusterSamplerUniform :426-527, ClusterSamplerPop :530-559.
Pinned by tests/golden/midx.npz (outputs of the unmodified reference, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch


def kmeans(X: torch.Tensor, K_or_center, max_iter: int = 300):
    """Lloyd iterations exactly as sampler.py:9-36 (fp32, same torch CPU ops for the distance matrix so the
    assignments are the reference's); returns (C, assign, loss, iterations_run)."""
    N = X.size(0)
    if isinstance(K_or_center, int):
        K = K_or_center
        C = X[torch.randperm(N)[:K]]                                              # :13
    else:
        K = K_or_center.size(0)
        C = K_or_center
    prev = np.inf
    it = 0
    for it in range(max_iter):
        dist = torch.sum(X * X, dim=-1, keepdim=True) - 2 * (X @ C.T) + torch.sum(C * C, dim=-1).unsqueeze(0)   # :19-20
        assign = dist.argmin(-1)                                                  # :21
        loss = torch.sum(torch.square(X - C[assign, :])).item()                   # :24
        if (prev - loss) < prev * 1e-6:                                           # :27
            break
        prev = loss
        count = torch.bincount(assign, minlength=K).to(X.dtype)                   # :30  assign_m.sum(0)
        sums = torch.zeros(K, X.size(1), dtype=X.dtype).index_add_(0, assign, X)  # :31  assign_m.T @ X
        C = sums / count.unsqueeze(-1)
        empty = count < .5                                                        # :32-34
        ndead = int(empty.sum().item())
        C[empty] = X[torch.randperm(N)[:ndead]]
    return C, assign, loss, it + 1


def construct_index(codes, K: int):
    """sampler.py:39-45: stable sort of the bucket codes -> (indices [N], indptr [K+1])."""
    codes = np.asarray(codes, dtype=np.int64)
    indices = np.argsort(codes, kind="stable").astype(np.int64)
    indptr = np.zeros(K + 1, dtype=np.int64)
    np.cumsum(np.bincount(codes, minlength=K), out=indptr[1:])
    return indices, indptr


def bucket_cdf(weight, indices, indptr):
    """sampler.py:300-306 / 417-423: cp = weight[indices], then per bucket cumsum / cumsum[-1]; also the bucket
    totals (= the entries of wkk, :297-300,408)."""
    w = np.asarray(weight, dtype=np.float32)[np.asarray(indices)]
    cp = w.copy()
    nb = len(indptr) - 1
    total = np.zeros(nb, dtype=np.float32)
    for c in range(nb):
        s, e = int(indptr[c]), int(indptr[c + 1])
        if e > s:
            cs = np.cumsum(cp[s:e], dtype=np.float32)
            total[c] = cs[-1]
            cp[s:e] = cs / cs[-1]
    return cp, total


def sample_item_uniform(k01, p01, u, indices, indptr):
    """sample_item without cp (sampler.py:337-343): floor(count * u) inside the bucket, ids + 1."""
    k01 = np.asarray(k01)
    cnt = (indptr[k01 + 1] - indptr[k01])
    idx = np.floor(cnt.astype(np.float32) * np.asarray(u, dtype=np.float32)).astype(np.int32)
    return indices[idx + indptr[k01]] + 1, np.asarray(p01)


def sample_item_with_pop(k01, p01, u, cp, indices, indptr, p):
    """_sample_item_with_pop (sampler.py:348-365) element by element: the row searched is cp[start..last] followed by
    repeats of cp[last]; searchsorted(left); min(item_idx, last); indices[...] WITHOUT +1; p indexed by POSITION + 1."""
    k01 = np.asarray(k01)
    neg = np.empty(k01.shape, dtype=np.int64)
    prob = np.empty(k01.shape, dtype=np.float32)
    start = indptr[k01]
    last = indptr[k01 + 1] - 1
    maxlen = int((last - start + 1).max())
    uu = np.asarray(u, dtype=np.float32)
    for pos in np.ndindex(*k01.shape):
        rng = np.minimum(start[pos] + np.arange(maxlen), last[pos])
        idx = int(np.searchsorted(cp[rng], uu[pos], side="left"))
        idx = min(idx, int(last[pos]))
        neg[pos] = indices[idx + start[pos]]
        prob[pos] = np.float32(np.asarray(p01, dtype=np.float32)[pos]) + np.log(np.float32(p[idx + start[pos] + 1]))
    return neg, prob


def midx_build(emb: torch.Tensor, K: int, c0, c1, max_iter: int, norm=None):
    """MIDXSamplerUniform.update / MIDXSamplerPop._update given the per-item weight ``norm`` (None = uniform)."""
    e1, e2 = torch.chunk(emb, 2, dim=-1)
    c0, cd0, _, _ = kmeans(e1, c0 if c0 is not None else K, max_iter)
    c1, cd1, _, _ = kmeans(e2, c1 if c1 is not None else K, max_iter)
    cd01 = cd0 * K + cd1
    indices, indptr = construct_index(cd01.numpy(), K * K)
    out = {"c0": c0, "c1": c1, "cd0": torch.cat([-cd0.new_ones(1), cd0]) + 1, "cd1": torch.cat([-cd1.new_ones(1), cd1]) + 1,
           "indices": indices, "indptr": indptr}
    if norm is None:
        out["wkk"] = np.diff(indptr).astype(np.float32).reshape(K, K)
    else:
        cp, total = bucket_cdf(norm, indices, indptr)
        out.update(cp=cp, wkk=total.reshape(K, K), p=np.concatenate([np.ones(1, np.float32), np.asarray(norm, np.float32)]))
    return out


def midx_item_p(query: torch.Tensor, pos: torch.Tensor, c0, c1, cd0, cd1, p=None):
    """MIDXSamplerUniform.compute_item_p (sampler.py:367-393), 2-D query."""
    pos_ = pos.unsqueeze(1) if pos.dim() == 1 else pos
    q0, q1 = query.chunk(2, dim=-1)
    c0_ = torch.cat([c0.new_zeros(1, c0.size(1)), c0]); c1_ = torch.cat([c1.new_zeros(1, c1.size(1)), c1])
    r = torch.einsum("bld,bd->bl", c0_[cd0[pos_]], q0) + torch.einsum("bld,bd->bl", c1_[cd1[pos_]], q1)
    if p is not None:
        r = r + torch.log(torch.as_tensor(p)[pos_])
    return r.view_as(pos)
