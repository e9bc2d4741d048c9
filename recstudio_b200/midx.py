"""Model-based samplers: k-means, construct_index and the MIDX / Cluster samplers of
recstudio/ann/sampler.py:9-45,261-527 with their per-epoch index build (``Sampler.update``) on the CUDA
kernels of csrc/midx.cu.

What is replaced, and what is kept as the reference's own small torch ops:

* ``kmeans``           -> rsb200_kmeans_assign / rsb200_kmeans_update: no [N,K] distance matrix, no [N,K]
                          one-hot ``assign_m`` (the reference builds both, sampler.py:19-23,31);
* ``construct_index``  -> rsb200_index_build (stable radix sort + bucket offsets);
* ``_update``          -> ``wkk`` from the bucket totals of the index (the reference: ``cd0m.T @ cd1m``, a
                          [K,N]x[N,K] GEMM against one-hot matrices) and ``cp`` from rsb200_segment_cdf (the
                          reference: a Python loop over K^2 buckets);
* ``_sample_item_with_pop`` -> rsb200_segment_search (the reference gathers [num_q, neg, max_bucket]);
* ``forward`` / ``compute_item_p``: [num_q, K]-sized matmuls, softmax and ``torch.multinomial`` exactly as the
  reference issues them (same RNG call order), on the device the query lives on.

Classes subclass ``iface.Sampler`` (the real ``recstudio.ann.sampler.Sampler`` when importable).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib, iface
from ._lib import check, lib, ptr, stream_ptr


def _need_cuda(t: Tensor, what: str):
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise _lib.Rsb200Error("%s must live on a CUDA device: recstudio_b200 has no CPU fallback" % what)


# ------------------------------------------------------------------------------------------ kernels
def kmeans_assign(X: Tensor, C: Tensor):
    """(assign int64 [N], loss 0-dim float64 tensor) for points X [N, d] (row stride may exceed d) and centers C."""
    _need_cuda(X, "kmeans points")
    if X.dim() != 2 or X.stride(1) != 1 or X.dtype != torch.float32:
        raise _lib.Rsb200Error("kmeans points must be fp32 [N, d] with unit column stride")
    C = C.to(X.device, torch.float32).contiguous()
    N, d = X.shape
    K = C.shape[0]
    assign = torch.empty(N, dtype=torch.int64, device=X.device)
    loss = torch.empty(1, dtype=torch.float64, device=X.device)
    cn = torch.empty(K, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(lib().rsb200_kmeans_assign(ptr(X), X.stride(0), N, d, ptr(C), K, ptr(cn), ptr(assign), ptr(loss), stream_ptr()),
              "kmeans_assign")
    return assign, loss[0]


def kmeans_update(X: Tensor, assign: Tensor, K: int):
    """(sums [K, d], counts [K]) of the points per cluster."""
    _need_cuda(X, "kmeans points")
    N, d = X.shape
    sums = torch.empty(K, d, dtype=torch.float32, device=X.device)
    counts = torch.empty(K, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(lib().rsb200_kmeans_update(ptr(X), X.stride(0), N, d, ptr(assign), K, ptr(sums), ptr(counts), stream_ptr()),
              "kmeans_update")
    return sums, counts


def kmeans(X: Tensor, K_or_center, max_iter: int = 300, verbose: bool = False):
    """``kmeans`` (recstudio/ann/sampler.py:9-36).  Returns ``(C, assign, None, loss)``: the third element of the
    reference's tuple is the dense one-hot ``assign_m`` [N, K], which this implementation never builds (its only
    uses, the centroid sums and ``wkk``, come from the kernels).  Initial centers and the re-seeding of empty clusters
    draw ``torch.randperm(N)`` on the CPU generator exactly where the reference does (also when no cluster is empty),
    so a run seeded like the reference's sees the same centers and leaves the generator in the same state."""
    _need_cuda(X, "kmeans points")
    N = X.size(0)
    if isinstance(K_or_center, int):
        K = K_or_center
        C = X[torch.randperm(N)[:K].to(X.device)]
    else:
        K = K_or_center.size(0)
        C = K_or_center
    prev_loss = np.inf
    assign, loss = None, np.inf
    for it in range(max_iter):
        assign, loss_t = kmeans_assign(X, C)
        loss = loss_t.item()
        if verbose:
            print(f"step:{it:<3d}, loss:{loss:.3f}")
        if (prev_loss - loss) < prev_loss * 1e-6:
            break
        prev_loss = loss
        sums, cluster_count = kmeans_update(X, assign, K)
        C = sums / cluster_count.unsqueeze(-1)
        empty_idx = cluster_count < .5
        ndead = int(empty_idx.sum().item())
        C[empty_idx] = X[torch.randperm(N)[:ndead].to(X.device)]
    return C, assign, None, loss


def construct_index(cd01: Tensor, K: int):
    """``construct_index`` (sampler.py:39-45): ``indices`` = stable argsort of the bucket codes, ``indptr`` [K + 1]."""
    _need_cuda(cd01, "bucket codes")
    codes = cd01.to(torch.int64).contiguous()
    N = codes.numel()
    indices = torch.empty(N, dtype=torch.int64, device=codes.device)
    indptr = torch.empty(K + 1, dtype=torch.int64, device=codes.device)
    nbytes = int(lib().rsb200_index_workspace_bytes(N, K))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=codes.device)
    with torch.cuda.device(codes.device):
        check(lib().rsb200_index_build(ptr(codes), N, K, ptr(indices), ptr(indptr), ptr(ws), nbytes, stream_ptr()), "index_build")
    return indices, indptr


def segment_cdf(weight: Tensor, indices: Tensor, indptr: Tensor):
    """(cp [N], total [buckets]): normalised cumulative ``weight`` inside every bucket of the index."""
    weight = weight.to(torch.float32).contiguous()
    nb = indptr.numel() - 1
    cp = torch.empty(indices.numel(), dtype=torch.float32, device=weight.device)
    total = torch.empty(nb, dtype=torch.float32, device=weight.device)
    with torch.cuda.device(weight.device):
        check(lib().rsb200_segment_cdf(ptr(weight), ptr(indices), ptr(indptr), nb, ptr(cp), ptr(total), stream_ptr()), "segment_cdf")
    return cp, total


def segment_search(k01: Tensor, u: Tensor, cp: Tensor, indices: Tensor, indptr: Tensor, p: Tensor):
    """``_sample_item_with_pop`` (sampler.py:348-365): (neg_items, log(neg_probs)) with k01's shape."""
    k = k01.to(torch.int64).contiguous()
    u = u.to(torch.float32).contiguous()
    neg = torch.empty_like(k)
    logp = torch.empty(k.shape, dtype=torch.float32, device=k.device)
    with torch.cuda.device(k.device):
        check(lib().rsb200_segment_search(ptr(k), ptr(u), k.numel(), ptr(cp), ptr(indices), ptr(indptr), indptr.numel() - 1,
                                          ptr(p), ptr(neg), ptr(logp), stream_ptr()), "segment_search")
    return neg, logp


def _pop_transform(pop_count, mode: int) -> Tensor:
    pop_count = torch.as_tensor(pop_count, dtype=torch.float)
    if mode == 0:
        return torch.log(pop_count + 1)
    if mode == 1:
        return torch.log(pop_count + 1) + 1e-6
    if mode == 2:
        return pop_count ** 0.75
    return pop_count


# ------------------------------------------------------------------------------------------ samplers
class _IndexedSampler(iface.Sampler):
    """Shared pieces of the MIDX / Cluster samplers: bucket index, per-bucket CDF, the final item draw."""

    def _is_cosine(self):
        return isinstance(self.scorer, iface.CosineScorer)

    def _is_euclid(self):
        return isinstance(self.scorer, iface.EuclideanScorer)

    def _norm(self, item_embs: Tensor) -> Optional[Tensor]:
        """per-item weight of the final draw: None = uniform inside the bucket (sampler.py:296-298)."""
        if self._is_euclid():
            return torch.exp(-0.5 * torch.sum(item_embs ** 2, dim=-1))
        return None

    def _build(self, codes: Tensor, num_buckets: int, item_embs: Tensor):
        self.indices, self.indptr = construct_index(codes, num_buckets)
        norm = self._norm(item_embs)
        if norm is None:
            for name in ("p", "cp"):
                if hasattr(self, name):
                    delattr(self, name)
            return (self.indptr[1:] - self.indptr[:-1]).to(torch.float32)      # bucket sizes = one-hot GEMM of the reference
        self.p = torch.cat([norm.new_ones(1), norm], dim=0)                     # :301 (avoids log 0 for the padding id)
        self.cp, total = segment_cdf(norm, self.indices, self.indptr)           # :302-306
        return total

    def sample_item(self, k01, p01, pos=None):                                  # :335-346
        if not hasattr(self, "cp"):
            item_cnt = self.indptr[k01 + 1] - self.indptr[k01]
            item_idx = torch.floor(item_cnt * torch.rand_like(item_cnt.float())).int()
            neg_items = self.indices[item_idx + self.indptr[k01]] + 1
            return neg_items, p01
        return self._sample_item_with_pop(k01, p01)

    def _sample_item_with_pop(self, k01, p01):                                  # :348-365
        u = torch.rand_like(self.indptr[k01].float())
        neg_items, logp = segment_search(k01, u, self.cp, self.indices, self.indptr, self.p)
        return neg_items, p01 + logp


class FusedMIDXSamplerUniform(_IndexedSampler):
    """MIDXSamplerUniform (sampler.py:261-393): two K-way codebooks over the two halves of the item vectors,
    K^2 buckets, uniform draw inside the bucket."""

    def __init__(self, num_items, num_clusters, scorer_fn=None):
        super().__init__(num_items, scorer_fn)
        self.K = num_clusters

    def update(self, item_embs, max_iter=30):
        _need_cuda(item_embs, "item_embs")
        item_embs = item_embs.detach()
        if self._is_cosine():
            item_embs = F.normalize(item_embs, dim=-1)
        embs1, embs2 = torch.chunk(item_embs, 2, dim=-1)
        self.c0, cd0, _, _ = kmeans(embs1, self.c0 if hasattr(self, "c0") else self.K, max_iter)
        self.c1, cd1, _, _ = kmeans(embs2, self.c1 if hasattr(self, "c1") else self.K, max_iter)
        self.c0_ = torch.cat([self.c0.new_zeros(1, self.c0.size(1)), self.c0], dim=0)
        self.c1_ = torch.cat([self.c1.new_zeros(1, self.c1.size(1)), self.c1], dim=0)
        self.cd0 = torch.cat([-cd0.new_ones(1), cd0], dim=0) + 1
        self.cd1 = torch.cat([-cd1.new_ones(1), cd1], dim=0) + 1
        cd01 = cd0 * self.K + cd1
        total = self._build(cd01, self.K ** 2, item_embs)
        self.wkk = total.view(self.K, self.K)          # = cd0m.T @ (cd1m * norm)   (:297,300,408)

    def _draw_buckets(self, flat_query: Tensor, num_neg: int):
        """Two-stage bucket draw (sampler.py:313-327): first-codebook code ~ softmax(q0.c0) weighted by the mass the second
        codebook can reach from it (``wkk``), then the second code given the first.  Returns (bucket [Q, n], logit [Q, n]) where
        the logit q0.c0[k0] + q1.c1[k1] is the un-normalised log-probability of the bucket.  RNG: one
        ``multinomial(.., n, replacement=True)`` over [Q, K], then one ``multinomial(.., 1)`` over [Q*n, K] -- the
        reference's calls in the reference's order."""
        half0, half1 = flat_query.chunk(2, dim=-1)
        logit1 = half1 @ self.c1.T                                  # [Q, K]
        prob1 = torch.softmax(logit1, dim=-1)
        logit0 = half0 @ self.c0.T
        prob0 = torch.softmax(logit0, dim=-1)
        first_w = (prob1 @ self.wkk.T) * prob0                      # mass reachable through each first code
        code0 = torch.multinomial(first_w, num_neg, replacement=True)
        second_w = self.wkk[code0, :] * prob1.unsqueeze(1)          # [Q, n, K]
        code1 = torch.multinomial(second_w.view(-1, second_w.size(-1)), 1).squeeze(-1).view(*second_w.shape[:-1])
        logit = torch.gather(logit0, -1, code0) + torch.gather(logit1, -1, code1)
        return code0 * self.K + code1, logit

    def forward(self, query, num_neg, pos_items=None):
        with torch.no_grad():
            if self._is_cosine():
                query = F.normalize(query, dim=-1)
            bucket, logit = self._draw_buckets(query.view(-1, query.size(-1)), num_neg)
            neg_items, neg_prob = self.sample_item(bucket, logit)
            lead = query.shape[:-1]
            neg_items, neg_prob = neg_items.view(*lead, -1), neg_prob.view(*lead, -1)
            if pos_items is None:
                return neg_items, neg_prob
            return self.compute_item_p(query, pos_items), neg_items, neg_prob

    def compute_item_p(self, query, pos_items):
        """Proposal log-probability (up to the per-query constant) of given items (sampler.py:367-393): the logit of the
        item's bucket, plus log of its in-bucket weight when the final draw is weighted.  Index 0 (padding) maps to the
        all-zero centre prepended in ``c0_`` / ``c1_``."""
        ids = pos_items.unsqueeze(1) if pos_items.dim() == 1 else pos_items
        cent0 = self.c0_[self.cd0[ids], :]                          # [B, L, d/2]
        cent1 = self.c1_[self.cd1[ids], :]
        half0, half1 = query.chunk(2, dim=-1)
        if query.dim() == ids.dim():                                # one query per row, L candidate items
            logit = (torch.bmm(cent0, half0.unsqueeze(-1)) + torch.bmm(cent1, half1.unsqueeze(-1))).squeeze(-1)
        else:                                                       # [B, Lq, d] queries against [B, L] items
            logit = torch.bmm(half0, cent0.transpose(1, 2)) + torch.bmm(half1, cent1.transpose(1, 2))
            ids = ids.unsqueeze(1)
        if hasattr(self, "p"):
            logit = logit + torch.log(self.p[ids])
        return logit.view_as(pos_items)


class FusedMIDXSamplerPop(FusedMIDXSamplerUniform):
    """MIDXSamplerPop (sampler.py:396-423): popularity-weighted draw inside the bucket."""

    def __init__(self, pop_count, num_clusters, scorer=None, mode=1):
        pop_count = torch.as_tensor(pop_count)
        super().__init__(pop_count.shape[0], num_clusters, scorer)
        self.pop_count = torch.nn.Parameter(_pop_transform(pop_count, mode), requires_grad=False)

    def _norm(self, item_embs):
        if not self._is_euclid():
            return self.pop_count.data
        return self.pop_count.data * torch.exp(-0.5 * torch.sum(item_embs ** 2, dim=-1))


class FusedClusterSamplerUniform(_IndexedSampler):
    """ClusterSamplerUniform (sampler.py:426-527): one K-way codebook over the whole item vector."""

    def __init__(self, num_items, num_clusters, scorer_fn=None):
        super().__init__(num_items, scorer_fn)
        self.K = num_clusters

    def update(self, item_embs, max_iter=30):
        _need_cuda(item_embs, "item_embs")
        item_embs = item_embs.detach()
        if self._is_cosine():
            item_embs = F.normalize(item_embs, dim=-1)
        self.c, cd, _, _ = kmeans(item_embs, self.K, max_iter)                  # :436 (always re-seeded, unlike MIDX)
        self.c_ = torch.cat([self.c.new_zeros(1, self.c.size(1)), self.c], dim=0)
        self.cd = torch.cat([-cd.new_ones(1), cd], dim=0) + 1
        self.wkk = self._build(cd, self.K, item_embs)                            # = cdm.sum(0) / (cdm * norm).sum(0)

    def forward(self, query, num_neg, pos_items=None):
        with torch.no_grad():
            if self._is_cosine():
                query = F.normalize(query, dim=-1)
            logit_all = query.view(-1, query.size(-1)) @ self.c.T                      # [Q, K]   (sampler.py:462-466)
            bucket = torch.multinomial(torch.softmax(logit_all, dim=-1), num_neg, replacement=True)
            neg_items, neg_prob = self.sample_item(bucket, torch.gather(logit_all, -1, bucket), pos_items)
            lead = query.shape[:-1]
            neg_items, neg_prob = neg_items.view(*lead, -1), neg_prob.view(*lead, -1)
            if pos_items is None:
                return neg_items, neg_prob
            return self.compute_item_p(query, pos_items), neg_items, neg_prob

    def compute_item_p(self, query, pos_items):
        """sampler.py:473-490: logit of the item's cluster (+ log of its in-cluster weight).  Unlike the reference, the
        weight term is reshaped to the items' shape, so 1-D positives work with a popularity table (the reference adds
        [B] + [B, 1] there and then fails in ``view_as``)."""
        shape = pos_items.shape
        ids = pos_items.view(-1, 1) if pos_items.dim() == 1 else pos_items
        cent = self.c_[self.cd[ids], :]
        if query.dim() == ids.dim():
            logit = torch.bmm(cent, query.unsqueeze(-1)).squeeze(-1)
        else:
            logit = torch.bmm(query, cent.transpose(1, 2))
        logit = logit.reshape(*shape)
        if hasattr(self, "p"):
            logit = logit + torch.log(self.p[pos_items])
        return logit


class FusedClusterSamplerPop(FusedClusterSamplerUniform):
    """ClusterSamplerPop (sampler.py:530-559)."""

    def __init__(self, pop_count, num_clusters, scorer=None, mode=1):
        pop_count = torch.as_tensor(pop_count)
        super().__init__(pop_count.shape[0], num_clusters, scorer)
        self.pop_count = torch.nn.Parameter(_pop_transform(pop_count, mode), requires_grad=False)

    def _norm(self, item_embs):
        if not self._is_euclid():
            return self.pop_count.data
        return self.pop_count.data * torch.exp(-0.5 * torch.sum(item_embs ** 2, dim=-1))
