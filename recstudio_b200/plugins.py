"""Drop-in replacements for the reference's plugin surfaces on the retriever hot path.

Every class subclasses the reference class it replaces (``iface.py`` binds the real
``recstudio`` classes) so that ``BaseRetriever(config,
item_encoder=, query_encoder=, scorer=, sampler=, loss=)`` accepts them unchanged
(recstudio/model/basemodel/baseretriever.py:14-46, recommender.py:48-54):

  FusedEmbedding(nn.Embedding)                E1/E2   baseretriever.py:84,104
  FusedUniformSampler(Sampler)                S1      recstudio/ann/sampler.py:81-114
  FusedPopularSampler(Sampler)                S2      recstudio/ann/sampler.py:224-258
  FusedInnerProductScorer / FusedEuclideanScorer      recstudio/model/scorer.py:5-17,28-34
  FusedBPRLoss / FusedSampledSoftmaxLoss(PairwiseLoss) recstudio/model/loss_func.py:50-63,80-90
  FusedRetrieverMixin                          R1/T1   baseretriever.py:142-192,374-404

Each plugin is a correct standalone op (its own CUDA kernel + analytic backward) so any
mix with reference plugins still works; ``FusedRetrieverMixin.training_step`` takes the
single fused path (rsb200_pair_step) when it recognises the combination.  There is no
CPU path: every op raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
from torch import Tensor

from . import _lib, fused, iface, sampling
from ._lib import LOSS_BPR, LOSS_SSM, SCORE_EUCLID, SCORE_IP, check, lib, ptr, stream_ptr


def _need_cuda(t: Tensor, what: str):
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise _lib.Rsb200Error("%s must live on a CUDA device: recstudio_b200 has no CPU fallback" % what)


# ============================================================================ E1 / E2
class _GatherFn(torch.autograd.Function):
    """F.embedding forward (gather) / embedding backward.  The backward follows ``emb.grad_mode``:
    'dense'  embedding_dense_backward (scatter-add into a zero-filled [N, d] gradient, row 0 skipped) -- the reference;
    'sparse' the touched rows only, as a coalesced sparse-COO gradient (rsb200_rows_coalesce);
    'rows'   the touched rows are left in ``emb._extra_row_grads`` for FusedRowOptimizer (no ``.grad`` at all)."""

    @staticmethod
    def forward(ctx, weight: Tensor, ids: Tensor, emb):
        _need_cuda(weight, "embedding table")
        ids_c = ids.to(device=weight.device, dtype=torch.int64).contiguous()
        out = torch.empty(ids_c.shape + (weight.shape[1],), dtype=torch.float32, device=weight.device)
        with torch.cuda.device(weight.device):
            check(lib().rsb200_gather_rows(ptr(weight), weight.shape[0], weight.shape[1], ptr(ids_c), ids_c.numel(),
                                           ptr(out), stream_ptr()), "gather_rows")
        ctx.save_for_backward(ids_c)
        ctx.wshape = tuple(weight.shape)
        ctx.emb = emb
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        (ids_c,) = ctx.saved_tensors
        g = g.contiguous()
        mode = getattr(ctx.emb, "grad_mode", "dense")
        if mode != "dense":
            from .sharded import CudaOps
            rows, vals = CudaOps.coalesce_rows(ids_c.reshape(-1), g.reshape(-1, ctx.wshape[1]), ctx.wshape[0], skip_row0=True)
            if mode == "rows":
                ctx.emb.__dict__.setdefault("_extra_row_grads", []).append((rows, vals))
                return None, None, None
            gi = torch.sparse_coo_tensor(rows.unsqueeze(0), vals, size=ctx.wshape, is_coalesced=True, check_invariants=False)
            return gi, None, None
        dw = torch.zeros(ctx.wshape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib().rsb200_scatter_add_rows(ptr(dw), ctx.wshape[0], ctx.wshape[1], ptr(ids_c), ids_c.numel(),
                                                ptr(g), stream_ptr()), "scatter_add_rows")
        return dw, None, None


class FusedEmbedding(torch.nn.Embedding):
    """``nn.Embedding(num, d, padding_idx=0)`` with CUDA gather / scatter kernels.

    It IS an ``nn.Embedding`` (init.py:6,24,37 dispatches on that; ``_get_item_vector``
    takes its O(1) ``weight[1:]`` path only for nn.Embedding, baseretriever.py:122-123),
    ``.weight`` stays a live Parameter for ``state_dict`` / optimizers / ``model.to()``.
    Accepts index tensors of any rank (SASRec calls it on [B, L], sasrec.py:42).
    ``grad_mode`` ('dense' | 'sparse' | 'rows') selects the form of the backward, see ``_GatherFn``; a FusedRetriever
    sets it from its ``fused_grad`` when the table is shared with a sequence encoder.
    """
    grad_mode = "dense"

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = 0, **kw):
        if embedding_dim % 4 != 0:
            raise _lib.Rsb200Error("embedding_dim must be a multiple of 4 (16-byte rows)")
        if padding_idx not in (0, None):
            raise _lib.Rsb200Error("only padding_idx=0 (the reference's convention) is supported")
        super().__init__(num_embeddings, embedding_dim, padding_idx=0, **kw)

    def forward(self, ids: Tensor) -> Tensor:
        return _GatherFn.apply(self.weight, ids, self)


# ============================================================================ S1 / S2
class FusedUniformSampler(iface.Sampler):
    """UniformSampler (sampler.py:81-114).  Indices are bit-identical to the reference's
    ``torch.randint(1, N, (Q, n), device=cuda)`` for the same generator state, and the
    log-probabilities are int64 zeros exactly like ``torch.zeros_like(ids)``."""

    def forward(self, query, num_neg: int, pos_items: Optional[Tensor] = None, device=None):
        if isinstance(query, int):
            num_queries = query
            device = pos_items.device if pos_items is not None else device
            shape = (num_queries,)
        else:
            num_queries = int(np.prod(query.shape[:-1]))
            device = query.device
            shape = tuple(query.shape[:-1])
        neg, _ = sampling.uniform_draw(self.num_items + 1, num_queries, num_neg, device)
        neg = neg.reshape(*shape, -1)
        neg_prob = torch.zeros_like(neg)
        if pos_items is not None:
            return torch.zeros_like(pos_items), neg, neg_prob
        return neg, neg_prob

    def compute_item_p(self, query, pos_items):
        return torch.zeros_like(pos_items)

    # fused-path hook: int32 ids only, no log-prob tensors (they are identically zero)
    def fused_draw(self, num_queries: int, num_neg: int, device):
        _, neg32 = sampling.uniform_draw(self.num_items + 1, num_queries, num_neg, device, want_i64=False, want_i32=True)
        return neg32, None


class FusedPopularSampler(iface.Sampler):
    """PopularSamplerModel (sampler.py:224-258) including its two quirks: ``pop_count[0] = 1``
    (padding id 0 is drawable) and ``pop_prob[-1] = 1.0`` applied after the cumsum.  The
    tables are built with the very same torch ops as the reference constructor; the draw is
    ``searchsorted(table, torch.rand(...))`` restated with a guide table (identical results)."""

    def __init__(self, pop_count, scorer=None, mode: int = 0):
        super().__init__(pop_count.shape[0], scorer)
        with torch.no_grad():
            pop_count = torch.tensor(pop_count, dtype=torch.float)
            if mode == 0:
                pop_count = torch.log(pop_count + 1)
            elif mode == 1:
                pop_count = torch.log(pop_count + 1) + 1e-6
            elif mode == 2:
                pop_count = pop_count ** 0.75
            pop_count[0] = 1
            self.register_buffer("pop_prob", pop_count / pop_count.sum())
            self.register_buffer("table", torch.cumsum(self.pop_prob, dim=0))
            self.pop_prob[-1] = 1.0
        self._guide = None
        self._guide_bits = 0

    def _guide_for(self, device):
        _need_cuda(self.table, "sampler table (call model.to(cuda))")
        if self._guide is None or self._guide.device != self.table.device:
            self._guide, self._guide_bits = sampling.build_guide(self.table)
        return self._guide, self._guide_bits

    def forward(self, query, num_neg: int, pos_items: Optional[Tensor] = None):
        num_queries = int(np.prod(query.shape[:-1]))
        guide, bits = self._guide_for(query.device)
        neg, _, neg_prob = sampling.popular_draw(self.table, self.pop_prob, num_queries, num_neg, guide, bits)
        neg = neg.reshape(*query.shape[:-1], -1)
        neg_prob = neg_prob.reshape(*query.shape[:-1], -1)
        if pos_items is not None:
            return self.compute_item_p(query, pos_items), neg, neg_prob
        return neg, neg_prob

    def compute_item_p(self, query, pos_items):
        return sampling.popular_logq(self.pop_prob, pos_items)

    def fused_draw(self, num_queries: int, num_neg: int, device):
        guide, bits = self._guide_for(device)
        _, neg32, logq = sampling.popular_draw(self.table, self.pop_prob, num_queries, num_neg, guide, bits,
                                               want_i64=False, want_i32=True)
        return neg32, logq


class FusedMaskedUniformSampler(iface.Sampler):
    """MaskedUniformSampler (sampler.py:187-214): uniform negatives that exclude each user's history.
    ``query`` is [B, d] (one draw row per user) or [B, L, d] (``L`` queries per user, result [B, L, n]);
    ``user_hist`` is the 0-padded int64 history matrix of the batch.  Ids are bit-identical to the reference
    on a CUDA device for the same generator state (rsb200_sample_uniform_masked); the log-probabilities are
    the reference's ``-log(ones_like(ids))`` (float32 negative zeros)."""

    def forward(self, query, num_neg: int, pos_items: Optional[Tensor] = None, user_hist: Optional[Tensor] = None):
        if user_hist is None:
            raise ValueError("MaskedUniformSampler needs `user_hist` (train with excluding_hist=True)")
        if query.dim() == 2:
            per_user, shape = num_neg, (user_hist.shape[0], num_neg)
        elif query.dim() == 3:
            per_user, shape = query.size(1) * num_neg, (user_hist.shape[0], query.size(1), num_neg)
        else:
            raise ValueError("`query` need to be 2-dimensional or 3-dimensional.")
        _need_cuda(user_hist, "user_hist")
        neg, _ = sampling.masked_uniform_draw(self.num_items + 1, user_hist, per_user)
        neg = neg.reshape(shape)
        neg_prob = self.compute_item_p(query, neg)
        if pos_items is not None:
            return self.compute_item_p(query, pos_items), neg, neg_prob
        return neg, neg_prob

    def compute_item_p(self, query, pos_items):
        return -torch.log(torch.ones_like(pos_items))

    def fused_draw(self, num_queries: int, num_neg: int, device, user_hist: Optional[Tensor] = None):
        if user_hist is None or user_hist.shape[0] != num_queries:
            raise ValueError("MaskedUniformSampler.fused_draw needs the batch's user_hist [B, H]")
        _, neg32 = sampling.masked_uniform_draw(self.num_items + 1, user_hist.to(device), num_neg, want_i64=False, want_i32=True)
        return neg32, None


# ============================================================================ Q1 / Q2
class _ScoreDenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query: Tensor, items: Tensor, kind: int):
        q2 = query.reshape(-1, query.shape[-1]).contiguous()
        B, d = q2.shape
        it = items.reshape(B, -1, d).contiguous()
        n = it.shape[1]
        out = torch.empty(B, n, dtype=torch.float32, device=q2.device)
        with torch.cuda.device(q2.device):
            check(lib().rsb200_score_dense(kind, ptr(q2), ptr(it), B, n, d, ptr(out), stream_ptr()), "score_dense")
        ctx.save_for_backward(q2, it)
        ctx.kind, ctx.qshape, ctx.ishape = kind, query.shape, items.shape
        return out.reshape(items.shape[:-1])

    @staticmethod
    def backward(ctx, g: Tensor):
        q2, it = ctx.saved_tensors
        B, n, d = it.shape
        g2 = g.reshape(B, n).contiguous()
        dq = torch.empty_like(q2)
        di = torch.empty_like(it)
        with torch.cuda.device(q2.device):
            check(lib().rsb200_score_dense_bwd(ctx.kind, ptr(q2), ptr(it), ptr(g2), B, n, d, ptr(dq), ptr(di),
                                               stream_ptr()), "score_dense_bwd")
        return dq.reshape(ctx.qshape), di.reshape(ctx.ishape), None


def _score_forward(kind: int, query: Tensor, items: Tensor) -> Tensor:
    """Shape dispatch of InnerProductScorer.forward (scorer.py:9-17): equal leading size =>
    per-row branch (CUDA kernel), else the full-score [B,D]x[N,D]^T branch (a plain library
    GEMM: cuBLAS through torch.matmul)."""
    _need_cuda(query, "query")
    _need_cuda(items, "items")
    if query.size(0) == items.size(0):
        if query.dim() == items.dim() or query.dim() + 1 == items.dim():
            return _ScoreDenseFn.apply(query, items, kind)
        raise _lib.Rsb200Error("unsupported scorer shapes %s x %s" % (tuple(query.shape), tuple(items.shape)))
    if kind == SCORE_IP:
        return torch.matmul(query, items.T)
    out = 2 * torch.matmul(query, items.T)
    out = out - torch.sum(torch.square(items), dim=-1)
    out = out - torch.sum(torch.square(query), dim=-1, keepdim=True)
    return out


class FusedInnerProductScorer(iface.InnerProductScorer):
    fused_kind = SCORE_IP

    def forward(self, query, items):
        return _score_forward(SCORE_IP, query, items)


class FusedEuclideanScorer(iface.EuclideanScorer):
    fused_kind = SCORE_EUCLID

    def forward(self, query, items):
        return _score_forward(SCORE_EUCLID, query, items)


# ============================================================================ L1 / L2
class _PairLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos_score: Tensor, neg_score: Tensor, lqp, lqn, kind: int):
        _need_cuda(pos_score, "pos_score")
        if pos_score.dim() != 1 or neg_score.dim() != 2 or neg_score.shape[0] != pos_score.shape[0]:
            raise _lib.Rsb200Error("fused pairwise losses take pos_score [B] and neg_score [B, n] "
                                   "(got %s, %s)" % (tuple(pos_score.shape), tuple(neg_score.shape)))
        ps, ns = pos_score.contiguous().float(), neg_score.contiguous().float()
        B, n = ns.shape
        dev = ps.device
        lqp = None if lqp is None else lqp.to(torch.float32).contiguous()
        lqn = None if lqn is None else lqn.to(torch.float32).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        d_pos = torch.empty_like(ps)
        d_neg = torch.empty_like(ns)
        part = torch.empty(B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().rsb200_pair_loss(kind, ptr(ps), ptr(ns), ptr(lqp), ptr(lqn), B, n, ptr(loss), ptr(d_pos),
                                         ptr(d_neg), ptr(part), stream_ptr()), "pair_loss")
        ctx.save_for_backward(d_pos, d_neg)
        return loss[0]

    @staticmethod
    def backward(ctx, g: Tensor):
        d_pos, d_neg = ctx.saved_tensors
        return g * d_pos, g * d_neg, None, None, None


class FusedBPRLoss(iface.PairwiseLoss):
    """BPRLoss(dns=False) (loss_func.py:50-59): -mean_{b,j} logsigmoid(s+_b - s-_bj)."""
    fused_kind = LOSS_BPR

    def __init__(self, dns: bool = False):
        super().__init__()
        if dns:
            raise _lib.Rsb200Error("BPRLoss(dns=True) is outside the fused path; use the reference class")
        self.dns = False

    def forward(self, label, pos_score, log_pos_prob, neg_score, log_neg_prob):
        return _PairLossFn.apply(pos_score, neg_score, None, None, LOSS_BPR)


class FusedSampledSoftmaxLoss(iface.PairwiseLoss):
    """SampledSoftmaxLoss (loss_func.py:80-90) for 1-D positives."""
    fused_kind = LOSS_SSM

    def forward(self, label, pos_score, log_pos_prob, neg_score, log_neg_prob):
        return _PairLossFn.apply(pos_score, neg_score, log_pos_prob, log_neg_prob, LOSS_SSM)


class FusedSoftmaxLoss(iface.FullScoreLoss):
    """SoftmaxLoss (loss_func.py:39-42), 1-D positives.  Standalone: one pass over a materialised
    ``all_score`` row block; inside FusedRetrieverMixin the [B, N-1] matrix is never built
    (rsb200_fullsoftmax_fwd_bwd)."""
    fused_kind = _lib.LOSS_FULL

    def forward(self, label, pos_score, all_score):
        if all_score.dim() != pos_score.dim() + 1:
            raise _lib.Rsb200Error("FusedSoftmaxLoss handles 1-D positives (all_score [B, N-1]); "
                                   "use the reference class for the padded 2-D branch")
        return _PairLossFn.apply(pos_score, all_score, None, None, _lib.LOSS_FULL)


class _FullSoftmaxFn(torch.autograd.Function):
    """L3 + Q1-full: loss = mean_b(logsumexp_i q_b.w_i - q_b.w_pos_b) with dQ and dense dW computed in
    the same call; the [B, N-1] score matrix never exists."""

    @staticmethod
    def forward(ctx, query: Tensor, w_item: Tensor, pos: Tensor):
        _need_cuda(query, "query")
        q = query.contiguous().float()
        B, d = q.shape
        N = w_item.shape[0]
        dev = q.device
        pos = pos.to(dev, torch.int64).contiguous()
        nbytes = int(lib().rsb200_fullsoftmax_workspace_bytes(B, N, d))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dq = torch.empty_like(q)
        dw = torch.empty(N, d, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().rsb200_fullsoftmax_fwd_bwd(ptr(q), ptr(w_item.detach()), ptr(pos), N, B, d, ptr(loss), ptr(dq), ptr(dw),
                                                   ptr(ws), nbytes, stream_ptr()), "fullsoftmax_fwd_bwd")
        ctx.save_for_backward(dq, dw)
        return loss[0]

    @staticmethod
    def backward(ctx, g: Tensor):
        dq, dw = ctx.saved_tensors
        return g * dq, g * dw, None


# ============================================================================ R1: the fused step
class _FusedStepFn(torch.autograd.Function):
    """forward = PHASE_COUNT|SCAN|FWD (loss + gradient coefficients), backward = PHASE_SCATTER
    (gradient rows), scaled by autograd's grad_output read on the device (no host sync)."""

    @staticmethod
    def forward(ctx, w_item: Tensor, w_user: Tensor, host, ws, user, pos, neg32, lqp, lqn, loss_kind, score_kind):
        loss = fused.pair_step(ws, w_item, w_user, user, pos, neg32, loss_kind, score_kind, logq_pos=lqp,
                               logq_neg=lqn, phases=_lib.PHASE_COUNT | _lib.PHASE_SCAN | _lib.PHASE_FWD,
                               dense_item_grad=w_item, dense_user_grad=w_user)   # sink buffers unused before SCATTER
        ctx.host, ctx.ws = host, ws
        ctx.args = (user, pos, neg32, lqp, lqn, loss_kind, score_kind)
        ctx.save_for_backward(w_item, w_user)
        return loss.clone()

    @staticmethod
    def backward(ctx, g: Tensor):
        w_item, w_user = ctx.saved_tensors
        ws, host = ctx.ws, ctx.host
        user, pos, neg32, lqp, lqn, loss_kind, score_kind = ctx.args
        g = g.reshape(1).to(torch.float32).contiguous()
        mode = host.fused_grad
        common = dict(logq_pos=lqp, logq_neg=lqn, phases=_lib.PHASE_SCATTER, grad_scale_dev=g)
        if mode == "dense":       # reference-compatible: dense [N, d] gradients for dense optimizers
            gi, gu = torch.zeros_like(w_item), torch.zeros_like(w_user)
            fused.pair_step(ws, w_item, w_user, user, pos, neg32, loss_kind, score_kind, dense_item_grad=gi,
                            dense_user_grad=gu, **common)
            return gi, gu, None, None, None, None, None, None, None, None, None
        if mode == "apply":       # nothing is computed here: FusedRowOptimizer.step() runs PHASE_SCATTER with the update
            ws.pending_apply = (w_item, w_user, user, pos, neg32, loss_kind, score_kind, common)   # fused into its epilogue
            return None, None, None, None, None, None, None, None, None, None, None
        d = w_item.shape[1]
        iv = torch.empty(ws.cap_item, d, dtype=torch.float32, device=w_item.device)
        uv = torch.empty(ws.cap_user, d, dtype=torch.float32, device=w_item.device)
        fused.pair_step(ws, w_item, w_user, user, pos, neg32, loss_kind, score_kind, item_vals=iv, user_vals=uv, **common)
        if mode == "rows":        # consumed by a row optimizer straight from the workspace (no host sync)
            ws.row_grads = (ws.item_rows, iv, ws.user_rows, uv, ws.totals)
            return None, None, None, None, None, None, None, None, None, None, None
        tot = ws.totals.tolist()  # sparse COO hand-off needs the row counts on the host
        ri, ru = tot[1], tot[3]
        ii, ui = ws.item_rows[:ri].unsqueeze(0).clone(), ws.user_rows[:ru].unsqueeze(0).clone()
        gi = torch.sparse_coo_tensor(ii, iv[:ri], size=w_item.shape, is_coalesced=True, check_invariants=False)
        gu = torch.sparse_coo_tensor(ui, uv[:ru], size=w_user.shape, is_coalesced=True, check_invariants=False)
        # AccumulateGrad shallow-copies sparse gradients and loses the coalesced flag (rows ARE unique and
        # ascending); _restore_coalesced (a post-accumulate hook) puts it back when .grad is exactly this tensor
        host._fused_sparse_ptrs = {ii.data_ptr(), ui.data_ptr()}
        return gi, gu, None, None, None, None, None, None, None, None, None


class _GraphedStepFn(torch.autograd.Function):
    """``_FusedStepFn`` for ``fused_graph=True`` (two embedding towers, in-kernel UniformSampler draw, fused_grad='rows'):
    forward and backward each replay ONE captured CUDA graph (fused.GraphedPairStep(split=True)), so a training step
    issued through the reference's trainer loop (recommender.py:594-646) costs two graph launches instead of ~10 kernel
    launches.  The gradient rows stay in the workspace for FusedRowOptimizer, exactly as in 'rows' mode."""

    @staticmethod
    def forward(ctx, w_item: Tensor, w_user: Tensor, host, step, user, pos):
        loss = step.forward(user, pos)
        ctx.host, ctx.step = host, step
        return loss.clone()

    @staticmethod
    def backward(ctx, g: Tensor):
        step, ws = ctx.step, ctx.step.ws
        step.backward(g.to(torch.float32))
        ws.row_grads = (ws.item_rows, ws.item_vals, ws.user_rows, ws.user_vals, ws.totals)
        return None, None, None, None, None, None


class _FusedHeadFn(torch.autograd.Function):
    """The fused step for an arbitrary query encoder (SASRec, DSSM ...): the [B, d] encoder output is
    the query 'table' (row 0 = padding, query b at row b + 1), so the same kernels produce the loss,
    the item-row gradients and d loss / d query, which flows back into the encoder through autograd."""

    @staticmethod
    def forward(ctx, w_item: Tensor, query: Tensor, host, ws, pos, neg32, lqp, lqn, loss_kind, score_kind):
        B, d = query.shape
        wq = torch.cat([query.new_zeros(1, d), query.detach().float()], dim=0).contiguous()
        uid = torch.arange(1, B + 1, device=query.device, dtype=torch.int64)
        loss = fused.pair_step(ws, w_item, wq, uid, pos, neg32, loss_kind, score_kind, logq_pos=lqp, logq_neg=lqn,
                               phases=_lib.PHASE_COUNT | _lib.PHASE_SCAN | _lib.PHASE_FWD,
                               dense_item_grad=w_item, dense_user_grad=wq)
        ctx.host, ctx.ws = host, ws
        ctx.args = (wq, uid, pos, neg32, lqp, lqn, loss_kind, score_kind)
        ctx.save_for_backward(w_item)
        return loss.clone()

    @staticmethod
    def backward(ctx, g: Tensor):
        (w_item,) = ctx.saved_tensors
        ws, host = ctx.ws, ctx.host
        wq, uid, pos, neg32, lqp, lqn, loss_kind, score_kind = ctx.args
        B, d = wq.shape[0] - 1, wq.shape[1]
        g = g.reshape(1).to(torch.float32).contiguous()
        common = dict(logq_pos=lqp, logq_neg=lqn, phases=_lib.PHASE_SCATTER, grad_scale_dev=g)
        nones = (None,) * 8
        if host.fused_grad == "apply":
            raise _lib.Rsb200Error("fused_grad='apply' fuses the optimizer into the step of two embedding-table towers; "
                                   "use 'rows' with an arbitrary query encoder")
        if host.fused_grad == "dense":
            gi, gq = torch.zeros_like(w_item), torch.zeros_like(wq)
            fused.pair_step(ws, w_item, wq, uid, pos, neg32, loss_kind, score_kind, dense_item_grad=gi, dense_user_grad=gq, **common)
            return (gi, gq[1:]) + nones
        iv = torch.empty(ws.cap_item, d, dtype=torch.float32, device=w_item.device)
        uv = torch.empty(ws.cap_user, d, dtype=torch.float32, device=w_item.device)
        fused.pair_step(ws, w_item, wq, uid, pos, neg32, loss_kind, score_kind, item_vals=iv, user_vals=uv, **common)
        dquery = uv[:B]                      # every query row 1..B is touched exactly once, rows come out ascending
        if host.fused_grad == "rows":
            ws.row_grads = (ws.item_rows, iv, None, None, ws.totals)
            return (None, dquery) + nones
        ri = int(ws.totals[1].item())
        ii = ws.item_rows[:ri].unsqueeze(0).clone()
        gi = torch.sparse_coo_tensor(ii, iv[:ri], size=w_item.shape, is_coalesced=True, check_invariants=False)
        host._fused_sparse_ptrs = {ii.data_ptr()}
        return (gi, dquery) + nones


_FUSABLE_SAMPLERS = (FusedUniformSampler, FusedPopularSampler, FusedMaskedUniformSampler)
_FUSABLE_METHODS = ("none", "dns", "sir", "toprand", "top&rand", "brute")


def score_ids(kind: int, query: Tensor, w_item: Tensor, ids: Tensor) -> Tensor:
    """score_func(query[b], W[ids[b, j]]) for a [B, n] id matrix without building the [B, n, d] tensor
    (rsb200_score_ids): the pool scoring of the dns / sir sampling methods (baseretriever.py:323-324)."""
    _need_cuda(query, "query")
    q = query.detach().contiguous().float()
    ids = ids.to(q.device, torch.int64).contiguous()
    B, n = ids.shape
    out = torch.empty(B, n, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        check(lib().rsb200_score_ids(kind, ptr(q), ptr(w_item.detach()), w_item.shape[0], w_item.shape[1], ptr(ids), B, n,
                                     ptr(out), stream_ptr()), "score_ids")
    return out
_LOSS_KIND = {FusedBPRLoss: LOSS_BPR, FusedSampledSoftmaxLoss: LOSS_SSM}
_SCORE_KIND = {FusedInnerProductScorer: SCORE_IP, FusedEuclideanScorer: SCORE_EUCLID}


class FusedRetrieverMixin:
    """Mix in FRONT of ``BaseRetriever``: ``training_step`` takes the
    single fused CUDA path when the plugin combination is one the kernels implement
    (Embedding x {Uniform, Popular} x {IP, Euclid} x {BPR, SampledSoftmax},
    ``sampling_method == 'none'``, 1-D ``item_id``, no history exclusion); anything else
    falls through to the reference's own ``training_step`` on top of the standalone plugins.

    ``fused_grad``: 'dense'  -> dense ``weight.grad`` (what the reference's dense Adam expects),
                    'sparse' -> coalesced sparse-COO ``weight.grad`` (sgd / adagrad / sparse_adam),
                    'rows'   -> gradients stay in the workspace for a row optimizer (no host sync),
                    'apply'  -> no gradient rows at all: ``FusedRowOptimizer.step()`` runs the scatter with the
                                optimizer update in its epilogue (RSB200_SINK_APPLY).
    """
    fused_grad = "dense"

    # --- BaseRetriever.sampling (baseretriever.py:248-369) on the CUDA ops -----------------------------------
    def _pool_scores(self, query, pool):
        """scores of a [B, n0] candidate pool: ids -> rows -> score in one kernel when the item tower is a
        plain table and the scorer is IP / Euclid, else through the (standalone) plugins like the reference."""
        sk = _SCORE_KIND.get(type(self.score_func))
        if sk is not None and isinstance(self.item_encoder, torch.nn.Embedding) and query.dim() == 2 and pool.dim() == 2 \
                and len(getattr(self, "item_fields", [self.fiid])) == 1 and self.item_encoder.weight.is_cuda:
            return score_ids(sk, query, self.item_encoder.weight, pool)
        with torch.no_grad():
            return self.score_func(query.detach(), self.item_encoder(self._get_item_feat(pool)))

    def sampling(self, batch, num_neg, method="none", excluding_hist=False, t=1, return_query=False, query=None):
        """BaseRetriever.sampling (baseretriever.py:248-369).  'none', 'toprand' and 'brute' ARE the reference's code
        (``super().sampling``; its ``self.topk`` resolves to the fused full-catalog top-k below).  Only the arms whose
        work the kernels replace are restated, with the reference's semantics and RNG call order:
        'dns' / 'sir' score the pool of num_neg[0] sampled ids with rsb200_score_ids (no [B, n0, d] tensor, :331-355);
        'top&rand' draws its random half with the fused Philox kernel (same ids as the reference's torch.randint, :289-299)."""
        if method not in ("dns", "sir", "top&rand"):
            return super().sampling(batch, num_neg, method, excluding_hist, t, return_query, query)
        pos_items = batch.get(self.fiid, None)
        if pos_items is not None and pos_items.dim() == 1:
            pos_items = pos_items.view(-1, 1)                                     # :255-258
        if isinstance(num_neg, int):
            num_neg = [num_neg, num_neg]
        elif isinstance(num_neg, (list, tuple)):
            assert len(num_neg) == 2, "length of negative_count must be 2 when it's list type for retriever_dns sampler."
            assert num_neg[0] >= num_neg[1], "the first element of negative_count must be larger than the second element."
        else:
            raise TypeError("num_neg only support int and List/Tuple type.")

        if method == "top&rand":
            user_hist = batch.get("user_hist", None)
            if user_hist is None:
                user_hist = batch.get(self.fiid, None)                            # :260-262
            num_neg_0 = num_neg[1] // 2
            _, neg_id, query = self.topk(batch, k=num_neg_0, user_h=user_hist, return_query=True)
            num_queries = int(np.prod(query.shape[:-1]))
            rand_id, _ = sampling.uniform_draw(self.item_vector.size(0) + 1, num_queries, num_neg[1] - num_neg_0, query.device)
            neg_id = torch.cat((neg_id, rand_id), dim=-1)
            log_neg_prob = torch.zeros_like(neg_id)
            log_pos_prob = None if pos_items is None else torch.zeros_like(pos_items)
        else:
            if pos_items is not None:
                log_pos_prob, neg_id_pool, _, query = self._sample(batch, num_neg[0], excluding_hist, True)
            else:
                neg_id_pool, _, query = self._sample(batch, num_neg[0], excluding_hist, True)
            scores_on_pool_items = self._pool_scores(query, neg_id_pool)
            if method == "dns":
                _, topk_id = torch.topk(scores_on_pool_items, num_neg[1])
                neg_id = torch.gather(neg_id_pool, -1, topk_id)
                log_neg_prob = torch.zeros_like(neg_id)
                log_pos_prob = None if pos_items is None else torch.zeros_like(pos_items)
            else:
                if pos_items is not None:
                    with torch.no_grad():
                        log_pos_prob = self.score_func(query.detach(), self.item_encoder(self._get_item_feat(batch)))
                probs_on_pool_items = torch.softmax(scores_on_pool_items + torch.finfo(torch.float32).eps, dim=-1)
                resampled_id = torch.multinomial(probs_on_pool_items, num_neg[1], replacement=True)
                neg_id = torch.gather(neg_id_pool, dim=-1, index=resampled_id)
                log_neg_prob = torch.gather(scores_on_pool_items, dim=-1, index=resampled_id)

        if pos_items is not None:
            log_pos_prob = log_pos_prob.view_as(batch.get(self.fiid))
            result = (log_pos_prob.detach(), neg_id, log_neg_prob.detach())
        else:
            result = (None, neg_id, log_neg_prob.detach())
        return (result, query) if return_query else (result, None)

    def _fused_combo(self, batch):
        """(loss_kind, score_kind, two_tables) when the kernels implement this plugin combination, else None.
        two_tables: both towers are plain embedding tables (BPR-style); otherwise the query encoder is an
        arbitrary module (SASRec ...) and only the head is fused."""
        cfg = self.config["train"] if hasattr(self, "config") else {}
        method, excl = cfg.get("sampling_method", "none"), bool(cfg.get("excluding_hist", False))
        if method not in _FUSABLE_METHODS:
            return None
        if method in ("none", "dns", "sir"):
            if not isinstance(self.sampler, iface.Sampler):
                return None
            if isinstance(self.sampler, FusedMaskedUniformSampler) and (not excl or batch.get("user_hist", None) is None):
                return None              # the masked sampler needs excluding_hist=True and the batch's user_hist
        lk, sk = _LOSS_KIND.get(type(self.loss_fn)), _SCORE_KIND.get(type(self.score_func))
        if lk is None or sk is None:
            return None
        if not isinstance(self.item_encoder, torch.nn.Embedding):
            return None
        if len(getattr(self, "item_fields", [self.fiid])) != 1 or batch[self.fiid].dim() != 1:
            return None
        if not self.item_encoder.weight.is_cuda:
            raise _lib.Rsb200Error("the fused retriever needs its tables on a CUDA device (no CPU fallback)")
        two_tables = isinstance(self.query_encoder, torch.nn.Embedding) and self.fuid in batch
        return lk, sk, two_tables

    def _fused_ws(self, B: int, n: int, num_query_rows: int):
        cache = self.__dict__.setdefault("_fused_ws_cache", {})
        wi = self.item_encoder.weight
        key = (B, n, wi.shape, num_query_rows, wi.device, self.fused_grad)
        if key not in cache:
            cache.clear()       # one live workspace per model: batch shape is stable within an epoch
            cache[key] = fused.PairWorkspace(wi.shape[0], num_query_rows, B, n, wi.shape[1], wi.device,
                                             sink="dense" if self.fused_grad == "dense" else "compact", alloc_vals=False)
        return cache[key]

    def _restore_coalesced(self, param):
        g = param.grad
        if g is not None and g.is_sparse and not g.is_coalesced() and \
                g._indices().data_ptr() in self.__dict__.get("_fused_sparse_ptrs", ()):
            g._coalesced_(True)

    def _fused_full_softmax(self, batch):
        """config-4 path: full-catalog SoftmaxLoss with plain embedding towers and no sampler
        (baseretriever.py:177-186 + loss_func.py:39-42)."""
        if type(self.loss_fn) is not FusedSoftmaxLoss or self.sampler is not None:
            return None
        if type(self.score_func) is not FusedInnerProductScorer or not isinstance(self.item_encoder, torch.nn.Embedding):
            return None
        if batch[self.fiid].dim() != 1 or len(getattr(self, "item_fields", [self.fiid])) != 1:
            return None
        wi = self.item_encoder.weight
        if not wi.is_cuda or wi.shape[1] > 128:
            return None
        # decided from static properties BEFORE the encoder runs: a [B, d] query needs a table / pooled sequence encoder
        # (running the encoder first and then falling back would encode twice and consume its dropout stream twice)
        if not (isinstance(self.query_encoder, torch.nn.Embedding) or getattr(self.query_encoder, "pools_to_2d", False)):
            return None
        query = self.query_encoder(self._get_query_feat(batch))         # any query encoder (keeps its own autograd)
        if query.dim() != 2:
            raise _lib.Rsb200Error("query encoder declared a [B, d] output but returned %s" % (tuple(query.shape),))
        return _FullSoftmaxFn.apply(query, wi, batch[self.fiid])

    def training_step(self, batch):
        full = self._fused_full_softmax(batch)
        if full is not None:
            return full
        combo = self._fused_combo(batch)
        if combo is None:
            return super().training_step(batch)
        loss_kind, score_kind, two_tables = combo
        wi = self.item_encoder.weight
        if self.fused_grad == "sparse" and not self.__dict__.get("_fused_hooks", False):
            wi.register_post_accumulate_grad_hook(self._restore_coalesced)
            if two_tables:
                self.query_encoder.weight.register_post_accumulate_grad_hook(self._restore_coalesced)
            self.__dict__["_fused_hooks"] = True
        pos = batch[self.fiid].to(wi.device, non_blocking=True).contiguous()
        nc = self.neg_count
        B, n = pos.numel(), int(nc[1] if isinstance(nc, (list, tuple)) else nc)
        cfg = self.config["train"] if hasattr(self, "config") else {}
        method, excl = cfg.get("sampling_method", "none"), bool(cfg.get("excluding_hist", False))
        query = None
        in_kernel_draw = method == "none" and type(self.sampler) in _FUSABLE_SAMPLERS
        if not in_kernel_draw:
            # candidate selection (dns / sir / toprand / top&rand / brute) runs on the CUDA ops of sampling(), and any
            # other Sampler (MIDX / Cluster / a reference sampler) draws through its own forward(); the selected ids
            # and their proposal log-probabilities then take the same fused step as given negatives.
            # sampling() also returns the query it encoded (with its autograd graph), as the reference's forward uses it.
            dev_batch = {k: (v.to(wi.device, non_blocking=True) if isinstance(v, Tensor) else v) for k, v in batch.items()}
            (lqp, neg32, lqn), query = self.sampling(dev_batch, nc, method, excl, return_query=True)
            neg32 = neg32.reshape(B, -1).contiguous()
            if neg32.shape[1] != n:
                raise _lib.Rsb200Error("sampling() returned %d negatives per query, expected %d" % (neg32.shape[1], n))
            lqp = lqp.reshape(B).float() if (lqp is not None and lqp.is_floating_point()) else None
            lqn = lqn.reshape(B, n).float().contiguous() if lqn.is_floating_point() else None
        if two_tables:
            wu = self.query_encoder.weight
            user = batch[self.fuid].to(wi.device, non_blocking=True).contiguous()
            ws = self._fused_ws(B, n, wu.shape[0])
        else:
            # the item table may also be read by the query encoder (SASRec shares the module, sasrec.py:107): its
            # encoder-side gradient takes the same form as the head's (dense | sparse rows | rows for the row optimizer)
            if isinstance(self.item_encoder, FusedEmbedding):
                self.item_encoder.grad_mode = {"dense": "dense", "sparse": "sparse", "rows": "rows"}.get(self.fused_grad, "dense")
            if query is None:
                query = self.query_encoder(self._get_query_feat(batch))  # [B, d], keeps its own autograd graph
            if query.dim() != 2 or query.shape[0] != B:
                if not in_kernel_draw:
                    raise _lib.Rsb200Error("fused sampling methods need a [B, d] query encoder output")
                return super().training_step(batch)
            ws = self._fused_ws(B, n, B + 1)
        if two_tables and in_kernel_draw and getattr(self, "fused_graph", False) and self.fused_grad == "rows" \
                and type(self.sampler) is FusedUniformSampler:
            gs = self.__dict__.get("_fused_graphed")
            key = (wi.data_ptr(), wu.data_ptr(), int(loss_kind), int(score_kind))
            if gs is None or gs.ws is not ws or gs.key != key:
                gs = fused.GraphedPairStep(ws, wi.detach(), wu.detach(), loss_kind, score_kind, split=True)
                self.__dict__["_fused_graphed"] = gs
            return _GraphedStepFn.apply(wi, wu, self, gs, user, pos)
        if in_kernel_draw:
            if isinstance(self.sampler, FusedMaskedUniformSampler):
                neg32, lqn = self.sampler.fused_draw(B, n, wi.device, user_hist=batch["user_hist"])
            else:
                neg32, lqn = self.sampler.fused_draw(B, n, wi.device)
            lqp = self.sampler.compute_item_p(None, pos) if (lqn is not None and loss_kind == LOSS_SSM) else None
        if loss_kind != LOSS_SSM:
            lqp = lqn = None
        if two_tables:
            return _FusedStepFn.apply(wi, wu, self, ws, user, pos, neg32, lqp, lqn, loss_kind, score_kind)
        return _FusedHeadFn.apply(wi, query, self, ws, pos, neg32, lqp, lqn, loss_kind, score_kind)

    def training_epoch_end(self, output_list):
        """recommender.py:249-277, preceded by the id check of the epoch's fused steps: ids outside their table are scored
        as the padding row instead of raising like F.embedding -- report them once per epoch (one host sync)."""
        for ws in self.__dict__.get("_fused_ws_cache", {}).values():
            fused.check_ids(ws)
        return super().training_epoch_end(output_list)

    def topk(self, batch, k, user_h=None, return_query=False):
        """BaseRetriever.topk (baseretriever.py:374-397) on rsb200_topk_full when the scorer is
        InnerProduct / Euclid over a plain item table and no ANN index is configured; the
        reference's torch.topk(k + H) -> mask -> topk(k) otherwise."""
        sk = _SCORE_KIND.get(type(self.score_func))
        item_w = getattr(self.item_encoder, "weight", None)
        fusable = (sk is not None and not getattr(self, "use_index", False) and isinstance(self.item_encoder, torch.nn.Embedding)
                   and len(getattr(self, "item_fields", [self.fiid])) == 1 and item_w is not None and item_w.is_cuda)
        more = user_h.size(1) if user_h is not None else 0
        if fusable and (k + more + 8 > 1024 or k > item_w.shape[0] - 1):
            fusable = False                      # candidate set too large for the in-smem final sort: reference path
        if not fusable:
            return super().topk(batch, k, user_h, return_query)
        from . import topk as _topk
        query = self.query_encoder(self._get_query_feat(batch))
        if query.dim() != 2:
            return super().topk(batch, k, user_h, return_query)
        # item_vector is weight[1:] (baseretriever.py:122-123,131-140); scoring the live table is
        # identical after _update_item_vector(), which the trainer calls before every eval epoch
        score, ids = _topk.topk_full(query, item_w, k, user_h, sk)
        if return_query:
            return score, ids, query
        return score, ids

    def fused_last_neg_id(self):
        """int32 [B, n] negatives drawn by the last fused training_step (for inspection / parity tests)."""
        gs = self.__dict__.get("_fused_graphed")
        if gs is not None:
            return gs.neg32
        cache = self.__dict__.get("_fused_ws_cache", {})
        for ws in cache.values():
            return ws._keepalive[2]
        return None
