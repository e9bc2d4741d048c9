"""Row optimizer for ``fused_grad='rows'`` / ``'apply'``.

The reference steps a dense optimizer over every table row each iteration
(``optim.Adam(self.parameters())``, recommender.py:427-428,462-463,646).  Here the gradient rows
stay in the fused step's workspace and ``FusedRowOptimizer.step()`` updates only those rows with
rsb200_rows_update; the row count is read on the device, so a training iteration issues no host
synchronisation at all.  With ``fused_grad='apply'`` the update is fused into the scatter epilogue
(RSB200_SINK_APPLY): the gradient rows (2.9 GB at config 2) are neither written nor re-read.  The trainer only duck-types ``zero_grad()`` / ``step()``
(recommender.py:598-600,645-646), and the documented hook for this is ``_get_optimizers``
("If you want to use multi learner, please override `_get_optimizers`", recommender.py:403-406).

Semantics (stated, not hidden): 'sgd' and 'adagrad' equal torch.optim.SGD / Adagrad fed with the same
sparse gradient; 'sparse_adam' equals torch.optim.SparseAdam (moments of untouched rows do not decay),
which is NOT the reference's dense Adam -- parity with the reference is therefore defined on loss and
gradients, not on post-step weights.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

KINDS = {"sgd": 0, "adagrad": 1, "sparse_adam": 2}


class FusedRowOptimizer(torch.optim.Optimizer):
    """Touched-row SGD / Adagrad / SparseAdam over the embedding tables of a fused retriever.  A ``torch.optim.Optimizer``
    (so torch's lr schedulers accept it: ``param_groups[0]['lr']`` is read at every step) whose ``step()`` consumes the
    gradient rows left by the fused training step instead of ``.grad``; ``state_dict()`` / ``load_state_dict()`` carry the
    moments and the step count through a checkpoint."""

    def __init__(self, model, learner: str = "sparse_adam", lr: float = 1e-3, betas=(0.9, 0.999), eps: float = None):
        if learner not in KINDS:
            raise ValueError("learner must be one of %s" % sorted(KINDS))
        params = [model.item_encoder.weight]
        if isinstance(model.query_encoder, torch.nn.Embedding):
            params.append(model.query_encoder.weight)
        super().__init__(params, {"lr": float(lr)})
        self.model, self.learner, self.kind, self.betas = model, learner, KINDS[learner], betas
        self.eps = float(eps) if eps is not None else (1e-10 if learner == "adagrad" else 1e-8)
        self.step_count = 0
        self.row_state = {}

    @property
    def lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    def zero_grad(self, set_to_none: bool = True):
        # row gradients are overwritten by the next fused step; encoder-side row gradients of a shared table
        # (FusedEmbedding.grad_mode == 'rows') are consumed by step()
        super().zero_grad(set_to_none=True)
        for emb in self._tables():
            if getattr(emb, "_extra_row_grads", None):
                emb._extra_row_grads.clear()

    def _tables(self):
        out = [self.model.item_encoder]
        if isinstance(self.model.query_encoder, torch.nn.Embedding):
            out.append(self.model.query_encoder)
        return out

    def _names(self):
        return {"item_encoder.weight": self.model.item_encoder.weight,
                **({"query_encoder.weight": self.model.query_encoder.weight}
                   if isinstance(self.model.query_encoder, torch.nn.Embedding) else {})}

    def _state_for(self, w: torch.Tensor):
        key = w.data_ptr()
        if key not in self.row_state:
            s1 = torch.zeros_like(w) if self.kind >= 1 else None
            s2 = torch.zeros_like(w) if self.kind == 2 else None
            self.row_state[key] = (s1, s2)
        return self.row_state[key]

    def state_dict(self):
        st = {}
        for name, w in self._names().items():
            s1, s2 = self.row_state.get(w.data_ptr(), (None, None))
            st[name] = {"state1": s1, "state2": s2}
        return {"learner": self.learner, "step_count": self.step_count, "lr": self.lr, "betas": tuple(self.betas), "eps": self.eps,
                "state": st}

    def load_state_dict(self, sd):
        if sd.get("learner") != self.learner:
            raise ValueError("optimizer state of learner %r cannot be loaded into %r" % (sd.get("learner"), self.learner))
        self.step_count = int(sd["step_count"])
        self.param_groups[0]["lr"] = float(sd["lr"])
        self.betas, self.eps = tuple(sd["betas"]), float(sd["eps"])
        for name, w in self._names().items():
            ent = sd["state"].get(name)
            if ent is None:
                continue
            s1, s2 = self._state_for(w)
            if s1 is not None and ent["state1"] is not None:
                s1.copy_(ent["state1"])
            if s2 is not None and ent["state2"] is not None:
                s2.copy_(ent["state2"])

    def _merge_extra(self, emb, rows, vals, count_dev):
        """head rows (+ count on the device) and the encoder-side row gradients of the same table -> one coalesced
        (rows, vals, count tensor): a row touched through both paths must see ONE optimizer update with the summed gradient"""
        from .sharded import CudaOps
        extra = emb._extra_row_grads
        r = int(count_dev.item())                         # the merge needs the head's row count on the host
        ids = torch.cat([rows[:r]] + [e[0] for e in extra])
        vv = torch.cat([vals[:r]] + [e[1] for e in extra])
        extra.clear()
        mr, mv = CudaOps.coalesce_rows(ids, vv, emb.weight.shape[0], skip_row0=True)
        cnt = torch.tensor([mr.numel()], dtype=torch.int32, device=mr.device)
        return mr.contiguous(), mv.contiguous(), cnt

    @torch.no_grad()
    def step(self, closure=None):
        cache = self.model.__dict__.get("_fused_ws_cache", {})
        ws = next(iter(cache.values()), None)
        pending = getattr(ws, "pending_apply", None) if ws is not None else None
        lr = self.lr
        if pending is not None:
            # fused_grad='apply': PHASE_SCATTER with the update in its epilogue -- the gradient rows are never written
            from . import fused
            self.step_count += 1
            w_item, w_user, user, pos, neg32, loss_kind, score_kind, common = pending
            (i1, i2), (u1, u2) = self._state_for(w_item), self._state_for(w_user)
            bc1 = 1.0 - self.betas[0] ** self.step_count
            bc2 = 1.0 - self.betas[1] ** self.step_count
            spec = {"kind": self.kind, "lr": lr, "beta1": self.betas[0], "beta2": self.betas[1], "eps": self.eps,
                    "step_size": lr * (bc2 ** 0.5) / bc1, "item_state1": i1, "item_state2": i2,
                    "user_state1": u1, "user_state2": u2}
            fused.pair_step(ws, w_item, w_user, user, pos, neg32, loss_kind, score_kind, apply=spec, **common)
            ws.pending_apply = None
            return
        grads = getattr(ws, "row_grads", None) if ws is not None else None
        if grads is None:
            raise _lib.Rsb200Error("FusedRowOptimizer.step(): no row gradients; run training_step(...).backward() "
                                   "with fused_grad='rows' first")
        self.step_count += 1
        item_rows, item_vals, user_rows, user_vals, totals = grads
        tables = [(self.model.item_encoder, item_rows, item_vals, totals[1:2])]
        if user_rows is not None:
            tables.append((self.model.query_encoder, user_rows, user_vals, totals[3:4]))
        for emb, rows, vals, cnt in tables:
            w = emb.weight
            if getattr(emb, "_extra_row_grads", None):
                rows, vals, cnt = self._merge_extra(emb, rows, vals, cnt)
            s1, s2 = self._state_for(w)
            with torch.cuda.device(w.device):
                check(lib().rsb200_rows_update(self.kind, ptr(w), ptr(s1), ptr(s2), w.shape[0], w.shape[1], ptr(rows), ptr(vals),
                                               cnt.data_ptr(), min(rows.numel(), vals.shape[0]), self.step_count, lr,
                                               self.betas[0], self.betas[1], self.eps, stream_ptr()), "rows_update")
            self._keep = (rows, vals, cnt)
        ws.row_grads = None
