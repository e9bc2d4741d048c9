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


class FusedRowOptimizer:
    def __init__(self, model, learner: str = "sparse_adam", lr: float = 1e-3, betas=(0.9, 0.999), eps: float = None):
        if learner not in KINDS:
            raise ValueError("learner must be one of %s" % sorted(KINDS))
        self.model, self.kind, self.lr, self.betas = model, KINDS[learner], float(lr), betas
        self.eps = float(eps) if eps is not None else (1e-10 if learner == "adagrad" else 1e-8)
        self.step_count = 0
        self.state = {}
        self.param_groups = [{"lr": self.lr}]          # schedulers / loggers look at this

    def zero_grad(self, set_to_none: bool = True):
        pass                                           # row gradients are overwritten by the next fused step

    def _state_for(self, w: torch.Tensor):
        key = w.data_ptr()
        if key not in self.state:
            s1 = torch.zeros_like(w) if self.kind >= 1 else None
            s2 = torch.zeros_like(w) if self.kind == 2 else None
            self.state[key] = (s1, s2)
        return self.state[key]

    @torch.no_grad()
    def step(self):
        cache = self.model.__dict__.get("_fused_ws_cache", {})
        ws = next(iter(cache.values()), None)
        pending = getattr(ws, "pending_apply", None) if ws is not None else None
        if pending is not None:
            # fused_grad='apply': PHASE_SCATTER with the update in its epilogue -- the gradient rows are never written
            from . import fused
            self.step_count += 1
            w_item, w_user, user, pos, neg32, loss_kind, score_kind, common = pending
            (i1, i2), (u1, u2) = self._state_for(w_item), self._state_for(w_user)
            bc1 = 1.0 - self.betas[0] ** self.step_count
            bc2 = 1.0 - self.betas[1] ** self.step_count
            spec = {"kind": self.kind, "lr": self.lr, "beta1": self.betas[0], "beta2": self.betas[1], "eps": self.eps,
                    "step_size": self.lr * (bc2 ** 0.5) / bc1, "item_state1": i1, "item_state2": i2,
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
        tables = [(self.model.item_encoder.weight, item_rows, item_vals, 1)]
        if user_rows is not None:
            tables.append((self.model.query_encoder.weight, user_rows, user_vals, 3))
        for w, rows, vals, tot_idx in tables:
            s1, s2 = self._state_for(w)
            cnt_ptr = totals.data_ptr() + 4 * tot_idx
            with torch.cuda.device(w.device):
                check(lib().rsb200_rows_update(self.kind, ptr(w), ptr(s1), ptr(s2), w.shape[0], w.shape[1], ptr(rows), ptr(vals),
                                               cnt_ptr, min(rows.numel(), vals.shape[0]), self.step_count, self.lr,
                                               self.betas[0], self.betas[1], self.eps, stream_ptr()), "rows_update")
        ws.row_grads = None
