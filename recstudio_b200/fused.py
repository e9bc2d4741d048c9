"""Host side of the fused retriever training step (rsb200_pair_step).

``PairWorkspace`` owns the device buffers (torch tensors, allocated once per
problem shape); ``pair_step`` fills the C argument block and enqueues the
phases on the current CUDA stream.  Nothing here computes: all arithmetic is in
librsb200.so.

Reference path replaced: BaseRetriever.forward (sampler branch) +
training_step + loss.backward()
(recstudio/model/basemodel/baseretriever.py:142-176,399-404, recommender.py:638).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import (LOSS_BPR, LOSS_SSM, PHASE_ALL, SCORE_EUCLID, SCORE_IP, SINK_COMPACT, SINK_DENSE,  # noqa: F401
                   PairArgs, PairSizes, check, lib, ptr, stream_ptr)


class PairWorkspace:
    """Device buffers of one (num_items, num_users, B, n, d) problem.

    sink = 'compact': gradients come back as (rows[R] int64 ascending, vals[R, d]) per
    table -- a coalesced sparse-COO gradient; capacity is the worst case
    min(touches, rows).  sink = 'dense': the caller passes dense [rows, d] gradient
    buffers to ``pair_step`` (reference-compatible ``weight.grad``).
    """

    def __init__(self, num_items: int, num_users: int, B: int, n: int, d: int, device,
                 sink: str = "compact", want_scores: bool = False, stage_entries: bool = False, cap_item: Optional[int] = None,
                 alloc_vals: bool = True, grouping: Optional[int] = None, bin_shift: Optional[int] = None):
        _lib.require_cuda()
        self.device = torch.device(device)
        self.shape = (num_items, num_users, B, n, d)
        self.sink = sink
        sz = PairSizes()
        check(lib().rsb200_pair_workspace_sizes(num_items, num_users, B, n, d, C.byref(sz)), "pair_workspace_sizes")
        self.sizes = sz
        dev = self.device
        i32, u32, i64, f32 = torch.int32, torch.int32, torch.int64, torch.float32   # torch has no uint32 arithmetic: same bytes

        def buf(nelem, dtype):
            return torch.empty(max(int(nelem), 1), dtype=dtype, device=dev)

        # grouping of the item-side touches: 1 = bins (csrc/bins.cu, the default wherever the library offers a bin size),
        # 0 = N-bucket counting sort (csrc/group.cu); RSB200_GROUPING overrides for A/B runs
        if grouping is None:
            grouping = int(os.environ.get("RSB200_GROUPING", "1"))
        if stage_entries or not sz.bin_shift:
            grouping = 0
        self.grouping = int(grouping)
        self.bin_shift = int(bin_shift if bin_shift is not None else os.environ.get("RSB200_BIN_SHIFT", sz.bin_shift)) if self.grouping else 0
        if self.grouping:
            self.bin_shift_user = int(sz.bin_shift_user)
            self.nbins = -(-num_items // (1 << self.bin_shift))
            self.nbins_user = -(-num_users // (1 << self.bin_shift_user))
            nb = self.nbins + self.nbins_user
            self.bin_cnt, self.bin_off = buf(nb, u32), buf(nb + 2, u32)
            self.bin_cursor, self.bin_status = buf(nb * 8, u32), buf(nb, i64)
            self.bin_ticket, self.bin_heavy = torch.zeros(2, dtype=u32, device=dev), buf(sz.bin_heavy, u32)
            self.off_item = self.slot_neg = self.slot_pos = self.urow_item = None
            self.off_user = self.slot_user = self.urow_user = self.scan_tmp = None
        else:
            self.bin_shift_user = 0
            self.off_item = buf(sz.off_item, u32)
            self.slot_neg = buf(sz.slot_neg, u32)
            self.slot_pos = buf(sz.slot_pos, u32)
            self.off_user = buf(sz.off_user, u32)
            self.slot_user = buf(sz.slot_user, u32)
            self.scan_tmp = buf(sz.scan_tmp, i64)
        self.neg32_buf = buf(sz.neg32_buf, i32)
        self.ent_item = buf(sz.ent_item, i64)
        self.ent_user = buf(sz.ent_user, i64)
        self.cap_item = int(cap_item) if cap_item is not None else int(sz.cap_item)
        self.cap_user = int(sz.cap_user)
        if not self.grouping:
            self.urow_item = buf(self.cap_item, u32)
            self.urow_user = buf(self.cap_user, u32)
        self.q_buf = buf(sz.q_buf, f32)
        self.dq_buf = buf(sz.dq_buf, f32)
        self.loss_part = buf(sz.loss_part, f32)
        self.lse = buf(sz.lse, f32)
        self.err_flag = torch.zeros(1, dtype=i32, device=dev)
        self.totals = torch.zeros(4, dtype=i32, device=dev)
        self.loss = torch.zeros(1, dtype=f32, device=dev)
        self.item_rows = buf(self.cap_item, i64)
        self.user_rows = buf(self.cap_user, i64)
        if sink == "compact" and alloc_vals:
            self.item_vals = torch.empty(self.cap_item, d, dtype=f32, device=dev)
            self.user_vals = torch.empty(self.cap_user, d, dtype=f32, device=dev)
        else:
            self.item_vals = None
            self.user_vals = None
        self.pos_score = torch.empty(max(B, 1), dtype=f32, device=dev) if want_scores else None
        self.neg_score = torch.empty(max(B, 1), max(n, 1), dtype=f32, device=dev) if want_scores else None
        self.cstage = torch.empty(max(B * n, 1), dtype=f32, device=dev) if stage_entries else None   # variant 7 only

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in vars(self).values() if isinstance(t, torch.Tensor))


def pair_step(ws: PairWorkspace, w_item: torch.Tensor, w_user: torch.Tensor, user: torch.Tensor,
              pos: torch.Tensor, neg: torch.Tensor, loss_kind: int, score_kind: int,
              logq_pos: Optional[torch.Tensor] = None, logq_neg: Optional[torch.Tensor] = None,
              phases: int = PHASE_ALL, grad_scale: float = 1.0, accumulate: bool = False,
              dense_item_grad: Optional[torch.Tensor] = None, dense_user_grad: Optional[torch.Tensor] = None,
              variant: int = 0, grad_scale_dev: Optional[torch.Tensor] = None,
              item_vals: Optional[torch.Tensor] = None, user_vals: Optional[torch.Tensor] = None,
              apply: Optional[dict] = None, draw: Optional[dict] = None):
    """Enqueue the selected phases of the fused step on the current stream.

    ``draw`` (binned grouping only): fuse the UniformSampler draw into the step -- ``neg`` is then the int32 [B, n] OUTPUT
    buffer, filled by rsb200_pair_draw_count with the ids ``torch.randint(1, num_items, (B, n), device=cuda)`` would return
    for the generator state ``{'state_dev': int64[2] device tensor}`` (advanced on the device) or ``{'seed': s, 'offset': o}``,
    and PHASE_COUNT (taken inside that kernel) is dropped from ``phases``.  The caller advances the torch generator.

    Returns the 0-dim loss tensor (a view of ``ws.loss``).  Gradients are left in
    ``ws.item_rows/item_vals`` and ``ws.user_rows/user_vals`` (first ``ws.totals[1]`` /
    ``ws.totals[3]`` rows) or in the dense buffers.
    """
    num_items, num_users, B, n, d = ws.shape
    for t, name in ((w_item, "w_item"), (w_user, "w_user"), (user, "user"), (pos, "pos"), (neg, "neg")):
        if not t.is_cuda or not t.is_contiguous():
            raise _lib.Rsb200Error("%s must be a contiguous CUDA tensor (no CPU fallback)" % name)
    if w_item.shape != (num_items, d) or w_user.shape != (num_users, d) or w_item.dtype != torch.float32:
        raise _lib.Rsb200Error("table shape/dtype does not match the workspace")
    if user.shape != (B,) or pos.shape != (B,) or tuple(neg.shape) != (B, n):
        raise _lib.Rsb200Error("batch shape does not match the workspace: user %s pos %s neg %s, expected B=%d n=%d"
                               % (tuple(user.shape), tuple(pos.shape), tuple(neg.shape), B, n))
    if user.dtype != torch.int64 or pos.dtype != torch.int64:
        raise _lib.Rsb200Error("user / pos ids must be int64 (what the reference's batches hold)")
    a = PairArgs()
    a.w_item, a.w_user, a.user, a.pos = ptr(w_item), ptr(w_user), ptr(user), ptr(pos)
    if neg.dtype == torch.int64:
        a.neg_i64, a.neg_i32 = ptr(neg), 0
    elif neg.dtype == torch.int32:
        a.neg_i64, a.neg_i32 = 0, ptr(neg)
    else:
        raise _lib.Rsb200Error("neg ids must be int64 or int32")
    if logq_pos is not None:
        logq_pos = logq_pos.to(torch.float32).contiguous()     # UniformSampler hands out int64 zeros
    if logq_neg is not None:
        logq_neg = logq_neg.to(torch.float32).contiguous()
    a.logq_pos, a.logq_neg = ptr(logq_pos), ptr(logq_neg)
    a.loss, a.pos_score, a.neg_score = ptr(ws.loss), ptr(ws.pos_score), ptr(ws.neg_score)
    a.cstage = ptr(ws.cstage) if variant == 7 else 0          # variant 7 (A/B): staged entries + permute pass, see rsb200.h
    dense = ws.sink == "dense" and apply is None
    keep_states = ()
    if apply is not None:
        # RSB200_SINK_APPLY: PHASE_SCATTER applies the optimizer update in its epilogue (no gradient rows)
        keep_states = tuple(apply.get(k) for k in ("item_state1", "item_state2", "user_state1", "user_state2"))
        a.w_item_rw, a.w_user_rw = ptr(w_item), ptr(w_user)
        a.item_state1, a.item_state2, a.user_state1, a.user_state2 = (ptr(t) for t in keep_states)
        a.opt_kind = int(apply["kind"])
        a.opt_lr, a.opt_beta1, a.opt_beta2 = float(apply["lr"]), float(apply["beta1"]), float(apply["beta2"])
        a.opt_eps, a.opt_step_size = float(apply["eps"]), float(apply["step_size"])
        a.item_vals, a.user_vals = 0, 0
    elif dense:
        if dense_item_grad is None or dense_user_grad is None:
            raise _lib.Rsb200Error("dense sink needs dense_item_grad / dense_user_grad")
        if tuple(dense_item_grad.shape) != (num_items, d) or tuple(dense_user_grad.shape) != (num_users, d):
            raise _lib.Rsb200Error("dense gradient buffers have the wrong shape")
        a.item_vals, a.user_vals = ptr(dense_item_grad), ptr(dense_user_grad)
    else:
        # explicit per-step value buffers (autograd hand-off) or the workspace's persistent ones
        iv = item_vals if item_vals is not None else ws.item_vals
        uv = user_vals if user_vals is not None else ws.user_vals
        if (phases & _lib.PHASE_SCATTER) and (iv is None or uv is None or iv.shape[0] < ws.cap_item
                                              or uv.shape[0] < ws.cap_user):
            raise _lib.Rsb200Error("compact sink needs item_vals [>=cap_item, d] / user_vals [>=cap_user, d]")
        a.item_vals, a.user_vals = ptr(iv), ptr(uv)
    if grad_scale_dev is not None:
        if not grad_scale_dev.is_cuda or grad_scale_dev.dtype != torch.float32 or grad_scale_dev.numel() != 1:
            raise _lib.Rsb200Error("grad_scale_dev must be a 1-element float32 CUDA tensor")
    a.grad_scale_dev = ptr(grad_scale_dev)
    a.item_rows, a.user_rows, a.totals = ptr(ws.item_rows), ptr(ws.user_rows), ptr(ws.totals)
    a.off_item, a.off_user, a.neg32_buf = ptr(ws.off_item), ptr(ws.off_user), ptr(ws.neg32_buf)
    a.slot_neg, a.slot_pos, a.slot_user = ptr(ws.slot_neg), ptr(ws.slot_pos), ptr(ws.slot_user)
    a.ent_item, a.ent_user = ptr(ws.ent_item), ptr(ws.ent_user)
    a.urow_item, a.urow_user = ptr(ws.urow_item), ptr(ws.urow_user)
    a.q_buf, a.dq_buf, a.loss_part, a.lse = ptr(ws.q_buf), ptr(ws.dq_buf), ptr(ws.loss_part), ptr(ws.lse)
    a.scan_tmp, a.err_flag = ptr(ws.scan_tmp), ptr(ws.err_flag)
    a.grouping, a.bin_shift, a.bin_shift_user = ws.grouping, ws.bin_shift, ws.bin_shift_user
    if ws.grouping:
        a.bin_cnt, a.bin_off, a.bin_cursor = ptr(ws.bin_cnt), ptr(ws.bin_off), ptr(ws.bin_cursor)
        a.bin_status, a.bin_ticket, a.bin_heavy = ptr(ws.bin_status), ptr(ws.bin_ticket), ptr(ws.bin_heavy)
    a.num_items, a.num_users, a.B, a.n, a.d = num_items, num_users, B, n, d
    a.cap_item, a.cap_user, a.scan_tmp_elems = ws.cap_item, ws.cap_user, (ws.scan_tmp.numel() if ws.scan_tmp is not None else 0)
    a.grad_scale = float(grad_scale)
    a.loss_kind, a.score_kind = int(loss_kind), int(score_kind)
    a.sink = _lib.SINK_APPLY if apply is not None else (SINK_DENSE if dense else SINK_COMPACT)
    a.accumulate, a.variant = int(bool(accumulate)), int(variant)
    with torch.cuda.device(w_item.device):
        if draw is not None:
            if not ws.grouping or neg.dtype != torch.int32:
                raise _lib.Rsb200Error("draw= needs the binned grouping and an int32 [B, n] output buffer for the ids")
            from . import sampling
            sm, mt = sampling._policy(w_item.device)
            st = draw.get("state_dev")
            check(lib().rsb200_pair_draw_count(C.byref(a), ptr(st), int(draw.get("seed", 0)) & ((1 << 64) - 1), int(draw.get("offset", 0)),
                                               sm, mt, ptr(neg), stream_ptr()), "pair_draw_count")
            phases = int(phases) & ~_lib.PHASE_COUNT
        check(lib().rsb200_pair_step(C.byref(a), int(phases), stream_ptr()), "pair_step")
    # keep temporaries referenced by the async launch alive until the stream catches up
    ws._keepalive = (logq_pos, logq_neg, neg, user, pos, grad_scale_dev) + keep_states
    return ws.loss[0]


def check_ids(ws: PairWorkspace):
    """Raise if any step on this workspace saw a user / item / negative id outside its table (host-synchronising).
    Such ids are never dereferenced: they are scored as the padding row 0 and receive no gradient -- where the reference's
    F.embedding raises an index error / device assert -- so a corrupted batch or an off-by-one ``num_items`` shows up here."""
    if int(ws.err_flag.item()):
        ws.err_flag.zero_()
        raise _lib.Rsb200Error("fused step: an id of the batch was outside [0, rows) of its table (it was scored as the padding row)")


def sparse_grads(ws: PairWorkspace):
    """Host-synchronising read-out of the compact gradients:
    ((item_rows[R], item_vals[R,d]), (user_rows[Ru], user_vals[Ru,d])).  Also runs ``check_ids``."""
    check_ids(ws)
    tot = ws.totals.tolist()
    ri, ru = tot[1], tot[3]
    return (ws.item_rows[:ri], ws.item_vals[:ri]), (ws.user_rows[:ru], ws.user_vals[:ru])


class GraphedPairStep:
    """The whole fused step -- UniformSampler draw + bin histogram, SCAN, FWD, SCATTER -- captured once in a CUDA graph
    and replayed: the kernels between the two big streams are a few microseconds each, so launch gaps are ~3 % of the
    step when they are issued one by one.  The draw reads the generator state from device memory, so every replay draws
    exactly what ``torch.randint(1, N, (B, n), device=cuda)`` would return for the torch generator's current state, and
    the generator is advanced accordingly; if anything else consumed the generator in between, the state is re-uploaded
    before the replay.  Gradients are left in ``ws`` as for ``pair_step`` (compact sink).

    ``split=True`` captures TWO graphs, for an autograd hand-off: ``forward(user, pos)`` replays draw | COUNT | SCAN | FWD and
    returns the loss, ``backward(grad_output)`` replays SCATTER scaled by the upstream gradient (read on the device).
    ``FusedRetrieverMixin.training_step`` uses it when the retriever is built with ``fused_graph=True``."""

    def __init__(self, ws: PairWorkspace, w_item: torch.Tensor, w_user: torch.Tensor, loss_kind: int, score_kind: int,
                 generator: Optional[torch.Generator] = None, split: bool = False):
        from . import sampling
        num_items, num_users, B, n, d = ws.shape
        dev = ws.device
        self.ws, self.dev, self.shape, self.split = ws, dev, (num_items, B, n), split
        self.key = (w_item.data_ptr(), w_user.data_ptr(), int(loss_kind), int(score_kind))
        self.gen = sampling._generator(dev, generator)
        self.inc = sampling.counter_offset(B * n, dev)
        self.user = torch.zeros(B, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(B, dtype=torch.int64, device=dev)
        self.neg32 = torch.empty(B, n, dtype=torch.int32, device=dev)
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)
        self.gscale = torch.ones(1, dtype=torch.float32, device=dev)
        self._tracked = None
        sm, mt = sampling._policy(dev)
        if split and (ws.item_vals is None or ws.user_vals is None):
            ws.item_vals = torch.empty(ws.cap_item, d, dtype=torch.float32, device=dev)
            ws.user_vals = torch.empty(ws.cap_user, d, dtype=torch.float32, device=dev)
        fwd_phases = _lib.PHASE_COUNT | _lib.PHASE_SCAN | _lib.PHASE_FWD

        def body(phases):
            if phases & _lib.PHASE_COUNT:
                if ws.grouping:    # draw + bin histogram in one kernel (rsb200_pair_draw_count)
                    pair_step(ws, w_item, w_user, self.user, self.pos, self.neg32, loss_kind, score_kind, phases=phases,
                              draw={"state_dev": self.state}, grad_scale_dev=self.gscale if split else None)
                    return
                with torch.cuda.device(dev):
                    check(lib().rsb200_sample_uniform_dev(ptr(self.state), num_items, B, n, sm, mt, 0, ptr(self.neg32), stream_ptr()),
                          "sample_uniform_dev")
            pair_step(ws, w_item, w_user, self.user, self.pos, self.neg32, loss_kind, score_kind, phases=phases,
                      grad_scale_dev=self.gscale if split else None)

        parts = [fwd_phases, _lib.PHASE_SCATTER] if split else [PHASE_ALL]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                 # warm-up outside the capture (lazy module loading, attributes)
            for ph in parts:
                body(ph)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        l0 = lib().rsb200_launch_count()
        self.graphs = []
        for ph in parts:
            g = torch.cuda.CUDAGraph()
            # thread_local: other threads (e.g. NCCL's watchdog polling events) must not invalidate the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                body(ph)
            self.graphs.append(g)
        self.graph = self.graphs[0]
        self.launches_per_step = int(lib().rsb200_launch_count() - l0)

    def _sync_state(self):
        seed, off = self.gen.initial_seed(), self.gen.get_offset()
        if self._tracked != (seed, off):              # first call, or someone else used the generator: upload its state
            self.state.copy_(torch.tensor([seed - (1 << 64) if seed >= (1 << 63) else seed, off], dtype=torch.int64))
        self.gen.set_offset(off + self.inc)
        self._tracked = (seed, off + self.inc)

    def __call__(self, user: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        self.user.copy_(user, non_blocking=True)
        self.pos.copy_(pos, non_blocking=True)
        self._sync_state()
        for g in self.graphs:
            g.replay()
        return self.ws.loss[0]

    def forward(self, user: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        """split mode: draw | COUNT | SCAN | FWD; the loss is ``ws.loss[0]``"""
        self.user.copy_(user, non_blocking=True)
        self.pos.copy_(pos, non_blocking=True)
        self._sync_state()
        self.graphs[0].replay()
        return self.ws.loss[0]

    def backward(self, grad_output: torch.Tensor):
        """split mode: SCATTER scaled by autograd's upstream gradient; rows in ``ws.item_rows / item_vals / user_rows / user_vals``"""
        self.gscale.copy_(grad_output.reshape(1), non_blocking=True)
        self.graphs[1].replay()
