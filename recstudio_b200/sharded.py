"""Row-sharded item table across the GPUs of one box (BASELINE config 5).

The reference has no table sharding at all (its only multi-GPU mode replicates every
parameter each step, recstudio/utils/data_parallel.py:151).  Here each rank owns a contiguous
block of item rows; one training step of a rank is:

  1. draw its own negatives (global ids) and collect the UNIQUE item rows its batch touches;
  2. remote-row lookup: all-to-all the ids to their owners, owners gather the rows
     (rsb200_gather_rows), all-to-all the rows back                     [the exchange step]
  3. run the unchanged fused step (rsb200_pair_step) on the compact local table of fetched rows;
  4. all-to-all the gradient rows back to their owners, which sum the contributions of all
     ranks per owned row (rsb200_rows_coalesce).
  The (small) user table is replicated; its gradient rows are all-gathered so every replica
  applies the same update.

Because unique ids come out sorted and ownership is contiguous, ids are already grouped by
owner: no permutation is needed on either side.  The exchange plan (who sends how many rows to
whom) is plain torch tensor code with no device assumptions, so the same functions run under
``gloo`` on CPU tensors in the tests; the arithmetic (gather / fused step / coalesce) is CUDA only
and is injected through ``ops``.

This is the literal design of the north star (row lookup over NVLink): 512 B per touched row
each way.  ``owner_compute_step`` below is the cheaper formulation of SURVEY.md 8(e): ship the
QUERIES (and 4-byte ids), let every owner score / accumulate the negatives it holds in place
(rsb200_shard_step, csrc/shard.cu) and exchange only per-query scalars and ``d loss / d query``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


# --------------------------------------------------------------------------------------- plan
def rows_per_rank(num_rows: int, world: int) -> int:
    return (num_rows + world - 1) // world


def owner_counts(sorted_ids: torch.Tensor, per_rank: int, world: int) -> torch.Tensor:
    """How many of the ascending ``sorted_ids`` each rank owns (contiguous ownership)."""
    bounds = torch.arange(1, world + 1, device=sorted_ids.device, dtype=sorted_ids.dtype) * per_rank
    ends = torch.searchsorted(sorted_ids, bounds, right=False)
    return torch.diff(ends, prepend=ends.new_zeros(1))


def exchange_counts(send_counts: torch.Tensor, group=None) -> torch.Tensor:
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts, group=group)
    return recv


def all_to_all_var(x: torch.Tensor, send_counts, recv_counts, group=None) -> torch.Tensor:
    """all_to_all_single with per-peer row counts (host lists)."""
    out = x.new_empty((int(sum(recv_counts)),) + tuple(x.shape[1:]))
    dist.all_to_all_single(out, x.contiguous(), output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                           group=group)
    return out


# --------------------------------------------------------------------------------------- CUDA ops
class CudaOps:
    """The arithmetic of the sharded step: CUDA kernels only."""

    @staticmethod
    def gather_rows(weight: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda()
        out = torch.empty(ids.numel(), weight.shape[1], dtype=torch.float32, device=weight.device)
        with torch.cuda.device(weight.device):
            _lib.check(_lib.lib().rsb200_gather_rows(_lib.ptr(weight), weight.shape[0], weight.shape[1], _lib.ptr(ids),
                                                     ids.numel(), _lib.ptr(out), _lib.stream_ptr()), "gather_rows")
        return out

    @staticmethod
    def coalesce_rows(ids: torch.Tensor, vals: torch.Tensor, num_rows: int, skip_row0: bool = False):
        """(ids[M], vals[M,d]) -> (unique ascending rows[R], summed vals[R,d])"""
        _lib.require_cuda()
        dev, M, d = vals.device, ids.numel(), vals.shape[1]
        cap = max(1, min(M, num_rows))
        i32 = torch.int32
        off = torch.empty(num_rows + 1, dtype=i32, device=dev)
        slot = torch.empty(max(M, 1), dtype=i32, device=dev)
        ent = torch.empty(max(M, 1), dtype=torch.int64, device=dev)
        urow = torch.empty(cap, dtype=i32, device=dev)
        tmp = torch.empty(num_rows // 4096 + 3, dtype=torch.int64, device=dev)
        totals = torch.zeros(2, dtype=i32, device=dev)
        err = torch.zeros(1, dtype=i32, device=dev)
        rows = torch.empty(cap, dtype=torch.int64, device=dev)
        out = torch.empty(cap, d, dtype=torch.float32, device=dev)
        ids = ids.contiguous(); vals = vals.contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rsb200_rows_coalesce(
                _lib.ptr(ids), _lib.ptr(vals), M, num_rows, d, int(skip_row0), _lib.ptr(rows), _lib.ptr(out),
                _lib.SINK_COMPACT, 0, _lib.ptr(totals), _lib.ptr(off), _lib.ptr(slot), _lib.ptr(ent), _lib.ptr(urow), cap,
                _lib.ptr(tmp), tmp.numel(), _lib.ptr(err), _lib.stream_ptr()), "rows_coalesce")
        r = int(totals[1].item())
        if int(err.item()):
            raise _lib.Rsb200Error("rows_coalesce: row id out of range")
        return rows[:r], out[:r]


# --------------------------------------------------------------------------------------- table
class ShardedRows:
    """One rank's block of a row-sharded fp32 table and the two exchanges around it."""

    def __init__(self, num_rows: int, dim: int, device, group=None, ops=CudaOps, init_std: float = 0.0, seed: int = 0):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.num_rows, self.dim = num_rows, dim
        self.per_rank = rows_per_rank(num_rows, self.world)
        self.row0 = self.rank * self.per_rank
        self.local_rows = max(0, min(self.per_rank, num_rows - self.row0))
        self.ops = ops
        self.weight = torch.zeros(max(self.local_rows, 1), dim, dtype=torch.float32, device=device)
        if init_std > 0:
            g = torch.Generator(device=device).manual_seed(seed + self.rank)
            self.weight.normal_(0, init_std, generator=g)
        if self.rank == 0:
            self.weight[0] = 0            # global row 0 is the padding row

    # unique ascending global ids -> their rows, fetched from the owners
    def lookup(self, uniq_ids: torch.Tensor):
        send = owner_counts(uniq_ids, self.per_rank, self.world)
        recv = exchange_counts(send, self.group)
        send_l, recv_l = send.tolist(), recv.tolist()
        want = all_to_all_var(uniq_ids, send_l, recv_l, self.group)              # ids other ranks ask me for
        rows = self.ops.gather_rows(self.weight, want - self.row0)
        got = all_to_all_var(rows, recv_l, send_l, self.group)                   # rows in uniq_ids order
        return got, (send_l, recv_l)

    # gradient rows for ascending global ids -> summed per owned row on the owner
    def push(self, ids: torch.Tensor, vals: torch.Tensor):
        send = owner_counts(ids, self.per_rank, self.world)
        recv = exchange_counts(send, self.group)
        send_l, recv_l = send.tolist(), recv.tolist()
        r_ids = all_to_all_var(ids, send_l, recv_l, self.group)
        r_vals = all_to_all_var(vals, send_l, recv_l, self.group)
        return self.ops.coalesce_rows(r_ids - self.row0, r_vals, max(self.local_rows, 1), skip_row0=False)


def sharded_training_step(items: ShardedRows, w_user: torch.Tensor, user: torch.Tensor, pos: torch.Tensor,
                          neg: torch.Tensor, loss_kind: int, score_kind: int, logq_pos: Optional[torch.Tensor] = None,
                          logq_neg: Optional[torch.Tensor] = None, fused_step=None):
    """One data-parallel step of a rank over the row-sharded item table.

    ``neg`` holds GLOBAL item ids [B, n].  Returns (global mean loss, (owned item rows, summed grads) with
    LOCAL row ids, (user rows, grads) identical on every rank).  ``fused_step(w_item, w_user, user, pos, neg32,
    lqp, lqn, grad_scale) -> (loss, item_rows, item_vals, user_rows, user_vals)`` is the single-GPU fused step;
    gradients are scaled by 1/world so that the sum over ranks is the gradient of the global mean loss.
    """
    world = items.world
    touched = torch.cat([neg.reshape(-1).to(torch.int64), pos.to(torch.int64)])
    uniq, inverse = torch.unique(touched, sorted=True, return_inverse=True)
    rows, _ = items.lookup(uniq)
    # compact local table: position 0 is a dummy padding row, fetched row i sits at position i + 1
    w_local = torch.cat([rows.new_zeros(1, items.dim), rows], dim=0)
    neg_c = (inverse[:neg.numel()].reshape(neg.shape) + 1)
    pos_c = (inverse[neg.numel():] + 1)
    # global id 0 (padding) must keep the padding semantics: map it to compact position 0
    if uniq.numel() > 0 and int(uniq[0].item()) == 0:
        neg_c = torch.where(neg == 0, torch.zeros_like(neg_c), neg_c)
        pos_c = torch.where(pos == 0, torch.zeros_like(pos_c), pos_c)
    loss, i_rows, i_vals, u_rows, u_vals = fused_step(w_local, w_user, user, pos_c.contiguous(), neg_c.to(torch.int32).contiguous(),
                                                      logq_pos, logq_neg, 1.0 / world)
    g_ids = uniq[i_rows - 1]                                  # ascending compact positions -> ascending global ids
    own_rows, own_vals = items.push(g_ids, i_vals)
    # replicated user table: every rank needs every rank's user-gradient rows
    # (padded to the batch size so that the all_gather is fixed-size; padding rows carry id 0 = dropped)
    B = user.numel()
    pr = u_rows.new_zeros(B); pr[:u_rows.numel()] = u_rows
    pv = u_vals.new_zeros(B, u_vals.shape[1]); pv[:u_rows.numel()] = u_vals
    all_rows = [torch.empty_like(pr) for _ in range(world)]
    all_vals = [torch.empty_like(pv) for _ in range(world)]
    dist.all_gather(all_rows, pr, group=items.group)
    dist.all_gather(all_vals, pv, group=items.group)
    ur, uv = items.ops.coalesce_rows(torch.cat(all_rows), torch.cat(all_vals), w_user.shape[0], skip_row0=True)
    loss = loss.detach().clone() / world
    dist.all_reduce(loss, group=items.group)
    return loss, (own_rows, own_vals), (ur, uv)


def make_fused_step(B: int, n: int, d: int, num_users: int, device):
    """The single-GPU fused step bound to a workspace sized for the worst-case compact table."""
    from . import fused

    def step(w_local, w_user, user, pos_c, neg32, lqp, lqn, grad_scale):
        ws = fused.PairWorkspace(w_local.shape[0], num_users, B, n, d, device)
        loss = fused.pair_step(ws, w_local.contiguous(), w_user, user, pos_c, neg32, step.loss_kind, step.score_kind,
                               logq_pos=lqp, logq_neg=lqn, grad_scale=grad_scale)
        (ri, vi), (ru, vu) = fused.sparse_grads(ws)
        return loss, ri, vi, ru, vu
    step.loss_kind, step.score_kind = _lib.LOSS_BPR, _lib.SCORE_IP
    return step


# ------------------------------------------------------------------- owner-compute ("ship queries")
def _all_gather_cat(x: torch.Tensor, group=None) -> torch.Tensor:
    """[m, ...] per rank -> [world * m, ...] in rank order (works under nccl and gloo)."""
    world = dist.get_world_size(group)
    out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
    if x.is_cuda:
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)        # one NCCL all-gather, no staging copies
    else:
        dist.all_gather(list(out.chunk(world, dim=0)), x.contiguous(), group=group)
    return out


class OwnerComputeCuda:
    """One owner's workspace and the four kernel phases of ``rsb200_shard_step`` (csrc/shard.cu).

    Nothing here talks to other ranks: ``owner_compute_step`` (or a test that loops over the owners
    on one device) performs the three exchanges between the phases.

    ``grouping``: 1 = binned grouping of the owned touches (csrc/bins.cu, default), 0 = counting sort over the
    owner's rows (round 1).  ``expected_touches``: how many gradient touches this owner receives per step
    (sizes the bins; default G * (n + 1) / world, the uniform expectation)."""

    def __init__(self, num_items: int, row0: int, local_rows: int, weight: torch.Tensor, world: int, rank: int,
                 G: int, n: int, with_logq: bool = False, grouping: Optional[int] = None, bin_shift: Optional[int] = None,
                 expected_touches: Optional[int] = None, slots: int = 1):
        import os
        _lib.require_cuda()
        dev, d = weight.device, weight.shape[1]
        self.weight, self.world, self.rank, self.G, self.n, self.d = weight, world, rank, G, n, d
        self.num_items, self.row0, self.local_rows = num_items, row0, local_rows
        i32, f32, i64 = torch.int32, torch.float32, torch.int64
        E = lambda *shape, dtype=f32: torch.empty(*shape, dtype=dtype, device=dev)
        L = _lib.lib()
        if grouping is None:
            grouping = int(os.environ.get("RSB200_GROUPING", "1"))
        if bin_shift is None:
            touches = expected_touches if expected_touches is not None else max(1, G * (n + 1) // max(world, 1))
            bin_shift = int(L.rsb200_bin_shift(max(local_rows, 1), int(touches), max(G, 1)))
        if not bin_shift or G * d >= (1 << 30):
            grouping = 0
        self.grouping, self.bin_shift = int(grouping), (int(bin_shift) if grouping else 0)
        self.cap = max(1, min(G * (n + 1), local_rows))
        # ``slots`` > 1: ping-pong copies of what PREP_NEG produces (compacted negatives, their log Q, per-query counts, the
        # bin histogram), so that the negatives of step k + 1 can be prepared on a side stream while step k is still running
        self.slots = int(slots)
        self._neg_c = [E(max(G * n, 1), dtype=i32) for _ in range(self.slots)]
        self._lq_c = [E(max(G * n, 1)) if with_logq else None for _ in range(self.slots)]
        self._ncount = [E(max(G, 1), dtype=i32) for _ in range(self.slots)]
        self.neg_c, self.lq_c, self.ncount = self._neg_c[0], self._lq_c[0], self._ncount[0]
        self.pos_local = E(max(G, 1), dtype=i32)
        self.ent = E(max(G * (n + 1), 1), dtype=i64)
        self.loss_part, self.lse = E(max(G, 1)), E(max(G, 1))
        if self.grouping:
            nb = -(-max(local_rows, 1) // (1 << self.bin_shift))
            self.nbins = nb
            self._bin_cnt = [E(nb + 1, dtype=i32) for _ in range(self.slots)]      # + PREP_NEG's work counter
            self.bin_cnt, self.bin_off, self.bin_cursor = self._bin_cnt[0], E(nb + 1, dtype=i32), E(nb * 8, dtype=i32)
            self.bin_status, self.bin_ticket = E(nb, dtype=i64), torch.zeros(1, dtype=i32, device=dev)
            self.bin_heavy = E(int(L.rsb200_bin_heavy_elems()), dtype=i32)
            self.slot_neg = self.slot_pos = self.off = self.urow = self.scan_tmp = None
            self.scan_elems = 0
        else:
            self.slot_neg, self.slot_pos = E(max(G * n, 1), dtype=i32), E(max(G, 1), dtype=i32)
            self.off, self.urow = E(local_rows + 1, dtype=i32), E(self.cap, dtype=i32)
            self.scan_elems = int(L.rsb200_scan_tmp_elems(local_rows))
            self.scan_tmp = E(self.scan_elems, dtype=i64)
        self.err, self.totals = torch.zeros(1, dtype=i32, device=dev), torch.zeros(2, dtype=i32, device=dev)
        self.sp2 = torch.zeros(2, max(G, 1), dtype=f32, device=dev)         # [0] positive scores, [1] log Q(pos) (popularity regen)
        self.sp = self.sp2[0]
        self.stats_all, self.dq = E(world, max(G, 1), 2), E(max(G, 1), d)
        self.loss = E(1)
        self.item_rows, self.item_vals = E(self.cap, dtype=i64), E(self.cap, d)
        self._args = None
        self._sum_lqp = False

    def bind(self, q_all, pos_all, neg_all, loss_kind, score_kind, logq_pos=None, logq_neg=None, grad_scale: float = 1.0,
             regen_state: Optional[torch.Tensor] = None, pop: Optional["PopularSlice"] = None, slot: int = 0):
        """``neg_all`` [G, n] int32 GLOBAL ids -- or ``None`` with ``regen_state`` [world, 2] int64 (seed, philox offset of
        every rank's CUDA generator): the owner then recomputes every rank's draw itself (nothing id-sized crosses NVLink):
        ``torch.randint(1, N, (B, n))`` (UniformSampler), or with ``pop`` (this owner's ``PopularSlice``)
        ``searchsorted(table, torch.rand(B, n))`` of the PopularSamplerModel incl. its log-probabilities."""
        G, n, d = self.G, self.n, self.d
        assert q_all.shape == (G, d) and q_all.dtype == torch.float32 and pos_all.shape == (G,) and pos_all.dtype == torch.int64
        if regen_state is None:
            assert neg_all.shape == (G, n) and neg_all.dtype == torch.int32 and pop is None
        else:
            assert neg_all is None and regen_state.shape == (self.world, 2) and regen_state.dtype == torch.int64 \
                and regen_state.is_cuda and logq_neg is None and G % self.world == 0
            if pop is not None and not self.grouping:
                raise _lib.Rsb200Error("owner-side regeneration of the popularity draw needs the binned grouping")
        if (logq_neg is not None or pop is not None) and self.lq_c is None:
            raise _lib.Rsb200Error("OwnerComputeCuda was built without the logq workspace (with_logq=True)")
        self._keep = [t.contiguous() if t is not None else None for t in (q_all, pos_all, neg_all, logq_pos, logq_neg, regen_state)]
        q_all, pos_all, neg_all, logq_pos, logq_neg, regen_state = self._keep
        self._pop = pop
        self.neg_c, self.lq_c, self.ncount = self._neg_c[slot], self._lq_c[slot], self._ncount[slot]
        if self.grouping:
            self.bin_cnt = self._bin_cnt[slot]
        a = _lib.ShardArgs()
        P = _lib.ptr
        a.w_local, a.q_all, a.pos, a.neg = P(self.weight), P(q_all), P(pos_all), (P(neg_all) if neg_all is not None else None)
        if regen_state is not None:
            dev = self.weight.device
            sm, mt, _, _ = _lib.device_info(dev.index if dev.index is not None else torch.cuda.current_device())
            a.regen_state, a.regen_B, a.regen_sm_count, a.regen_max_threads_per_sm = P(regen_state), G // self.world, sm, mt
        # SampledSoftmax with a regenerated popularity draw: log Q(pos) comes from the positive's owner (summed with sp)
        self._sum_lqp = pop is not None and int(loss_kind) == _lib.LOSS_SSM
        if pop is not None:
            a.regen_kind = 1
            a.pop_table_local, a.pop_prob_local = P(pop.table), P(pop.prob)
            a.pop_guide_local, a.pop_guide_bits, a.pop_guide_k0 = P(pop.guide), pop.guide_bits, pop.k0
            a.pop_cdf_lo, a.pop_cdf_hi = pop.cdf_lo, pop.cdf_hi
            if self._sum_lqp:
                a.lq_pos_out = P(self.sp2[1])
                logq_pos = self.sp2[1]
        a.logq_pos = P(logq_pos) if logq_pos is not None else None
        a.logq_neg = P(logq_neg) if logq_neg is not None else None
        a.grad_scale_dev = None
        a.sp, a.stats_all, a.dq, a.loss = P(self.sp), P(self.stats_all), P(self.dq), P(self.loss)
        a.item_rows, a.item_vals, a.totals = P(self.item_rows), P(self.item_vals), P(self.totals)
        a.neg_c, a.slot_neg = P(self.neg_c), P(self.slot_neg)
        a.lq_c = P(self.lq_c) if self.lq_c is not None else None
        a.ncount, a.pos_local, a.slot_pos = P(self.ncount), P(self.pos_local), P(self.slot_pos)
        a.off, a.urow, a.ent = P(self.off), P(self.urow), P(self.ent)
        a.loss_part, a.lse, a.scan_tmp, a.err_flag = P(self.loss_part), P(self.lse), P(self.scan_tmp), P(self.err)
        a.G, a.n, a.d, a.num_items, a.row0, a.local_rows = G, n, d, self.num_items, self.row0, self.local_rows
        a.cap, a.scan_tmp_elems, a.grad_scale = self.cap, self.scan_elems, float(grad_scale)
        a.world, a.rank, a.loss_kind, a.score_kind = self.world, self.rank, int(loss_kind), int(score_kind)
        a.sink, a.accumulate = _lib.SINK_COMPACT, 0
        a.grouping, a.bin_shift = self.grouping, self.bin_shift
        if self.grouping:
            a.bin_cnt, a.bin_off, a.bin_cursor = P(self.bin_cnt), P(self.bin_off), P(self.bin_cursor)
            a.bin_status, a.bin_ticket, a.bin_heavy = P(self.bin_status), P(self.bin_ticket), P(self.bin_heavy)
        self._args = a

    def _run(self, phases: int, what: str, args=None):
        with torch.cuda.device(self.weight.device):
            _lib.check(_lib.lib().rsb200_shard_step(C.byref(args if args is not None else self._args), phases, _lib.stream_ptr()), what)

    def take_args(self):
        """the argument block of the last ``bind`` (detached: a later ``bind`` does not change it).  With ``slots`` > 1 this is
        how the negatives of the NEXT step are prepared ahead: ``bind(..., slot=s)``, ``a = take_args()``, and later
        ``prep_neg(a)`` on a side stream while the current step runs on the buffers of the other slot."""
        a, self._args = self._args, None
        return a, list(self._keep)

    def prep(self) -> torch.Tensor:            # -> sp[G] (or [2, G] with log Q(pos))   (all-reduce SUM next)
        self._run(_lib.SHARD_PREP, "shard_step(PREP)")
        return self.sp2[:, :self.G] if self._sum_lqp else self.sp[:self.G]

    def prep_neg(self, args=None):             # the half of PREP that does not need q_all / pos: may overlap their all-gather
        self._run(_lib.SHARD_PREP_NEG, "shard_step(PREP_NEG)", args)

    def prep_pos(self) -> torch.Tensor:        # the other half (+ the bin scan); same return value as prep()
        self._run(_lib.SHARD_PREP_POS, "shard_step(PREP_POS)")
        return self.sp2[:, :self.G] if self._sum_lqp else self.sp[:self.G]

    def fwd(self) -> torch.Tensor:             # -> this owner's stats slice [G, 2]   (all-gather next)
        self._run(_lib.SHARD_FWD, "shard_step(FWD)")
        return self.stats_all[self.rank, :self.G]

    def finish(self):                          # -> (loss[1], dq[G, d])   (all-reduce SUM of dq next)
        self._run(_lib.SHARD_FINISH, "shard_step(FINISH)")
        return self.loss, self.dq[:self.G]

    def scatter(self):                         # -> (rows[cap], vals[cap, d], totals): R = totals[1] on the device
        self._run(_lib.SHARD_SCATTER, "shard_step(SCATTER)")
        return self.item_rows, self.item_vals, self.totals

    def check(self):
        if int(self.err.item()):
            raise _lib.Rsb200Error("shard_step: item id outside [0, num_items)")


class PopularSlice:
    """One owner's share of a ``PopularSamplerModel`` (recstudio/ann/sampler.py:224-258) for the owner-side
    regeneration of the draw: the slices ``table[row0 : row0 + L]`` / ``pop_prob[row0 : row0 + L]`` of the sampler's
    buffers (built ONCE, on the CPU, by the very torch ops of the reference constructor, so every rank holds bit-identical
    values), the two CDF values that bound the owner's share of the unit interval, and a guide table over that share."""

    def __init__(self, table_cpu: torch.Tensor, prob_cpu: torch.Tensor, row0: int, local_rows: int, device):
        N = table_cpu.numel()
        assert prob_cpu.numel() == N and 0 <= row0 and row0 + local_rows <= N and local_rows >= 1
        self.row0, self.local_rows = row0, local_rows
        self.cdf_lo = float(table_cpu[row0 - 1]) if row0 > 0 else float("-inf")
        self.cdf_hi = float(table_cpu[row0 + local_rows - 1]) if row0 + local_rows < N else float("inf")
        self.table = table_cpu[row0:row0 + local_rows].to(device).contiguous()
        self.prob = prob_cpu[row0:row0 + local_rows].to(device).contiguous()
        bits = 1
        while (1 << bits) * 8 < N and bits < 24:           # ~8 table entries per guide bucket, as the single-GPU sampler
            bits += 1
        self.guide_bits = bits
        K = 1 << bits
        lo = max(self.cdf_lo, 0.0)
        hi = min(self.cdf_hi, 1.0)
        self.k0 = int(lo * K)
        k1 = min(int(hi * K), K - 1)
        length = k1 - self.k0 + 2
        self.guide = torch.empty(length, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().rsb200_popular_build_guide_range(_lib.ptr(self.table), local_rows, bits, self.k0, length,
                                                                   _lib.ptr(self.guide), _lib.stream_ptr()), "build_guide_range")

    @staticmethod
    def tables(pop_count, mode: int = 0):
        """(table, pop_prob) of PopularSamplerModel(pop_count, mode) on the CPU: sampler.py:225-241 op for op."""
        with torch.no_grad():
            pc = torch.as_tensor(pop_count).to(torch.float)
            if mode == 0:
                pc = torch.log(pc + 1)
            elif mode == 1:
                pc = torch.log(pc + 1) + 1e-6
            elif mode == 2:
                pc = pc ** 0.75
            pc[0] = 1
            prob = pc / pc.sum()
            table = torch.cumsum(prob, dim=0)
            prob[-1] = 1.0
        return table, prob


def uniform_regen_state(device, B: int, n: int, group=None, generator=None) -> torch.Tensor:
    """[world, 2] int64 (seed, philox offset) of every rank's CUDA generator, all-gathered (16 bytes per rank), and this
    rank's generator advanced by exactly what ``torch.randint(1, N, (B, n), device=cuda)`` would have consumed: the
    owners regenerate the UniformSampler draws from these states instead of receiving the ids."""
    from . import sampling
    device = torch.device(device)
    gen = sampling._generator(device, generator)
    seed, off = gen.initial_seed(), gen.get_offset()
    mine = torch.tensor([[seed - (1 << 64) if seed >= (1 << 63) else seed, off]], dtype=torch.int64).pin_memory()
    gen.set_offset(off + sampling.counter_offset(B * n, device))
    return _all_gather_cat(mine.to(device, non_blocking=True), group)


class DrawStates:
    """The generator states of all ranks, resident on the device: what ``uniform_regen_state`` returns, without the
    per-step host-to-device copy and all-gather.  ``resync()`` (collective) gathers every rank's current (seed, offset);
    ``next()`` hands out the [world, 2] state tensor for this step's draw of B x n numbers per rank and advances the
    device copy and this rank's torch generator by what the draw consumes -- valid as long as nothing else consumes the
    generators between steps (``next()`` checks the local one and raises, asking for a ``resync()`` on all ranks)."""

    def __init__(self, device, B: int, n: int, group=None, generator=None):
        from . import sampling
        self.device, self.group = torch.device(device), group
        self.gen = sampling._generator(self.device, generator)
        self.inc = sampling.counter_offset(B * n, self.device)
        self.B, self.n = B, n
        self._inc_dev = torch.tensor([0, self.inc], dtype=torch.int64, device=self.device)
        self.resync()

    def resync(self):
        seed, off = self.gen.initial_seed(), self.gen.get_offset()
        mine = torch.tensor([[seed - (1 << 64) if seed >= (1 << 63) else seed, off]], dtype=torch.int64).to(self.device)
        self.state = _all_gather_cat(mine, self.group)
        self._expect = (seed, off)
        self._pending = False

    def next(self) -> torch.Tensor:
        if (self.gen.initial_seed(), self.gen.get_offset()) != self._expect:
            raise _lib.Rsb200Error("DrawStates: the CUDA generator was used outside the sharded step; call resync() on every rank")
        if self._pending:                      # the previous step's PREP is already enqueued on this stream: advance after it
            self.state += self._inc_dev
        self._pending = True
        seed, off = self._expect
        self.gen.set_offset(off + self.inc)
        self._expect = (seed, off + self.inc)
        return self.state


def owner_compute_step(engine, q: torch.Tensor, pos: torch.Tensor, neg: Optional[torch.Tensor], loss_kind: int, score_kind: int,
                       logq_pos: Optional[torch.Tensor] = None, logq_neg: Optional[torch.Tensor] = None, group=None,
                       gathered=None, regen_state: Optional[torch.Tensor] = None, pop: Optional[PopularSlice] = None):
    """One data-parallel step over the row-sharded item table WITHOUT moving rows.

    Every rank passes its own B queries (vectors ``q`` [B, d], GLOBAL ids ``pos`` [B], ``neg`` [B, n] int32);
    ``engine`` holds this rank's row block (``OwnerComputeCuda`` on GPUs).  With ``regen_state``
    (``uniform_regen_state(...)``) the negatives are not passed at all: every owner regenerates every rank's draw --
    UniformSampler, or PopularSamplerModel when ``pop`` (this owner's ``PopularSlice``) is given.  Returns
    ``(loss, (rows, vals, totals), dq_all)``: the global mean loss (identical on every rank), the gradient rows
    of the rows THIS rank owns (LOCAL ids, ``R = totals[1]`` valid rows) and ``d loss / d query`` of all
    ``G = world x B`` queries (rank r's queries are ``dq_all[r*B:(r+1)*B]``).  No host synchronisation."""
    if gathered is None and regen_state is not None:          # negatives regenerated by the owners: no id exchange
        q_all, pos_all, neg_all, lqp, lqn = _all_gather_cat(q, group), _all_gather_cat(pos, group), None, None, None
    elif gathered is None:
        q_all, pos_all, neg_all = _all_gather_cat(q, group), _all_gather_cat(pos, group), _all_gather_cat(neg, group)
        lqp = _all_gather_cat(logq_pos, group) if logq_pos is not None else None
        lqn = _all_gather_cat(logq_neg, group) if logq_neg is not None else None
    else:
        q_all, pos_all, neg_all, lqp, lqn = gathered
    if regen_state is not None:
        kw = {"pop": pop} if pop is not None else {}
        engine.bind(q_all, pos_all, None, loss_kind, score_kind, lqp, lqn, regen_state=regen_state, **kw)
    else:
        engine.bind(q_all, pos_all, neg_all, loss_kind, score_kind, lqp, lqn)
    sp = engine.prep()
    dist.all_reduce(sp, group=group)                                   # every positive has exactly one owner
    mine = engine.fwd()
    exchange_stats(engine, mine, group)
    loss, dq = engine.finish()
    work = dist.all_reduce(dq, group=group, async_op=True)             # runs on the communicator's stream ...
    rows = engine.scatter()                                            # ... while the owner-local scatter runs here
    work.wait()
    return loss, rows, dq


def exchange_stats(engine, mine: torch.Tensor, group=None):
    """all-gather of the per-owner [G, 2] statistics into engine.stats_all[world, G, 2]"""
    world = dist.get_world_size(group)
    st = engine.stats_all
    if st.is_cuda and st.shape[1] == engine.G:
        dist.all_gather_into_tensor(st, mine.clone(), group=group)
    else:
        dist.all_gather([st[r, :engine.G] for r in range(world)], mine.clone(), group=group)


def owner_compute_training_step(items: ShardedRows, engine, w_user: torch.Tensor, user: torch.Tensor, pos: torch.Tensor,
                                neg: Optional[torch.Tensor], loss_kind: int, score_kind: int, logq_pos=None, logq_neg=None,
                                regen_state: Optional[torch.Tensor] = None):
    """Same contract as ``sharded_training_step`` (replicated user table), on the owner-compute path.  ``neg=None`` with
    ``regen_state=uniform_regen_state(...)``: UniformSampler negatives regenerated by the owners (no id exchange)."""
    q = items.ops.gather_rows(w_user, user)
    user_all = _all_gather_cat(user, items.group)
    loss, (rows, vals, totals), dq_all = owner_compute_step(engine, q, pos, neg.to(torch.int32) if neg is not None else None,
                                                            loss_kind, score_kind, logq_pos, logq_neg, group=items.group,
                                                            regen_state=regen_state)
    r = int(totals[1].item())
    ur, uv = items.ops.coalesce_rows(user_all, dq_all, w_user.shape[0], skip_row0=True)
    return loss, (rows[:r], vals[:r]), (ur, uv)
