"""TEST INFRASTRUCTURE ONLY -- CPU restatement of one owner's share of the owner-compute step of
the row-sharded table (recstudio_b200/csrc/shard.cu), phase by phase, with the interface of
``recstudio_b200.sharded.OwnerComputeCuda`` so that the gloo test can inject it.

The reference has no sharded step; what pins this file is that the SUM over owners of what it
computes must equal the reference step on the whole table (``oracle.retriever.training_step_aten``,
which follows baseretriever.py:142-176,399-404 / scorer.py:5-34 / loss_func.py:55-59,80-90 op for
op): tests/test_sharded_gloo.py checks exactly that.  Plain per-query loops in float64 -- small
cases only.
"""
import numpy as np
import torch

BPR, SSM = 0, 1
IP, EUCLID = 0, 1


class OwnerComputeOracle:
    def __init__(self, num_items, row0, local_rows, weight, world, rank, G, n, with_logq=False):
        self.w = weight.double().numpy()
        self.num_items, self.row0, self.local_rows = num_items, row0, local_rows
        self.world, self.rank, self.G, self.n, self.d = world, rank, G, n, weight.shape[1]
        self.stats_all = torch.zeros(world, G, 2, dtype=torch.float32)

    def _score(self, q, v):
        return float(q @ v) if self.score_kind == IP else -float(((q - v) ** 2).sum())

    def bind(self, q_all, pos_all, neg_all, loss_kind, score_kind, logq_pos=None, logq_neg=None, grad_scale=1.0):
        self.q = q_all.double().numpy(); self.pos = pos_all.numpy(); self.neg = neg_all.numpy().astype(np.int64)
        self.lqp = logq_pos.double().numpy() if logq_pos is not None else np.zeros(self.G)
        self.lqn = logq_neg.double().numpy() if logq_neg is not None else np.zeros((self.G, self.n))
        self.loss_kind, self.score_kind = loss_kind, score_kind
        denom = self.G * (self.n if loss_kind == BPR else 1)
        self.coef, self.lscale = grad_scale / denom, 1.0 / denom

    def _owned(self, gid):
        return self.row0 <= gid < self.row0 + self.local_rows

    def prep(self):
        self.sp = torch.zeros(self.G, dtype=torch.float32)
        for b in range(self.G):
            if self._owned(self.pos[b]):
                self.sp[b] = self._score(self.q[b], self.w[self.pos[b] - self.row0])
        return self.sp

    def fwd(self):
        sp = self.sp.double().numpy()                       # all-reduced by the caller
        self.raw = np.zeros((self.G, self.d)); self.touch = []
        for b in range(self.G):
            own = [(j, g - self.row0) for j, g in enumerate(self.neg[b]) if self._owned(g)]
            s = np.array([self._score(self.q[b], self.w[l]) for _, l in own])
            if self.loss_kind == BPR:
                c = self.coef / (1.0 + np.exp(-(s - sp[b]))) if len(own) else s
                self.stats_all[self.rank, b, 0] = float(c.sum())
                self.stats_all[self.rank, b, 1] = float(np.logaddexp(0.0, s - sp[b]).sum())
                wgt = c
            else:
                z = s - np.array([self.lqn[b, j] for j, _ in own])
                m = z.max() if len(own) else -np.inf
                wgt = np.exp(z - m) if len(own) else z
                self.stats_all[self.rank, b, 0] = float(m)
                self.stats_all[self.rank, b, 1] = float(wgt.sum())
                c = z                                        # entry keeps the logit
            for (j, l), ww, cc in zip(own, wgt, c):
                self.raw[b] += ww * self.w[l]
                self.touch.append((b, l, cc))
        return self.stats_all[self.rank]

    def finish(self):
        st = self.stats_all.double().numpy(); sp = self.sp.double().numpy()
        dq = np.zeros((self.G, self.d)); loss = 0.0
        self.lse = np.zeros(self.G); self.cpos = np.zeros(self.G)
        for b in range(self.G):
            if self.loss_kind == BPR:
                CS = st[:, b, 0].sum(); cpos = -CS; mul = 1.0; cs_own = st[self.rank, b, 0]
                loss += st[:, b, 1].sum() * self.lscale
            else:
                z0 = sp[b] - self.lqp[b]
                terms = [st[o, b, 0] + np.log(st[o, b, 1]) for o in range(self.world) if st[o, b, 0] != -np.inf]
                lse = np.logaddexp.reduce(terms + [z0])
                self.lse[b] = lse
                cpos = (np.exp(z0 - lse) - 1.0) * self.coef
                m_own = st[self.rank, b, 0]
                mul = 0.0 if m_own == -np.inf else np.exp(m_own - lse) * self.coef
                cs_own = st[self.rank, b, 1] * mul
                loss += (lse - z0) * self.lscale
            self.cpos[b] = cpos
            a = self.raw[b] * mul
            lp = self.pos[b] - self.row0
            own = self._owned(self.pos[b])
            if self.score_kind == IP:
                dq[b] = a + (cpos * self.w[lp] if own else 0.0)
            else:
                dq[b] = 2.0 * (a - cs_own * self.q[b]) + (2.0 * cpos * (self.w[lp] - self.q[b]) if own else 0.0)
        self.dq = torch.from_numpy(dq).float()
        return torch.tensor([loss], dtype=torch.float32), self.dq

    def scatter(self):
        g = np.zeros((self.local_rows, self.d)); hit = np.zeros(self.local_rows, dtype=bool)
        touches = [(b, l, (c if self.loss_kind == BPR else np.exp(c - self.lse[b]) * self.coef)) for b, l, c in self.touch]
        touches += [(b, self.pos[b] - self.row0, self.cpos[b]) for b in range(self.G) if self._owned(self.pos[b])]
        for b, l, c in touches:
            if l + self.row0 == 0:
                continue                                     # padding row: no gradient
            hit[l] = True
            g[l] += c * self.q[b] if self.score_kind == IP else 2.0 * c * (self.q[b] - self.w[l])
        rows = np.nonzero(hit)[0]
        return (torch.from_numpy(rows), torch.from_numpy(g[rows]).float(),
                torch.tensor([len(touches), len(rows)], dtype=torch.int32))

    def check(self):
        assert ((self.neg >= 0) & (self.neg < self.num_items)).all()
