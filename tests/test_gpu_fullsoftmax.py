"""L3: full-catalog SoftmaxLoss (rsb200_fullsoftmax_fwd_bwd) against the reference's golden
vectors, the ATen restatement, and size-independent properties at the config-4 shape."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import retriever as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


def _close(got, want, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got - want).max() <= rtol * scale, (np.abs(got - want).max(), scale)


def _run(w_item, w_user, user, pos):
    from recstudio_b200 import plugins
    wi = torch.as_tensor(w_item).to(DEV).requires_grad_(True)
    wu = torch.as_tensor(w_user).to(DEV).requires_grad_(True)
    q = torch.nn.functional.embedding(torch.as_tensor(user).to(DEV), wu, padding_idx=0)
    loss = plugins._FullSoftmaxFn.apply(q, wi, torch.as_tensor(pos).to(DEV))
    loss.backward()
    return loss.item(), wi.grad.cpu().numpy(), wu.grad.cpu().numpy()


def test_golden_full_softmax_step():
    g = load_golden("step_full_softmax")
    loss, di, du = _run(g["w_item"], g["w_user"], g["user"], g["pos"])
    assert abs(loss - g["loss"].item()) <= RTOL * abs(g["loss"].item())
    _close(di, g["d_item"]); _close(du, g["d_user"])
    assert np.all(di[0] == 0)


@pytest.mark.parametrize("shape", [(300, 40, 64, 7), (1000, 60, 128, 130), (2049, 33, 100, 257), (129, 20, 32, 300)])
def test_random_vs_aten(shape):
    N, U, d, B = shape
    g = torch.Generator().manual_seed(N + B)
    wi = torch.randn(N, d, generator=g) * 0.4; wi[0] = 0
    wu = torch.randn(U, d, generator=g) * 0.4; wu[0] = 0
    user = torch.randint(1, U, (B,), generator=g)
    pos = torch.randint(1, N, (B,), generator=g)
    pos[::9] = 0                                     # padding positives: no gradient to row 0
    ref = R.full_softmax_step_aten(wi, wu, user, pos)
    loss, di, du = _run(wi, wu, user, pos)
    assert abs(loss - ref["loss"].item()) <= RTOL * abs(ref["loss"].item())
    _close(di, ref["d_item"].numpy()); _close(du, ref["d_user"].numpy())


def test_standalone_softmax_loss_plugin():
    from recstudio_b200 import plugins
    g = load_golden("step_full_softmax")
    ps = torch.from_numpy(g["pos_score"]).to(DEV).requires_grad_(True)
    al = torch.from_numpy(g["all_score"]).to(DEV).requires_grad_(True)
    loss = plugins.FusedSoftmaxLoss()(None, ps, al)
    loss.backward()
    a = torch.from_numpy(g["pos_score"]).requires_grad_(True); b = torch.from_numpy(g["all_score"]).requires_grad_(True)
    want = R.softmax_loss(a, b); want.backward()
    assert abs(loss.item() - want.item()) <= RTOL * abs(want.item())
    _close(ps.grad.cpu(), a.grad); _close(al.grad.cpu(), b.grad)
    ga = load_golden("appendix_a")
    for kind in ("ip", "eu"):
        v = plugins.FusedSoftmaxLoss()(None, torch.from_numpy(ga[f"{kind}_pos"]).to(DEV), torch.from_numpy(ga[f"{kind}_neg"]).to(DEV))
        assert abs(v.item() - ga[f"{kind}_softmax"].item()) < 5e-6


def test_fused_retriever_full_softmax_path():
    from recstudio_b200 import plugins, retriever
    U, N, d, B = 64, 3000, 64, 200
    m = retriever.build_synthetic(U, N, d, 0, device=DEV, init_std=0.3)
    m.loss_fn = plugins.FusedSoftmaxLoss(); m.sampler = None
    gen = torch.Generator().manual_seed(3)
    batch = {"user_id": torch.randint(1, U, (B,), generator=gen).to(DEV), "item_id": torch.randint(1, N, (B,), generator=gen).to(DEV),
             "rating": torch.ones(B, device=DEV)}
    loss = m.training_step(batch); loss.backward()
    ref = R.full_softmax_step_aten(m.item_encoder.weight.detach().cpu(), m.query_encoder.weight.detach().cpu(),
                                   batch["user_id"].cpu(), batch["item_id"].cpu())
    assert abs(loss.item() - ref["loss"].item()) <= RTOL * abs(ref["loss"].item())
    _close(m.item_encoder.weight.grad.cpu(), ref["d_item"].numpy()); _close(m.query_encoder.weight.grad.cpu(), ref["d_user"].numpy())


def test_config4_shape_properties():
    """1,000,001 x 128, B = 1024: loss = ln(N-1) at tiny init; rows of dS sum to 0 so that
    sum_rows dW = 0-weighted checks hold: column sums of dW equal -(1/B) sum_b q_b + (1/B) sum_b q_b = 0,
    and dQ_b = E_softmax[w] - w_pos; compared against torch on the same device for 8 queries."""
    N, U, d, B = 1_000_001, 5001, 128, 1024
    gen = torch.Generator(device=DEV).manual_seed(1)
    wi = torch.randn(N, d, device=DEV, generator=gen) * 0.1; wi[0] = 0
    wu = torch.randn(U, d, device=DEV, generator=gen) * 0.1; wu[0] = 0
    user = torch.randint(1, U, (B,), device=DEV, generator=gen)
    pos = torch.randint(1, N, (B,), device=DEV, generator=gen)
    from recstudio_b200 import plugins
    wi.requires_grad_(True)
    q = wu[user].clone().requires_grad_(True)
    loss = plugins._FullSoftmaxFn.apply(q, wi, pos)
    loss.backward()
    with torch.no_grad():
        s = q[:8] @ wi[1:].T
        lse = torch.logsumexp(s, -1)
        p = torch.softmax(s, -1)
        want_dq = (p @ wi[1:] - wi[pos[:8]]) / B
        assert (q.grad[:8] - want_dq).abs().max().item() <= 1e-5 * want_dq.abs().max().item()
        full_lse = torch.cat([torch.logsumexp(q[i:i + 128] @ wi[1:].T, -1) for i in range(0, B, 128)])
        want_loss = (full_lse - (q * wi[pos]).sum(-1)).mean()
        assert abs(loss.item() - want_loss.item()) <= 1e-5 * abs(want_loss.item())
        colsum = wi.grad.double().sum(0).abs().max().item()          # sum_i dS_bi = 0 for every query
        assert colsum <= 1e-4 * wi.grad.double().abs().sum(0).max().item()
        assert float(wi.grad[0].abs().sum()) == 0.0
