"""S1 / S2: the sampled indices must be BIT-IDENTICAL to the reference's ATen calls on
CUDA (torch.randint / torch.rand + searchsorted) for the same generator state, and to the
numpy Philox oracle.  Runs live against the installed torch (version recorded in the
assertion messages): the arithmetic lives in ATen, a dependency that is not under
/root/reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import philox, samplers as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _policy():
    p = torch.cuda.get_device_properties(0)
    return p.multi_processor_count, p.max_threads_per_multi_processor


@pytest.mark.parametrize("shape", [(1, 1), (2, 4), (512, 1), (77, 33), (1184, 1025), (8192, 1024)])
@pytest.mark.parametrize("num_items", [10, 1575, 10_000_001])
def test_uniform_matches_torch_randint(shape, num_items):
    from recstudio_b200 import sampling
    Q, n = shape
    for seed, burn in ((2022, 0), (7, 3)):
        torch.manual_seed(seed)
        for _ in range(burn):                       # move the generator off offset 0
            torch.rand(1000, device=DEV)
        gen = torch.cuda.default_generators[0]
        off0 = gen.get_offset()
        want = torch.randint(1, num_items, (Q, n), device=DEV)        # recstudio/ann/sampler.py:102-104
        off_ref = gen.get_offset()
        gen.set_offset(off0)
        got64, got32 = sampling.uniform_draw(num_items, Q, n, DEV, want_i64=True, want_i32=True)
        assert gen.get_offset() == off_ref, f"generator advance differs (torch {torch.__version__})"
        assert got64.dtype == torch.int64 and torch.equal(got64, want), f"torch {torch.__version__}"
        assert torch.equal(got32.long(), want)
        sm, mt = _policy()
        ora = philox.torch_cuda_randint(seed, off0, 1, num_items, Q * n, sm, mt).reshape(Q, n)
        assert np.array_equal(ora, want.cpu().numpy()), "numpy Philox oracle disagrees with torch CUDA"
        assert philox.torch_cuda_counter_offset(Q * n, sm, mt) == off_ref - off0


def test_uniform_sampler_oracle_contract():
    sm, mt = _policy()
    torch.manual_seed(11)
    want = torch.randint(1, 1000, (5, 9), device=DEV).cpu().numpy()
    lp, neg, ln, off = S.uniform_sampler(11, 0, 1000, 5, 9, np.array([1, 2, 3, 4, 5]), sm, mt)
    assert np.array_equal(neg, want) and off == torch.cuda.default_generators[0].get_offset()
    assert lp.dtype == np.int64 and not lp.any() and not ln.any()


@pytest.mark.parametrize("shape", [(64, 33), (8192, 128), (1000, 1024)])
@pytest.mark.parametrize("use_guide", [False, True])
def test_popular_matches_torch(shape, use_guide):
    from recstudio_b200 import sampling
    g = load_golden("popular")
    Q, n = shape
    for tag in ("big_m0", "big_m2", "small_m0"):
        table = torch.from_numpy(g[f"{tag}_table"]).to(DEV)
        prob = torch.from_numpy(g[f"{tag}_prob"]).to(DEV)
        guide, bits = (sampling.build_guide(table) if use_guide else (None, 0))
        torch.manual_seed(99)
        gen = torch.cuda.default_generators[0]
        seeds = torch.rand(Q, n, device=DEV)                              # sampler.py:246
        want = torch.searchsorted(table, seeds)                           # sampler.py:247
        want = want.clamp(max=table.numel() - 1)
        want_lq = torch.log(prob[want])                                   # sampler.py:257-258
        off_ref = gen.get_offset()
        gen.set_offset(0)
        ids, ids32, lq = sampling.popular_draw(table, prob, Q, n, guide=guide, guide_bits=bits, want_i32=True)
        assert gen.get_offset() == off_ref
        assert torch.equal(ids, want), f"{tag} guide={use_guide} torch {torch.__version__}"
        assert torch.equal(ids32.long(), want)
        assert torch.equal(lq, want_lq), "log-probabilities must be bit-identical (same libdevice logf)"
        sm, mt = _policy()
        o_seeds = philox.torch_cuda_rand(99, 0, Q * n, sm, mt)
        assert np.array_equal(o_seeds, seeds.cpu().numpy().reshape(-1)), "numpy oracle of torch.rand disagrees"
        o_idx, _ = S.popular_draw(g[f"{tag}_table"], g[f"{tag}_prob"], o_seeds)
        assert np.array_equal(o_idx.reshape(Q, n), want.cpu().numpy())


def test_popular_large_table_guide_equals_bisection():
    """100M-row class table (C5 shape scaled to 20M here): the guided search must return
    exactly what the plain bisection returns, including id 0 and the tail."""
    from recstudio_b200 import sampling
    N = 20_000_000
    torch.manual_seed(0)
    cnt = torch.empty(N, device=DEV).exponential_(1.0).mul_(3).floor_()
    pc = torch.log(cnt + 1); pc[0] = 1
    prob = (pc / pc.sum()).float()
    table = torch.cumsum(prob.double(), 0).float()
    guide, bits = sampling.build_guide(table)
    torch.manual_seed(5)
    a, _, la = sampling.popular_draw(table, prob, 4096, 512, guide=guide, guide_bits=bits)
    torch.manual_seed(5)
    b, _, lb = sampling.popular_draw(table, prob, 4096, 512)
    assert torch.equal(a, b) and torch.equal(la, lb)
    torch.manual_seed(5)
    want = torch.searchsorted(table, torch.rand(4096, 512, device=DEV)).clamp(max=N - 1)
    assert torch.equal(a, want)


def test_popular_logq_matches_compute_item_p():
    from recstudio_b200 import sampling
    g = load_golden("popular")
    prob = torch.from_numpy(g["big_m0_prob"]).to(DEV)
    ids = torch.randint(0, prob.numel(), (100, 7), device=DEV)
    assert torch.equal(sampling.popular_logq(prob, ids), torch.log(prob[ids]))
