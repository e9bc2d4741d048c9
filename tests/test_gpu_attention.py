"""A1: the tcgen05 attention core.  bf16 operands with fp32 accumulation cannot meet the 1e-5
of the fp32 paths: the tolerance is 2e-2 relative to the largest output of an
fp32 evaluation of the reference formula on the same (bf16-rounded) inputs, and 3e-2 against the
fp32 formula on the unrounded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _lib():
    from recstudio_b200 import _lib
    return _lib


@pytest.mark.parametrize("N,K", [(16, 16), (64, 64), (256, 64), (64, 256), (256, 256), (112, 48)])
def test_tcgen05_gemm_plumbing(N, K):
    """descriptors / TMEM / commit / tcgen05.ld against torch on exactly representable bf16 inputs"""
    L = _lib()
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = (torch.randint(-8, 9, (128, K), generator=g).float() / 4).to(DEV)       # exact in bf16, sums exact in fp32
    B = (torch.randint(-8, 9, (N, K), generator=g).float() / 4).to(DEV)
    D = torch.full((128, N), float("nan"), device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.check(L.lib().rsb200_tc_gemm_test(L.ptr(A), L.ptr(B), L.ptr(D), N, K, L.ptr(err), L.stream_ptr()), "tc_gemm_test")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "tensor-core barrier timed out"
    assert torch.equal(D, A @ B.T)


def _ref_attention(q, k, v, hist, heads, causal):
    B, Lq, d = q.shape
    dh = d // heads
    qh, kh, vh = (t.reshape(B, Lq, heads, dh).transpose(1, 2) for t in (q, k, v))      # [B, H, L, dh]
    s = (qh @ kh.transpose(-1, -2)) / dh ** 0.5
    mask = torch.zeros(B, 1, Lq, Lq, dtype=torch.bool, device=q.device)
    if causal:
        mask |= torch.triu(torch.ones(Lq, Lq, dtype=torch.bool, device=q.device), 1)        # sasrec.py:47
    if hist is not None:
        mask |= (hist == 0)[:, None, None, :]                                                  # sasrec.py:44,53
    s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, d), torch.logsumexp(s, -1)


@pytest.mark.parametrize("Lq,causal", [(200, 1), (200, 0), (37, 1), (128, 1), (256, 1), (129, 0)])
def test_attention_forward(Lq, causal):
    L = _lib()
    B, heads, dh = 5, 2, 64
    d = heads * dh
    gen = torch.Generator(device=DEV).manual_seed(Lq)
    q, k, v = (torch.randn(B, Lq, d, device=DEV, generator=gen) for _ in range(3))
    seqlen = torch.randint(1, Lq + 1, (B,), device=DEV, generator=gen)
    hist = (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None]).long() * 7              # right-padded ids
    out = torch.full((B, Lq, d), float("nan"), device=DEV)
    lse = torch.empty(B, heads, Lq, device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.check(L.lib().rsb200_attn_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(hist), B, Lq, heads, dh, causal, L.ptr(out),
                                    L.ptr(lse), L.ptr(err), L.stream_ptr()), "attn_fwd")
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    rb = lambda t: t.bfloat16().float()
    want_b, lse_b = _ref_attention(rb(q), rb(k), rb(v), hist, heads, causal)
    want_f, _ = _ref_attention(q, k, v, hist, heads, causal)
    # rows whose every key is masked (bidirectional + all-padding never happens: seqlen >= 1) are NaN in torch
    valid = ~torch.isnan(want_b).any(-1)
    scale = want_f[valid].abs().max().item()
    assert (out[valid] - want_b[valid]).abs().max().item() <= 2e-2 * scale
    assert (out[valid] - want_f[valid]).abs().max().item() <= 3e-2 * scale
    lv = torch.isfinite(lse_b)
    assert (lse[lv] - lse_b[lv]).abs().max().item() <= 2e-2
