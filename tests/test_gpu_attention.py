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
    L.check(L.lib().rsb200_attn_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(hist), B, Lq, heads, dh, causal, 0.0, 0, L.ptr(out),
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


@pytest.mark.parametrize("Lq,causal", [(200, 1), (64, 1), (256, 0), (130, 1)])
def test_attention_backward(Lq, causal):
    from recstudio_b200 import attention
    B, heads, dh = 4, 2, 64
    d = heads * dh
    gen = torch.Generator(device=DEV).manual_seed(100 + Lq)
    rb = lambda t: t.bfloat16().float()
    q, k, v = (rb(torch.randn(B, Lq, d, device=DEV, generator=gen)) for _ in range(3))
    seqlen = torch.randint(1, Lq + 1, (B,), device=DEV, generator=gen)
    hist = (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None]).long() * 3
    go = rb(torch.randn(B, Lq, d, device=DEV, generator=gen))
    valid_rows = (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None]) if not causal else torch.ones(B, Lq, dtype=torch.bool, device=DEV)
    go = go * valid_rows[..., None]
    qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = attention.fused_attention(qf, kf, vf, hist, heads, bool(causal))
    out.backward(go)
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    want, _ = _ref_attention(qr, kr, vr, hist, heads, causal)
    want = torch.nan_to_num(want)
    want.backward(go)
    for got, ref, name in ((qf.grad, qr.grad, "dq"), (kf.grad, kr.grad, "dk"), (vf.grad, vr.grad, "dv")):
        ref = torch.nan_to_num(ref)
        scale = ref.abs().max().item()
        err = (got - ref).abs().max().item()
        assert err <= 3e-2 * scale, (name, err, scale)


@pytest.mark.parametrize("Lq,causal,p", [(200, 1, 0.5), (96, 0, 0.2), (256, 1, 0.5)])
def test_attention_dropout_inside_the_kernel(Lq, causal, p):
    """Attention-probability dropout (the reference's default: dropout_rate 0.5, seq/config/sasrec.yaml:5) runs in the
    tcgen05 kernels.  The keep mask is a stateless hash; rebuilt on the host (attention.dropout_keep_mask) it gives a torch
    evaluation with the SAME mask: out = (softmax(.) o M / (1 - p)) V, and dq / dk / dv through autograd."""
    from recstudio_b200 import attention
    B, heads, dh = 3, 2, 64
    d = heads * dh
    gen = torch.Generator(device=DEV).manual_seed(7 + Lq)
    rb = lambda t: t.bfloat16().float()
    q, k, v = (rb(torch.randn(B, Lq, d, device=DEV, generator=gen)) for _ in range(3))
    seqlen = torch.randint(Lq // 2, Lq + 1, (B,), device=DEV, generator=gen)
    hist = (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None]).long() * 3
    go = rb(torch.randn(B, Lq, d, device=DEV, generator=gen))
    valid_rows = (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None]) if not causal else torch.ones(B, Lq, dtype=torch.bool, device=DEV)
    go = go * valid_rows[..., None]
    key = 0x1234_5678_9ABC_DEF1
    keep = attention.dropout_keep_mask(key, p, B, heads, Lq, DEV)
    rate = keep.float().mean().item()
    assert abs(rate - (1 - p)) < 0.01, rate                                  # ~1e5 .. 4e5 Bernoulli draws
    assert abs(keep[0].float().mean().item() - keep[-1].float().mean().item()) < 0.02 and not torch.equal(keep[0], keep[1])
    qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = attention.fused_attention(qf, kf, vf, hist, heads, bool(causal), p_drop=p, drop_key=key)
    out.backward(go)
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    qh, kh, vh = (t.reshape(B, Lq, heads, dh).transpose(1, 2) for t in (qr, kr, vr))
    sc = (qh @ kh.transpose(-1, -2)) / dh ** 0.5
    mask = torch.zeros(B, 1, Lq, Lq, dtype=torch.bool, device=DEV)
    if causal:
        mask |= torch.triu(torch.ones(Lq, Lq, dtype=torch.bool, device=DEV), 1)
    mask |= (hist == 0)[:, None, None, :]
    pr = torch.softmax(sc.masked_fill(mask, float("-inf")), -1)
    pr = torch.nan_to_num(pr) * keep.float() / (1 - p)                       # F.dropout semantics on the probabilities
    want = (pr @ vh).transpose(1, 2).reshape(B, Lq, d)
    want.backward(go)
    scale = want.abs().max().item()
    assert (out - want.detach()).abs().max().item() <= 3e-2 * scale
    for got, ref, name in ((qf.grad, qr.grad, "dq"), (kf.grad, kr.grad, "dk"), (vf.grad, vr.grad, "dv")):
        ref = torch.nan_to_num(ref)
        assert (got - ref).abs().max().item() <= 3e-2 * ref.abs().max().item(), name
    # p = 0 is the plain kernel; a fresh key per call comes from torch's generator (same seed => same masks)
    torch.manual_seed(5)
    a1 = attention.fused_attention(q, k, v, hist, heads, bool(causal), p_drop=p)
    a2 = attention.fused_attention(q, k, v, hist, heads, bool(causal), p_drop=p)
    torch.manual_seed(5)
    b1 = attention.fused_attention(q, k, v, hist, heads, bool(causal), p_drop=p)
    assert torch.equal(a1, b1) and not torch.equal(a1, a2)


def test_sasrec_stock_dropout_config_takes_the_fused_path():
    """dropout 0.5 in training mode no longer falls back to nn.TransformerEncoder (VERDICT r1 missing-3)."""
    from recstudio_b200 import attention, plugins
    item = plugins.FusedEmbedding(501, 128, padding_idx=0).to(DEV)
    enc = attention.FusedSASRecQueryEncoder(fiid="item_id", embed_dim=128, max_seq_len=50, n_head=2, hidden_size=128, dropout=0.5,
                                            activation="gelu", layer_norm_eps=1e-12, n_layer=2, item_encoder=item).to(DEV)
    enc.train()
    assert enc._use_fused(50, torch.device(DEV))
    ids = torch.randint(1, 501, (4, 50), device=DEV)
    out = enc({"in_item_id": ids, "seqlen": torch.full((4,), 50, device=DEV)})
    out.sum().backward()
    assert out.shape == (4, 128) and torch.isfinite(out).all()
    assert torch.isfinite(enc.transformer_layer.layers[0].self_attn.in_proj_weight.grad).all()
    enc.eval()
    with torch.no_grad():
        e1, e2 = enc({"in_item_id": ids, "seqlen": torch.full((4,), 50, device=DEV)}), enc({"in_item_id": ids, "seqlen": torch.full((4,), 50, device=DEV)})
    assert torch.equal(e1, e2)                                               # no dropout in eval


@pytest.mark.parametrize("bidirectional", [False, True])
def test_sasrec_query_encoder_matches_reference_path(bidirectional):
    """FusedSASRecQueryEncoder (fused tcgen05 core) vs the reference's nn.TransformerEncoder path on the
    same weights (sasrec.py:37-67): output and parameter gradients, config-3 shape class (L = 200, d = 128)."""
    from recstudio_b200 import attention, plugins
    torch.manual_seed(0)
    N, d, Lq, B = 5000, 128, 200, 16
    item = plugins.FusedEmbedding(N, d).to(DEV)
    enc = attention.FusedSASRecQueryEncoder("item_id", d, Lq, 2, 128, 0.0, "gelu", 1e-12, 2, item, bidirectional=bidirectional).to(DEV)
    with torch.no_grad():
        item.weight.normal_(0, 0.5); item.weight[0] = 0
        enc.position_emb.weight.normal_(0, 0.5)
    gen = torch.Generator(device=DEV).manual_seed(1)
    seqlen = torch.randint(1, Lq + 1, (B,), device=DEV, generator=gen)
    ids = torch.randint(1, N, (B, Lq), device=DEV, generator=gen)
    ids = ids * (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None])
    batch = {"in_item_id": ids, "seqlen": seqlen}
    enc.train()
    assert enc._use_fused(Lq, ids.device)
    out = enc(batch)
    g = torch.randn_like(out)
    enc.zero_grad(); item.zero_grad()
    out.backward(g)
    grads = {n: p.grad.clone() for n, p in enc.named_parameters() if p.grad is not None}
    enc._use_fused = lambda L, dev: False                 # the reference's own path
    ref = enc(batch)
    enc.zero_grad(); item.zero_grad()
    ref.backward(g)
    assert out.shape == (B, d)
    assert (out - ref).abs().max().item() <= 3e-2 * ref.abs().max().item()
    for n, p in enc.named_parameters():
        if p.grad is None:
            continue
        r = p.grad
        if r.abs().max().item() == 0:
            continue
        assert (grads[n] - r).abs().max().item() <= 6e-2 * r.abs().max().item(), n
    # state_dict keys are those of the reference module
    keys = set(enc.state_dict().keys())
    assert "position_emb.weight" in keys and "transformer_layer.layers.0.self_attn.in_proj_weight" in keys
    assert "transformer_layer.layers.1.norm2.bias" in keys


@pytest.mark.parametrize("mode", ["dense", "sparse"])
def test_sasrec_fused_head_training_step(mode):
    """BASELINE config-3 shape class: SASRec encoder (L = 200, d = 128, 2 layers x 2 heads) + fused
    sampled-softmax head.  The head must reproduce the reference head on the SAME query vectors
    (fp32, 1e-5) and hand d loss / d query back to the encoder."""
    from oracle import retriever as R
    from recstudio_b200 import retriever
    N, d, Lq, B, n = 20_001, 128, 200, 32, 300
    m = retriever.build_sasrec_synthetic(N, d, n, max_seq_len=Lq, fused_grad=mode, device=DEV, init_std=0.1)
    gen = torch.Generator(device=DEV).manual_seed(5)
    seqlen = torch.randint(1, Lq + 1, (B,), device=DEV, generator=gen)
    ids = torch.randint(1, N, (B, Lq), device=DEV, generator=gen) * (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None])
    batch = {"in_item_id": ids, "seqlen": seqlen, "item_id": torch.randint(1, N, (B,), device=DEV, generator=gen),
             "rating": torch.ones(B, device=DEV)}
    m.train()
    torch.manual_seed(9)
    loss = m.training_step(batch)
    loss.backward()
    neg = m.fused_last_neg_id().long()
    torch.manual_seed(9)
    assert torch.equal(neg, torch.randint(1, N, (B, n), device=DEV))
    # reference head on the very same query vectors (re-encode: deterministic, dropout 0)
    with torch.no_grad():
        query = m.query_encoder(batch)
    wi = m.item_encoder.weight.detach().cpu()
    q_cpu = torch.cat([torch.zeros(1, d), query.cpu()])
    ref = R.training_step_aten(wi, q_cpu, torch.arange(1, B + 1), batch["item_id"].cpu(), neg.cpu(), loss=R.SSM, scorer=R.IP)
    assert abs(loss.item() - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
    # the encoder received gradient through d loss / d query; the shared item table got head + encoder gradients
    g_in = m.query_encoder.transformer_layer.layers[0].self_attn.in_proj_weight.grad
    assert g_in is not None and torch.isfinite(g_in).all() and g_in.abs().max().item() > 0
    gi = m.item_encoder.weight.grad
    gi = gi.to_dense() if gi.is_sparse else gi
    assert torch.isfinite(gi).all() and float(gi[0].abs().sum()) == 0.0
    # head-only part of the item gradient: rows that are NOT in any input sequence come from the head alone
    head_only = torch.ones(N, dtype=torch.bool); head_only[ids.unique().cpu()] = False
    want = ref["d_item"]
    sel = head_only & (want.abs().sum(-1) > 0)
    assert (gi.cpu()[sel] - want[sel]).abs().max().item() <= 1e-5 * want.abs().max().item()


def test_sasrec_rows_mode_steps_the_shared_table_with_head_and_encoder_gradients():
    """fused_grad='rows' with a sequence encoder that SHARES the item table (sasrec.py:107): the table receives gradient
    through the head (rows in the fused workspace) AND through the encoder's [B, L] gather (FusedEmbedding.grad_mode =
    'rows').  FusedRowOptimizer must apply both -- one update per row with the summed gradient -- so one SGD step equals
    the reference-compatible dense mode + torch.optim.SGD, and no dense [N, d] .grad is left behind."""
    from recstudio_b200 import retriever, rowopt
    N, d, Lq, B, n, lr = 5_001, 128, 64, 16, 40, 0.5
    models = {}
    for mode in ("dense", "rows"):
        m = retriever.build_sasrec_synthetic(N, d, n, max_seq_len=Lq, fused_grad=mode, device=DEV, init_std=0.1, seed=3)
        m.config["train"].update({"learner": "sgd", "learning_rate": lr, "weight_decay": 0, "scheduler": None})
        m.val_check = False                      # what fit() sets before it asks for the optimizers (recommender.py:129)
        models[mode] = m
    gen = torch.Generator(device=DEV).manual_seed(5)
    seqlen = torch.randint(1, Lq + 1, (B,), device=DEV, generator=gen)
    ids = torch.randint(1, N, (B, Lq), device=DEV, generator=gen) * (torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None])
    batch = {"in_item_id": ids, "seqlen": seqlen, "item_id": torch.randint(1, N, (B,), device=DEV, generator=gen),
             "rating": torch.ones(B, device=DEV)}
    models["rows"].load_state_dict(models["dense"].state_dict())      # biases are drawn from the CPU generator at construction
    before = models["dense"].item_encoder.weight.detach().clone()
    assert torch.equal(before, models["rows"].item_encoder.weight)
    for mode, m in models.items():
        m.train()
        opts = m._get_optimizers()
        assert (mode == "rows") == isinstance(opts[0]["optimizer"], rowopt.FusedRowOptimizer)
        torch.manual_seed(11)
        for o in opts:
            o["optimizer"].zero_grad()
        loss = m.training_step(dict(batch))
        loss.backward()
        if mode == "rows":
            assert m.item_encoder.weight.grad is None            # no dense [N, d] gradient accumulates
        for o in opts:
            o["optimizer"].step()
    da = models["dense"].item_encoder.weight.detach() - before
    db = models["rows"].item_encoder.weight.detach() - before
    touched_by_encoder = torch.zeros(N, dtype=torch.bool, device=DEV); touched_by_encoder[ids.unique()] = True
    assert da[touched_by_encoder].abs().max().item() > 0         # the encoder-side gradient is really there ...
    assert (da - db).abs().max().item() <= 2e-4 * da.abs().max().item()     # ... and the row optimizer applied it (bf16 attention noise)
    assert float(models["rows"].item_encoder.weight[0].abs().sum()) == 0.0
    # the transformer parameters were stepped by the reference optimizer in both modes
    pa = models["dense"].query_encoder.transformer_layer.layers[0].linear1.weight
    pb = models["rows"].query_encoder.transformer_layer.layers[0].linear1.weight
    assert (pa - pb).abs().max().item() <= 2e-4 * pa.abs().max().item()


def test_bert4rec_masked_training_step():
    """BERT4Rec on the fused kernels (bert4rec.py:8-58): bidirectional tcgen05 attention, 'mask' pooling (one query per
    masked position), the item table extended by the mask-token row, full-catalog SoftmaxLoss through
    rsb200_fullsoftmax_fwd_bwd.  The loss must equal torch's logsumexp loss on the SAME pooled queries (1e-5), the rows of
    the table that no input sequence touches must receive exactly the softmax head's gradient, and the encoder must
    get its gradient through d loss / d query."""
    from recstudio_b200 import retriever
    N, d, Lq, B = 5_000, 128, 64, 24
    m = retriever.build_bert4rec_synthetic(N, d, max_seq_len=Lq, device=DEV, init_std=0.1)
    assert m.item_encoder.weight.shape[0] == N + 1 and m.sampler is None and m.query_encoder.bidirectional
    gen = torch.Generator(device=DEV).manual_seed(3)
    seqlen = torch.randint(2, Lq + 1, (B,), device=DEV, generator=gen)
    real = torch.arange(Lq, device=DEV)[None, :] < seqlen[:, None]
    ids = torch.randint(1, N, (B, Lq), device=DEV, generator=gen) * real
    # _reconstruct_train_data (bert4rec.py:46-58): mask a fifth of the real positions, targets = the original tokens
    masked = (torch.rand(B, Lq, device=DEV, generator=gen) < 0.2) & real
    masked[:, 0] |= ~masked.any(1)                                   # at least one masked position per sequence
    target = ids[masked]
    ids = ids.masked_fill(masked, N)                                 # the mask token is id N (the extra table row)
    batch = {"in_item_id": ids, "seqlen": seqlen, "mask_token": masked, "item_id": target,
             "rating": torch.ones(target.numel(), device=DEV)}
    m.train()
    loss = m.training_step(batch)
    assert type(loss.grad_fn).__name__.startswith("_FullSoftmaxFn")
    loss.backward()
    with torch.no_grad():
        query = m.query_encoder(batch)                               # [number of masked positions, d]
    assert query.shape == (int(masked.sum()), d)
    w = m.item_encoder.weight.detach()
    scores = query.double() @ w[1:].double().T
    want = (torch.logsumexp(scores, -1) - (query.double() * w[target].double()).sum(-1)).mean()
    assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
    gi = m.item_encoder.weight.grad
    assert torch.isfinite(gi).all() and float(gi[0].abs().sum()) == 0.0
    p = torch.softmax(scores, -1)
    p[torch.arange(target.numel(), device=DEV), target - 1] -= 1.0
    head = torch.zeros(N + 1, d, dtype=torch.float64, device=DEV)
    head[1:] = p.T @ query.double() / target.numel()
    untouched = torch.ones(N + 1, dtype=torch.bool, device=DEV); untouched[ids.unique()] = False; untouched[0] = False
    assert (gi[untouched].double() - head[untouched]).abs().max().item() <= 1e-5 * head.abs().max().item()
    g_in = m.query_encoder.transformer_layer.layers[0].self_attn.in_proj_weight.grad
    assert g_in is not None and torch.isfinite(g_in).all() and g_in.abs().max().item() > 0
