"""FusedSASRecQueryEncoder against the REFERENCE's own SASRecQueryEncoder (sasrec.py:8-67), through a golden that
tests/golden/make_golden.py::golden_sasrec_encoder produced by running the unmodified reference module on the CPU:
same state_dict keys, same pooled outputs and parameter gradients.
  * CPU (not gpu): the encoder's non-fused branch (what it runs where the tcgen05 core does not apply) is the reference to 1e-5
    -- so the GPU tests that compare the fused core with that branch (test_gpu_attention.py) are anchored to the reference;
  * GPU: the fused tcgen05 path directly against the golden, at the bf16 tolerance of the core (operands are bf16)."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def _build(device, item_encoder, bidirectional):
    from recstudio_b200 import attention
    g = load_golden("sasrec_encoder")
    N, d = g["w:item_encoder.weight"].shape
    L = g["w:position_emb.weight"].shape[0]
    enc = attention.FusedSASRecQueryEncoder("item_id", d, L, 2, 128, 0.0, "gelu", 1e-12, 2, item_encoder(N, d),
                                            bidirectional=bidirectional)
    state = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w:")}
    assert set(state) == set(enc.state_dict().keys())            # the reference's checkpoint loads as is
    enc.load_state_dict(state)
    enc = enc.to(device)
    batch = {"in_item_id": torch.from_numpy(g["ids"]).to(device), "seqlen": torch.from_numpy(g["seqlen"]).to(device)}
    return g, enc, batch


def _check(g, enc, batch, tag, out_tol, grad_tol):
    enc.train()
    enc.zero_grad()
    out = enc(batch)
    gout = torch.from_numpy(g["g"]).to(out.device)
    (out * gout).sum().backward()
    ref = torch.from_numpy(g["out_" + tag]).to(out.device)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= out_tol * ref.abs().max().item()
    params = dict(enc.named_parameters())
    for name in g["watch"].tolist():
        r = torch.from_numpy(g["grad_%s:%s" % (tag, name)]).to(out.device)
        got = params[name].grad
        got = got.to_dense() if got.is_sparse else got
        assert (got - r).abs().max().item() <= grad_tol * r.abs().max().item(), name


@pytest.mark.parametrize("tag", ["causal", "bidir"])
def test_non_fused_branch_is_the_reference_encoder(tag):
    g, enc, batch = _build(torch.device("cpu"), lambda N, d: torch.nn.Embedding(N, d, padding_idx=0), tag == "bidir")
    assert not enc._use_fused(batch["in_item_id"].size(1), batch["in_item_id"].device)
    _check(g, enc, batch, tag, 1e-5, 1e-4)
    if tag == "causal":
        enc.eval()
        with torch.no_grad():
            out = enc(batch)
        np.testing.assert_allclose(out.numpy(), g["out_eval_causal"], rtol=0, atol=1e-5 * float(np.abs(g["out_eval_causal"]).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["causal", "bidir"])
def test_fused_core_against_the_reference_encoder_golden(tag):
    from recstudio_b200 import plugins
    dev = torch.device("cuda", 0)
    g, enc, batch = _build(dev, lambda N, d: plugins.FusedEmbedding(N, d), tag == "bidir")
    assert enc._use_fused(batch["in_item_id"].size(1), batch["in_item_id"].device)
    _check(g, enc, batch, tag, 5e-2, 1e-1)            # bf16 operands in both layers' cores; the fp32 branch above holds 1e-5
