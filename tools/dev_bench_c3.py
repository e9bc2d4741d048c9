"""Developer timing of BASELINE config 3: SASRec (L = 200, 1M items x 128, 2 layers x 2 heads) with the
tcgen05 attention core + fused sampled-softmax head, against the reference's nn.TransformerEncoder path
with the same fused head, and the attention core alone against torch SDPA (bf16 / fp32)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recstudio_b200 import attention, retriever  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device("cuda:0")
    N, d, L, B, n = 1_000_001, 128, 200, 1024, 1024
    out = {}
    m = retriever.build_sasrec_synthetic(N, d, n, max_seq_len=L, fused_grad="sparse", device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    seqlen = torch.randint(1, L + 1, (B,), device=dev, generator=gen)
    ids = torch.randint(1, N, (B, L), device=dev, generator=gen) * (torch.arange(L, device=dev)[None, :] < seqlen[:, None])
    batch = {"in_item_id": ids, "seqlen": seqlen, "item_id": torch.randint(1, N, (B,), device=dev, generator=gen),
             "rating": torch.ones(B, device=dev)}
    m.train()

    def step():
        m.zero_grad(set_to_none=True)
        m.training_step(batch).backward()
    out["sasrec_step_fused_ms"] = timeit(step)
    m.query_encoder._use_fused = lambda Lq, device: False
    out["sasrec_step_reference_encoder_ms"] = timeit(step)
    out["sequences_per_s_fused"] = B / out["sasrec_step_fused_ms"] * 1e3
    # attention core alone
    q, k, v = (torch.randn(B, L, d, device=dev, requires_grad=True) for _ in range(3))
    hist = ids

    def core():
        attention.fused_attention(q, k, v, hist, 2, True).sum().backward()

    def sdpa(dtype):
        def f():
            qq, kk, vv = (t.reshape(B, L, 2, 64).transpose(1, 2).to(dtype) for t in (q, k, v))
            mask = torch.triu(torch.ones(L, L, dtype=torch.bool, device=dev), 1)[None, None] | (hist == 0)[:, None, None, :]
            torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, attn_mask=~mask).sum().backward()
        return f
    out["attn_core_fwd_bwd_ms"] = timeit(core)
    out["torch_sdpa_fp32_ms"] = timeit(sdpa(torch.float32))
    out["torch_sdpa_bf16_ms"] = timeit(sdpa(torch.bfloat16))
    flops = B * 2 * (4 * L * L * 64) * 3.5          # fwd 2 GEMMs + bwd 5 GEMMs, per (b, h)
    out["attn_core_tflops"] = flops / out["attn_core_fwd_bwd_ms"] / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
