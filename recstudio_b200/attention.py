"""A1: the SASRec / BERT4Rec query encoder with the attention core on tcgen05 tensor cores.

``FusedSASRecQueryEncoder`` mirrors ``recstudio.model.seq.sasrec.SASRecQueryEncoder``
(sasrec.py:8-67): same constructor arguments, same sub-modules and therefore the same
``state_dict`` keys (``position_emb``, ``transformer_layer.layers.N.self_attn.in_proj_weight`` ...),
same forward contract (batch dict with ``in_<fiid>`` and ``seqlen`` -> [B, D] after 'last' pooling).

Inside, each post-norm ``TransformerEncoderLayer`` is evaluated as
    qkv  = x W_in^T + b                       plain library GEMM (cuBLAS via torch)
    attn = softmax(QK^T/sqrt(dh) + masks) V   rsb200_attn_fwd / rsb200_attn_bwd (bf16, tcgen05/TMEM)
    x    = LN(x + attn W_o^T + b); x = LN(x + FFN(x))   cuBLAS + torch elementwise
with the causal mask ``triu(ones(L, L), 1)`` (unless bidirectional) and the key-padding mask
``hist == 0`` of the reference (sasrec.py:44-53) applied inside the kernel.

Precision: bf16 operands / fp32 accumulation -> ~1e-2 relative to the reference's fp32;
this is the one place on the path where tensor cores are used.  Attention-probability dropout is
not implemented in the kernel: when the module is in training mode with ``dropout > 0`` the
reference's own ``nn.TransformerEncoder`` path runs instead (exact semantics are kept).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, hist, heads: int, causal: bool):
        if not q.is_cuda:
            raise _lib.Rsb200Error("fused attention needs CUDA tensors (no CPU fallback)")
        q, k, v = (t.contiguous().float() for t in (q, k, v))
        B, L, d = q.shape
        dh = d // heads
        hist_c = None if hist is None else hist.to(torch.int64).contiguous()
        out = torch.empty_like(q)
        lse = torch.empty(B, heads, L, dtype=torch.float32, device=q.device)
        err = torch.zeros(1, dtype=torch.int32, device=q.device)
        with torch.cuda.device(q.device):
            check(lib().rsb200_attn_fwd(ptr(q), ptr(k), ptr(v), ptr(hist_c), B, L, heads, dh, int(bool(causal)), ptr(out),
                                        ptr(lse), ptr(err), stream_ptr()), "attn_fwd")
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.hist, ctx.heads, ctx.causal, ctx.err = hist_c, heads, causal, err
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        q, k, v, out, lse = ctx.saved_tensors
        B, L, d = q.shape
        g = g.contiguous().float()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        with torch.cuda.device(q.device):
            check(lib().rsb200_attn_bwd(ptr(q), ptr(k), ptr(v), ptr(out), ptr(g), ptr(lse), ptr(ctx.hist), B, L, ctx.heads,
                                        d // ctx.heads, int(bool(ctx.causal)), ptr(dq), ptr(dk), ptr(dv), ptr(ctx.err),
                                        stream_ptr()), "attn_bwd")
        return dq, dk, dv, None, None, None


def fused_attention(q: Tensor, k: Tensor, v: Tensor, hist, heads: int, causal: bool) -> Tensor:
    """q, k, v: [B, L, heads * 64] fp32; hist: [B, L] item ids (0 = padding key) or None."""
    return _AttnFn.apply(q, k, v, hist, heads, causal)


def attention_supported(embed_dim: int, n_head: int, seq_len: int) -> bool:
    return embed_dim % n_head == 0 and embed_dim // n_head == 64 and 1 <= seq_len <= 256


class FusedSASRecQueryEncoder(torch.nn.Module):
    """Drop-in for ``SASRecQueryEncoder`` (recstudio/model/seq/sasrec.py:8-67)."""

    def __init__(self, fiid, embed_dim, max_seq_len, n_head, hidden_size, dropout, activation, layer_norm_eps, n_layer,
                 item_encoder, bidirectional=False, training_pooling_type="last", eval_pooling_type="last") -> None:
        super().__init__()
        self.fiid = fiid
        self.item_encoder = item_encoder
        self.bidirectional = bidirectional
        self.training_pooling_type = training_pooling_type
        self.eval_pooling_type = eval_pooling_type
        self.n_head = n_head
        self.position_emb = torch.nn.Embedding(max_seq_len, embed_dim)
        layer = torch.nn.TransformerEncoderLayer(d_model=embed_dim, nhead=n_head, dim_feedforward=hidden_size,
                                                 dropout=dropout, activation=activation, layer_norm_eps=layer_norm_eps,
                                                 batch_first=True, norm_first=False)
        self.transformer_layer = torch.nn.TransformerEncoder(encoder_layer=layer, num_layers=n_layer)
        self.dropout = torch.nn.Dropout(p=dropout)
        self.p_drop = float(dropout)

    @property
    def pools_to_2d(self) -> bool:
        """whether forward() yields a [rows, d] matrix in the current mode (FusedRetrieverMixin decides from this, before the
        encoder runs, whether the fused full-softmax step applies)"""
        return (self.training_pooling_type if self.training else self.eval_pooling_type) in ("mask", "last", "sum", "mean", "concat")

    # one post-norm encoder layer (torch.nn.TransformerEncoderLayer.forward, norm_first=False) with the fused core
    def _layer(self, lyr, x: Tensor, hist: Tensor) -> Tensor:
        at = lyr.self_attn
        qkv = F.linear(x, at.in_proj_weight, at.in_proj_bias)
        q, k, v = qkv.chunk(3, dim=-1)
        a = fused_attention(q, k, v, hist, self.n_head, not self.bidirectional)
        x = lyr.norm1(x + lyr.dropout1(F.linear(a, at.out_proj.weight, at.out_proj.bias)))
        ff = lyr.linear2(lyr.dropout(lyr.activation(lyr.linear1(x))))
        return lyr.norm2(x + lyr.dropout2(ff))

    def _use_fused(self, L: int, device) -> bool:
        d = self.position_emb.embedding_dim
        if device.type != "cuda" or not attention_supported(d, self.n_head, L):
            return False
        return not (self.training and self.p_drop > 0.0)          # attention-prob dropout: reference path

    def encode(self, user_hist: Tensor) -> Tensor:
        """[B, L] item ids -> [B, L, D] (sasrec.py:38-53)."""
        L = user_hist.size(1)
        positions = torch.arange(L, dtype=torch.long, device=user_hist.device).unsqueeze(0).expand_as(user_hist)
        x = self.dropout(self.item_encoder(user_hist) + self.position_emb(positions))
        if self._use_fused(L, user_hist.device):
            for lyr in self.transformer_layer.layers:
                x = self._layer(lyr, x, user_hist)
            return x
        mask4padding = user_hist == 0
        if not self.bidirectional:
            attention_mask = torch.triu(torch.ones((L, L), dtype=torch.bool, device=user_hist.device), 1)
        else:
            attention_mask = torch.zeros((L, L), dtype=torch.bool, device=user_hist.device)
        return self.transformer_layer(src=x, mask=attention_mask, src_key_padding_mask=mask4padding)

    def forward(self, batch, need_pooling=True):
        user_hist = batch["in_" + self.fiid]
        out = self.encode(user_hist)
        if not need_pooling:
            return out
        ptype = self.training_pooling_type if self.training else self.eval_pooling_type
        return pool_sequence(out, batch["seqlen"], ptype, batch.get("mask_token") if ptype == "mask" else None)


POOLING_TYPES = ("origin", "mask", "concat", "sum", "mean", "max", "last")


def pool_sequence(seq_out: Tensor, seqlen: Tensor, pooling_type: str = "last", mask_token: Optional[Tensor] = None):
    """``SeqPoolingLayer.forward`` (recstudio/model/module/layers.py:256-314, keepdim=False, 3-D input) on the encoder
    output [B, L, D]: 'last' = position seqlen-1 (SASRec), 'mask' = the rows selected by the boolean [B, L] ``mask_token``
    (BERT4Rec training: one row per masked position, bert4rec.py:46-58), 'origin' / 'concat' / 'sum' / 'mean' / 'max' =
    the reference's reductions over the first ``seqlen`` positions with padded positions zeroed first."""
    if pooling_type not in POOLING_TYPES:
        raise ValueError("pooling_type can only be one of %s but %s is given." % (list(POOLING_TYPES), pooling_type))
    B, L, D = seq_out.shape
    if pooling_type == "mask":
        assert mask_token is not None, "mask_token can be None when pooling_type is 'mask'."
        return seq_out[mask_token]
    if pooling_type == "last":
        idx = (seqlen - 1).view(-1, 1, 1).expand(-1, -1, D)
        return seq_out.gather(dim=1, index=idx).squeeze(1)
    valid = torch.arange(L, device=seq_out.device).view(1, L, 1) < seqlen.view(-1, 1, 1)
    kept = seq_out.masked_fill(~valid, 0.0)
    if pooling_type == "origin":
        return kept
    if pooling_type == "concat":
        return kept.reshape(B, -1)
    if pooling_type == "max":
        return kept.max(dim=1)                      # the reference returns torch's (values, indices) pair here
    total = kept.sum(dim=1)
    return total if pooling_type == "sum" else total / seqlen.view(-1, 1)
