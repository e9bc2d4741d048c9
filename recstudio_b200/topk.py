"""T1: full-catalog top-k with history mask (rsb200_topk_full).

Replaces ``BaseRetriever.topk`` without ANN index
(recstudio/model/basemodel/baseretriever.py:374-397): scores every item row
1..num_items-1, masks the ids in the user's history, returns the k best as
(score [Be, k] f32 descending, ids [Be, k] i64 1-based).  Ties: lower id first.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

_ws_cache = {}


def topk_full(query: torch.Tensor, w_item: torch.Tensor, k: int, user_hist: Optional[torch.Tensor] = None,
              score_kind: int = _lib.SCORE_IP):
    """query [Be, d] f32 cuda, w_item [num_items, d] (row 0 = padding), user_hist [Be, H] i64 or None."""
    _lib.require_cuda()
    if not query.is_cuda or not w_item.is_cuda:
        raise _lib.Rsb200Error("topk_full needs CUDA tensors (no CPU fallback)")
    q = query.detach().to(torch.float32).contiguous()
    w = w_item.detach()
    if not w.is_contiguous() or w.dtype != torch.float32:
        raise _lib.Rsb200Error("the item table must be a contiguous float32 tensor")
    Be, d = q.shape
    num_items = w.shape[0]
    H = 0
    hist = None
    if user_hist is not None and user_hist.numel() > 0:
        hist = user_hist.to(device=q.device, dtype=torch.int64).contiguous()
        H = hist.shape[1]
    nbytes = int(lib().rsb200_topk_workspace_bytes(Be, num_items, k, H))
    if nbytes == 0:
        raise _lib.Rsb200Error("bad top-k shape (Be=%d, num_items=%d, k=%d, H=%d)" % (Be, num_items, k, H))
    key = (q.device, nbytes)
    ws = _ws_cache.get(key)
    if ws is None:
        _ws_cache.clear()
        ws = _ws_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    score = torch.empty(Be, k, dtype=torch.float32, device=q.device)
    ids = torch.empty(Be, k, dtype=torch.int64, device=q.device)
    with torch.cuda.device(q.device):
        check(lib().rsb200_topk_full(int(score_kind), ptr(q), ptr(w), num_items, d, Be, int(k), ptr(hist), H,
                                     ptr(score), ptr(ids), ptr(ws), nbytes, stream_ptr()), "topk_full")
    return score, ids
