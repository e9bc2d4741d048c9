"""S1 / S2 draws on the torch CUDA generator's Philox stream.

``uniform_draw`` returns exactly what ``torch.randint(1, num_items, (Q, n), device=cuda)``
would return for the generator's current (seed, offset) and advances the generator by the
same amount (recstudio/ann/sampler.py:102-104); ``popular_draw`` does the same for
``searchsorted(table, torch.rand(Q, n, device=cuda))`` (sampler.py:246-247).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _generator(device: torch.device, generator: Optional[torch.Generator]):
    if generator is not None:
        return generator
    torch.cuda.init()           # default_generators is empty until CUDA is initialised (lazy init)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch.cuda.default_generators[idx]


def _policy(device: torch.device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    sm, mt, _, _ = _lib.device_info(idx)
    return sm, mt


def counter_offset(numel: int, device: torch.device) -> int:
    sm, mt = _policy(device)
    return int(lib().rsb200_philox_counter_offset(int(numel), sm, mt))


def uniform_draw(num_items: int, num_queries: int, num_neg: int, device, want_i64: bool = True,
                 want_i32: bool = False, generator: Optional[torch.Generator] = None):
    """ids uniform on [1, num_items-1]; returns (neg_i64 | None, neg_i32 | None)."""
    _lib.require_cuda()
    device = torch.device(device)
    gen = _generator(device, generator)
    seed, offset = gen.initial_seed(), gen.get_offset()
    sm, mt = _policy(device)
    o64 = torch.empty(num_queries, num_neg, dtype=torch.int64, device=device) if want_i64 else None
    o32 = torch.empty(num_queries, num_neg, dtype=torch.int32, device=device) if want_i32 else None
    with torch.cuda.device(device):
        check(lib().rsb200_sample_uniform(seed, offset, int(num_items), int(num_queries), int(num_neg), sm, mt,
                                          ptr(o64), ptr(o32), stream_ptr()), "sample_uniform")
    gen.set_offset(offset + counter_offset(num_queries * num_neg, device))
    return o64, o32


def masked_uniform_draw(num_items: int, user_hist: torch.Tensor, per_user: int, want_i64: bool = True,
                        want_i32: bool = False, generator: Optional[torch.Generator] = None):
    """``uniform_sample_masked_hist`` (recstudio/ann/sampler.py:117-147) on the CUDA generator's stream:
    ``per_user`` ids per history row, uniform over ``[1, num_items-1]`` minus the row's non-zero history.
    ``num_items`` counts the padding row.  Returns (neg_i64 | None, neg_i32 | None) of shape [rows, per_user]."""
    _lib.require_cuda()
    if not user_hist.is_cuda or user_hist.dim() != 2 or user_hist.dtype != torch.int64:
        raise _lib.Rsb200Error("user_hist must be a CUDA int64 [num_users, hist_len] tensor (no CPU fallback)")
    device = user_hist.device
    user_hist = user_hist.contiguous()
    rows, hlen = user_hist.shape
    gen = _generator(device, generator)
    seed, offset = gen.initial_seed(), gen.get_offset()
    sm, mt = _policy(device)
    o64 = torch.empty(rows, per_user, dtype=torch.int64, device=device) if want_i64 else None
    o32 = torch.empty(rows, per_user, dtype=torch.int32, device=device) if want_i32 else None
    adj = torch.empty(rows * hlen, dtype=torch.int64, device=device)
    cnt = torch.empty(rows, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().rsb200_sample_uniform_masked(seed, offset, int(num_items), ptr(user_hist), rows, hlen, int(per_user),
                                                 sm, mt, ptr(adj), ptr(cnt), ptr(o64), ptr(o32), stream_ptr()),
              "sample_uniform_masked")
    gen.set_offset(offset + counter_offset(rows * per_user, device))
    return o64, o32


def build_guide(table: torch.Tensor, guide_bits: Optional[int] = None):
    """guide[k] = searchsorted(table, k / 2^bits): O(1)-expected replacement of the bisection."""
    n = table.numel()
    if guide_bits is None:
        # about 8 table entries per guide bucket: the draw kernel then needs no bisection step at all (it counts the
        # <= 8 entries of the bracket with independent loads), i.e. two dependent random accesses per draw instead
        # of log2(N).  N = 10 M: 2^21 entries (8 MB, guide + table + pop_prob stay inside the 126 MB L2);
        # N = 100 M: 2^24 entries (64 MB; nothing fits L2 at that size anyway).
        guide_bits = max(1, min(24, ((max(n, 2) - 1) // 8).bit_length()))
    guide = torch.empty((1 << guide_bits) + 1, dtype=torch.int32, device=table.device)
    with torch.cuda.device(table.device):
        check(lib().rsb200_popular_build_guide(ptr(table), n, guide_bits, ptr(guide), stream_ptr()), "build_guide")
    return guide, guide_bits


def popular_draw(table: torch.Tensor, pop_prob: torch.Tensor, num_queries: int, num_neg: int,
                 guide: Optional[torch.Tensor] = None, guide_bits: int = 0, want_i64: bool = True,
                 want_i32: bool = False, want_logq: bool = True, generator: Optional[torch.Generator] = None):
    """returns (neg_i64 | None, neg_i32 | None, logq | None)."""
    _lib.require_cuda()
    device = table.device
    gen = _generator(device, generator)
    seed, offset = gen.initial_seed(), gen.get_offset()
    sm, mt = _policy(device)
    o64 = torch.empty(num_queries, num_neg, dtype=torch.int64, device=device) if want_i64 else None
    o32 = torch.empty(num_queries, num_neg, dtype=torch.int32, device=device) if want_i32 else None
    lq = torch.empty(num_queries, num_neg, dtype=torch.float32, device=device) if want_logq else None
    with torch.cuda.device(device):
        check(lib().rsb200_sample_popular(seed, offset, ptr(table), ptr(pop_prob), table.numel(), int(num_queries),
                                          int(num_neg), sm, mt, ptr(guide), int(guide_bits), ptr(o64), ptr(o32),
                                          ptr(lq), stream_ptr()), "sample_popular")
    gen.set_offset(offset + counter_offset(num_queries * num_neg, device))
    return o64, o32, lq


def popular_logq(pop_prob: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """log(pop_prob[ids]) (PopularSamplerModel.compute_item_p, sampler.py:257-258)."""
    ids = ids.contiguous()
    out = torch.empty(ids.shape, dtype=torch.float32, device=ids.device)
    with torch.cuda.device(ids.device):
        check(lib().rsb200_popular_logq(ptr(pop_prob), pop_prob.numel(), ptr(ids), ids.numel(), ptr(out), stream_ptr()),
              "popular_logq")
    return out
