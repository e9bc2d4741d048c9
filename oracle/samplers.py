"""CPU restatement of the reference's static negative samplers (S1, S2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Integer work (the draw) is numpy; the one floating-point construction (the
popularity table) uses the same torch CPU ops the reference's constructor
runs, because torch's fp32 ``sum`` / ``cumsum`` (double accumulator on CPU)
are not reproducible bit-for-bit with numpy reductions.
"""
from __future__ import annotations

import numpy as np
import torch

from . import philox


# --------------------------------------------------------------------------- S1
def uniform_sampler(seed: int, offset: int, num_items: int, num_queries: int,
                    num_neg: int, pos_items, sm_count: int, max_threads_per_sm: int):
    """``UniformSampler.forward`` on a CUDA device (recstudio/ann/sampler.py:86-114).

    ``num_items`` is the table size INCLUDING padding row 0 (what the reference
    passes to ``UniformSampler(num_items)``; ``Sampler.__init__`` stores
    ``num_items - 1``, sampler.py:51).  Negatives are
    ``torch.randint(1, self.num_items + 1, (num_queries, num_neg))`` = ids in
    ``[1, num_items - 1]`` (sampler.py:102-104).  ``log_pos_prob`` and
    ``log_neg_prob`` are ``zeros_like`` of int64 id tensors -> int64 zeros
    (sampler.py:113-114).

    Returns (log_pos_prob | None, neg_id int64 [num_queries, num_neg],
    log_neg_prob int64 zeros, new_offset).
    """
    numel = num_queries * num_neg
    ids = philox.torch_cuda_randint(seed, offset, 1, num_items, numel,
                                    sm_count, max_threads_per_sm)
    neg = ids.reshape(num_queries, num_neg)
    new_offset = offset + philox.torch_cuda_counter_offset(numel, sm_count, max_threads_per_sm)
    log_neg = np.zeros_like(neg)
    log_pos = None if pos_items is None else np.zeros_like(np.asarray(pos_items))
    return log_pos, neg, log_neg, new_offset


# --------------------------------------------------------------------------- S2
def popular_tables(pop_count, mode: int = 0):
    """``PopularSamplerModel.__init__`` (recstudio/ann/sampler.py:225-241).

    Returns (pop_prob f32 [N], table f32 [N]) as numpy arrays, INCLUDING the
    two quirks of the reference:
      (a) ``pop_count[0] = 1`` after the transform, so padding id 0 has mass;
      (b) ``pop_prob[-1] = 1.0`` is written AFTER the cumsum, so the last
          item's log-probability is log(1) = 0 while ``table`` is unaffected.
    """
    with torch.no_grad():
        pc = torch.tensor(np.asarray(pop_count), dtype=torch.float)   # :228
        if mode == 0:
            pc = torch.log(pc + 1)                                    # :230
        elif mode == 1:
            pc = torch.log(pc + 1) + 1e-6                             # :232
        elif mode == 2:
            pc = pc ** 0.75                                           # :234
        pc[0] = 1                                                     # :237
        pop_prob = pc / pc.sum()                                      # :239
        table = torch.cumsum(pop_prob, dim=0)                         # :240
        pop_prob[-1] = 1.0                                            # :241
    return pop_prob.numpy().copy(), table.numpy().copy()


def popular_draw(table: np.ndarray, pop_prob: np.ndarray, seeds: np.ndarray):
    """``searchsorted(table, seeds)`` (left bisect: first i with table[i] >= u)
    and ``log(pop_prob[idx])`` (sampler.py:247,257-258).  ``seeds`` fp32 in [0,1).

    If a seed exceeds ``table[-1]`` (fp32 cumsum may end below 1.0) the
    reference returns ``N`` and then indexes out of range; like the product we
    clamp to ``N - 1`` and the tests avoid/flag that case explicitly.
    """
    idx = np.searchsorted(table, seeds.astype(np.float32), side="left").astype(np.int64)
    idx = np.minimum(idx, table.shape[0] - 1)
    logq = torch.log(torch.from_numpy(pop_prob)[torch.from_numpy(idx)]).numpy()
    return idx, logq


def popular_sampler(seed: int, offset: int, table: np.ndarray, pop_prob: np.ndarray,
                    num_queries: int, num_neg: int, pos_items,
                    sm_count: int, max_threads_per_sm: int):
    """``PopularSamplerModel.forward`` on CUDA (sampler.py:243-255): seeds are
    ``torch.rand(num_queries, num_neg, device=cuda)``."""
    numel = num_queries * num_neg
    seeds = philox.torch_cuda_rand(seed, offset, numel, sm_count, max_threads_per_sm)
    idx, logq = popular_draw(table, pop_prob, seeds)
    new_offset = offset + philox.torch_cuda_counter_offset(numel, sm_count, max_threads_per_sm)
    log_pos = None
    if pos_items is not None:
        pos = torch.as_tensor(np.asarray(pos_items))
        log_pos = torch.log(torch.from_numpy(pop_prob)[pos]).numpy()
    return log_pos, idx.reshape(num_queries, num_neg), logq.reshape(num_queries, num_neg), new_offset


# --------------------------------------------------------------------------- S3
def _upper_bound(row: np.ndarray, val: int) -> int:
    """ATen ``upper_bound`` (aten/src/ATen/native/cuda/Bucketization.cu; same loop on CPU):
    the exact bisection, so NON-monotone rows (duplicate history items) give the reference's answer."""
    start, end = 0, row.shape[0]
    while start < end:
        mid = start + ((end - start) >> 1)
        if not (row[mid] > val):
            start = mid + 1
        else:
            end = mid
    return start


def masked_uniform_from_seeds(num_items: int, user_hist: np.ndarray, seeds: np.ndarray) -> np.ndarray:
    """``uniform_sample_masked_hist`` (recstudio/ann/sampler.py:117-147) given the
    ``torch.rand(num_user, n_q * num_neg)`` seeds.  ``num_items`` is ``Sampler.num_items``
    (real items, padding excluded).  Integer work in numpy; the one fp32 product
    ``rand * (num_items - count)`` is an fp32 multiply exactly as torch promotes it."""
    user_hist = np.asarray(user_hist, dtype=np.int64)
    seeds = np.asarray(seeds, dtype=np.float32)
    num_user, hist_len = user_hist.shape
    nz = np.count_nonzero(user_hist, axis=-1)                                        # :131
    span = (num_items - nz).astype(np.float32)[:, None]                              # int64 -> fp32 promotion
    neg = np.floor(seeds * span).astype(np.int64) + 1                                # :133
    sorted_hist = np.sort(user_hist, axis=-1)                                        # :134
    offset = np.arange(hist_len, dtype=np.int64)[None, :] - (hist_len - nz)[:, None]  # :135-136
    offset[offset < 0] = 0                                                           # :137
    adj = sorted_hist - offset                                                       # :138
    out = np.empty_like(neg)
    for b in range(num_user):                                                        # :139-141
        pad = hist_len - nz[b]
        for j in range(neg.shape[1]):
            out[b, j] = neg[b, j] + _upper_bound(adj[b], int(neg[b, j])) - pad
    return out


def masked_uniform_sampler(seed: int, offset: int, num_items_with_pad: int, user_hist: np.ndarray, per_user: int,
                           sm_count: int, max_threads_per_sm: int):
    """``MaskedUniformSampler.forward`` on CUDA (sampler.py:187-214): seeds are
    ``torch.rand(num_user, per_user, device=cuda)``.  Returns (neg [num_user, per_user], new_offset)."""
    num_user = user_hist.shape[0]
    numel = num_user * per_user
    seeds = philox.torch_cuda_rand(seed, offset, numel, sm_count, max_threads_per_sm).reshape(num_user, per_user)
    neg = masked_uniform_from_seeds(num_items_with_pad - 1, user_hist, seeds)
    return neg, offset + philox.torch_cuda_counter_offset(numel, sm_count, max_threads_per_sm)
