"""Philox4x32-10 and the element<->counter mapping of PyTorch's CUDA generator.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws negatives with ``torch.randint(1, N, (B, n), device=cuda)``
(recstudio/ann/sampler.py:102-104) and ``torch.rand(B, n, device=cuda)``
(sampler.py:246).  The arithmetic lives in ATen, not in RecStudio:

* ``ATen/native/cuda/DistributionTemplates.h:50-62``  calc_execution_policy
* ``ATen/native/cuda/DistributionTemplates.h:65-89``  grid-stride kernel
* ``ATen/native/cuda/DistributionTemplates.h:318-346`` random_from_to (u32 path
  while range < 2**28, else the two-u32 -> u64 path)
* ``ATen/native/cuda/DistributionTemplates.h:485-506`` uniform_ (0,1] -> [0,1)
* ``ATen/core/TransformationHelper.h`` uniform_int_from_to: ``V % range + base``
* cuRAND ``curand_init(seed, subsequence, offset)`` / ``curand4`` /
  ``curand_uniform4`` for Philox4_32_10.

numpy only; vectorised so that 8.4 M draws take about a second.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)

BLOCK = 256          # block_size_bound, DistributionTemplates.h:34
UNROLL = 4           # sizeof(uint4)/sizeof(uint32_t), DistributionTemplates.h:117


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Ten rounds of Philox-4x32.  All arguments broadcastable uint32 arrays.

    Returns the four uint32 output words (x, y, z, w) in cuRAND's order.
    """
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def torch_cuda_grid(numel: int, sm_count: int, max_threads_per_sm: int) -> int:
    """grid.x of calc_execution_policy (DistributionTemplates.h:50-58)."""
    blocks = (numel + BLOCK - 1) // BLOCK
    return min(sm_count * (max_threads_per_sm // BLOCK), blocks)


def torch_cuda_counter_offset(numel: int, sm_count: int, max_threads_per_sm: int) -> int:
    """How far one randint/rand call advances the generator offset (:60)."""
    if numel == 0:
        return 0
    grid = torch_cuda_grid(numel, sm_count, max_threads_per_sm)
    return ((numel - 1) // (BLOCK * grid * UNROLL) + 1) * 4


def torch_cuda_raw_u32(seed: int, offset: int, numel: int, sm_count: int,
                       max_threads_per_sm: int) -> np.ndarray:
    """The uint32 that element ``li`` of a CUDA distribution kernel consumes.

    Thread ``idx`` runs ``curand_init(seed, idx, offset)``: key = seed,
    counter = (offset/4 as u64, idx as u64).  In round ``r`` of the grid-stride
    loop it calls ``curand4`` once (counter low half += r) and element
    ``li = idx + T*(4r + ii)`` takes output word ``ii`` (T = 256*grid).
    """
    assert offset % 4 == 0, "torch keeps the philox offset a multiple of 4"
    grid = torch_cuda_grid(numel, sm_count, max_threads_per_sm)
    T = BLOCK * grid
    li = np.arange(numel, dtype=np.int64)
    r = li // (UNROLL * T)
    rem = li % (UNROLL * T)
    ii = rem // T
    idx = rem % T
    ctr_lo = (offset // 4) + r                      # u64 in (c0, c1)
    c0 = (ctr_lo & 0xFFFFFFFF).astype(np.uint32)
    c1 = (ctr_lo >> 32).astype(np.uint32)
    c2 = (idx & 0xFFFFFFFF).astype(np.uint32)
    c3 = (idx >> 32).astype(np.uint32)
    x, y, z, w = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.where(ii == 0, x, np.where(ii == 1, y, np.where(ii == 2, z, w)))
    return out.astype(np.uint32)


def torch_cuda_randint(seed: int, offset: int, low: int, high: int, numel: int,
                       sm_count: int, max_threads_per_sm: int) -> np.ndarray:
    """``torch.randint(low, high, (numel,), device='cuda')`` as int64.

    32-bit path only (range < 2**28), DistributionTemplates.h:335-345 with
    uniform_int_from_to = ``(uint32 % range) + base``.
    """
    rng = high - low
    assert 0 < rng < (1 << 28), "64-bit draw path (range >= 2**28) not restated"
    u = torch_cuda_raw_u32(seed, offset, numel, sm_count, max_threads_per_sm)
    return (u.astype(np.int64) % rng) + low


def curand_uniform_from_u32(u: np.ndarray) -> np.ndarray:
    """cuRAND ``_curand_uniform``: ``x * 2^-32 + 2^-33`` in fp32, range (0, 1]."""
    two_m32 = np.float32(2.3283064365386963e-10)
    xf = u.astype(np.float32)                       # cvt.rn.f32.u32
    return xf * two_m32 + np.float32(two_m32 / np.float32(2.0))


def torch_cuda_rand(seed: int, offset: int, numel: int, sm_count: int,
                    max_threads_per_sm: int) -> np.ndarray:
    """``torch.rand(numel, device='cuda')`` fp32: (0,1] folded to [0,1) by
    mapping 1.0 -> 0.0 (DistributionTemplates.h:495-502)."""
    u = torch_cuda_raw_u32(seed, offset, numel, sm_count, max_threads_per_sm)
    v = curand_uniform_from_u32(u)
    return np.where(v == np.float32(1.0), np.float32(0.0), v).astype(np.float32)
