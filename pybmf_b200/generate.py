"""Synthetic Boolean matrices generated ON the device.

The reference builds its inputs on the host: `matmul(U, V.T, boolean=True)` of two random factors
(PyBMF/generators/BaseGenerator.py:202-221) followed by `add_noise` (PyBMF/utils/generator_utils.py:30-49).  At the
Netflix-shaped config that is a 1e8-nnz scipy matrix per rank (~10 s and 3 GB of host memory each).  Here the same recipe
runs on bit rows in HBM with a counter-based generator, so a rank generates only ITS rows of the one logical matrix and the
result does not depend on how many ranks there are.  It is not bit-identical to numpy's Mersenne-twister draws; tests pin
the densities, the shard independence and a whole fit against the CPU restatement run on the downloaded bits.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from . import _native, device
from .engine import ShardPlan, dist_ctx


class DeviceBits:
    """Rows [r0, r1) of a logical m x n Boolean matrix as bit rows on the current device."""

    def __init__(self, bits, m, n, r0, r1):
        self.bits, self.m, self.n, self.r0, self.r1 = bits, int(m), int(n), int(r0), int(r1)
        self.shape = (self.m, self.n)

    def to_csr(self) -> sp.csr_matrix:
        """This rank's rows as a host csr int64 (for cross-checks; moves m_loc * n / 8 bytes)."""
        from .utils import _bits_to_csr
        return _bits_to_csr(self.bits, self.r1 - self.r0, self.n)


def _rows_of(m, rank=None, world=None):
    if rank is None or world is None:
        rank, world = dist_ctx()
    return ShardPlan(m, world).rows(rank)


def random_bits(m, n, p, seed, stream_id=0, rank=None, world=None) -> DeviceBits:
    """Bernoulli(p) m x n matrix (this rank's rows)."""
    _native.require_gpu()
    r0, r1 = _rows_of(m, rank, world)
    words = device.words_for(n)
    bits = device.zeros((max(r1 - r0, 1), words), torch.int64)
    if r1 > r0:
        _native.call("bmf_random_bits", bits, r1 - r0, n, words, r0, int(seed), int(stream_id), float(p))
    return DeviceBits(bits, m, n, r0, r1)


def add_noise_bits(X: DeviceBits, noise=(0.0, 0.0), seed=0) -> DeviceBits:
    """`add_noise` (PyBMF/utils/generator_utils.py:30-49) in place: ones dropped with probability noise[0], then entries set
    with probability noise[1]."""
    if X.r1 > X.r0:
        _native.call("bmf_noise_bits", X.bits, X.r1 - X.r0, X.n, X.bits.shape[1], X.r0, int(seed), float(noise[0]),
                     float(noise[1]))
    return X


def planted_bits(m, n, k_true, d_u, d_v, p_fn, p_fp, seed, rank=None, world=None) -> DeviceBits:
    """The planted recipe of pybmf_b200.synth.planted on the device: U* ~ Bern(d_u)^{m x k}, V* ~ Bern(d_v)^{n x k},
    X = (U* o V*^T) with ones dropped with probability p_fn and entries set with probability p_fp."""
    _native.require_gpu()
    assert 1 <= k_true <= 64, "k_true <= 64 (one usage word per row)"
    r0, r1 = _rows_of(m, rank, world)
    rows = r1 - r0
    words = device.words_for(n)
    uw = device.zeros((max(rows, 1), 2), torch.int64)           # usage words (k bits; an even word count for alignment)
    vt = device.zeros((k_true, words), torch.int64)             # rows of V*^T, identical on every rank
    _native.call("bmf_random_bits", vt, k_true, n, words, 0, int(seed), 2, float(d_v))
    bits = device.zeros((max(rows, 1), words), torch.int64)
    if rows > 0:
        _native.call("bmf_random_bits", uw, rows, k_true, 2, r0, int(seed), 1, float(d_u))
        _native.call("bmf_bool_product", uw[:, :1].contiguous(), rows, 1, vt, k_true, words, bits)
        _native.call("bmf_noise_bits", bits, rows, n, words, r0, int(seed), float(p_fn), float(p_fp))
    return DeviceBits(bits, m, n, r0, r1)


def transpose_bits(bits, rows, ncols):
    """Bit matrix [rows][words(ncols)] -> its transpose [ncols][words(rows)] on the device."""
    words_t = device.words_for(rows)
    out = device.zeros((max(ncols, 1), words_t), torch.int64)
    step = 64 * 65535                                           # rows per launch (grid.y limit), a multiple of 64
    for a in range(0, rows, step):
        b = min(rows, a + step)
        _native.call("bmf_transpose_bits", bits[a:b], b - a, ncols, bits.shape[1], out[:, a // 64:], words_t)
    return out
