"""Seeded synthetic Boolean matrices for tests and benchmarks (host side, numpy/scipy).

`planted` is the recipe SURVEY.md section 8(d) fixes for BASELINE.json configs
c2-c4 (MovieLens-1M-shaped and Netflix-shaped matrices): a Boolean product of two
random sparse factors, thinned by false negatives and sprinkled with false
positives, all drawn from ONE legacy `np.random.RandomState(seed)` in a fixed
call order so that every run (here, on the GPU box, on every rank) sees the same X.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def planted(m, n, k_true, d_u, d_v, p_fn, p_fp, seed):
    """Return X (csr int64, m x n) = noisy (U* o V*^T)."""
    rng = np.random.RandomState(seed)
    U = sp.random(m, k_true, density=d_u, format="csr", random_state=rng)
    U.data[:] = 1
    V = sp.random(n, k_true, density=d_v, format="csr", random_state=rng)
    V.data[:] = 1
    X = (U @ V.T).tocsr()
    X.data[:] = 1
    keep = rng.rand(X.nnz) >= p_fn
    X.data = X.data * keep
    X.eliminate_zeros()
    nfp = int(p_fp * m * n)
    r = rng.randint(0, m, nfp)
    c = rng.randint(0, n, nfp)
    F = sp.csr_matrix((np.ones(nfp), (r, c)), shape=(m, n))
    X = (X + F).tocsr()
    X.data[:] = 1
    X.sort_indices()
    return X.astype(np.int64)


def config_c2():
    """BASELINE.json configs[1]/[2]: 6040 x 3706 at ~4.5 % density."""
    return planted(6040, 3706, 20, 0.048, 0.048, 0.10, 0.005, seed=6040)


def config_c4(rows=None):
    """BASELINE.json configs[3]: 480189 x 17770 at ~1.2 % density (rows= takes a row slice)."""
    X = planted(480189, 17770, 40, 0.0175, 0.0175, 0.10, 0.001, seed=20240)
    if rows is not None:
        X = X[rows[0]:rows[1]]
    return X


def product_factors(m, n, k, seed, p=2.0 / 64):
    """BASELINE.json configs[4]: U ~ Bern(p)^{m x k}, V ~ Bern(p)^{n x k} as dense uint8."""
    rng = np.random.RandomState(seed)
    U = (rng.rand(m, k) < p).astype(np.uint8)
    V = (rng.rand(n, k) < p).astype(np.uint8)
    return U, V


def random_binary(m, n, density, seed):
    rng = np.random.RandomState(seed)
    return sp.csr_matrix((rng.rand(m, n) < density).astype(np.int64))
