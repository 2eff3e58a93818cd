"""GreConD+'s pattern expansion on the GPU -- `expansion` / `_expansion` of PyBMF/models/GreConDPlus.py:207-308
(SURVEY.md section 8f, rank 1), same names, arguments and return values.

`_expansion(X_gt, X_old, u, v, w_fp, w_fn, axis)` is get_vector's shape seen from the other side: instead of asking
which rows should use a given pattern, it asks which ONE row (axis = 1) or column (axis = 0) gains most from joining
the pattern u x v.  Per row i outside u:  delta_i = coverage_score(x_i, old_i | v) - coverage_score(x_i, old_i), evaluated
with the reference's fp64 expression (PyBMF/utils/metrics.py:201) on integer TP / FP counts; rows already in u get
exactly 0.  The counts come from one streaming pass over the bit rows (bmf_expand_scores); column-wise expansion runs
the same kernel on the transposed bit matrices.  `expansion()` keeps all four bit matrices on the device for the
whole loop and moves one (score, index) pair per call.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
from scipy.sparse import lil_matrix

from . import _native, device
from . import utils as U_


def _vec_bits(vec, length):
    """0/1 vector (any container, (len, 1) / (1, len) / flat) -> device bit row [1, words_for(length)]."""
    v = vec.toarray() if sp.issparse(vec) else np.asarray(vec)
    v = (np.asarray(v).reshape(-1) != 0).astype(np.uint8)
    assert v.size == length, "u and v should match the shape of X"
    return torch.from_numpy(device.dense_to_words(v.reshape(1, -1))).to(device.dev())


class _ExpansionState:
    """X_gt and X_old as bit rows in both orientations, resident on the device."""

    def __init__(self, X_gt, X_old):
        _native.require_gpu()
        G, O = U_._pattern(X_gt), U_._pattern(X_old)
        assert G.shape == O.shape, "X_gt and X_old should have the same shape"
        self.m, self.n = G.shape
        self.x = U_._bits_on_device(G)
        self.o = U_._bits_on_device(O)
        self.xt = U_._bits_on_device(G, transposed=True)
        self.ot = U_._bits_on_device(O, transposed=True)
        self.delta = device.zeros((max(self.m, self.n),), torch.float64)
        self.best = device.zeros((2,), torch.int64)

    def score(self, u_bits, v_bits, w_fp, w_fn, axis):
        """(max delta, first argmax) over rows (axis = 1) or columns (axis = 0) -- GreConDPlus.py:275-308."""
        w_fn = 1 - w_fp if w_fn is None else w_fn
        if axis == 1:
            x, o, rows, pat, exc = self.x, self.o, self.m, v_bits, u_bits
        elif axis == 0:
            x, o, rows, pat, exc = self.xt, self.ot, self.n, u_bits, v_bits
        else:
            raise UnboundLocalError("cannot access local variable '_u' where it is not associated with a value")
        _native.call("bmf_expand_scores", x, o, rows, x.shape[1], pat, exc, float(w_fp), float(w_fn), self.delta, self.best)
        b = self.best.cpu().numpy()
        return np.float64(b[0:1].view(np.float64)[0]), int(b[1])


def _expansion(X_gt, X_old, u, v, w_fp, w_fn, axis):
    """Row-wise (axis = 1) or column-wise (axis = 0) expansion score of the pattern u x v -- GreConDPlus.py:267-308.
    Returns (score, index) = (d_scores.max(), d_scores.argmax())."""
    st = _ExpansionState(X_gt, X_old)
    return st.score(_vec_bits(u, st.m), _vec_bits(v, st.n), w_fp, w_fn, axis)


def expansion(X_gt, X_old, u, v, w_fp, w_fn):
    """Grow the pattern (u, v) one row or column at a time while that raises the coverage score --
    GreConDPlus.py:207-264.  Returns the expansion parts (u_exp, v_exp) as (m, 1) / (n, 1) lil matrices."""
    st = _ExpansionState(X_gt, X_old)
    m, n = st.m, st.n
    u_bits, v_bits = _vec_bits(u, m), _vec_bits(v, n)
    u_exp, v_exp = lil_matrix((m, 1)), lil_matrix((n, 1))
    n_iter = 0
    is_improving = True
    while is_improving:
        r_score, r_index = st.score(u_bits, v_bits, w_fp, w_fn, axis=1)
        c_score, c_index = st.score(u_bits, v_bits, w_fp, w_fn, axis=0)
        if r_score > c_score and r_score > 0:
            u_bits[0, r_index >> 6] |= _bit(r_index)
            u_exp[r_index] = 1
        elif c_score > r_score and c_score > 0:
            v_bits[0, c_index >> 6] |= _bit(c_index)
            v_exp[c_index] = 1
        else:
            is_improving = False
        n_iter += 1
    from .models import _say
    _say(f"[I]     expansion() finished after {n_iter} iterations.")
    return u_exp, v_exp


def _bit(i):
    """Bit i & 63 as a signed 64-bit Python int (torch.int64 words)."""
    b = 1 << (i & 63)
    return b - (1 << 64) if b >= (1 << 63) else b
