"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of PyBMF's Asso hot path.

This module is the *checker*, never the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  The shipped path (`pybmf_b200`) never falls back to it.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this restatement
against (a) the stored known-answer table of examples/ex01_6_logs.ipynb:253-361
and (b) outputs of the genuine reference run in the authoring container
(`oracle/make_golden.py` -> `tests/golden/*.npz`).

Every function cites the reference lines it restates (paths relative to
/root/reference).  Integer quantities are kept as exact integers; every
floating-point expression is written with the same operations, in the same
order, as the reference evaluates them (three separate ufuncs, no FMA).
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def as_dense01(X) -> np.ndarray:
    """Dense uint8 0/1 view of an ndarray / scipy sparse matrix (non-zero -> 1)."""
    if hasattr(X, "toarray"):
        X = X.toarray()
    X = np.asarray(X)
    return (X != 0).astype(np.uint8)


def resolve_weights(w_fp, w_fn):
    """`w_fn = 1 - w_fp if w_fn is None` -- PyBMF/utils/metrics.py:200."""
    return w_fp, (1 - w_fp if w_fn is None else w_fn)


# ----------------------------------------------------------------------------
# Boolean product and confusion counts
# ----------------------------------------------------------------------------
def bool_product(U, V) -> np.ndarray:
    """OR-AND product `U o V^T` (m x n, uint8).

    Restates get_prediction -> matmul(U, V.T, sparse=True, boolean=True):
    integer product then `.minimum(1)` -- PyBMF/utils/common.py:98-107,
    PyBMF/utils/boolean_utils.py:71-78.
    """
    U = as_dense01(U).astype(np.int64)
    V = as_dense01(V).astype(np.int64)
    assert U.shape[1] == V.shape[1], "U and V should be multiplicable"
    return np.minimum(U @ V.T, 1).astype(np.uint8)


def confusion(gt, pd, axis=None):
    """(TP, FP, FN) as exact int64 -- PyBMF/utils/metrics.py:56-68,75-76.

    TP = sum(gt AND pd); FP = sum(max(pd - gt, 0)); FN = FP with roles swapped.
    """
    gt = as_dense01(gt).astype(np.int64)
    pd = as_dense01(pd).astype(np.int64)
    tp = (gt * pd).sum(axis=axis)
    fp = np.maximum(pd - gt, 0).sum(axis=axis)
    fn = np.maximum(gt - pd, 0).sum(axis=axis)
    return tp, fp, fn


def coverage_score_from_counts(tp, fp, w_fp, w_fn):
    """`- w_fp * FP + w_fn * TP` -- PyBMF/utils/metrics.py:201 (literal order)."""
    w_fp, w_fn = resolve_weights(w_fp, w_fn)
    return -w_fp * np.asarray(fp) + w_fn * np.asarray(tp)


def rates_from_counts(tp, fp, fn, size, sum_pd=None):
    """All scalar metrics of PyBMF/utils/metrics.py:79-139 from three integers.

    Formulas are the reference's literal ones (FPR = 1 - TNR, not FP/denom;
    ERR = 1 - ACC; each rate is 0 when its denominator is 0).
    `size` = m*n (reconstruction) or the number of triplets (prediction).
    """
    tp = int(tp); fp = int(fp); fn = int(fn); size = int(size)
    tn = size - tp - fp - fn
    sum_gt = tp + fn
    sum_pd = tp + fp if sum_pd is None else int(sum_pd)
    sum_inv_gt = size - sum_gt
    f = np.float64
    tpr = f(tp) / f(sum_gt) if sum_gt > 0 else 0
    tnr = f(tn) / f(sum_inv_gt) if sum_inv_gt > 0 else 0
    fpr = 1 - tnr
    fnr = 1 - tpr
    ppv = f(tp) / f(sum_pd) if sum_pd > 0 else 0
    acc = f(tp + tn) / size
    err = 1 - acc
    denom = ppv + tpr
    f1 = 2 * ppv * tpr / denom if denom > 0 else 0
    return {"TP": tp, "FP": fp, "TN": tn, "FN": fn, "TPR": tpr, "TNR": tnr, "FPR": fpr,
            "FNR": fnr, "PPV": ppv, "ACC": acc, "ERR": err, "F1": f1,
            "Recall": tpr, "Precision": ppv, "Accuracy": acc, "Error": err}


# ----------------------------------------------------------------------------
# association matrix and candidate basis
# ----------------------------------------------------------------------------
def assoc_counts(X) -> np.ndarray:
    """`X.T @ X` co-occurrence counts (n x n int64) -- PyBMF/models/Asso.py:207."""
    Xd = as_dense01(X)
    Xf = Xd.astype(np.float32 if Xd.shape[0] < (1 << 24) else np.float64)
    return np.rint(Xf.T @ Xf).astype(np.int64)


def build_assoc(X) -> np.ndarray:
    """assoc[i, :] = cnt[i, :] / s[i] if s[i] > 0 else 0 -- Asso.py:207-212."""
    cnt = assoc_counts(X)
    s = as_dense01(X).astype(np.int64).sum(axis=0)
    out = np.zeros(cnt.shape, dtype=np.float64)
    nz = s > 0
    out[nz] = cnt[nz].astype(np.float64) / s[nz].astype(np.float64)[:, None]
    return out


def build_basis(assoc, tau):
    """basis = (assoc > tau), all-zero rows dropped, order kept -- Asso.py:231-234,
    PyBMF/utils/common.py:75 (strict `>`).  Returns (rows uint8 nb x n, source column ids)."""
    B = (np.asarray(assoc) > tau).astype(np.uint8)
    keep = np.flatnonzero(B.sum(axis=1) != 0)
    return B[keep], keep


# ----------------------------------------------------------------------------
# cover-gain scoring
# ----------------------------------------------------------------------------
def cover_counts(X, C, B):
    """P[i,j] = |X_i & ~C_i & B_j|, N[i,j] = |~X_i & ~C_i & B_j| (int64, m x nb).

    This is the integer content of get_vector's `add(X_old, pattern)` followed by
    TP/FP per row -- Asso.py:173-179, metrics.py:56-68: adding the pattern of
    candidate j to row i turns P[i,j] uncovered ones into TPs and N[i,j]
    uncovered zeros into FPs.
    """
    X = as_dense01(X); C = as_dense01(C); B = as_dense01(B)
    unc = (1 - C)
    Xu = (X * unc).astype(np.float32)
    Nu = ((1 - X) * unc).astype(np.float32)
    Bt = B.T.astype(np.float32)
    assert X.shape[1] < (1 << 24)
    P = np.rint(Xu @ Bt).astype(np.int64)
    N = np.rint(Nu @ Bt).astype(np.int64)
    return P, N


def score_candidates(X, C, B, w_fp, w_fn):
    """Scores of every candidate row of B against (X, covered mask C).

    Restates the body of the hot loop Asso.py:83-95 -> get_vector Asso.py:144-188
    for all candidates at once.  Returns (score[nb] float64, use[m, nb] bool,
    P, N, tp_old[m], fp_old[m]).
    """
    w_fp, w_fn = resolve_weights(w_fp, w_fn)
    tp_old, fp_old, _ = confusion(X, C, axis=1)
    P, N = cover_counts(X, C, B)
    s_old = -w_fp * fp_old + w_fn * tp_old                      # metrics.py:201
    nb = B.shape[0]
    score = np.zeros(nb, dtype=np.float64)
    use = np.zeros((X.shape[0], nb), dtype=bool)
    for j in range(nb):
        s_new = -w_fp * (fp_old + N[:, j]) + w_fn * (tp_old + P[:, j])
        u = s_new > s_old                                      # Asso.py:181 (strict)
        use[:, j] = u
        score[j] = s_old[~u].sum() + s_new[u].sum()            # Asso.py:184-186
    return score, use, P, N, tp_old, fp_old


def get_vector(X, C, b, w_fp, w_fn):
    """One candidate row `b` -- Asso.py:144-188.  Returns (score, u bool[m])."""
    score, use, *_ = score_candidates(X, C, np.asarray(b).reshape(1, -1), w_fp, w_fn)
    return score[0], use[:, 0]


class NoCandidateError(TypeError):
    """The reference's D2 defect: early_stop(msg=...) calls _early_stop without
    `verbose` (PyBMF/models/BaseModelTools.py:338-341 vs :346) -> TypeError.
    `args[1]` carries the factors/logs as they stood when the reference would have raised."""


def asso_fit(X, k, tau, w_fp=0.5, w_fn=None, tol=0):
    """Asso.fit on the training matrix -- Asso.py:48-140, BaseModelTools.py:299-405.

    Returns dict(U, V, logs, basis_left).  `logs` is a list of per-step dicts
    with the columns of logs['updates'] (Asso.py:121-132).  Reproduces quirk D1
    (error <= tol truncates the factor just added, BaseModelTools.py:326-328 with
    the 0-based k of Asso.py:135) and D2 (TypeError when nothing improves).
    """
    X = as_dense01(X)
    m, n = X.shape
    w_fp, w_fn = resolve_weights(w_fp, w_fn)
    B, _src = build_basis(build_assoc(X), tau)
    kcols = k if k is not None else 1                          # BaseModelTools.py:283-288
    U = np.zeros((m, kcols), dtype=np.uint8)
    V = np.zeros((n, kcols), dtype=np.uint8)
    logs = []
    step = 0
    best_score = 0
    is_improving = True
    size = m * n
    sum_gt = int(X.sum())
    while is_improving:
        best_score = 0 if step == 0 else best_score            # Asso.py:71
        if B.shape[0] == 0:                                    # Asso.py:75-77
            raise NoCandidateError("Candidate list is empty", {"U": U, "V": V, "logs": logs})
        C = bool_product(U, V)                                 # Asso.py:80
        score, use, *_ = score_candidates(X, C, B, w_fp, w_fn)
        best_idx = None
        for j in range(B.shape[0]):                            # Asso.py:94 strict >
            if score[j] > best_score:
                best_score, best_idx = score[j], j
        if best_idx is None:                                   # Asso.py:98-100
            raise NoCandidateError("No pattern found.", {"U": U, "V": V, "logs": logs})
        col = use[:, best_idx].astype(np.uint8)
        row = B[best_idx].copy()
        if U.shape[1] < step + 1:                              # BaseModelTools.py:378-379
            U = np.hstack([U, np.zeros((m, step + 1 - U.shape[1]), np.uint8)])
            V = np.hstack([V, np.zeros((n, step + 1 - V.shape[1]), np.uint8)])
        U[:, step] = col
        V[:, step] = row
        B = np.delete(B, best_idx, axis=0)                     # Asso.py:106-107
        C = bool_product(U, V)                                 # Asso.py:110
        tp, fp, fn = confusion(X, C)
        rec = {"k": step, "score": float(best_score),
               "score_0.5": float(-0.5 * fp + 0.5 * tp),       # Asso.py:119
               "desc_len": float(1 * (float(U.sum()) + float(V.sum())) + 1 * fp + 1 * fn),
               "shape": [int(col.sum()), int(row.sum())]}
        r = rates_from_counts(tp, fp, fn, size)
        for name in ("TP", "TPR", "FP", "FPR", "FN", "FNR", "ERR", "ACC",
                     "Recall", "Precision", "F1"):
            rec[name] = r[name]
        logs.append(rec)
        is_improving = True
        if r["ERR"] <= tol:                                    # Asso.py:135 (D1)
            U = U[:, :step]
            V = V[:, :step]
        if k is not None and step + 1 >= k:                    # Asso.py:136
            is_improving = False
        step += 1
    return {"U": U, "V": V, "logs": logs, "basis_left": B, "sum_gt": sum_gt}


def asso_iter_fit(X, U, V, k, w_fp=0.5, w_fn=None):
    """AssoIter._fit -- PyBMF/models/AssoIter.py:45-100.

    `k` is the *requested* rank imported from the source model (AssoIter.py:28).
    U is refined in place semantics: returns dict(U, trace, refinements) where
    trace is a list of (column, accepted) and refinements the logged rows.
    """
    X = as_dense01(X)
    U = as_dense01(U).copy()
    V = as_dense01(V)
    m, n = X.shape
    size = m * n
    w_fp, w_fn = resolve_weights(w_fp, w_fn)
    C = bool_product(U, V)
    tp, fp, fn = confusion(X, C)
    best_score = -w_fp * fp + w_fn * tp                        # AssoIter.py:52
    best_error = rates_from_counts(tp, fp, fn, size)["ERR"]
    n_stop = 0
    trace, refinements = [], []
    is_improving = True
    while is_improving:
        for c in range(k):
            if k > U.shape[1]:                                 # lil fancy indexing U[:, idx], AssoIter.py:85-86
                raise IndexError("index (%d) out of range" % (k - 1))
            idx = [i for i in range(k) if i != c]
            C_old = bool_product(U[:, idx], V[:, idx])         # AssoIter.py:86
            score, u = get_vector(X, C_old, V[:, c], w_fp, w_fn)
            U[:, c] = u                                        # AssoIter.py:60 (always)
            C = bool_product(U, V)
            tp, fp, fn = confusion(X, C)
            r = rates_from_counts(tp, fp, fn, size)
            if r["ERR"] < best_error:                          # AssoIter.py:64
                best_error, best_score = r["ERR"], score
                refinements.append({"k": c, "score": float(best_score), "error": float(best_error),
                                    "Recall": r["Recall"], "Precision": r["Precision"],
                                    "Accuracy": r["Accuracy"], "F1": r["F1"]})
                trace.append((c, True))
                n_stop = 0
            else:
                n_stop += 1
                trace.append((c, False))
                if n_stop == k:                                # AssoIter.py:74-77
                    is_improving = False
                    break
    return {"U": U, "trace": trace, "refinements": refinements}


# ----------------------------------------------------------------------------
# integer ("exact") weight form, used to check the tensor-core formulation
# ----------------------------------------------------------------------------
def integer_weights(w_fp, w_fn, max_int=127, max_shift=30):
    """Return (a, b, s) with w_fp = a / 2**s and w_fn = b / 2**s exactly, 0 <= a,b <= max_int,
    or None.  When this exists every product/sum of metrics.py:201 is exact in fp64,
    so `s_new > s_old`  <=>  b*P - a*N > 0 and score = 2**-s * integer."""
    w_fp, w_fn = resolve_weights(w_fp, w_fn)
    for s in range(max_shift + 1):
        a = w_fp * (1 << s)
        b = w_fn * (1 << s)
        if a == int(a) and b == int(b):
            a, b = int(a), int(b)
            if 0 <= a <= max_int and 0 <= b <= max_int:
                return a, b, s
            return None
    return None


def integer_gains(X, C, B, a, b):
    """G[j] = sum_i relu(b*P[i,j] - a*N[i,j]) (int64) -- the single signed contraction."""
    P, N = cover_counts(X, C, B)
    D = b * P - a * N
    return np.maximum(D, 0).sum(axis=0)
