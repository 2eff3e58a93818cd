"""TEST INFRASTRUCTURE ONLY -- Asso / AssoIter on bit rows, driven from Python over oracle/asso_c.c.

Same role and rules as oracle/asso_oracle.py (the checker, never the product; only tests/, smoke() and bench.py's
CPU arm may import it).  The greedy loop below restates PyBMF/models/Asso.py:62-140 step for step; the per-pair
arithmetic lives in the C file (AND + POPCNT, OpenMP).  It reaches BASELINE configs c2-c4 at FULL size, which the
dense numpy restatement cannot (34 GB operands at c4).

Parity status: PINNED -- tests/test_oracle_golden.py compares `asso_fit` here with the numpy restatement, the
ex01_6 known-answer table and the genuine reference's golden outputs (tests/golden/*.npz).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import time

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libasso_oracle.so")
_lib = None

_p, _i64, _f64, _int = C.c_void_p, C.c_int64, C.c_double, C.c_int


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from oracle import build_c
            build_c.build()
        L = C.CDLL(LIB_PATH)
        L.bmfo_threads.restype = _int
        L.bmfo_has_avx512.restype = _int
        L.bmfo_set_threads.argtypes = [_int]
        L.bmfo_assoc_counts.argtypes = [_p, _i64, _i64, _p]
        L.bmfo_basis.argtypes = [_p, _i64, _f64, _p, _i64, _p, _p]
        L.bmfo_score_all.argtypes = [_p, _p, _i64, _i64, _p, _p, _p, _i64, _p, _p, _f64, _f64, _int, _int, _p, _p, _p]
        L.bmfo_apply.argtypes = [_p, _p, _i64, _i64, _p, _i64, _p, _p, _f64, _f64, _int, _int, _p, _p]
        L.bmfo_confusion.argtypes = [_p, _p, _i64, _i64, _p, _p, _p]
        L.bmfo_bool_product.argtypes = [_p, _i64, _i64, _p, _i64, _i64, _i64, _p]
        L.bmfo_pack_csr.argtypes = [_p, _p, _i64, _int, _p, _i64]
        for name in ("bmfo_set_threads", "bmfo_assoc_counts", "bmfo_basis", "bmfo_score_all", "bmfo_apply",
                     "bmfo_confusion", "bmfo_bool_product", "bmfo_pack_csr"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def words_for(ncols):
    w = (int(ncols) + 63) // 64
    w += w & 1
    return max(w, 2)


def integer_weights(w_fp, w_fn, max_int=127, max_shift=30):
    """(a, b, s) with w_fp = a/2^s, w_fn = b/2^s exactly (then metrics.py:201 is exact in fp64), else None."""
    for s in range(max_shift + 1):
        a, b = w_fp * (1 << s), w_fn * (1 << s)
        if a == int(a) and b == int(b):
            a, b = int(a), int(b)
            return (a, b, s) if (0 <= a <= max_int and 0 <= b <= max_int and (a | b)) else None
    return None


def pack_csr(X, transposed=False):
    """Non-zero pattern of X -> uint64 bit rows [m][words(n)] (or X^T: [n][words(m)])."""
    X = sp.csr_matrix(X)
    if X.nnz and np.count_nonzero(X.data) != X.nnz:
        X = X.copy()
        X.eliminate_zeros()
    m, n = X.shape
    rows, cols = (n, m) if transposed else (m, n)
    words = words_for(cols)
    bits = np.zeros((max(rows, 1), words), dtype=np.uint64)
    ip = np.ascontiguousarray(X.indptr.astype(np.int64))
    ix = np.ascontiguousarray(X.indices.astype(np.int32))
    lib().bmfo_pack_csr(_ptr(ip), _ptr(ix), m, 1 if transposed else 0, _ptr(bits), words)
    return bits


def count_ones(bits):
    """Number of set bits of a uint64 bit matrix (chunked so that the unpacked view stays small)."""
    flat = np.ascontiguousarray(bits).view(np.uint8).reshape(-1)
    total = 0
    for a in range(0, flat.size, 1 << 26):
        total += int(np.unpackbits(flat[a:a + (1 << 26)]).sum(dtype=np.int64))
    return total


def unpack_rows(bits, ncols):
    b = np.ascontiguousarray(bits).view(np.uint8).reshape(bits.shape[0], -1)
    return np.unpackbits(b, axis=1, bitorder="little")[:, :ncols]


def column_bytes(vec01):
    """Canonical bytes of a 0/1 vector: bit i of the vector is bit i&7 of byte i>>3 (ceil(len/8) bytes)."""
    return np.packbits(np.asarray(vec01, dtype=np.uint8), bitorder="little").tobytes()


def result_digest(steps, U_cols, V_cols):
    """The digest bench.py prints (`result_digest`) and tests/golden/c*_digest.json hold:
    per step the winner (ORIGINAL column index of the chosen association row), the bits of the float64 score,
    #used rows and the cumulative TP / FP of the cover, plus SHA-256 of the packed U and V columns in factor order."""
    hu, hv = hashlib.sha256(), hashlib.sha256()
    for u in U_cols:
        hu.update(column_bytes(u))
    for v in V_cols:
        hv.update(column_bytes(v))
    return {"winners": [int(s["winner"]) for s in steps],
            "score_bits": [np.float64(s["score"]).tobytes().hex() for s in steps],
            "used": [int(s["used"]) for s in steps],
            "tp": [int(s["tp"]) for s in steps], "fp": [int(s["fp"]) for s in steps],
            "u_sha256": hu.hexdigest(), "v_sha256": hv.hexdigest()}


class BitState:
    """X, covered mask, candidate basis and per-row counters as bit rows / int32 vectors (host memory)."""

    def __init__(self, X, tau, threads=None):
        L = lib()
        if threads:
            L.bmfo_set_threads(int(threads))
        X = sp.csr_matrix(X)
        self.m, self.n = X.shape
        self.words = words_for(self.n)
        self.x = pack_csr(X)
        self.sum_x = count_ones(self.x)
        t0 = time.perf_counter()
        xt = pack_csr(X, transposed=True)
        cnt = np.zeros((self.n, self.n), dtype=np.int32)
        L.bmfo_assoc_counts(_ptr(xt), self.n, xt.shape[1], _ptr(cnt))              # Asso.py:207
        self.assoc_seconds = time.perf_counter() - t0
        self.cnt = cnt
        self.basis = np.zeros((self.n, self.words), dtype=np.uint64)
        self.alive = np.zeros(self.n, dtype=np.uint8)
        self.pop = np.zeros(self.n, dtype=np.int32)
        L.bmfo_basis(_ptr(cnt), self.n, float(tau), _ptr(self.basis), self.words, _ptr(self.alive), _ptr(self.pop))
        self.c = np.zeros_like(self.x)
        self.tpo = np.zeros(self.m, dtype=np.int32)
        self.fpo = np.zeros(self.m, dtype=np.int32)

    def score_all(self, w_fp, w_fn, iw):
        gp = np.zeros(self.n, dtype=np.int64)
        gn = np.zeros(self.n, dtype=np.int64)
        gd = np.zeros(self.n, dtype=np.int64)
        wa, wb = (iw[0], iw[1]) if iw else (0, 0)
        lib().bmfo_score_all(_ptr(self.x), _ptr(self.c), self.m, self.words, _ptr(self.basis), _ptr(self.alive),
                             _ptr(self.pop), self.n, _ptr(self.tpo), _ptr(self.fpo), float(w_fp), float(w_fn), wa, wb,
                             _ptr(gp), _ptr(gn), _ptr(gd))
        return gp, gn, gd

    def apply(self, j, w_fp, w_fn, iw):
        used = np.zeros(self.m, dtype=np.uint8)
        tot = np.zeros(3, dtype=np.int64)
        wa, wb = (iw[0], iw[1]) if iw else (0, 0)
        b = np.ascontiguousarray(self.basis[j])
        lib().bmfo_apply(_ptr(self.x), _ptr(self.c), self.m, self.words, _ptr(b), int(self.pop[j]), _ptr(self.tpo),
                         _ptr(self.fpo), float(w_fp), float(w_fn), wa, wb, _ptr(used), _ptr(tot))
        return used, tot


def asso_fit(X, k, tau, w_fp=0.5, w_fn=None, tol=0, threads=None, progress=None):
    """Asso.fit on bit rows -- PyBMF/models/Asso.py:48-140 (+ BaseModelTools.py:299-405 for the early stops).

    Returns dict(steps, U_cols, V_cols, digest, error): one entry of `steps` per greedy step with the columns of
    logs['updates'] that derive from integers.  Candidate scores (Asso.py:184-186):
      integer weights a/2^s, b/2^s : score_j = (b*TP - a*FP + sum_i relu(b*P - a*N)) * 2^-s   -- exact, equals the reference;
      other weights                : score_j = (-w_fp)*(FP + sum_use N) + w_fn*(TP + sum_use P) -- exact integer totals, the
                                     reference adds per-row rounded values (last-ulp differences, see asso_oracle.py).
    Quirks D1 (error <= tol drops the factor just added) and D2 (no improving candidate -> TypeError) are reported,
    not raised: `error` names what the reference would have raised."""
    w_fn = 1 - w_fp if w_fn is None else w_fn
    iw = integer_weights(float(w_fp), float(w_fn))
    st = BitState(X, tau, threads)
    m, n = st.m, st.n
    size = m * n
    steps = []
    kept = [None] * (k if k is not None else 1)                                      # BaseModelTools.py:283-288: m x k zeros
    tp_tot = fp_tot = 0
    best_score = 0
    step = 0
    error = ""
    while True:
        best_score = 0 if step == 0 else best_score                                  # Asso.py:71
        if not st.alive.any():
            error = "TypeError"                                                      # Asso.py:75-77 (D2)
            break
        t0 = time.perf_counter()
        gp, gn, gd = st.score_all(w_fp, w_fn, iw)
        dt = time.perf_counter() - t0
        if iw:
            a, b, s = iw
            score = (b * tp_tot - a * fp_tot + gd).astype(np.float64) * (1.0 / float(1 << s))
        else:
            score = (-w_fp) * (fp_tot + gn).astype(np.float64) + w_fn * (tp_tot + gp).astype(np.float64)
        best_idx = -1
        cur = best_score
        alive_ids = np.flatnonzero(st.alive)
        sc = score[alive_ids]
        if len(sc):                                                                  # Asso.py:94: strict >, first maximum wins
            top = int(np.argmax(sc))
            if sc[top] > cur:
                best_idx, cur = int(alive_ids[top]), float(sc[top])
        if best_idx < 0:
            error = "TypeError"                                                      # Asso.py:98-100 (D2)
            break
        best_score = cur
        used, tot = st.apply(best_idx, w_fp, w_fn, iw)                               # Asso.py:103-110
        st.alive[best_idx] = 0                                                       # Asso.py:106-107
        tp_tot += int(tot[1])
        fp_tot += int(tot[2])
        vrow = unpack_rows(st.basis[best_idx:best_idx + 1], n)[0]
        while len(kept) < step + 1:
            kept.append(None)
        kept[step] = (used, vrow)
        fn = st.sum_x - tp_tot
        tn = size - tp_tot - fp_tot - fn
        err = 1 - np.float64(tp_tot + tn) / size                                     # metrics.py ACC / ERR
        steps.append({"k": step, "winner": best_idx, "score": float(best_score), "used": int(tot[0]),
                      "rowsum": int(st.pop[best_idx]), "tp": tp_tot, "fp": fp_tot, "fn": fn, "err": float(err),
                      "score_seconds": dt})
        if progress:
            progress(steps[-1])
        if err <= tol:                                                               # Asso.py:135 (D1): truncate to `step` columns
            kept = kept[:step]
            st.c[:] = 0
            st.tpo[:] = 0
            st.fpo[:] = 0
            tp_tot = fp_tot = 0
            live = [e for e in kept if e is not None]
            if live:
                kw = (len(live) + 63) // 64
                uw = np.zeros((m, kw), dtype=np.uint64)
                vt = np.zeros((len(live), st.words), dtype=np.uint64)
                for l, (u, v) in enumerate(live):
                    uw[:, l >> 6] |= u.astype(np.uint64) << np.uint64(l & 63)
                    vt[l] = pack_csr(sp.csr_matrix(v.reshape(1, -1)))[0]
                lib().bmfo_bool_product(_ptr(uw), m, kw, _ptr(vt), len(live), st.words, -1, _ptr(st.c))
                cnts = np.zeros(3, dtype=np.int64)
                lib().bmfo_confusion(_ptr(st.x), _ptr(st.c), m, st.words, _ptr(cnts), _ptr(st.tpo), _ptr(st.fpo))
                tp_tot, fp_tot = int(cnts[0]), int(cnts[1])
        if k is not None and step + 1 >= k:                                          # Asso.py:136
            break
        step += 1
    zero_u, zero_v = np.zeros(m, np.uint8), np.zeros(n, np.uint8)
    U_cols = [e[0] if e is not None else zero_u for e in kept]
    V_cols = [e[1] if e is not None else zero_v for e in kept]
    return {"steps": steps, "U_cols": U_cols, "V_cols": V_cols, "error": error, "sum_x": st.sum_x,
            "assoc_seconds": st.assoc_seconds, "candidates": int((st.pop > 0).sum()),
            "digest": result_digest(steps, U_cols, V_cols)}


def asso_iter_fit(X, U, V, k, w_fp=0.5, w_fn=None):
    """AssoIter._fit on bit rows -- PyBMF/models/AssoIter.py:45-100.  U (m x kU), V (n x kU) dense 0/1.
    Returns dict(U, trace, scores, errors) like oracle/asso_oracle.py::asso_iter_fit."""
    L = lib()
    w_fn = 1 - w_fp if w_fn is None else w_fn
    iw = integer_weights(float(w_fp), float(w_fn))
    wa, wb = (iw[0], iw[1]) if iw else (0, 0)
    X = sp.csr_matrix(X)
    m, n = X.shape
    size = m * n
    words = words_for(n)
    x = pack_csr(X)
    sum_x = count_ones(x)
    U = (np.asarray(U) != 0).astype(np.uint8).copy()
    V = (np.asarray(V) != 0).astype(np.uint8)
    kU = U.shape[1]
    kw = max((kU + 63) // 64, 1)
    vt = pack_csr(sp.csr_matrix(V.T)) if kU else np.zeros((1, words), np.uint64)
    vpop = V.sum(axis=0).astype(np.int64)

    def u_words():
        uw = np.zeros((m, kw), dtype=np.uint64)
        for l in range(kU):
            uw[:, l >> 6] |= U[:, l].astype(np.uint64) << np.uint64(l & 63)
        return uw

    def counts(skip=-1):
        c = np.zeros((m, words), dtype=np.uint64)
        uw = u_words()
        L.bmfo_bool_product(_ptr(uw), m, kw, _ptr(vt), kU, words, skip, _ptr(c))
        cn = np.zeros(3, dtype=np.int64)
        tpo, fpo = np.zeros(m, np.int32), np.zeros(m, np.int32)
        L.bmfo_confusion(_ptr(x), _ptr(c), m, words, _ptr(cn), _ptr(tpo), _ptr(fpo))
        return c, cn, tpo, fpo

    def err_of(tp, fp, fn):
        tn = size - tp - fp - fn
        return 1 - np.float64(tp + tn) / size

    _c, cn, _t, _f = counts()
    best_score = -w_fp * np.int64(cn[1]) + w_fn * np.int64(cn[0])                  # AssoIter.py:52
    best_error = err_of(int(cn[0]), int(cn[1]), int(cn[2]))
    n_stop = 0
    trace, scores, errors = [], [], []
    improving = True
    while improving:
        for col in range(k):
            if k > kU:                                                             # AssoIter.py:85-86 fancy index
                raise IndexError("index (%d) out of range" % (k - 1))
            c, _cn, tpo, fpo = counts(skip=col)                                    # AssoIter.py:86-87
            used = np.zeros(m, dtype=np.uint8)
            tot = np.zeros(3, dtype=np.int64)
            b = np.ascontiguousarray(vt[col])
            L.bmfo_apply(_ptr(x), _ptr(c), m, words, _ptr(b), int(vpop[col]), _ptr(tpo), _ptr(fpo), float(w_fp),
                         float(w_fn), wa, wb, _ptr(used), _ptr(tot))               # get_vector for V[:, col]
            U[:, col] = used                                                       # AssoIter.py:60 (always)
            tp, fp = int(tpo.sum(dtype=np.int64)), int(fpo.sum(dtype=np.int64))
            fn = sum_x - tp
            score = -w_fp * np.int64(fp) + w_fn * np.int64(tp)
            error = err_of(tp, fp, fn)
            if error < best_error:                                                 # AssoIter.py:64
                best_error, best_score = error, score
                trace.append((col, True))
                scores.append(float(score))
                errors.append(float(error))
                n_stop = 0
            else:
                n_stop += 1
                trace.append((col, False))
                if n_stop == k:                                                    # AssoIter.py:74-77
                    improving = False
                    break
    return {"U": U, "trace": trace, "scores": scores, "errors": errors}
