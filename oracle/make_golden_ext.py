"""TEST INFRASTRUCTURE ONLY -- golden vectors for the SURVEY section 8(f) widenings, from the GENUINE reference
(imported through oracle/ref_shim.py; authoring container only):

  tests/golden/expansion.npz : GreConDPlus._expansion (both axes) and expansion() -- PyBMF/models/GreConDPlus.py:207-308
  tests/golden/assoopt.npz   : AssoOpt.set_optimal_row for every row            -- PyBMF/models/AssoOpt.py:69-80

    python oracle/make_golden_ext.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ref_shim  # noqa: E402
from pybmf_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def dense01(M):
    return (np.asarray(M.todense() if hasattr(M, "todense") else M) != 0).astype(np.uint8)


def expansion_cases():
    ref_shim.load()
    from PyBMF.models.GreConDPlus import _expansion, expansion
    from PyBMF.utils import get_residual
    rec = {}
    cases = []
    rng = np.random.RandomState(3)
    for ci, (m, n, w_fp, w_fn) in enumerate([(60, 45, 0.5, None), (130, 70, 0.2, None), (90, 140, 0.3, 0.6), (64, 128, 0.25, 0.75)]):
        X = synth.planted(m, n, 4, 0.25, 0.25, 0.1, 0.03, seed=20 + ci)
        # a partial cover U o V^T and the residual the reference passes as X_old (GreConDPlus.py:56)
        U = sp.csr_matrix((rng.rand(m, 2) < 0.2).astype(np.int64))
        V = sp.csr_matrix((rng.rand(n, 2) < 0.2).astype(np.int64))
        X_old = sp.csr_matrix(get_residual(X=X, U=sp.lil_matrix(U), V=sp.lil_matrix(V)))
        # the starting pattern: a random one (cases 0, 1) or a formal-concept core like GreConD's (a few columns of one data
        # row and every row that has all of them), which the expansion then grows
        if ci in (0, 1):
            u = sp.lil_matrix((rng.rand(m, 1) < 0.15).astype(np.int64))
            v = sp.lil_matrix((rng.rand(n, 1) < 0.15).astype(np.int64))
        else:
            Xd = dense01(X)
            seed_row = int(np.argmax(Xd.sum(axis=1)))
            cols = np.flatnonzero(Xd[seed_row])[: 3 + ci]
            vv = np.zeros((n, 1), dtype=np.int64); vv[cols] = 1
            uu = (Xd[:, cols].sum(axis=1) == len(cols)).astype(np.int64).reshape(-1, 1)
            u, v = sp.lil_matrix(uu), sp.lil_matrix(vv)
        with ref_shim.quiet():
            r_score, r_index = _expansion(X, X_old, u, v, w_fp, w_fn, axis=1)
            c_score, c_index = _expansion(X, X_old, u, v, w_fp, w_fn, axis=0)
            u_exp, v_exp = expansion(X_gt=X, X_old=X_old, u=u, v=v, w_fp=w_fp, w_fn=w_fn)
        p = "c%d_" % ci
        rec[p + "X"] = dense01(X); rec[p + "X_old"] = dense01(X_old)
        rec[p + "u"] = dense01(u).ravel(); rec[p + "v"] = dense01(v).ravel()
        rec[p + "w"] = np.array([w_fp, np.nan if w_fn is None else w_fn])
        rec[p + "row"] = np.array([float(r_score), float(r_index)]); rec[p + "col"] = np.array([float(c_score), float(c_index)])
        rec[p + "u_exp"] = dense01(u_exp).ravel(); rec[p + "v_exp"] = dense01(v_exp).ravel()
        cases.append(ci)
        print("expansion case %d: row (%.4f, %d) col (%.4f, %d) |u_exp| %d |v_exp| %d" % (
            ci, r_score, r_index, c_score, c_index, rec[p + "u_exp"].sum(), rec[p + "v_exp"].sum()))
    rec["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "expansion.npz"), **rec)


def assoopt_cases():
    ref_shim.load()
    from PyBMF.models import Asso, AssoOpt
    rec = {}
    cases = []
    for ci, (m, n, k, tau, w_fp, w_fn) in enumerate([(60, 50, 4, 0.3, 1, 1), (45, 70, 5, 0.25, 0.5, 0.5), (40, 130, 3, 0.3, 0.2, 0.8)]):
        X = synth.planted(m, n, k, 0.25, 0.25, 0.1, 0.03, seed=40 + ci)
        with ref_shim.quiet():
            base = Asso(tau=tau, k=k, w_fp=0.5)
            base.fit(X, **ref_shim.FIT_KW)
            opt = AssoOpt(model=base, w_fp=w_fp, w_fn=w_fn)
            opt.load_dataset(X_train=X)
            best = np.array([opt.set_optimal_row(i) for i in range(m)], dtype=np.int64)
        p = "c%d_" % ci
        rec[p + "X"] = dense01(X); rec[p + "U"] = dense01(base.U); rec[p + "V"] = dense01(base.V)
        rec[p + "k"] = np.array(k); rec[p + "tau"] = np.array(tau); rec[p + "w"] = np.array([w_fp, w_fn], dtype=np.float64)
        rec[p + "best"] = best
        cases.append(ci)
        print("assoopt case %d: V%s best trials %s..." % (ci, rec[p + "V"].shape, best[:8]))
    rec["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "assoopt.npz"), **rec)




def grecond_case():
    """A model OTHER than Asso through the routed utils: the genuine GreConD run, its factors
    and, per step, what its call sites compute with get_prediction / get_residual / ERR / weighted_error /
    description_length (PyBMF/models/GreConD.py:51-66, MEBF.py:115-158, Panda.py:79-82)."""
    ref_shim.load()
    from PyBMF.models import GreConD
    from PyBMF.utils import get_prediction, get_residual, ERR, weighted_error, description_length, coverage_score
    X = synth.planted(150, 110, 5, 0.2, 0.2, 0.1, 0.02, seed=77)
    with ref_shim.quiet():
        mdl = GreConD(k=6, tol=0)
        mdl.fit(X, **ref_shim.FIT_KW)
    U, V = dense01(mdl.U), dense01(mdl.V)
    rec = {"X": dense01(X), "U": U, "V": V}
    err, werr, dl, rs_sum, cs, pd_sum = [], [], [], [], [], []
    with ref_shim.quiet():
        for t in range(U.shape[1]):
            Ut, Vt = sp.lil_matrix(mdl.U[:, : t + 1]), sp.lil_matrix(mdl.V[:, : t + 1])
            X_pd = get_prediction(U=Ut, V=Vt, boolean=True)
            X_rs = get_residual(X=X, U=Ut, V=Vt)
            err.append(float(ERR(gt=X, pd=X_pd)))
            werr.append(float(weighted_error(gt=X, pd=X_pd, w_fp=0.3, w_fn=0.7)))
            dl.append(float(description_length(gt=X, U=Ut, V=Vt, w_model=1.0, w_fp=1.0, w_fn=1.0)))
            cs.append(float(coverage_score(gt=X, pd=X_pd, w_fp=0.5)))
            rs_sum.append(float(X_rs.sum()))
            pd_sum.append(float(X_pd.sum()))
        rec["X_pd"] = dense01(get_prediction(U=mdl.U, V=mdl.V, boolean=True))
        rec["X_rs"] = dense01(get_residual(X=X, U=mdl.U, V=mdl.V))
    for name, vals in (("ERR", err), ("weighted_error", werr), ("desc_len", dl), ("coverage", cs), ("rs_sum", rs_sum), ("pd_sum", pd_sum)):
        rec["step_" + name] = np.array(vals, dtype=np.float64)
    df = mdl.logs["updates"]
    for c in ("Recall", "Precision", "Accuracy", "F1"):
        rec["log_" + c] = np.array([float(v) for v in df[("train", 0, c)]], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "grecond.npz"), **rec)
    print("grecond: U%s ERR %s" % (U.shape, np.round(err, 4)))


if __name__ == "__main__":
    expansion_cases()
    assoopt_cases()
    grecond_case()
