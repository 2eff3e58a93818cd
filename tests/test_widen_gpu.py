"""GPU: the SURVEY section 8(f) widenings against golden vectors of the GENUINE reference
(oracle/make_golden_ext.py -> tests/golden/{expansion,assoopt,grecond}.npz):
  * GreConDPlus._expansion / expansion   (PyBMF/models/GreConDPlus.py:207-308)
  * AssoOpt.set_optimal_row              (PyBMF/models/AssoOpt.py:69-80)
  * another model's call sites (GreConD / MEBF / Panda) through the routed utils."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import GOLDEN  # noqa: E402

FIT_KW = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)


@pytest.fixture(scope="module")
def M():
    from pybmf_b200 import _native, models
    _native.require_gpu()
    models.SILENT = True
    return models


def _w(arr):
    return float(arr[0]), (None if arr[1] != arr[1] else float(arr[1]))


def test_expansion_scores_match_reference_both_axes(M):
    from pybmf_b200.expansion import _expansion
    g = np.load(os.path.join(GOLDEN, "expansion.npz"))
    for ci in g["cases"]:
        p = "c%d_" % ci
        X, X_old = sp.csr_matrix(g[p + "X"]), sp.csr_matrix(g[p + "X_old"])
        u, v = sp.lil_matrix(g[p + "u"].reshape(-1, 1)), sp.lil_matrix(g[p + "v"].reshape(-1, 1))
        w_fp, w_fn = _w(g[p + "w"])
        for axis, key in ((1, "row"), (0, "col")):
            score, index = _expansion(X, X_old, u, v, w_fp, w_fn, axis=axis)
            assert (float(score), int(index)) == (float(g[p + key][0]), int(g[p + key][1])), (ci, axis)   # bit-exact fp64


def test_expansion_loop_matches_reference(M):
    from pybmf_b200.expansion import expansion
    g = np.load(os.path.join(GOLDEN, "expansion.npz"))
    for ci in g["cases"]:
        p = "c%d_" % ci
        w_fp, w_fn = _w(g[p + "w"])
        u_exp, v_exp = expansion(X_gt=sp.csr_matrix(g[p + "X"]), X_old=sp.csr_matrix(g[p + "X_old"]),
                                 u=sp.lil_matrix(g[p + "u"].reshape(-1, 1)), v=sp.lil_matrix(g[p + "v"].reshape(-1, 1)),
                                 w_fp=w_fp, w_fn=w_fn)
        assert sp.isspmatrix_lil(u_exp) and u_exp.shape == (g[p + "X"].shape[0], 1) and v_exp.shape == (g[p + "X"].shape[1], 1)
        assert np.array_equal((u_exp.toarray().ravel() != 0).astype(np.uint8), g[p + "u_exp"]), ci
        assert np.array_equal((v_exp.toarray().ravel() != 0).astype(np.uint8), g[p + "v_exp"]), ci


def test_expansion_deltas_against_numpy_restatement(M):
    """Every per-row delta (not only the maximum) against a numpy restatement of GreConDPlus.py:275-308, ragged shapes."""
    from pybmf_b200 import _native, device
    from pybmf_b200 import utils as U_
    rng = np.random.RandomState(5)
    for (m, n, w_fp, w_fn) in [(1, 1, 0.5, 0.5), (70, 130, 0.2, 0.8), (257, 65, 0.3, 0.6), (33, 1000, 0.5, 0.5)]:
        X = (rng.rand(m, n) < 0.3).astype(np.int64)
        O = (rng.rand(m, n) < 0.2).astype(np.int64)
        u = (rng.rand(m) < 0.3).astype(np.int64)
        v = (rng.rand(n) < 0.3).astype(np.int64)
        tp, fp = (X * O).sum(1), np.maximum(O - X, 0).sum(1)
        s_old = -w_fp * fp + w_fn * tp
        Xn = np.minimum(O + np.outer(1 - u, v), 1)
        s_new = -w_fp * np.maximum(Xn - X, 0).sum(1) + w_fn * (X * Xn).sum(1)
        want = s_new - s_old
        xb, ob = U_._bits_on_device(sp.csr_matrix(X)), U_._bits_on_device(sp.csr_matrix(O))
        vb = torch.from_numpy(device.dense_to_words(v.reshape(1, -1))).cuda()
        ub = torch.from_numpy(device.dense_to_words(u.reshape(1, -1))).cuda()
        delta = device.zeros((m,), torch.float64)
        best = device.zeros((2,), torch.int64)
        _native.call("bmf_expand_scores", xb, ob, m, xb.shape[1], vb, ub, w_fp, w_fn, delta, best)
        got = delta.cpu().numpy()
        assert np.array_equal(got, want), (m, n)
        b = best.cpu().numpy()
        assert b[0:1].view(np.float64)[0] == want.max() and int(b[1]) == int(want.argmax())


def test_assoopt_set_optimal_row_matches_reference(M):
    g = np.load(os.path.join(GOLDEN, "assoopt.npz"))
    for ci in g["cases"]:
        p = "c%d_" % ci
        X = sp.csr_matrix(g[p + "X"])
        k, tau = int(g[p + "k"]), float(g[p + "tau"])
        w_fp, w_fn = float(g[p + "w"][0]), float(g[p + "w"][1])
        base = M.Asso(tau=tau, k=k, w_fp=0.5)
        base.fit(X, **FIT_KW)
        assert np.array_equal((base.U.toarray() != 0).astype(np.uint8), g[p + "U"])
        assert np.array_equal((base.V.toarray() != 0).astype(np.uint8), g[p + "V"])
        opt = M.AssoOpt(model=base, w_fp=w_fp, w_fn=w_fn)
        opt.load_dataset(X_train=X)
        best, _score = opt.optimal_rows()
        assert np.array_equal(best, g[p + "best"]), ci
        assert [opt.set_optimal_row(i) for i in (0, 3, X.shape[0] - 1)] == [int(g[p + "best"][i]) for i in (0, 3, X.shape[0] - 1)]
        # the whole fit: U refined from the trial indices (int2bin is MSB first), then the reference's defect D4
        with pytest.raises(AttributeError):
            opt.fit(X, **FIT_KW)
        bits = ((g[p + "best"][:, None] >> (k - 1 - np.arange(k))[None, :]) & 1).astype(np.uint8)
        assert np.array_equal((opt.U.toarray() != 0).astype(np.uint8), bits)


def test_assoopt_wide_k(M):
    """k = 11 (2048 trials per row), V^T larger than one warp's registers: against a numpy brute force."""
    from pybmf_b200 import _native, device
    from pybmf_b200 import utils as U_
    rng = np.random.RandomState(8)
    m, n, k = 40, 300, 11
    X = (rng.rand(m, n) < 0.15).astype(np.int64)
    V = (rng.rand(n, k) < 0.08).astype(np.int64)
    xb = U_._bits_on_device(sp.csr_matrix(X))
    vt = U_._bits_on_device(sp.csr_matrix(V.T))
    best = device.zeros((m,), torch.int64)
    score = device.zeros((m,), torch.float64)
    _native.call("bmf_optimal_rows", xb, m, xb.shape[1], vt, k, 0.3, 0.7, best, score)
    trials = ((np.arange(1 << k)[:, None] >> (k - 1 - np.arange(k))[None, :]) & 1)        # [2^k, k], MSB first
    PD = np.minimum(trials @ V.T, 1)                                                        # [2^k, n]
    for i in range(m):
        tp = (PD * X[i]).sum(1)
        fp = np.maximum(PD - X[i], 0).sum(1)
        sc = -0.3 * fp + 0.7 * tp
        assert int(best[i].item()) == int(np.argmax(sc)) and float(score[i].item()) == float(sc.max()), i


def test_grecond_call_sites_through_routed_utils(M):
    """A model other than Asso: the factors of a genuine GreConD run, pushed through the utils its call sites use
    (get_prediction, get_residual, ERR, weighted_error, description_length, coverage_score, evaluate's metrics)."""
    from pybmf_b200 import utils as U_
    g = np.load(os.path.join(GOLDEN, "grecond.npz"))
    X = sp.csr_matrix(g["X"])
    U, V = sp.lil_matrix(g["U"].astype(float)), sp.lil_matrix(g["V"].astype(float))
    for t in range(g["U"].shape[1]):
        Ut, Vt = sp.lil_matrix(U[:, : t + 1]), sp.lil_matrix(V[:, : t + 1])
        X_pd = U_.get_prediction(U=Ut, V=Vt, boolean=True)
        X_rs = U_.get_residual(X=X, U=Ut, V=Vt)
        assert sp.isspmatrix_csr(X_pd) and X_pd.dtype == np.int64 and sp.isspmatrix_lil(X_rs)
        assert float(U_.ERR(gt=X, pd=X_pd)) == g["step_ERR"][t]
        assert float(U_.weighted_error(gt=X, pd=X_pd, w_fp=0.3, w_fn=0.7)) == g["step_weighted_error"][t]
        assert float(U_.description_length(gt=X, U=Ut, V=Vt, w_model=1.0, w_fp=1.0, w_fn=1.0)) == g["step_desc_len"][t]
        assert float(U_.coverage_score(gt=X, pd=X_pd, w_fp=0.5)) == g["step_coverage"][t]
        assert float(X_rs.sum()) == g["step_rs_sum"][t] and float(X_pd.sum()) == g["step_pd_sum"][t]
        got = U_.get_metrics(gt=X, pd=X_pd, metrics=["Recall", "Precision", "Accuracy", "F1"])
        assert [float(v) for v in got] == [float(g["log_" + c][t]) for c in ("Recall", "Precision", "Accuracy", "F1")]
    assert np.array_equal((U_.get_prediction(U=U, V=V, boolean=True).toarray() != 0).astype(np.uint8), g["X_pd"])
    assert np.array_equal((U_.get_residual(X=X, U=U, V=V).toarray() != 0).astype(np.uint8), g["X_rs"])


# ---- on-device input generation -------------------------------------------------------
def test_random_bits_density_padding_and_shard_independence(M):
    from pybmf_b200 import device, generate
    m, n, p = 3000, 1237, 0.3
    full = generate.random_bits(m, n, p, seed=11, rank=0, world=1)
    A = device.bits_to_host(full.bits, n)
    assert A.shape == (m, n)
    dens = A.mean()
    assert abs(dens - p) < 4 * np.sqrt(p * (1 - p) / (m * n)), dens
    raw = full.bits.cpu().numpy().view(np.uint64)
    assert (raw[:, (n + 63) // 64:] == 0).all() and (raw[:, n // 64] >> np.uint64(n % 64) == 0).all()   # pad bits are 0
    parts = [generate.random_bits(m, n, p, seed=11, rank=r, world=4) for r in range(4)]
    cat = np.concatenate([device.bits_to_host(q.bits[: q.r1 - q.r0], n) for q in parts if q.r1 > q.r0])
    assert np.array_equal(cat, A)                                   # a rank's rows do not depend on the world size
    other = device.bits_to_host(generate.random_bits(m, n, p, seed=12, rank=0, world=1).bits, n)
    assert 0.35 < (other != A).mean() < 0.5                         # another seed is another matrix (2 p (1 - p) = 0.42)
    # columns are not correlated with rows (a counter bug would show as identical rows / columns)
    assert len({r.tobytes() for r in A[:200]}) == 200


def test_noise_bits_follow_add_noise_semantics(M):
    """generator_utils.add_noise: X = max(X - Bern(p_pos), 0) then X = min(X + Bern(p_neg), 1)."""
    from pybmf_b200 import device, generate
    m, n = 2000, 900
    X = generate.random_bits(m, n, 0.5, seed=1, rank=0, world=1)
    before = device.bits_to_host(X.bits, n).astype(np.int64)
    generate.add_noise_bits(X, noise=(0.25, 0.1), seed=5)
    after = device.bits_to_host(X.bits, n).astype(np.int64)
    ones, zeros = before == 1, before == 0
    kept = after[ones].mean()                                       # P(stays 1) = (1 - p_pos) + p_pos * p_neg
    made = after[zeros].mean()                                      # P(0 -> 1) = p_neg
    assert abs(kept - (0.75 + 0.25 * 0.1)) < 0.004 and abs(made - 0.1) < 0.003, (kept, made)
    Y = generate.random_bits(m, n, 0.5, seed=1, rank=0, world=1)
    generate.add_noise_bits(Y, noise=(1.0, 0.0), seed=5)
    assert int(device.bits_to_host(Y.bits, n).sum()) == 0


@pytest.mark.parametrize("rows,ncols", [(1, 1), (64, 64), (70, 130), (1000, 37), (129, 4100)])
def test_transpose_bits(M, rows, ncols):
    from pybmf_b200 import device, generate
    rng = np.random.RandomState(rows + ncols)
    A = (rng.rand(rows, ncols) < 0.3).astype(np.uint8)
    bits = torch.from_numpy(device.dense_to_words(A)).cuda()
    t = generate.transpose_bits(bits, rows, ncols)
    assert np.array_equal(device.bits_to_host(t, rows), A.T)


def test_fit_from_device_bits_equals_fit_from_host_csr_and_cpu_restatement(M):
    """Asso.fit() on a matrix that was GENERATED on the device (never a csr, nothing uploaded) equals the fit on the same
    matrix downloaded to a host csr, and the CPU restatement's fit on that csr."""
    from oracle import asso_oracle_c as OC
    from pybmf_b200 import generate
    from pybmf_b200.digest import DIGEST_KEYS, result_digest
    m, n = 5000, 1500
    Xd = generate.planted_bits(m, n, 12, 0.08, 0.08, 0.1, 0.01, seed=77)
    Xh = Xd.to_csr()
    dens = Xh.nnz / (m * n)
    assert 0.03 < dens < 0.2, dens
    a = M.Asso(tau=0.45, k=8, w_fp=0.5)
    a.fit(Xd, **FIT_KW)
    b = M.Asso(tau=0.45, k=8, w_fp=0.5)
    b.fit(Xh, **FIT_KW)
    da, db = result_digest(a), result_digest(b)
    want = OC.asso_fit(Xh, 8, 0.45, 0.5)["digest"]
    for key in DIGEST_KEYS:
        assert da[key] == db[key] == want[key], key
    assert len(da["winners"]) == 8
