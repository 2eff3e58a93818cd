"""CPU: pin the oracle restatement (oracle/asso_oracle.py) against
 (a) the known-answer table stored in /root/reference/examples/ex01_6_logs.ipynb:253-361,
 (b) outputs of the genuine reference (tests/golden/*.npz, made by oracle/make_golden.py),
 (c) the live reference when /root/reference is present (authoring container only)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, LOG_COLS, load_golden
from oracle import asso_oracle as O


def _run_oracle(c):
    try:
        return O.asso_fit(c["X"], c["k"], c["tau"], c["w_fp"], c["w_fn"]), ""
    except O.NoCandidateError as e:
        return e.args[1], "TypeError"


def test_known_answer_table_ex01_6():
    # numbers typed from the notebook's stored output, not from our fixture
    c = load_golden("ex01_6")
    assert int(c["X"].sum()) == 39251                        # ex01_6_logs.ipynb:100
    r, _ = _run_oracle(c)
    logs = r["logs"]
    assert [l["score"] for l in logs] == [817.5, 1564.5, 2182.5, 2797.5, 2953.0]
    assert [l["shape"] for l in logs] == [[71, 151], [63, 152], [24, 226], [24, 223], [23, 189]]
    assert [l["TP"] for l in logs] == [6178, 11713, 15043, 18334, 18938]
    assert [l["FP"] for l in logs] == [4543, 8584, 10678, 12739, 13032]
    assert logs[0]["desc_len"] == 37838.0 and logs[-1]["desc_len"] == 34491.0


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_outputs(name):
    c = load_golden(name)
    g = c["g"]
    r, err = _run_oracle(c)
    assert err == c["error"]
    assert r["U"].shape == g["U"].shape and np.array_equal(r["U"], g["U"])
    assert r["V"].shape == g["V"].shape and np.array_equal(r["V"], g["V"])
    for col in LOG_COLS:                                       # bit-exact, including the fp64 rates
        got = np.array([l[col] for l in r["logs"]], dtype=np.float64)
        assert np.array_equal(got, g["log_" + col]), col
    assert np.array_equal(np.array([l["shape"] for l in r["logs"]]).reshape(-1, 2), g["log_shape"])


@pytest.mark.parametrize("name", ["ex01_6", "c1_noisy", "planted_w02", "planted_w025"])
def test_oracle_assoiter_matches_reference(name):
    c = load_golden(name)
    g = c["g"]
    r, _ = _run_oracle(c)
    it = O.asso_iter_fit(c["X"], r["U"], r["V"], c["k"], c["w_fp"], c["w_fn"])
    trace = np.array([(col, int(acc)) for col, acc in it["trace"]], dtype=np.int64).reshape(-1, 2)
    assert np.array_equal(trace, g["iter_trace"])
    assert np.array_equal(it["U"], g["iter_U"])
    if "iter_score" in g.files:
        assert np.array_equal(np.array([x["score"] for x in it["refinements"]]), g["iter_score"])
        assert np.array_equal(np.array([x["error"] for x in it["refinements"]]), g["iter_error"])
        for col in ["Recall", "Precision", "Accuracy", "F1"]:
            assert np.array_equal(np.array([x[col] for x in it["refinements"]]), g["iter_" + col])
    else:
        assert it["refinements"] == []


def test_d1_truncation_quirk():
    c = load_golden("c1_clean")
    r, _ = _run_oracle(c)
    assert r["U"].shape == (1000, 4) and len(r["logs"]) == 5 and r["logs"][-1]["ERR"] == 0.0


def test_integer_weight_form_equals_float_form():
    rng = np.random.RandomState(3)
    X = (rng.rand(70, 50) < 0.3).astype(np.uint8)
    C = (rng.rand(70, 50) < 0.1).astype(np.uint8)
    B = (rng.rand(20, 50) < 0.3).astype(np.uint8)
    for w_fp, w_fn in [(0.5, None), (0.25, None), (0.375, 0.5), (1.0, 1.0)]:
        a, b, s = O.integer_weights(w_fp, w_fn)
        score, use, P, N, tpo, fpo = O.score_candidates(X, C, B, w_fp, w_fn)
        assert np.array_equal(use, (b * P - a * N) > 0)
        G = O.integer_gains(X, C, B, a, b)
        so = int((b * tpo - a * fpo).sum())
        assert np.array_equal(score, (so + G).astype(np.float64) / (1 << s))
    assert O.integer_weights(0.2, None) is None


def test_live_reference_small():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_shim.load()
    from PyBMF.models import Asso
    from pybmf_b200 import synth
    X = synth.planted(90, 70, 4, 0.25, 0.25, 0.1, 0.02, seed=11)
    with ref_shim.quiet():
        mdl = Asso(tau=0.3, k=3, w_fp=0.4)
        mdl.fit(X, **ref_shim.FIT_KW)
    r = O.asso_fit(X, 3, 0.3, 0.4)
    assert np.array_equal(r["U"], (mdl.U.toarray() != 0)) and np.array_equal(r["V"], (mdl.V.toarray() != 0))
    assert [l["score"] for l in r["logs"]] == [float(v) for v in mdl.logs["updates"][("train", 0, "score")]]
