"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the GENUINE
reference (/root/reference, imported through oracle/ref_shim.py) in the authoring
container.  The GPU box has no /root/reference, so the vectors are committed.

    python oracle/make_golden.py            # ~2 min of single-threaded reference time

Each file holds the packed input bits, the reference's U, V, the numeric columns of
logs['updates'] (PyBMF/models/Asso.py:121-132) and, where noted, the AssoIter trace.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import ref_shim  # noqa: E402
from pybmf_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
COLS = ["score", "score_0.5", "desc_len", "TP", "TPR", "FP", "FPR", "FN", "FNR", "ERR", "ACC",
        "Recall", "Precision", "F1"]


def dense01(M):
    return (np.asarray(M.todense() if hasattr(M, "todense") else M) != 0).astype(np.uint8)


def run_case(name, X, k, tau, w_fp, w_fn=None, with_iter=False, expect_error=None):
    P = ref_shim.load()
    from PyBMF.models import Asso, AssoIter
    rec = {"m": X.shape[0], "n": X.shape[1], "Xbits": np.packbits(dense01(X), axis=1),
           "k": -1 if k is None else k, "tau": tau, "w_fp": w_fp, "w_fn": np.nan if w_fn is None else w_fn}
    t0 = time.time()
    err = ""
    with ref_shim.quiet():
        model = Asso(tau=tau, k=k, w_fp=w_fp, w_fn=w_fn)
        try:
            model.fit(X, **ref_shim.FIT_KW)
        except Exception as e:  # D2
            err = type(e).__name__
    rec["error"] = err
    rec["ref_seconds"] = time.time() - t0
    rec["U"] = dense01(model.U)
    rec["V"] = dense01(model.V)
    if "updates" in model.logs:
        df = model.logs["updates"]
        for c in COLS:
            rec["log_" + c] = np.array([float(v) for v in df[("train", 0, c)]], dtype=np.float64)
        rec["log_shape"] = np.array([[int(a), int(b)] for a, b in df[("train", 0, "shape")]], dtype=np.int64)
    if with_iter and not err:
        import io, contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
            it = AssoIter(model=model, w_fp=w_fp, w_fn=w_fn)
            it.fit(X, **ref_shim.FIT_KW)
        trace = []
        for line in buf.getvalue().splitlines():
            if "Refined column" in line:
                trace.append((int(line.split("column i:")[1].split(",")[0]), 1))
            elif "Skipped column" in line:
                trace.append((int(line.split("column i:")[1].strip(" .")), 0))
        rec["iter_trace"] = np.array(trace, dtype=np.int64).reshape(-1, 2)
        rec["iter_U"] = dense01(it.U)
        if "refinements" in it.logs:
            df = it.logs["refinements"]
            rec["iter_score"] = np.array([float(v) for v in df[("train", 0, "score")]])
            rec["iter_error"] = np.array([float(v) for v in df[("train", 0, "error")]])
            for c in ["Recall", "Precision", "Accuracy", "F1"]:
                rec["iter_" + c] = np.array([float(v) for v in df[("train", 0, c)]])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-22s %.1fs  U%s err=%r steps=%d" % (name, rec["ref_seconds"], rec["U"].shape, err,
                                               len(rec.get("log_score", []))))


def block_diag(m, n, k, overlap, seed, noise=None, noise_seed=None):
    ref_shim.load()
    from PyBMF.generators import BlockDiagonalMatrixGenerator
    with ref_shim.quiet():
        g = BlockDiagonalMatrixGenerator(m=m, n=n, k=k, overlap=overlap)
        g.generate(seed=seed)
        if noise is not None:
            g.add_noise(noise, seed=noise_seed)
    return g.X


def main():
    os.makedirs(OUT, exist_ok=True)
    # the one known-answer table the reference ships: examples/ex01_6_logs.ipynb:81,399-400
    X = block_diag(300, 500, 5, [0.3, 0.2], 1000, [0.4, 0.1], 2000)
    run_case("ex01_6", X, k=5, tau=0.25, w_fp=0.5, with_iter=True)
    # BASELINE.json configs[0], variants A (noise-free -> quirk D1) and B
    X = block_diag(1000, 500, 5, [0.2, 0.1], 1000)
    run_case("c1_clean", X, k=5, tau=0.5, w_fp=0.5)
    X = block_diag(1000, 500, 5, [0.2, 0.1], 1000, [0.2, 0.02], 2000)
    run_case("c1_noisy", X, k=5, tau=0.5, w_fp=0.5, with_iter=True)
    # general (non-dyadic) weights and asymmetric dyadic weights on a small planted matrix
    X = synth.planted(240, 180, 6, 0.2, 0.2, 0.1, 0.02, seed=7)
    run_case("planted_w02", X, k=4, tau=0.15, w_fp=0.2, with_iter=True)
    run_case("planted_w025", X, k=4, tau=0.3, w_fp=0.25, with_iter=True)
    run_case("planted_w37", X, k=3, tau=0.4, w_fp=0.3, w_fn=0.6)
    # D2: more factors requested than improving steps exist -> TypeError
    X = block_diag(120, 90, 3, [0.0, 0.0], 5)
    run_case("d2_no_pattern", X, k=6, tau=0.5, w_fp=0.5)


if __name__ == "__main__":
    main()
