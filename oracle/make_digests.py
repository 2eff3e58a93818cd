"""TEST INFRASTRUCTURE ONLY -- config-level fixtures for BASELINE configs c2 / c3 / c4 at FULL size.

    python oracle/make_digests.py c2 c3 c2w02    # seconds (c2w02 = c2 with w_fp = 0.2: the general-weights path)
    python oracle/make_digests.py c4 [threads]   # ~1 h on 8 host cores (20 full greedy steps, 8.5e9 pairs each)

Runs the bit-packed C restatement (oracle/asso_c.c through oracle/asso_oracle_c.py, pinned against the numpy
restatement and the genuine reference's golden vectors) on the seeded synthetic inputs of pybmf_b200/synth.py
and writes tests/golden/<config>_digest.json: per greedy step the winner, the float64 score bits, #used rows,
cumulative TP / FP, and SHA-256 of the packed U / V columns -- the `result_digest` that bench.py prints at
every GPU count and that tests/test_asso_gpu.py compares Asso(k=20).fit() with.  Every step is a FULL rescoring
pass: nothing here shares code or shortcuts (incremental gains, tensor-core encodings) with the CUDA path.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import asso_oracle_c as OC  # noqa: E402
from pybmf_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def fit_digest(name, X, k, tau, w_fp, threads=None):
    t0 = time.time()

    def progress(s):
        print("[%s] step %2d winner %5d score %.1f used %d tp %d fp %d  (%.1fs scoring, %.0fs elapsed)" % (
            name, s["k"], s["winner"], s["score"], s["used"], s["tp"], s["fp"], s["score_seconds"], time.time() - t0),
            flush=True)
    r = OC.asso_fit(X, k, tau, w_fp, threads=threads, progress=progress)
    d = dict(r["digest"])
    d.update({"config": name, "m": int(X.shape[0]), "n": int(X.shape[1]), "nnz": int(X.nnz), "k": k, "tau": tau,
              "w_fp": w_fp, "candidates": r["candidates"], "sum_x": r["sum_x"], "error": r["error"],
              "rowsum": [s["rowsum"] for s in r["steps"]],
              "oracle_seconds": time.time() - t0, "oracle_assoc_seconds": r["assoc_seconds"],
              "oracle_score_seconds_per_step": float(np.mean([s["score_seconds"] for s in r["steps"]])),
              "oracle_threads": OC.lib().bmfo_threads(),
              "made_by": "oracle/make_digests.py (bit-packed C restatement, full rescoring every step)"})
    return d, r


def main():
    which = [a for a in sys.argv[1:] if not a.isdigit()] or ["c2", "c3"]
    threads = next((int(a) for a in sys.argv[1:] if a.isdigit()), None)
    os.makedirs(OUT, exist_ok=True)
    if "c2" in which or "c3" in which:
        X = synth.config_c2()
        d, r = fit_digest("c2", X, 20, 0.5, 0.5, threads)
        if "c2" in which:
            json.dump(d, open(os.path.join(OUT, "c2_digest.json"), "w"), indent=1)
        if "c3" in which:                                  # AssoIter(k=20) on top of the c2 model
            t0 = time.time()
            U, V = np.stack(r["U_cols"], 1), np.stack(r["V_cols"], 1)
            it = OC.asso_iter_fit(X, U, V, 20, 0.5)
            import hashlib
            h = hashlib.sha256()
            for c in range(it["U"].shape[1]):
                h.update(OC.column_bytes(it["U"][:, c]))
            d3 = {"config": "c3", "trace": [[int(a), int(b)] for a, b in it["trace"]],
                  "score_bits": [np.float64(s).tobytes().hex() for s in it["scores"]],
                  "error_bits": [np.float64(s).tobytes().hex() for s in it["errors"]],
                  "u_sha256": h.hexdigest(), "u_ones": int(it["U"].sum()), "oracle_seconds": time.time() - t0,
                  "made_by": "oracle/make_digests.py (AssoIter on the c2 digest's factors)"}
            json.dump(d3, open(os.path.join(OUT, "c3_digest.json"), "w"), indent=1)
            print("c3: %d column passes, %d accepted, %.1fs" % (len(it["trace"]), len(it["scores"]), time.time() - t0))
    if "c2w02" in which:                                   # general (non-dyadic) weights at the full c2 size
        X = synth.config_c2()
        d, _r = fit_digest("c2w02", X, 20, 0.5, 0.2, threads)
        json.dump(d, open(os.path.join(OUT, "c2w02_digest.json"), "w"), indent=1)
    if "c4" in which:
        X = synth.config_c4()
        d, _r = fit_digest("c4", X, 20, 0.5, 0.5, threads)
        json.dump(d, open(os.path.join(OUT, "c4_digest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
