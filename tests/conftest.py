import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    import numpy as np
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = int(g["n"])
    X = np.unpackbits(g["Xbits"], axis=1)[:, :n]
    k = int(g["k"])
    return {
        "g": g, "X": X, "k": None if k < 0 else k, "tau": float(g["tau"]), "w_fp": float(g["w_fp"]),
        "w_fn": None if g["w_fn"] != g["w_fn"] else float(g["w_fn"]), "error": str(g["error"]),
    }


GOLDEN_CASES = ["ex01_6", "c1_clean", "c1_noisy", "planted_w02", "planted_w025", "planted_w37",
                "d2_no_pattern"]
LOG_COLS = ["score", "score_0.5", "desc_len", "TP", "TPR", "FP", "FPR", "FN", "FNR", "ERR", "ACC",
            "Recall", "Precision", "F1"]
