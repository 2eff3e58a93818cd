"""Profiling target: ONE Asso(k).fit() at a BASELINE config through the public API (for ncu launch lists / --set full).
    python profiles/prof_fit.py [c4|c2] [k] [rescore]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pybmf_b200 import models, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rescore = sys.argv[3] if len(sys.argv) > 3 else "auto"
models.SILENT = True
X = synth.config_c4() if cfg == "c4" else synth.config_c2()
mdl = models.Asso(tau=0.5, k=k, w_fp=0.5, rescore=rescore)
mdl.fit(X, task="reconstruction", save_model=False, show_logs=False, show_result=False)
print("fit ok:", cfg, k, rescore, [s["winner"] for s in mdl.fit_steps_], "launches", mdl._dev_launches)
