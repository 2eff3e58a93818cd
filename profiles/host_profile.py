"""cProfile of the host side of Asso(k=20).fit() at BASELINE config c2 (launch-bound: the GPU work is < 2 ms).
    python profiles/host_profile.py > profile.txt"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pybmf_b200 import models, synth  # noqa: E402

models.SILENT = True
X = synth.config_c2()
KW = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)
for _ in range(2):
    models.Asso(tau=0.5, k=20, w_fp=0.5).fit(X, **KW)
torch.cuda.synchronize()
t0 = time.perf_counter()
models.Asso(tau=0.5, k=20, w_fp=0.5).fit(X, **KW)
torch.cuda.synchronize()
print("fit seconds (no profiler): %.4f" % (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
mdl = models.Asso(tau=0.5, k=20, w_fp=0.5)
mdl.fit(X, **KW)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
src = mdl
t0 = time.perf_counter()
it = models.AssoIter(model=src, w_fp=0.5)
it.fit(X, **KW)
torch.cuda.synchronize()
print("AssoIter fit seconds: %.4f" % (time.perf_counter() - t0))
