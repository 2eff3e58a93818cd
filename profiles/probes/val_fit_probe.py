"""Cost of supplying X_val / X_test to Asso.fit() at config c4 (round-1 finding: every step re-uploaded the split and
materialised U).  Now the splits are packed once per fit and counted against the device-resident cover.
    python profiles/probes/val_fit_probe.py"""
import gc
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np
import scipy.sparse as sp
import torch

from pybmf_b200 import models, synth

models.SILENT = True
X = synth.config_c4()
rng = np.random.RandomState(1)
mask = rng.rand(X.nnz) < 0.1                                   # a 10 % validation split of the stored entries
Xv = sp.csr_matrix((X.data[mask], X.indices[mask], np.concatenate([[0], np.cumsum(np.add.reduceat(mask, X.indptr[:-1]))])), shape=X.shape)
kw = dict(save_model=False, show_logs=False, show_result=False)
models.Asso(tau=0.5, k=1, w_fp=0.5).fit(X, task="reconstruction", **kw)
for label, args, task in (("train only", (X,), "reconstruction"), ("train + val (same shape, 10 % of the entries)", (X, Xv), "reconstruction"),
                          ("train + val + test", (X, Xv, Xv), "reconstruction"), ("train + val, task=prediction", (X, Xv), "prediction")):
    ts = []
    for _ in range(3):
        gc.collect()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mdl = models.Asso(tau=0.5, k=20, w_fp=0.5)
        mdl.fit(*args, task=task, **kw)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        cols = [c for c in mdl.logs["updates"].columns if c[0] == "val"]
        del mdl
    print("%-48s fit(k=20) %s s   val columns logged: %d" % (label, ["%.3f" % t for t in ts], len(cols)), flush=True)
