"""Which phase of fit() inflates in the slow repeats of bench.py's e2e loop?  Same loop as bench.timed_fits (digest + first read
of U / V between fits, previous model released before the clock starts, GC paused while it runs) with BMF_FIT_TRACE=1."""
import gc
import os
import sys
import time

sys.path.insert(0, os.getcwd())
os.environ["BMF_FIT_TRACE"] = "1"
import torch

from pybmf_b200 import models, synth
from pybmf_b200.digest import result_digest

models.SILENT = True
X = synth.config_c4()
kw = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)
models.Asso(tau=0.5, k=1, w_fp=0.5).fit(X, **kw)
torch.cuda.empty_cache() if "--empty-cache" in sys.argv else None
import numpy as np
scratch = np.empty_like(X.indices)
mdl = None
for i in range(8):
    mdl = None
    gc.collect()
    torch.cuda.synchronize()
    th = time.perf_counter()
    scratch[:] = X.indices                                   # host-only read of the index array (400 MB)
    host_ms = 1e3 * (time.perf_counter() - th)
    th = time.perf_counter()
    z = torch.zeros((480189, 278), dtype=torch.int64, device="cuda"); torch.cuda.synchronize(); del z
    zero_ms = 1e3 * (time.perf_counter() - th)
    print("   before fit %d: host memcpy of indices %.1f ms, torch.zeros(1.07 GB) %.1f ms" % (i, host_ms, zero_ms), file=sys.stderr)
    gc.disable()
    t0 = time.perf_counter()
    mdl = models.Asso(tau=0.5, k=20, w_fp=0.5)
    mdl.fit(X, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gc.enable()
    t1 = time.perf_counter()
    if i % 2 == 1 or "--always-read" in sys.argv:
        result_digest(mdl)
        _ = mdl.U, mdl.V
    print("fit %d: %.3f s   (host work after: %.3f s; reserved %.1f GB)" % (i, dt, time.perf_counter() - t1, torch.cuda.memory_reserved() / 1e9),
          file=sys.stderr, flush=True)
