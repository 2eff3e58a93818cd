import sys, time, cProfile, pstats, os
sys.path.insert(0, os.getcwd())
import torch
from pybmf_b200 import models, synth
models.SILENT = True
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
X = synth.config_c4() if wl == "c4" else synth.config_c2()
kw = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)
models.Asso(tau=0.5, k=2, w_fp=0.5).fit(X, **kw)   # warm
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
mdl = models.Asso(tau=0.5, k=int(sys.argv[2]) if len(sys.argv) > 2 else 5, w_fp=0.5)
mdl.fit(X, **kw)
torch.cuda.synchronize()
pr.disable()
print("fit seconds", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
