"""General-weights FP4 scorer at c4 with / without the fixed-point pre-decision (isolated launches, min of 3)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from pybmf_b200 import synth
from pybmf_b200.engine import CoverEngine
X = synth.config_c4()
eng = CoverEngine(X, 0.2, 0.8)
eng.build_basis(0.5)
ref = None
for flag in ("0", "1", "0"):
    os.environ["BMF_NO_FIXED_PREDECISION"] = flag
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); time.sleep(0.3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.score_all(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    if ref is None:
        ref = (eng.gain_p.clone(), eng.gain_n.clone())
    same = torch.equal(ref[0], eng.gain_p) and torch.equal(ref[1], eng.gain_n)
    print("no_fixed_predecision=%s kernel_ms=%.2f identical_gains=%s" % (flag, best, same), flush=True)
