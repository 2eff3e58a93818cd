"""FP4 cover-scoring GEMM at c4: raster group size sweep (env read per launch); min of 3 launches per setting."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from pybmf_b200 import synth
from pybmf_b200.engine import CoverEngine
X = synth.config_c4()
eng = CoverEngine(X, 0.5, 0.5)
eng.build_basis(0.5)
print("operand", eng.operand, flush=True)
for g in (16, 8, 12, 24, 32, 48, 16):
    os.environ["BMF_GROUP_M2"] = str(g)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); time.sleep(0.2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.score_all(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("group=%d kernel_ms=%.2f" % (g, best), flush=True)
