"""FP4 cover-scoring GEMM at c4: raster group size sweep (env read per launch); min of 3 launches per setting."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from pybmf_b200 import synth
from pybmf_b200.engine import CoverEngine
X = synth.config_c4()
w = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
eng = CoverEngine(X, w, 1 - w)
eng.build_basis(0.5)
print("operand", eng.operand, flush=True)
for g in (16, 4, 8, 12, 24, 32, 16):
    os.environ["BMF_GROUP_M2"] = str(g)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); time.sleep(0.2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.score_all(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("group=%d kernel_ms=%.2f" % (g, best), flush=True)
