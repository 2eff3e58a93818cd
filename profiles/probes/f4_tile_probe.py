"""Timing probe (wrong results by construction): MMA-only rate of the FP4 pair kernel for different UMMA N."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from pybmf_b200 import synth
from pybmf_b200.engine import CoverEngine
X = synth.config_c4(rows=(0, 453840))
eng = CoverEngine(X, 0.5, 0.5, assoc="tcgen05_i8")
eng.build_basis(0.5)
best = 1e9
for _ in range(3):
    torch.cuda.synchronize(); time.sleep(0.2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.score_all(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("kernel_ms=%.3f rows_pad=%d" % (best, eng.rows_plane.shape[0]), flush=True)
