"""Cover-scoring GEMM at c4 under different L2 eviction hints / raster group sizes (env read per launch).
plain run: kernel ms per config;  under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:gemm_i8_2sm`:
DRAM bytes per launch in the same order (first launch = association, then 3 launches per config)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from pybmf_b200 import models, synth
from pybmf_b200.engine import CoverEngine
w_fp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
X = synth.config_c4()
eng = CoverEngine(X, w_fp, 1 - w_fp)
eng.build_basis(0.5)
configs = [(h, g) for g in (16, 8) for h in (0, 1, 2, 3)] + [(1, 12), (1, 24), (1, 32)]
for hint, g in configs:
    os.environ["BMF_L2_HINT"] = str(hint)
    os.environ["BMF_GROUP_M2"] = str(g)
    eng.score_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.score_all(); eng.score_all()
    e1.record()
    torch.cuda.synchronize()
    print("hint=%d group=%d kernel_ms=%.2f" % (hint, g, e0.elapsed_time(e1) / 2), flush=True)
