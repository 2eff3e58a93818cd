"""One rank's share of the c4 scoring pass at N = 8 (60024 data rows) timed on ONE GPU for several raster groups
(BMF_GROUP_M2): is the N = 8 kernel slower per tile than the N = 1 kernel, and does the raster matter?
    python profiles/probes/shard_kernel_probe.py [rows]"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from pybmf_b200 import synth
from pybmf_b200.engine import CoverEngine

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 60024
X = synth.config_c4()
for nrows in (rows, 480189):
    Xs = X[:nrows]
    eng = CoverEngine(Xs, 0.5, 0.5, rescore="full")
    eng.build_basis(0.5)
    for g in ("8", "16", "24", "70"):
        os.environ["BMF_GROUP_M2"] = g
        eng.first_pass()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng._launch_scorer()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        tiles = eng.rows_plane.shape[0] // 496 * (eng.cand_pad // 256)
        print("rows %6d group %2s: %.3f ms (min %.3f)  %.2f us per tile round" % (
            nrows, g, sum(ts) / len(ts), min(ts), 1e3 * min(ts) / (-(-tiles // 74))), flush=True)
    del eng
    torch.cuda.empty_cache()
