"""Worker of tests/test_multigpu_gpu.py (one process per GPU, NCCL): Asso.fit() with the rows of X sharded over the
ranks must give, on EVERY rank, the factors / logs of the golden vectors (integer and general weights) and of the
oracle on a seeded MovieLens-1M-shaped slice.  Usage: torchrun --nproc-per-node N tests/multi_gpu_worker.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from conftest import LOG_COLS, load_golden
from oracle import asso_oracle as O
from pybmf_b200 import models, synth

rank, world, local = (int(os.environ[v]) for v in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
models.SILENT = True
KW = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)


def dense(A):
    return (np.asarray(A.todense()) != 0).astype(np.uint8)


for name in ("ex01_6", "c1_noisy", "planted_w02"):
    c = load_golden(name)
    g = c["g"]
    for scorer in ("tcgen05", "tcgen05_i8", "popc"):
        mdl = models.Asso(tau=c["tau"], k=c["k"], w_fp=c["w_fp"], w_fn=c["w_fn"], scorer=scorer)
        mdl.fit(sp.csr_matrix(c["X"]), **KW)
        assert np.array_equal(dense(mdl.U), g["U"]) and np.array_equal(dense(mdl.V), g["V"]), (name, scorer, rank)
        df = mdl.logs["updates"]
        for col in ("TP", "FP", "FN"):
            assert np.array_equal(np.array([float(v) for v in df[("train", 0, col)]]), g["log_" + col]), (name, col)
        if O.integer_weights(c["w_fp"], c["w_fn"]) is not None:
            assert np.array_equal(np.array([float(v) for v in df[("train", 0, "score")]]), g["log_score"]), name

# seeded c2-shaped slice (3000 x 3706): enough rows for several 256-row shards per rank
X = synth.config_c2()[:3000]
want = O.asso_fit(X, 6, 0.5, 0.5)
mdl = models.Asso(tau=0.5, k=6, w_fp=0.5)
mdl.fit(X, **KW)
assert np.array_equal(dense(mdl.U), want["U"]) and np.array_equal(dense(mdl.V), want["V"]), rank
assert [float(v) for v in mdl.logs["updates"][("train", 0, "score")]] == [l["score"] for l in want["logs"]]
dist.barrier()
dist.destroy_process_group()
print("rank %d of %d ok" % (rank, world))
