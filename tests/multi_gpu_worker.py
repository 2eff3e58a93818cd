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


import json

# golden vectors of the genuine reference; c1_clean exercises the truncation quirk D1 (rollback + collective in
# reset_cover on every rank), d2_no_pattern has 120 rows, so with >= 2 ranks at least one rank owns NO rows
for name in ("ex01_6", "c1_noisy", "c1_clean", "planted_w02", "d2_no_pattern"):
    c = load_golden(name)
    g = c["g"]
    for scorer, rescore in (("tcgen05", "auto"), ("tcgen05", "full"), ("tcgen05_i8", "auto"), ("popc", "full")):
        mdl = models.Asso(tau=c["tau"], k=c["k"], w_fp=c["w_fp"], w_fn=c["w_fn"], scorer=scorer, rescore=rescore)
        err = ""
        try:
            mdl.fit(sp.csr_matrix(c["X"]), **KW)
        except TypeError:
            err = "TypeError"
        assert err == c["error"], (name, scorer, err)
        assert np.array_equal(dense(mdl.U), g["U"]) and np.array_equal(dense(mdl.V), g["V"]), (name, scorer, rank)
        if "log_TP" not in g:
            continue
        df = mdl.logs["updates"]
        for col in ("TP", "FP", "FN"):
            assert np.array_equal(np.array([float(v) for v in df[("train", 0, col)]]), g["log_" + col]), (name, col)
        if O.integer_weights(c["w_fp"], c["w_fn"]) is not None:
            assert np.array_equal(np.array([float(v) for v in df[("train", 0, "score")]]), g["log_score"]), name

# D1 in the middle of a fit (rollback of speculative steps + cover rebuild on every rank), then D2
c = load_golden("c1_noisy")
try:
    O.asso_fit(c["X"], 5, 0.5, 0.5, tol=0.12)
    raise SystemExit("the oracle should have raised")
except O.NoCandidateError as e:
    want_state = e.args[1]
for rescore in ("auto", "full"):
    mdl = models.Asso(tau=0.5, k=5, tol=0.12, w_fp=0.5, rescore=rescore)
    try:
        mdl.fit(sp.csr_matrix(c["X"]), **KW)
        raise SystemExit("fit should have raised TypeError")
    except TypeError:
        pass
    assert np.array_equal(dense(mdl.U), want_state["U"]) and np.array_equal(dense(mdl.V), want_state["V"]), (rescore, rank)
    assert [float(v) for v in mdl.logs["updates"][("train", 0, "score")]] == [l["score"] for l in want_state["logs"]]

# BASELINE configs[1] at full size and rank against the CPU restatement's fixture
from pybmf_b200.digest import DIGEST_KEYS, result_digest
with open(os.path.join(ROOT, "tests", "golden", "c2_digest.json")) as fh:
    want = json.load(fh)
X = synth.config_c2()
for rescore in ("auto", "full"):
    mdl = models.Asso(tau=0.5, k=20, w_fp=0.5, rescore=rescore)
    mdl.fit(X, **KW)
    got = result_digest(mdl)
    for key in DIGEST_KEYS:
        assert got[key] == want[key], (key, rescore, rank)
# an input generated ON the devices (every rank only its own rows) against the same matrix fitted from a host csr
from pybmf_b200 import generate
Xd = generate.planted_bits(5000, 1500, 12, 0.08, 0.08, 0.1, 0.01, seed=77)                       # this rank's rows
Xfull = generate.planted_bits(5000, 1500, 12, 0.08, 0.08, 0.1, 0.01, seed=77, rank=0, world=1).to_csr()
a = models.Asso(tau=0.45, k=6, w_fp=0.5)
a.fit(Xd, **KW)
b = models.Asso(tau=0.45, k=6, w_fp=0.5)
b.fit(Xfull, **KW)
da, db = result_digest(a), result_digest(b)
for key in DIGEST_KEYS:
    assert da[key] == db[key], (key, rank)
dist.barrier()
dist.destroy_process_group()
print("rank %d of %d ok" % (rank, world))
