"""Why did bench.py's repeated e2e fits slow down at N=2 (0.19 -> 0.31 -> 0.42 s)?  Repeats Asso(k=20).fit() at c4 with and
without the host-side work bench.py does between fits (digest + first read of U / V), and with the previous model
released before / inside the timed region.   torchrun --nproc-per-node N profiles/probes/fit_repeat_probe.py"""
import gc
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist

from pybmf_b200 import models, synth
from pybmf_b200.digest import result_digest

rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
models.SILENT = True
X = synth.config_c4()
kw = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)


def one(tag, read_factors, release_before):
    global mdl
    if release_before:
        mdl = None
        gc.collect()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mdl = models.Asso(tau=0.5, k=20, w_fp=0.5)
    mdl.fit(X, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    if read_factors:
        result_digest(mdl)
        _ = mdl.U, mdl.V
    if rank == 0:
        print("%-40s fit %.3f s   host work after %.3f s" % (tag, dt, time.perf_counter() - t1), file=sys.stderr)


mdl = None
one("warm-up", False, True)
for i in range(3):
    one("plain %d" % i, False, False)
for i in range(4):
    one("read U,V after; old model freed inside %d" % i, True, False)
for i in range(4):
    one("read U,V after; old model freed before %d" % i, True, True)
if world > 1:
    dist.destroy_process_group()
