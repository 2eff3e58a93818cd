"""Wall-clock phases of Asso.fit() at c4 (BMF_FIT_TRACE=1), 1 rank or under torchrun.
    python profiles/fit_trace.py [k]      /      torchrun --nproc-per-node N profiles/fit_trace.py [k]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist
from pybmf_b200 import models, synth

rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
models.SILENT = True
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
X = synth.config_c4()
kw = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)
for trace in ("0", "0", "1", "1", "0", "0"):
    os.environ["BMF_FIT_TRACE"] = trace
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    models.Asso(tau=0.5, k=k, w_fp=0.5).fit(X, **kw)
    torch.cuda.synchronize()
    if rank == 0:
        print("fit(k=%d) world=%d trace=%s: %.3f s  (torch reserved %.1f GB, allocated %.1f GB)" % (
            k, world, trace, time.perf_counter() - t0, torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_allocated() / 1e9),
            file=sys.stderr)
if world > 1:
    dist.destroy_process_group()
