"""How should the 400 MB csr index array of config c4 reach the GPU?  Times, on the box:
  (a) pageable tensor.to(device)                        (what fit() does today, per chunk, overlapped with X^T X)
  (b) cudaHostRegister in place + async copy + unregister
  (c) threaded memcpy into a (cached) pinned staging buffer + async copy
    python profiles/probes/h2d_pinned_probe.py"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

n = 100_000_000
a = np.random.randint(0, 17770, size=n, dtype=np.int32)
torch.cuda.init()
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
rt = torch.cuda.cudart()


def t(fn, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        out.append(time.perf_counter() - t0)
    return ["%.1f ms" % (1e3 * v) for v in out]


src = torch.from_numpy(a)
print("(a) pageable .to(device)            ", t(lambda: src.to(dev)))
dst = torch.empty(n, dtype=torch.int32, device=dev)


def reg_copy():
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    t1 = time.perf_counter()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rt.cudaHostUnregister(a.ctypes.data)
    t3 = time.perf_counter()
    reg_copy.parts = (int(rc), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2))


print("(b) register + copy + unregister    ", t(reg_copy), "rc, register / copy / unregister ms:", reg_copy.parts)
t0 = time.perf_counter()
stage = torch.empty(n, dtype=torch.int32, pin_memory=True)
print("    pinned staging buffer allocation: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
stage_np = stage.numpy()
for threads in (1, 4, 8, 16):
    pool = ThreadPoolExecutor(threads)
    bounds = np.linspace(0, n, threads + 1, dtype=np.int64)

    def staged():
        t0 = time.perf_counter()
        list(pool.map(lambda ab: np.copyto(stage_np[ab[0]:ab[1]], a[ab[0]:ab[1]]), zip(bounds[:-1], bounds[1:])))
        t1 = time.perf_counter()
        dst.copy_(stage, non_blocking=True)
        torch.cuda.synchronize()
        staged.parts = (1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1))
    print("(c) %2d-thread memcpy to pinned + copy" % threads, t(staged), "memcpy / H2D ms: %.1f / %.1f" % staged.parts)
