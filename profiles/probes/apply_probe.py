"""cover_apply at the Netflix-shaped size (480189 x 17770 bits): register-fed (BMF_APPLY_RING=0) vs ring-fed kernel.
    python profiles/probes/apply_probe.py
Synthetic state: x ~ 1.2 % ones, cover ~ 0.4 %, a basis row of ~300 columns; no operand plane, no compaction (the streaming
part is what is timed: 2 x m x words x 8 bytes read).  Both forms must leave identical cover / counters / usage bits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from pybmf_b200 import _native, device

m, n = 480189, 17770
words = device.words_for(n)
g = torch.Generator(device="cuda"); g.manual_seed(3)


def rnd(shape, a):
    w = torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
    for _ in range(a - 1):
        w &= torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
    return w


def clip(t):
    if n % 64:
        t[:, n // 64] &= (1 << (n % 64)) - 1
    t[:, (n + 63) // 64:] = 0
    return t


x = clip(rnd((m, words), 6))                                     # 1.6 %
c0 = clip(rnd((m, words), 8)) & x                                # part of x already covered
basis = clip(rnd((4, words), 6))
alive0 = torch.ones((n,), dtype=torch.uint8, device="cuda")
win = torch.tensor([2], dtype=torch.int64, device="cuda")
bytes_read = 2.0 * m * words * 8
out = {}
for ring in ("0", "1"):
    os.environ["BMF_APPLY_RING"] = ring
    times = []
    for rep in range(6):
        c = c0.clone()
        tp = torch.zeros((m,), dtype=torch.int32, device="cuda")
        fp = torch.zeros((m,), dtype=torch.int32, device="cuda")
        ub = torch.zeros((device.words_for(m),), dtype=torch.int64, device="cuda")
        tot = torch.zeros((3,), dtype=torch.int64, device="cuda")
        alive = alive0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        _native.call("bmf_cover_apply", x, c, m, n, words, basis, alive, win, tp, fp, 1, 1, 0.5, 0.5, None, 128 * ((n + 127) // 128),
                     0, ub, tot)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    out[ring] = (c, tp, fp, ub, tot)
    best = min(times[1:])
    print("ring=%s  %s ms  best %.3f ms = %.2f TB/s  totals %s" % (ring, ["%.3f" % t for t in times], best, bytes_read / best / 1e9,
                                                                 tot.tolist()), flush=True)
print("identical:", all(bool(torch.equal(a, b)) for a, b in zip(out["0"], out["1"])))
