"""Profiling target: one Boolean product + one fused confusion pass at the largest c5 point (1M x 100k, k = 64).
    python profiles/prof_c5.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pybmf_b200 import _native, device

m, n, k = 1_000_000, 100_000, 64
words = device.words_for(n)
g = torch.Generator(device="cuda"); g.manual_seed(5)


def rnd(shape, ands):
    w = torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
    for _ in range(ands - 1):
        w &= torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
    return w


uw = rnd((m, 1), 5)
vt = rnd((k, words), 5)
vt[:, n // 64] &= (1 << (n % 64)) - 1
vt[:, (n + 63) // 64:] = 0
pd = device.zeros((m, words), torch.int64)
counts = device.zeros((3,), torch.int64)
for _ in range(2):
    _native.call("bmf_bool_product", uw, m, 1, vt, k, words, pd)
    _native.call("bmf_confusion_factors", pd, m, words, uw, 1, vt, k, 1, counts, None, None)
torch.cuda.synchronize()
print("ok", counts.tolist())
