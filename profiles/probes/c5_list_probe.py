"""c5 largest point (1M x 100k, k = 64): bit-scan vs selection-list forms of the two panel kernels, per count mode.
    python profiles/probes/c5_list_probe.py [usage_ands]      (usage density 2^-ands per bit; default 5 = 2 factors / row)
Prints ms and algorithmic TB/s (m * n / 8 bytes per launch) per variant; every variant must return the same counts / bits."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from pybmf_b200 import _native, device

ands = int(sys.argv[1]) if len(sys.argv) > 1 else 5
m, n, k = 1_000_000, 100_000, 64
words = device.words_for(n)
g = torch.Generator(device="cuda"); g.manual_seed(5)


def rnd(shape, a):
    w = torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
    for _ in range(a - 1):
        w &= torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
    return w


uw = rnd((m, 1), ands)
vt = rnd((k, words), 5)
vt[:, n // 64] &= (1 << (n % 64)) - 1
vt[:, (n + 63) // 64:] = 0
pd = device.zeros((m, words), torch.int64)
os.environ["BMF_PANEL_LIST"] = "0"
_native.call("bmf_bool_product", uw, m, 1, vt, k, words, pd)
x = pd.clone()
for c0 in range(0, m, 65536):
    x[c0:c0 + 65536] ^= rnd(x[c0:c0 + 65536].shape, 4)
x[:, n // 64] &= (1 << (n % 64)) - 1
x[:, (n + 63) // 64:] = 0
ref = device.zeros((3,), torch.int64)
_native.call("bmf_confusion_bits", x, pd, m, words, -1, ref, None, None)
ref = ref.tolist()
ones = ref[0] + ref[2]
bytes_one = m * words * 8.0
reps = 10


def timed(fn):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("usage density 2^-%d per bit, reference counts %s" % (ands, ref), flush=True)
pd2 = device.zeros((m, words), torch.int64)
for lst, dyn in (("0", "1"), ("1", "0"), ("1", "1")):
    os.environ["BMF_PANEL_LIST"] = lst
    os.environ["BMF_PANEL_DYNAMIC"] = dyn
    pd2.fill_(-1)
    ms = timed(lambda: _native.call("bmf_bool_product", uw, m, 1, vt, k, words, pd2))
    print("product   list=%s dynamic=%s %.3f ms  %.2f TB/s  equal=%s" % (lst, dyn, ms, bytes_one / ms / 1e9, bool(torch.equal(pd, pd2))), flush=True)
    for mode in ("0", "1", "2", "3"):
        if (lst == "0" and mode == "3") or (lst, dyn) == ("1", "0"):
            continue
        os.environ["BMF_CONFUSION_COUNT"] = mode
        for known in (ones, -1):
            c = device.zeros((3,), torch.int64)
            ms = timed(lambda: _native.call("bmf_confusion_factors", x, m, words, uw, 1, vt, k, known, c, None, None))
            print("confusion list=%s mode=%s gt=%-5s %.3f ms  %.2f TB/s  ok=%s" % (
                lst, mode, "known" if known >= 0 else "count", ms, bytes_one / ms / 1e9, c.tolist() == ref), flush=True)
