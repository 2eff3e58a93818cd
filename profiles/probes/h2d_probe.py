"""H2D of a 408 MB pageable int32 array: one copy vs chunked copies from several host threads on separate streams."""
import time, sys
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
n = 102_000_000
src = np.random.randint(0, 17770, n).astype(np.int32)
dev = torch.device("cuda", 0)
def one():
    t = torch.from_numpy(src).to(dev, non_blocking=True); torch.cuda.synchronize(); return t
def chunked(k):
    dst = torch.empty(n, dtype=torch.int32, device=dev)
    bounds = np.linspace(0, n, k + 1, dtype=np.int64)
    streams = [torch.cuda.Stream() for _ in range(k)]
    def work(i):
        with torch.cuda.stream(streams[i]):
            dst[bounds[i]:bounds[i + 1]].copy_(torch.from_numpy(src[bounds[i]:bounds[i + 1]]), non_blocking=True)
    with ThreadPoolExecutor(k) as pool:
        list(pool.map(work, range(k)))
    torch.cuda.synchronize()
    return dst
for name, fn in (("one", one), ("chunk2", lambda: chunked(2)), ("chunk4", lambda: chunked(4)), ("chunk8", lambda: chunked(8)), ("one", one)):
    fn()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    assert int(out[12345].item()) == int(src[12345])
    print("%s: %.1f ms (%.1f GB/s)" % (name, best * 1e3, src.nbytes / best / 1e9), flush=True)
