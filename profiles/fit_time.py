"""Wall-clock of Asso(k).fit() through the public API at a BASELINE config, per rescoring mode (run on the GPU box):
    python profiles/fit_time.py [c4|c2] [k] [repeats]
Prints one JSON line per mode: seconds of every repeat, the result digest's agreement with the fixture, launches."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pybmf_b200 import models, synth  # noqa: E402
from pybmf_b200.digest import digest_matches, result_digest  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
models.SILENT = True
X = synth.config_c4() if cfg == "c4" else synth.config_c2()
KW = dict(task="reconstruction", save_model=False, show_logs=False, show_result=False)
want = None
path = os.path.join(ROOT, "tests", "golden", cfg + "_digest.json")
if os.path.exists(path):
    want = json.load(open(path))
models.Asso(tau=0.5, k=1, w_fp=0.5).fit(X, **KW)                # warm-up (first-use costs)
for mode in ("auto", "full"):
    secs, plus = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mdl = models.Asso(tau=0.5, k=k, w_fp=0.5, rescore=mode)
        mdl.fit(X, **KW)
        torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
        d = result_digest(mdl)
        t1 = time.perf_counter()
        _ = mdl.U, mdl.V
        plus.append(secs[-1] + time.perf_counter() - t1)
    ok = None if want is None else digest_matches(d, want, steps=None if k == want["k"] else k)
    print(json.dumps({"config": cfg, "k": k, "rescore": mode, "fit_seconds": secs, "fit_plus_factors_seconds": plus,
                      "digest_ok": ok, "launches": mdl._dev_launches, "winners": d["winners"]}), flush=True)
