"""CPU-only checks: the C-ABI library builds, loads and exports every symbol the header declares;
host-side logic (integer weight detection, row shard plan, metric formulas, log bookkeeping);
and the multi-rank algebra under torch.distributed/gloo with world_size 2 (no GPU involved)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    from pybmf_b200 import _native
    from pybmf_b200 import build as B
    lib_path = B.build()
    assert os.path.exists(lib_path)
    lib = _native.load()
    assert lib.bmf_abi_version() == 1
    header = open(os.path.join(ROOT, "include", "pybmf_b200.h")).read()
    declared = set(re.findall(r"\b(bmf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 19
    raw = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(raw, name), "header declares %s but the library does not export it" % name
    assert declared - {"bmf_last_error"} == set(_native.SIGNATURES), "ctypes table and header disagree"


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    from pybmf_b200 import build as B
    sass = subprocess.run(["cuobjdump", "-sass", B.build()], capture_output=True, text=True).stdout
    assert "UTCIMMA" in sass, "tcgen05.mma kind::i8 missing from SASS"
    assert "UTCOMMA" in sass, "tcgen05.mma kind::mxf4 (block-scaled FP4) missing from SASS"
    assert "UTMALDG" in sass and "LDTM" in sass and "STTM" in sass and "UBLKCP" in sass   # TMA tiles, TMEM ld/st, bulk copy
    # the four MMAs of a pipeline stage must be issued back to back with uniform-register operands: an ELECT + R2UR loop
    # between them (what `if (lane == 0)` issue produces) made the issuing thread the bottleneck (profiles/r01c_fp4_issue_rate.md)
    lines = [ln for ln in sass.splitlines() if "/*" in ln and ";" in ln]
    ops = [ln.split("*/")[1].strip().split()[0] if "*/" in ln else "" for ln in lines]
    runs = best = 0
    for op in ops:
        runs = runs + 1 if op.startswith("UTCOMMA") else 0
        best = max(best, runs)
    assert best >= 4, "FP4 MMAs are no longer issued back to back (found runs of %d)" % best


def test_no_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pybmf_b200 import _native, models, utils
    with pytest.raises(_native.NativeError):
        _native.require_gpu()
    with pytest.raises(_native.NativeError):
        utils.matmul(np.eye(3, dtype=int), np.eye(3, dtype=int), boolean=True)
    models.SILENT = True
    with pytest.raises(_native.NativeError):
        models.Asso(tau=0.5, k=2).fit(np.eye(4, dtype=int), task="reconstruction", save_model=False,
                                      show_logs=False, show_result=False)


def test_product_path_never_imports_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "pybmf_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "pybmf_b200", fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn + " must not reference oracle/"


def test_integer_weights_agree_with_oracle():
    from oracle import asso_oracle as O
    from pybmf_b200.engine import integer_weights
    for w in [(0.5, 0.5), (0.25, 0.75), (0.375, 0.5), (1.0, 1.0), (0.2, 0.8), (0.3, 0.6), (0.0, 1.0), (2.0, 3.0),
              (0.5, 0.0078125), (200.0, 1.0), (0.1, 0.9)]:
        assert integer_weights(*w) == O.integer_weights(*w) or (w == (0.0, 1.0)), w
    assert integer_weights(0.5, 0.5) == (1, 1, 1)
    assert integer_weights(0.25, 0.75) == (1, 3, 2)
    assert integer_weights(0.2, 0.8) is None and integer_weights(200.0, 1.0) is None


def test_csr_row_views_and_stored_zero_scan():
    """host side of the row sharding: zero-copy row ranges and the (threaded) stored-zero scan"""
    import numpy as np
    import scipy.sparse as sp
    from pybmf_b200 import device
    X = sp.random(700, 300, density=0.1, format="csr", random_state=3)
    X.data[:] = 1
    for r0, r1 in ((0, 700), (0, 256), (256, 512), (512, 700), (700, 700), (5, 5)):
        V = device.csr_rows_view(X, r0, r1)
        assert V.shape == (r1 - r0, 300) and (V.toarray() == X[r0:r1].toarray()).all()
        if r1 > r0 and (r0, r1) != (0, 700):
            assert np.shares_memory(V.indices, X.indices)            # a view, not a copy
    assert not device.has_stored_zeros(X)
    big = sp.csr_matrix((np.ones(1 << 23, dtype=np.int64), np.zeros(1 << 23, dtype=np.int32), np.array([0, 1 << 23])),
                        shape=(1, 1))
    assert not device.has_stored_zeros(big)                          # threaded path
    big.data[(1 << 23) - 7] = 0
    assert device.has_stored_zeros(big)
    Y = X.copy()
    Y.data[11] = 0
    assert device.has_stored_zeros(Y) and device.drop_stored_zeros(Y).nnz == X.nnz - 1


def test_e2m1_codes_gate_the_fp4_path():
    """bmf_e2m1_code is a host function: which small integers the FP4 operand planes can hold exactly"""
    from pybmf_b200 import _native
    lib = _native.load()
    want = {0: 0, 1: 2, 2: 4, 3: 5, 4: 6, 6: 7}
    for v in range(-2, 130):
        assert lib.bmf_e2m1_code(v) == want.get(v, -1), v
    # decode the E2M1 bit patterns (sign 1, exponent 2, mantissa 1) to check the table itself
    def e2m1(code):
        e, m = (code >> 1) & 3, code & 1
        return (0.5 * m) if e == 0 else (1 + 0.5 * m) * 2 ** (e - 1)
    assert all(e2m1(c) == v for v, c in want.items())


def test_shard_plan():
    from pybmf_b200.engine import ROW_ALIGN, ShardPlan
    for m, world in [(480189, 8), (480189, 4), (480189, 2), (480189, 1), (1000, 8), (6040, 4), (3, 2), (256, 2)]:
        plan = ShardPlan(m, world)
        covered = 0
        for r in range(world):
            a, b = plan.rows(r)
            assert a == covered and b >= a and (a % ROW_ALIGN == 0 or a == m)
            covered = b
        assert covered == m
    assert ShardPlan(480189, 8).rows(0) == (0, 60160)


def test_rates_follow_reference_formulas():
    from oracle import asso_oracle as O
    from pybmf_b200 import utils
    rng = np.random.RandomState(0)
    for _ in range(50):
        tp, fp, fn = (int(v) for v in rng.randint(0, 1000, 3))
        size = tp + fp + fn + int(rng.randint(0, 5000))
        a, b = utils.rates(tp, fp, fn, size), O.rates_from_counts(tp, fp, fn, size)
        for k in b:
            assert float(a[k]) == float(b[k]), k
    z = utils.rates(0, 0, 0, 10)
    assert z["TPR"] == 0 and z["PPV"] == 0 and z["F1"] == 0 and z["ACC"] == 1.0
    # the reference's FPR is 1 - TNR, not FP / negatives (last-ulp difference)
    r = utils.rates(6178, 4543, 33073, 150000)
    assert r["FPR"] == 1 - r["TNR"]


def test_log_bookkeeping_matches_reference_layout():
    from pybmf_b200 import utils
    logs = {}
    cols = utils.header(["k"], levels=3) + [("train", 0, "score"), ("train", 0, "TP")]
    utils.record(logs, "updates", cols, [0, 1.5, np.array(7)])
    utils.record(logs, "updates", cols, [1, 2.5, np.array(9)])
    df = logs["updates"]
    assert list(df.columns) == [("", "", "time"), ("", "", "k"), ("train", 0, "score"), ("train", 0, "TP")]
    assert list(df[("train", 0, "score")]) == [1.5, 2.5] and len(df) == 2
    assert utils.header(["time", "k"], levels=3, depth=2) == [("", "time", ""), ("", "k", "")]


def test_bit_word_helpers_roundtrip():
    from pybmf_b200 import device
    rng = np.random.RandomState(1)
    for rows, cols in [(1, 1), (5, 64), (7, 65), (3, 200)]:
        A = (rng.rand(rows, cols) < 0.5).astype(np.uint8)
        W = device.dense_to_words(A)
        assert W.shape == (rows, device.words_for(cols)) and W.shape[1] % 2 == 0
        assert np.array_equal(device.words_to_dense(W, cols), A)
        assert (W.view(np.uint64)[0, 0] & np.uint64(1)) == A[0, 0]      # bit 0 of word 0 = column 0


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from oracle import asso_oracle as O
from pybmf_b200.engine import ShardPlan, all_reduce_sum, dist_ctx
from pybmf_b200 import synth
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist_ctx()
X = O.as_dense01(synth.planted(700, 90, 4, 0.25, 0.25, 0.1, 0.02, seed=9))
plan = ShardPlan(X.shape[0], world)
r0, r1 = plan.rows(rank)
Xl = X[r0:r1]
# association: per-shard partial counts, one integer all-reduce
cnt = torch.from_numpy(O.assoc_counts(Xl))
all_reduce_sum(cnt)
assert np.array_equal(cnt.numpy(), O.assoc_counts(X))
B, _ = O.build_basis(O.build_assoc(X), 0.4)
C = np.zeros_like(X)
U = np.zeros((X.shape[0], 0), np.uint8); V = np.zeros((X.shape[1], 0), np.uint8)
best = 0.0
for step in range(3):
    G = torch.from_numpy(O.integer_gains(Xl, C[r0:r1], B, 1, 1))
    all_reduce_sum(G)                                   # the ONE exchange of a greedy step
    G_all = O.integer_gains(X, C, B, 1, 1)
    assert np.array_equal(G.numpy(), G_all)             # integer sums are order independent: exact
    tpo, fpo, _ = O.confusion(X, C)
    score = (int(tpo) - int(fpo) + G.numpy()) * 0.5
    j = int(np.argmax(score))                           # numpy argmax = first maximum = lowest index
    assert score[j] > best
    best = score[j]
    _, u_loc = O.get_vector(Xl, C[r0:r1], B[j], 0.5, None)
    parts = [None, None]
    dist.all_gather_object(parts, u_loc)
    u = np.concatenate(parts)
    _, u_ref = O.get_vector(X, C, B[j], 0.5, None)
    assert np.array_equal(u, u_ref)
    C = C | (u[:, None].astype(np.uint8) & B[j][None, :])
    U = np.hstack([U, u[:, None].astype(np.uint8)]); V = np.hstack([V, B[j][:, None]])
    B = np.delete(B, j, axis=0)
ref = O.asso_fit(X, 3, 0.4, 0.5)
assert np.array_equal(U, ref["U"]) and np.array_equal(V, ref["V"])

# round 2: (a) the association exchange -- partial counts -> row blocks -> every rank thresholds ITS block -> all-gather
# (gloo has no reduce-scatter: the block is taken from the all-reduced sum; NCCL does reduce_scatter_tensor)
n = X.shape[1]
per = -(-n // world)
cnt_blk = cnt.numpy()[rank * per:min(n, (rank + 1) * per)]
s_blk = np.diag(cnt.numpy())[rank * per:min(n, (rank + 1) * per)].astype(np.float64)
rows_blk = np.zeros((per, n), np.uint8)
nzr = s_blk > 0
rows_blk[:len(s_blk)][nzr] = (cnt_blk[nzr].astype(np.float64) / s_blk[nzr][:, None] > 0.4)
gathered = [torch.zeros((per, n), dtype=torch.uint8) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(rows_blk))
B_all = torch.cat(gathered).numpy()[:n]
B_ref = (O.build_assoc(X) > 0.4).astype(np.uint8)
assert np.array_equal(B_all, B_ref)

# (b) the product's own exchange routine on CPU tensors: gains and the step's three counters travel in ONE buffer
from pybmf_b200.engine import CoverEngine
eng = CoverEngine.__new__(CoverEngine)
cp = 8
eng.world = world
eng.gbuf = torch.arange(2 * cp + 8, dtype=torch.int64) * (rank + 1)
eng.gred = torch.zeros(2 * cp + 8, dtype=torch.int64)
eng.tail, eng.tail_red = eng.gbuf[cp:cp + 8], eng.gred[cp:cp + 8]
for red_len in (cp + 8, 2 * cp + 8):                       # integer mode / general mode
    eng.red_len = red_len
    eng.gred.zero_()
    eng._reduce_gains()
    want = torch.arange(2 * cp + 8, dtype=torch.int64) * sum(r + 1 for r in range(world))
    assert torch.equal(eng.gred[:red_len], want[:red_len]) and int(eng.gred[red_len:].abs().sum()) == 0
    assert torch.equal(eng.gbuf, torch.arange(2 * cp + 8, dtype=torch.int64) * (rank + 1))      # local sums untouched
eng.gred.zero_()
eng._reduce_gains(tail_only=True)
assert torch.equal(eng.tail_red, want[cp:cp + 8]) and int(eng.gred[:cp].abs().sum()) == 0
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_row_sharded_greedy_world2_gloo(tmp_path):
    """world_size 2 over gloo: sharded partial gains + one all-reduce per step reproduce the
    unsharded Asso (the arithmetic per shard is the oracle's; the plumbing is the product's)."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_install_as_pybmf_aliases_the_reference_module_paths():
    """same module paths as the reference -- `from PyBMF.models import Asso` resolves to this package."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); import pybmf_b200; pybmf_b200.install_as_pybmf();"
            "from PyBMF.models import Asso, AssoIter, AssoOpt, TransposedModel;"
            "from PyBMF.utils import matmul, add, multiply, get_prediction, TP, FP, TN, FN, coverage_score, description_length;"
            "import pybmf_b200.models as m; assert Asso is m.Asso and AssoIter is m.AssoIter; print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-2000:]


def test_stored_zero_scan_sees_explicit_false_in_bool_matrices():
    """A bool csr can STORE False entries; the pattern kernels would treat them as ones (advisor finding, round 1)."""
    import scipy.sparse as sp
    from pybmf_b200 import device
    X = sp.csr_matrix((np.array([True, False, True]), (np.array([0, 0, 1]), np.array([0, 2, 1]))), shape=(2, 3))
    assert X.nnz == 3 and device.has_stored_zeros(X)
    clean = device.drop_stored_zeros(X)
    assert clean.nnz == 2 and not device.has_stored_zeros(clean)
    assert not device.has_stored_zeros(sp.csr_matrix(np.eye(3, dtype=bool)))
