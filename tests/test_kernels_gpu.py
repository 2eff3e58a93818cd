"""GPU: every C-ABI entry point of libbmf_b200.so against numpy / the oracle on seeded inputs.
Integer and bit work must be bit-exact."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import asso_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def nat():
    from pybmf_b200 import _native, device
    _native.require_gpu()
    return _native, device


def _rand01(rng, m, n, d):
    return (rng.rand(m, n) < d).astype(np.uint8)


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


@pytest.mark.parametrize("m,n,d", [(1, 1, 1.0), (7, 130, 0.3), (300, 500, 0.2), (1000, 64, 0.05), (65, 129, 0.0)])
def test_pack_csr_both_orientations(nat, m, n, d):
    _native, device = nat
    rng = np.random.RandomState(m * 1000 + n)
    A = _rand01(rng, m, n, d)
    X = device.to_csr_pattern(sp.csr_matrix(A))
    ip, ix = device.upload_csr(X)
    bits = device.pack_csr(ip, ix, m, n)
    assert np.array_equal(device.bits_to_host(bits, n), A)
    assert np.array_equal(bits.cpu().numpy(), device.dense_to_words(A))          # pad bits are zero
    bt = device.pack_csr(ip, ix, m, n, transposed=True)
    assert np.array_equal(device.bits_to_host(bt, m), A.T)


@pytest.mark.parametrize("rows,ncols,tile", [(5, 70, 128), (300, 500, 256), (257, 128, 256)])
def test_expand_bits_i8(nat, rows, ncols, tile):
    _native, device = nat
    rng = np.random.RandomState(rows + ncols)
    A = _rand01(rng, rows, ncols, 0.4)
    bits = _dev(device.dense_to_words(A))
    plane = device.expand_bits_i8(bits, rows, ncols, 3, -2, tile).cpu().numpy()
    want = np.zeros_like(plane)
    want[:rows, :ncols] = np.where(A == 1, 3, -2)
    assert plane.shape == (device.round_up(rows, tile), device.round_up(ncols, 128))
    assert np.array_equal(plane, want)
    K = _rand01(rng, rows, ncols, 0.3)                              # covered mask -> `masked` value
    for mv in (0, 5):
        masked = device.expand_bits_i8(bits, rows, ncols, 3, -2, tile, mask=_dev(device.dense_to_words(K)),
                                       masked=mv).cpu().numpy()
        want[:rows, :ncols][K == 1] = mv
        assert np.array_equal(masked, want)


@pytest.mark.parametrize("m,n", [(90, 70), (300, 500), (1000, 200)])
def test_assoc_counts_popc(nat, m, n):
    _native, device = nat
    rng = np.random.RandomState(m + n)
    A = _rand01(rng, m, n, 0.2)
    xt = _dev(device.dense_to_words(A.T))
    cnt = device.zeros((n, n), torch.int32)
    _native.call("bmf_assoc_counts_popc", xt, n, xt.shape[1], cnt, n)
    assert np.array_equal(cnt.cpu().numpy().astype(np.int64), O.assoc_counts(A))


@pytest.fixture
def gemm_variant(request, monkeypatch):
    """BMF_GEMM_VARIANT: 1 = one CTA per tile (cta_group::1), 2 = CTA pair (cta_group::2)."""
    monkeypatch.setenv("BMF_GEMM_VARIANT", str(request.param))
    return request.param


@pytest.mark.parametrize("gemm_variant", [1, 2], indirect=True)
@pytest.mark.parametrize("ma,nb,k", [(128, 256, 128), (128, 256, 512), (256, 512, 384), (384, 256, 1152),
                                      (256, 256, 128), (512, 768, 2048), (1280, 2304, 640), (19200, 2560, 256)])
def test_gemm_i8_tcgen05_exact(nat, ma, nb, k, gemm_variant):
    """tcgen05 kind::i8 + TMA + TMEM primitive on random signed int8 (catches descriptor/swizzle bugs)."""
    _native, device = nat
    if gemm_variant == 2 and ma % 256:
        with pytest.raises(ValueError):
            _native.call("bmf_gemm_i8_nt", device.zeros((ma, k), torch.int8), ma, device.zeros((nb, k), torch.int8), nb,
                         k, device.zeros((ma, nb), torch.int32), nb)
        return
    rng = np.random.RandomState(ma + nb + k)
    A = rng.randint(-128, 128, size=(ma, k)).astype(np.int8)
    B = rng.randint(-128, 128, size=(nb, k)).astype(np.int8)
    c = device.zeros((ma, nb), torch.int32)
    _native.call("bmf_gemm_i8_nt", _dev(A), ma, _dev(B), nb, k, c, nb)
    torch.cuda.synchronize()
    want = A.astype(np.int64) @ B.astype(np.int64).T
    assert np.array_equal(c.cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize("gemm_variant", [1, 2], indirect=True)
@pytest.mark.parametrize("m,n", [(300, 500), (700, 130)])
def test_assoc_counts_i8_matches_popc_and_oracle(nat, m, n, gemm_variant):
    _native, device = nat
    rng = np.random.RandomState(m * 3 + n)
    A = _rand01(rng, m, n, 0.15)
    xt = _dev(device.dense_to_words(A.T))
    plane = device.expand_bits_i8(xt, n, m, 1, 0, 256)
    n_pad = plane.shape[0]
    cnt = device.zeros((n_pad, n_pad), torch.int32)
    _native.call("bmf_assoc_counts_i8", plane, n, n_pad, plane.shape[1], cnt, n_pad)
    assert np.array_equal(cnt.cpu().numpy()[:n, :n].astype(np.int64), O.assoc_counts(A))


@pytest.mark.parametrize("tau", [-0.25, 0.0, 0.25, 0.5, 0.99, 1.0])
def test_basis_threshold(nat, tau):
    _native, device = nat
    rng = np.random.RandomState(17)
    m, n = 200, 150
    A = _rand01(rng, m, n, 0.2)
    A[:, 5] = 0                                                    # an empty column -> dead candidate
    cnt_h = O.assoc_counts(A)
    cnt = _dev(cnt_h.astype(np.int32))
    words = device.words_for(n)
    ld = device.round_up(n, 128)
    bits = device.zeros((n, words), torch.int64)
    plane = device.zeros((device.round_up(n, 128), ld), torch.int8)
    alive = device.zeros((n,), torch.uint8)
    pop = device.zeros((n,), torch.int32)
    _native.call("bmf_basis_threshold", cnt, n, n, float(tau), bits, words, plane, ld, alive, pop)
    want = (O.build_assoc(A) > tau).astype(np.uint8)
    assert np.array_equal(pop.cpu().numpy(), want.sum(axis=1))
    assert np.array_equal(device.bits_to_host(bits, n), want)
    assert np.array_equal(plane.cpu().numpy()[:n, :n], want.astype(np.int8))
    assert plane.cpu().numpy()[:, n:].sum() == 0 and plane.cpu().numpy()[n:].sum() == 0
    assert np.array_equal(alive.cpu().numpy(), (want.sum(axis=1) != 0).astype(np.uint8))


@pytest.mark.parametrize("tau", [-0.5, 0.0, 0.1, 1.0 / 3.0, 0.5, 0.75, 1.0, float("inf")])
@pytest.mark.parametrize("n", [37, 150, 333])
def test_basis_threshold_symmetric_tiles(nat, tau, n, monkeypatch):
    """UPPER-stored X^T X (junk below the diagonal, as the symmetric association GEMM leaves it): the 64 x 64 tile kernel and
    the row-per-warp kernel (BMF_BASIS_TILES=0), both deciding `(double)c / (double)s > tau` through the per-row minimal
    count, against the literal numpy division -- incl. ratios that are exactly tau (columns duplicated on purpose)."""
    _native, device = nat
    rng = np.random.RandomState(n)
    A = _rand01(rng, 90, n, 0.3)
    A[:, 5] = 0                                                    # dead candidate
    A[:, 7] = A[:, 3]                                              # ratio exactly 1.0
    A[:45, 9] = 1; A[45:, 9] = 0; A[:, 11] = 0; A[:15, 11] = 1     # 15 / 45 = 1 / 3
    cnt_h = O.assoc_counts(A).astype(np.int32)
    junk = np.tril(rng.randint(-5, 1000, size=cnt_h.shape), -1).astype(np.int32)
    n_pad = device.round_up(n, 64)
    up = np.zeros((n_pad, n_pad), np.int32)
    up[:n, :n] = np.triu(cnt_h) + junk
    want = (O.build_assoc(A) > tau).astype(np.uint8)
    words = device.words_for(n)
    for tiles in ("1", "0"):
        monkeypatch.setenv("BMF_BASIS_TILES", tiles)
        bits = device.zeros((n, words), torch.int64) - 1            # every word (pad word too) must be written
        alive = device.zeros((n,), torch.uint8) + 9
        pop = device.zeros((n,), torch.int32) - 1
        _native.call("bmf_basis_threshold_rows", _dev(up), n_pad, n, 0, n, 1, float(tau), bits, words, alive, pop)
        assert np.array_equal(device.bits_to_host(bits, n), want), tiles
        assert np.array_equal(bits.cpu().numpy(), device.dense_to_words(want)), tiles
        assert np.array_equal(pop.cpu().numpy(), want.sum(axis=1)), tiles
        assert np.array_equal(alive.cpu().numpy(), (want.sum(axis=1) != 0).astype(np.uint8)), tiles


def _cover_inputs(seed, m, n, nb_density=0.25):
    rng = np.random.RandomState(seed)
    X = _rand01(rng, m, n, 0.3)
    C = _rand01(rng, m, n, 0.15)
    B = _rand01(rng, n, n, nb_density)
    alive = (rng.rand(n) < 0.9).astype(np.uint8)
    return X, C, B, alive


@pytest.mark.parametrize("m,n", [(70, 50), (300, 200), (129, 257)])
@pytest.mark.parametrize("w", [(0.5, 0.5), (0.25, 0.75), (0.2, 0.8), (0.3, 0.6)])
def test_cover_score_popc(nat, m, n, w):
    _native, device = nat
    X, C, B, alive = _cover_inputs(m + n, m, n)
    w_fp, w_fn = w
    iw = O.integer_weights(w_fp, w_fn)
    wa, wb = (iw[0], iw[1]) if iw else (0, 0)
    words = device.words_for(n)
    tpo, fpo, _ = O.confusion(X, C, axis=1)
    gp = device.zeros((n,), torch.int64)
    gn = device.zeros((n,), torch.int64)
    _native.call("bmf_cover_score_popc", _dev(device.dense_to_words(X)), _dev(device.dense_to_words(C)), m, n, words,
                 _dev(device.dense_to_words(B)), _dev(alive), _dev(tpo.astype(np.int32)), _dev(fpo.astype(np.int32)),
                 wa, wb, w_fp, w_fn, gp, gn)
    score, use, P, N, _, _ = O.score_candidates(X, C, B, w_fp, w_fn)
    live = alive.astype(bool)
    # 64-candidate tiles with no live candidate are skipped entirely; live ones must be exact
    if iw:
        G = O.integer_gains(X, C, B, wa, wb)
        assert np.array_equal(gp.cpu().numpy()[live], G[live])
    else:
        assert np.array_equal(gp.cpu().numpy()[live], (P * use).sum(axis=0)[live])
        assert np.array_equal(gn.cpu().numpy()[live], (N * use).sum(axis=0)[live])


@pytest.mark.parametrize("gemm_variant", [1, 2], indirect=True)
@pytest.mark.parametrize("m,n,w", [(70, 50, (0.5, 0.5)), (300, 200, (0.25, 0.75)), (600, 700, (0.5, 0.5)),
                                   (1000, 500, (0.375, 0.5)), (5000, 2100, (0.5, 0.5))])
def test_cover_score_i8_tcgen05(nat, m, n, w, gemm_variant):
    _native, device = nat
    X, C, B, alive = _cover_inputs(m * 7 + n, m, n)
    wa, wb, _ = O.integer_weights(*w)
    words = device.words_for(n)
    rows = np.where(C == 1, 0, np.where(X == 1, wb, -wa)).astype(np.int8)
    ld = device.round_up(n, 128)
    rows_plane = np.zeros((device.round_up(m, 256), ld), np.int8)
    rows_plane[:m, :n] = rows
    cand_plane = np.zeros((device.round_up(n, 256), ld), np.int8)
    cand_plane[:n, :n] = B
    G = O.integer_gains(X, C, B, wa, wb)
    for sign in (1, -1):                                            # signed encodings
        gain = device.zeros((cand_plane.shape[0],), torch.int64)
        _native.call("bmf_cover_score_i8", _dev(cand_plane), cand_plane.shape[0], _dev(sign * rows_plane),
                     rows_plane.shape[0], ld, sign, None, 0, gain)
        got = gain.cpu().numpy()
        assert np.array_equal(got[:n], G) and got[n:].sum() == 0
    # zero-dominant encoding: uncovered one -> wa+wb, covered -> wa, uncovered zero -> 0, bias wa*|b_j|
    zplane = np.zeros_like(rows_plane)
    zplane[:m, :n] = np.where(C == 1, wa, np.where(X == 1, wa + wb, 0))
    pop = np.zeros(cand_plane.shape[0], np.int32)
    pop[:n] = B.sum(axis=1)
    gain = device.zeros((cand_plane.shape[0],), torch.int64)
    _native.call("bmf_cover_score_i8", _dev(cand_plane), cand_plane.shape[0], _dev(zplane), rows_plane.shape[0], ld, 1,
                 _dev(pop), wa, gain)
    got = gain.cpu().numpy()
    assert np.array_equal(got[:n], G) and got[n:].sum() == 0


def _pq_plane(device, X, C, ld):
    """numpy statement of the interleaved P/Q operand (include/pybmf_b200.h, bmf_expand_bits_pq)"""
    m, n = X.shape
    blocks = (m + 127) // 128
    plane = np.zeros((blocks * 256, ld), np.int8)
    for i in range(m):
        pr = (i // 128) * 256 + i % 128
        plane[pr, :n] = X[i] & (1 - C[i])
        plane[pr + 128, :n] = C[i]
    return plane


@pytest.mark.parametrize("gemm_variant", [1, 2], indirect=True)
@pytest.mark.parametrize("m,n,w", [(70, 50, (0.2, 0.8)), (300, 200, (0.3, 0.6)), (129, 257, (0.2, 0.8)),
                                   (1000, 500, (0.5, 0.5)), (3000, 1100, (0.15, 0.85))])
def test_cover_score_i8_general_tcgen05(nat, m, n, w, gemm_variant):
    """general weights on the tensor cores: P/Q side by side in the accumulator, fp64 row test in the epilogue"""
    _native, device = nat
    X, C, B, alive = _cover_inputs(m * 11 + n, m, n)
    w_fp, w_fn = w
    words = device.words_for(n)
    ld = device.round_up(n, 128)
    tpo, fpo, _ = O.confusion(X, C, axis=1)
    cand_pad = device.round_up(n, 256)
    cand_plane = np.zeros((cand_pad, ld), np.int8)
    cand_plane[:n, :n] = B
    pop = np.zeros(cand_pad, np.int32)
    pop[:n] = B.sum(axis=1)
    pq = device.empty((2 * device.round_up(m, 128), ld), torch.int8)
    _native.call("bmf_expand_bits_pq", _dev(device.dense_to_words(X)), _dev(device.dense_to_words(C)), m, n, words,
                 pq, ld)
    assert np.array_equal(pq.cpu().numpy(), _pq_plane(device, X, C, ld))
    gp = device.zeros((cand_pad,), torch.int64) + 7                 # overwritten, not accumulated
    gn = device.zeros((cand_pad,), torch.int64) + 7
    _native.call("bmf_cover_score_i8_general", _dev(cand_plane), cand_pad, pq, m, ld, _dev(pop),
                 _dev(tpo.astype(np.int32)), _dev(fpo.astype(np.int32)), w_fp, w_fn, gp, gn)
    score, use, P, N, _, _ = O.score_candidates(X, C, B, w_fp, w_fn)
    got_p, got_n = gp.cpu().numpy(), gn.cpu().numpy()
    assert np.array_equal(got_p[:n], (P * use).sum(axis=0)) and np.array_equal(got_n[:n], (N * use).sum(axis=0))
    assert got_p[n:].sum() == 0 and got_n[n:].sum() == 0


@pytest.mark.parametrize("m,n,w", [(70, 50, (0.2, 0.8)), (300, 200, (0.3, 0.6)), (121, 257, (0.2, 0.8)),
                                   (3000, 1100, (0.15, 0.85))])
@pytest.mark.parametrize("no_fix", ["0", "1"])
def test_cover_score_f4_general_tcgen05(nat, m, n, w, no_fix, monkeypatch):
    """general weights on the FP4 pipe: packed E2M1 P/Q planes in blocks of 120 rows, fp64 row test in the epilogue;
    the plane kept current by bmf_cover_apply_f4_general equals a fresh expansion.  no_fix = 1 switches off the fixed-point
    pre-decision of the row test (fp64 for every element): both must equal the oracle"""
    monkeypatch.setenv("BMF_NO_FIXED_PREDECISION", no_fix)
    _native, device = nat
    X, C, B, alive = _cover_inputs(m * 13 + n, m, n)
    alive[:] = 1
    w_fp, w_fn = w
    words = device.words_for(n)
    ld_bytes = device.round_up(n, 256) // 2
    cand_pad = device.round_up(n, 256)
    tpo, fpo, _ = O.confusion(X, C, axis=1)
    x_d, c_d, b_d = _dev(device.dense_to_words(X)), _dev(device.dense_to_words(C)), _dev(device.dense_to_words(B))
    cand_plane = device.empty((cand_pad, ld_bytes), torch.uint8)
    _native.call("bmf_expand_bits_f4", b_d, None, n, n, words, 2, 0, 0, cand_plane, cand_pad, ld_bytes)
    pop = np.zeros(cand_pad, np.int32)
    pop[:n] = B.sum(axis=1)
    pq = device.empty((2 * device.round_up(m, 120), ld_bytes), torch.uint8)
    _native.call("bmf_expand_bits_pq_f4", x_d, c_d, m, n, words, pq, ld_bytes)
    vals = np.zeros((pq.shape[0], n), np.int64)                      # numpy statement of the layout
    for i in range(m):
        pr = (i // 120) * 240 + i % 120
        vals[pr] = X[i] & (1 - C[i])
        vals[pr + 120] = C[i]
    assert np.array_equal(pq.cpu().numpy(), _pack_f4(vals, ld_bytes))
    gp = device.zeros((cand_pad,), torch.int64) + 7
    gn = device.zeros((cand_pad,), torch.int64) + 7
    tp_d, fp_d = _dev(tpo.astype(np.int32)), _dev(fpo.astype(np.int32))
    _native.call("bmf_cover_score_f4_general", cand_plane, cand_pad, pq, m, ld_bytes, _dev(pop), tp_d, fp_d, w_fp, w_fn,
                 gp, gn)
    score, use, P, N, _, _ = O.score_candidates(X, C, B, w_fp, w_fn)
    got_p, got_n = gp.cpu().numpy(), gn.cpu().numpy()
    assert np.array_equal(got_p[:n], (P * use).sum(axis=0)) and np.array_equal(got_n[:n], (N * use).sum(axis=0))
    assert got_p[n:].sum() == 0 and got_n[n:].sum() == 0
    j = int(np.argmax(score))
    ub = device.zeros((device.words_for(m),), torch.int64)
    tot = device.zeros((3,), torch.int64)
    _native.call("bmf_cover_apply_f4_general", x_d, c_d, m, n, words, b_d, _dev(alive), _dev(np.array([j], np.int64)),
                 tp_d, fp_d, w_fp, w_fn, pq, ld_bytes, ub, tot)
    u = use[:, j]
    assert list(tot.cpu().numpy()) == [int(u.sum()), int(P[u, j].sum()), int(N[u, j].sum())]
    fresh = torch.empty_like(pq)
    _native.call("bmf_expand_bits_pq_f4", x_d, c_d, m, n, words, fresh, ld_bytes)
    assert torch.equal(fresh, pq)


def test_cover_apply_general_updates_pq_plane(nat):
    _native, device = nat
    m, n = 333, 200
    X, C, B, alive = _cover_inputs(99, m, n)
    alive[:] = 1
    w_fp, w_fn = 0.2, 0.8
    j = 37
    words = device.words_for(n)
    ld = device.round_up(n, 128)
    tpo, fpo, _ = O.confusion(X, C, axis=1)
    score, use, P, N, _, _ = O.score_candidates(X, C, B[j:j + 1], w_fp, w_fn)
    u = use[:, 0]
    pq = _dev(_pq_plane(device, X, C, ld))
    c_d = _dev(device.dense_to_words(C))
    tp_d, fp_d = _dev(tpo.astype(np.int32)), _dev(fpo.astype(np.int32))
    ub = device.zeros((device.words_for(m),), torch.int64)
    tot = device.zeros((3,), torch.int64)
    _native.call("bmf_cover_apply_general", _dev(device.dense_to_words(X)), c_d, m, n, words,
                 _dev(device.dense_to_words(B)), _dev(alive), _dev(np.array([j], np.int64)), tp_d, fp_d, w_fp, w_fn,
                 pq, ld, ub, tot)
    Cn = C | (u[:, None].astype(np.uint8) & B[j][None, :])
    assert np.array_equal(device.bits_to_host(c_d, n), Cn)
    assert np.array_equal(pq.cpu().numpy(), _pq_plane(device, X, Cn, ld))
    assert list(tot.cpu().numpy()) == [int(u.sum()), int(P[u, 0].sum()), int(N[u, 0].sum())]


E2M1_CODE = {0: 0, 1: 2, 2: 4, 3: 5, 4: 6, 6: 7}


def _pack_f4(values, ld_bytes):
    """integer matrix -> packed E2M1 plane [rows, ld_bytes]: element k in byte k/2, low nibble for even k"""
    rows, cols = values.shape
    codes = np.zeros((rows, ld_bytes * 2), np.uint8)
    lut = np.zeros(8, np.uint8)
    for v, c in E2M1_CODE.items():
        lut[v] = c
    codes[:, :cols] = lut[values]
    return (codes[:, 0::2] | (codes[:, 1::2] << 4)).astype(np.uint8)


@pytest.mark.parametrize("ma,nb,k", [(256, 240, 256), (256, 480, 1024), (512, 240, 2304), (768, 720, 17920),
                                     (256, 496, 256), (512, 992, 1280), (768, 1488, 17920), (256, 7440, 512)])
def test_gemm_f4_tcgen05_exact(nat, ma, nb, k, monkeypatch):
    """tcgen05 kind::mxf4 with unit block scales on small-integer operands is EXACT (FP32 accumulate of integers)"""
    _native, device = nat
    rng = np.random.RandomState(ma + nb + k)
    A = (rng.rand(ma, k) < 0.4).astype(np.int64)                      # candidate-like operand {0,1}
    vals = np.array([0, 1, 2, 3, 4, 6])
    B = vals[rng.randint(0, 6, size=(nb, k))]
    B[: nb // 4] = 6                                                  # dense rows: sums up to 6*0.4*k
    assert _native.load().bmf_e2m1_code(5) == -1 and _native.load().bmf_e2m1_code(6) == 7
    ld_bytes = k // 2
    want = A @ B.T
    assert want.max() > 4000 or k < 2000
    a_d, b_d = _dev(_pack_f4(A, ld_bytes)), _dev(_pack_f4(B, ld_bytes))
    for no_super in ("0", "1"):                                       # rows % 496 == 0 -> super-tile kernel (256 + 240)
        if no_super == "1" and nb % 240:
            continue
        monkeypatch.setenv("BMF_F4_NO_SUPER", no_super)
        c = device.zeros((ma, nb), torch.int32) - 1
        _native.call("bmf_gemm_f4_nt", a_d, ma, b_d, nb, ld_bytes, c, nb, 0)
        assert np.array_equal(c.cpu().numpy().astype(np.int64), want)
        if no_super == "0" and nb % 496 == 0:                         # accumulate mode: c += a b^T
            _native.call("bmf_gemm_f4_nt", a_d, ma, b_d, nb, ld_bytes, c, nb, 1)
            assert np.array_equal(c.cpu().numpy().astype(np.int64), 2 * want)


@pytest.mark.parametrize("rows,ncols", [(5, 70), (300, 500), (241, 257)])
def test_expand_bits_f4(nat, rows, ncols):
    _native, device = nat
    rng = np.random.RandomState(rows + ncols)
    A = _rand01(rng, rows, ncols, 0.4)
    K = _rand01(rng, rows, ncols, 0.3)
    ld_bytes = device.round_up(ncols, 256) // 2
    rows_pad = device.round_up(rows, 240)
    plane = device.empty((rows_pad, ld_bytes), torch.uint8)
    _native.call("bmf_expand_bits_f4", _dev(device.dense_to_words(A)), _dev(device.dense_to_words(K)), rows, ncols,
                 device.words_for(ncols), 4, 0, 2, plane, rows_pad, ld_bytes)
    vals = np.zeros((rows_pad, ncols), np.int64)
    vals[:rows] = np.where(K == 1, 1, np.where(A == 1, 2, 0))
    assert np.array_equal(plane.cpu().numpy(), _pack_f4(vals, ld_bytes))


@pytest.mark.parametrize("row_pad", [240, 496])
@pytest.mark.parametrize("m,n,w", [(70, 50, (0.5, 0.5)), (300, 200, (0.25, 0.75)), (1000, 500, (0.75, 0.25)),
                                   (5000, 2100, (0.5, 0.5))])
def test_cover_score_f4_tcgen05(nat, m, n, w, row_pad):
    """FP4 scorer == integer gains, and the plane kept current by bmf_cover_apply_f4 == a fresh expansion"""
    _native, device = nat
    X, C, B, alive = _cover_inputs(m * 7 + n, m, n)
    alive[:] = 1
    wa, wb, _ = O.integer_weights(*w)
    lib = _native.load()
    c_one, c_cov = lib.bmf_e2m1_code(wa + wb), lib.bmf_e2m1_code(wa)
    assert c_one >= 0 and c_cov >= 0
    words = device.words_for(n)
    ld_bytes = device.round_up(n, 256) // 2
    rows_pad, cand_pad = device.round_up(m, row_pad), device.round_up(n, 256)   # 496 -> super-tile kernel
    x_d, c_d, b_d = _dev(device.dense_to_words(X)), _dev(device.dense_to_words(C)), _dev(device.dense_to_words(B))
    rows_plane = device.empty((rows_pad, ld_bytes), torch.uint8)
    cand_plane = device.empty((cand_pad, ld_bytes), torch.uint8)
    _native.call("bmf_expand_bits_f4", x_d, c_d, m, n, words, c_one, 0, c_cov, rows_plane, rows_pad, ld_bytes)
    _native.call("bmf_expand_bits_f4", b_d, None, n, n, words, 2, 0, 0, cand_plane, cand_pad, ld_bytes)
    pop = np.zeros(cand_pad, np.int32)
    pop[:n] = B.sum(axis=1)
    gain = device.zeros((cand_pad,), torch.int64) + 5
    _native.call("bmf_cover_score_f4", cand_plane, cand_pad, rows_plane, rows_pad, ld_bytes, _dev(pop), wa, gain)
    G = O.integer_gains(X, C, B, wa, wb)
    got = gain.cpu().numpy()
    assert np.array_equal(got[:n], G) and got[n:].sum() == 0
    # apply the best candidate and compare the updated plane with a fresh expansion
    j = int(np.argmax(G))
    tpo, fpo, _ = O.confusion(X, C, axis=1)
    ub = device.zeros((device.words_for(m),), torch.int64)
    tot = device.zeros((3,), torch.int64)
    _native.call("bmf_cover_apply_f4", x_d, c_d, m, n, words, b_d, _dev(alive), _dev(np.array([j], np.int64)),
                 _dev(tpo.astype(np.int32)), _dev(fpo.astype(np.int32)), wa, wb, rows_plane, ld_bytes, c_cov, ub, tot)
    fresh = device.empty((rows_pad, ld_bytes), torch.uint8)
    _native.call("bmf_expand_bits_f4", x_d, c_d, m, n, words, c_one, 0, c_cov, fresh, rows_pad, ld_bytes)
    assert torch.equal(fresh, rows_plane)


def test_select_first_max(nat):
    _native, device = nat
    n = 3000
    rng = np.random.RandomState(5)
    g = rng.randint(0, 50, size=n).astype(np.int64)
    g[[700, 1500, 2900]] = 99                                      # three-way tie: lowest live index wins
    alive = np.ones(n, np.uint8)
    alive[700] = 0
    rec = device.zeros((2,), torch.int64)
    args = (_dev(g), None, _dev(alive), n, 1, 1, 10, 0.5, 0.5, 0.5, 0, 0)
    _native.call("bmf_select_first_max", *args, 0.0, rec)
    r = rec.cpu().numpy()
    assert r[0] == 1500 and r[1:2].view(np.float64)[0] == (10 + 99) * 0.5
    _native.call("bmf_select_first_max", *args, 54.5, rec)          # not strictly greater than inherited best
    assert rec.cpu().numpy()[0] == -1
    # general mode: score from totals in fp64
    gp = rng.randint(0, 1000, size=n).astype(np.int64)
    gn = rng.randint(0, 1000, size=n).astype(np.int64)
    _native.call("bmf_select_first_max", _dev(gp), _dev(gn), _dev(alive), n, 0, 0, 0, 0.0, 0.2, 0.8, 12345, 678,
                 -1e300, rec)
    sc = -0.2 * (678 + gn).astype(np.float64) + 0.8 * (12345 + gp).astype(np.float64)
    sc[alive == 0] = -np.inf
    r = rec.cpu().numpy()
    assert r[0] == int(np.argmax(sc)) and r[1:2].view(np.float64)[0] == sc.max()


@pytest.mark.parametrize("w", [(0.5, 0.5), (0.2, 0.8)])
def test_cover_apply(nat, w):
    _native, device = nat
    m, n = 333, 200
    X, C, B, alive = _cover_inputs(99, m, n)
    alive[:] = 1
    w_fp, w_fn = w
    iw = O.integer_weights(w_fp, w_fn)
    wa, wb = (iw[0], iw[1]) if iw else (0, 0)
    j = 37
    words = device.words_for(n)
    ld = device.round_up(n, 128)
    tpo, fpo, _ = O.confusion(X, C, axis=1)
    score, use, P, N, _, _ = O.score_candidates(X, C, B[j:j + 1], w_fp, w_fn)
    u = use[:, 0]
    rows = np.zeros((device.round_up(m, 256), ld), np.int8)
    rows[:m, :n] = np.where(C == 1, 0, np.where(X == 1, max(wb, 1), -max(wa, 1)))
    rows_d = _dev(rows)
    c_d = _dev(device.dense_to_words(C))
    tp_d, fp_d = _dev(tpo.astype(np.int32)), _dev(fpo.astype(np.int32))
    alive_d = _dev(alive)
    ub = device.zeros((device.words_for(m),), torch.int64)
    tot = device.zeros((3,), torch.int64)
    win = _dev(np.array([j], np.int64))
    _native.call("bmf_cover_apply", _dev(device.dense_to_words(X)), c_d, m, n, words, _dev(device.dense_to_words(B)),
                 alive_d, win, tp_d, fp_d, wa, wb, w_fp, w_fn, rows_d, ld, 9, ub, tot)
    Cn = C | (u[:, None].astype(np.uint8) & B[j][None, :])
    assert np.array_equal(device.bits_to_host(c_d, n), Cn)
    assert np.array_equal(device.words_to_dense(ub.cpu().numpy().reshape(1, -1), m)[0], u.astype(np.uint8))
    tpn, fpn, _ = O.confusion(X, Cn, axis=1)
    assert np.array_equal(tp_d.cpu().numpy(), tpn) and np.array_equal(fp_d.cpu().numpy(), fpn)
    assert list(tot.cpu().numpy()) == [int(u.sum()), int(P[u, 0].sum()), int(N[u, 0].sum())]
    want_rows = rows.copy()
    want_rows[:m, :n][(Cn == 1) & (C == 0)] = 9                     # newly covered entries take covered_value
    assert np.array_equal(rows_d.cpu().numpy(), want_rows)
    assert alive_d.cpu().numpy()[j] == 0 and alive_d.cpu().numpy().sum() == n - 1
    # winner < 0 is a no-op
    before = c_d.clone()
    _native.call("bmf_cover_apply", _dev(device.dense_to_words(X)), c_d, m, n, words, _dev(device.dense_to_words(B)),
                 alive_d, _dev(np.array([-1], np.int64)), tp_d, fp_d, wa, wb, w_fp, w_fn, rows_d, ld, 9, ub, tot)
    assert torch.equal(before, c_d)


@pytest.mark.parametrize("m,n,k", [(50, 70, 1), (300, 500, 5), (257, 129, 64), (100, 200, 70), (40, 20000, 9)])
def test_bool_product_and_confusion(nat, m, n, k):
    _native, device = nat
    rng = np.random.RandomState(m + n + k)
    U = _rand01(rng, m, k, 0.2)
    V = _rand01(rng, n, k, 0.2)
    X = _rand01(rng, m, n, 0.3)
    kw = (k + 63) // 64
    uw = np.ascontiguousarray(device.dense_to_words(U, words=kw))
    vt = device.dense_to_words(V.T)
    words = device.words_for(n)
    pd = device.zeros((m, words), torch.int64)
    _native.call("bmf_bool_product", _dev(uw), m, kw, _dev(vt), k, words, pd)
    want = O.bool_product(U, V)
    assert np.array_equal(device.bits_to_host(pd, n), want)
    assert np.array_equal(pd.cpu().numpy(), device.dense_to_words(want))
    tp, fp, fn = O.confusion(X, want)
    rtp, rfp, _ = O.confusion(X, want, axis=1)
    for mode, ones in (("factors", -1), ("bits", -1), ("factors", int(X.sum())), ("bits", int(X.sum()))):
        counts = device.zeros((3,), torch.int64) + 12345            # overwritten, not accumulated
        row_tp = device.zeros((m,), torch.int32)
        row_fp = device.zeros((m,), torch.int32)
        if mode == "factors":
            _native.call("bmf_confusion_factors", _dev(device.dense_to_words(X)), m, words, _dev(uw), kw, _dev(vt), k,
                         ones, counts, row_tp, row_fp)
        else:
            _native.call("bmf_confusion_bits", _dev(device.dense_to_words(X)), pd, m, words, ones, counts, row_tp, row_fp)
        assert list(counts.cpu().numpy()) == [int(tp), int(fp), int(fn)]
        assert np.array_equal(row_tp.cpu().numpy(), rtp) and np.array_equal(row_fp.cpu().numpy(), rfp)
        if mode == "factors":                                       # totals only -> column-panel kernel (Harley-Seal)
            counts = device.zeros((3,), torch.int64) + 999
            _native.call("bmf_confusion_factors", _dev(device.dense_to_words(X)), m, words, _dev(uw), kw, _dev(vt), k,
                         ones, counts, None, None)
            assert list(counts.cpu().numpy()) == [int(tp), int(fp), int(fn)]


@pytest.mark.parametrize("m,n,k", [(1, 64, 1), (33, 16500, 64), (2500, 33000, 17), (10001, 4097, 100), (777, 1000, 128),
                                   (150001, 130, 5)])
def test_panel_kernels_match_row_stream_kernels(nat, m, n, k, monkeypatch):
    """column-panel product / confusion (V^T in shared memory, Harley-Seal counting) vs the row-stream
    kernels (BMF_NO_PANEL=1) and numpy; k > 100 does not fit shared memory and must fall back silently"""
    _native, device = nat
    rng = np.random.RandomState(m + n + k)
    kw = (k + 63) // 64
    U = _rand01(rng, m, k, 2.0 / max(k, 4))
    V = _rand01(rng, n, k, 0.05)
    uw = _dev(np.ascontiguousarray(device.dense_to_words(U, words=kw)))
    vt = _dev(device.dense_to_words(V.T))
    words = device.words_for(n)
    gt = torch.randint(-2 ** 63, 2 ** 63 - 1, (m, words), dtype=torch.int64, device="cuda")
    gt &= torch.randint(-2 ** 63, 2 ** 63 - 1, (m, words), dtype=torch.int64, device="cuda")
    if n % 64:
        gt[:, n // 64] &= (1 << (n % 64)) - 1
    gt[:, (n + 63) // 64:] = 0
    out = {}
    for tag, env in (("panel", "0"), ("rows", "1")):
        monkeypatch.setenv("BMF_NO_PANEL", env)
        pd = device.zeros((m, words), torch.int64) - 1
        _native.call("bmf_bool_product", uw, m, kw, vt, k, words, pd)
        c_known = device.zeros((3,), torch.int64)
        c_count = device.zeros((3,), torch.int64)
        _native.call("bmf_confusion_factors", gt, m, words, uw, kw, vt, k, -1, c_count, None, None)
        _native.call("bmf_confusion_factors", gt, m, words, uw, kw, vt, k, int(c_count[0] + c_count[2]), c_known, None, None)
        out[tag] = (pd, c_count.cpu().numpy(), c_known.cpu().numpy())
    assert torch.equal(out["panel"][0], out["rows"][0])
    assert np.array_equal(out["panel"][1], out["rows"][1]) and np.array_equal(out["panel"][2], out["rows"][2])
    assert np.array_equal(out["panel"][1], out["panel"][2])
    want = O.bool_product(U, V)
    assert np.array_equal(device.bits_to_host(out["panel"][0], n), want)
    G = device.bits_to_host(gt, n)
    tp, fp, fn = O.confusion(G, want)
    assert list(out["panel"][1]) == [int(tp), int(fp), int(fn)]


@pytest.mark.parametrize("density", [0.02, 0.07, 0.11, 0.3])
@pytest.mark.parametrize("k", [64, 37, 20])
def test_panel_selection_lists_and_count_modes(nat, density, k, monkeypatch):
    """selection-list panel kernels (<= 7 factors per row: straight-line ORs; more: the bit-scan path; one factor:
    |prediction| from the per-panel row counts; none: the row is skipped) in every count mode, and the bit-scan
    forms (BMF_PANEL_LIST=0), against numpy.  k = 20 takes the two-chunk panel."""
    _native, device = nat
    m, n = 1501, 20010
    rng = np.random.RandomState(int(density * 1000) + k)
    U = _rand01(rng, m, k, density)
    U[5] = 0
    U[6] = 0; U[6, k - 1] = 1
    U[7] = 1
    V = _rand01(rng, n, k, 0.05)
    G = _rand01(rng, m, n, 0.1)
    uw_h = np.ascontiguousarray(device.dense_to_words(U, words=1))
    want = O.bool_product(U, V)
    tp, fp, fn = (int(v) for v in O.confusion(G, want))
    vt = _dev(device.dense_to_words(V.T))
    gt = _dev(device.dense_to_words(G))
    words = device.words_for(n)
    uw = _dev(uw_h)
    for lists, dynamic in (("1", "1"), ("1", "0"), ("0", "1")):
        monkeypatch.setenv("BMF_PANEL_LIST", lists)
        monkeypatch.setenv("BMF_PANEL_DYNAMIC", dynamic)                # row blocks drawn from a counter / static split
        pd = device.zeros((m, words), torch.int64) - 1
        _native.call("bmf_bool_product", uw, m, 1, vt, k, words, pd)
        assert np.array_equal(device.bits_to_host(pd, n), want)
        for mode in ("0", "1", "2", "3"):
            monkeypatch.setenv("BMF_CONFUSION_COUNT", mode)
            for known in (-1, tp + fn):
                c = device.zeros((3,), torch.int64) + 7
                _native.call("bmf_confusion_factors", gt, m, words, uw, 1, vt, k, known, c, None, None)
                assert [int(v) for v in c.cpu().numpy()] == [tp, fp, fn], (lists, mode, known)


def test_c5_scale_panel_kernels_properties(nat, monkeypatch):
    """BASELINE configs[4] scale (200k x 100k bits, k = 64; the 1M-row point runs in bench.py --workload c5 with the same
    checks): the column-panel kernels equal the row-stream kernels bit for bit, the two confusion forms agree, and
    TP + FN = |X|, TP + FP = |product|."""
    _native, device = nat
    m, n, k = 200_000, 100_000, 64
    words = device.words_for(n)
    g = torch.Generator(device="cuda"); g.manual_seed(11)

    def rnd(shape, ands):
        w = torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
        for _ in range(ands - 1):
            w &= torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
        return w
    uw = rnd((m, 1), 5)                                             # ~2 of 64 factors per row
    vt = rnd((k, words), 5)
    vt[:, n // 64] &= (1 << (n % 64)) - 1
    vt[:, (n + 63) // 64:] = 0
    out = {}
    for tag, env in (("panel", "0"), ("rows", "1")):
        monkeypatch.setenv("BMF_NO_PANEL", env)
        pd = device.zeros((m, words), torch.int64)
        _native.call("bmf_bool_product", uw, m, 1, vt, k, words, pd)
        out[tag] = pd
    assert torch.equal(out["panel"], out["rows"])
    pd = out["panel"]
    del out
    x = pd ^ rnd((m, words), 4)                                     # ground truth = product with ~6 % of the bits flipped
    x[:, n // 64] &= (1 << (n % 64)) - 1
    x[:, (n + 63) // 64:] = 0
    c_bits = device.zeros((3,), torch.int64)
    _native.call("bmf_confusion_bits", x, pd, m, words, -1, c_bits, None, None)
    tp, fp, fn = (int(v) for v in c_bits.cpu().numpy())
    ones_x = tp + fn
    for env in ("0", "1"):
        monkeypatch.setenv("BMF_NO_PANEL", env)
        for known in (-1, ones_x):
            c = device.zeros((3,), torch.int64) + 3
            _native.call("bmf_confusion_factors", x, m, words, uw, 1, vt, k, known, c, None, None)
            assert [int(v) for v in c.cpu().numpy()] == [tp, fp, fn]
    cx = device.zeros((3,), torch.int64)
    _native.call("bmf_confusion_bits", x, x, m, words, -1, cx, None, None)
    cp = device.zeros((3,), torch.int64)
    _native.call("bmf_confusion_bits", pd, pd, m, words, -1, cp, None, None)
    assert int(cx[0].item()) == tp + fn and int(cp[0].item()) == tp + fp


def test_c5_largest_point_blocks_against_cpu_restatement(nat):
    """BASELINE configs[4] at its LARGEST point (1M x 100k, k = 64): the materialised product and the per-block TP / FP / FN
    of the fused confusion kernel against the CPU restatement (oracle/asso_c.c: bmfo_bool_product + bmfo_confusion) on
    four sampled 4096-row blocks (first, two interior, the ragged last one), plus the size-independent identities."""
    import ctypes as C
    from oracle import asso_oracle_c as OC
    _native, device = nat
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
    _native.call("bmf_bool_product", uw, m, 1, vt, k, words, pd)
    x = pd.clone()
    for c0 in range(0, m, 65536):                                   # ground truth = product with ~6 % of the bits flipped
        blk = x[c0:c0 + 65536]
        blk ^= rnd(blk.shape, 4)
    x[:, n // 64] &= (1 << (n % 64)) - 1
    x[:, (n + 63) // 64:] = 0
    total = device.zeros((3,), torch.int64)
    _native.call("bmf_confusion_factors", x, m, words, uw, 1, vt, k, -1, total, None, None)
    vt_h = vt.cpu().numpy().view(np.uint64)
    L = OC.lib()
    for r0 in (0, 117 * 4096 + 17, 700_001, m - 3000):
        r1 = min(m, r0 + 4096)
        rows = r1 - r0
        uw_h = np.ascontiguousarray(uw[r0:r1].cpu().numpy().view(np.uint64))
        want_pd = np.zeros((rows, words), dtype=np.uint64)
        L.bmfo_bool_product(OC._ptr(uw_h), rows, 1, OC._ptr(vt_h), k, words, -1, OC._ptr(want_pd))
        assert np.array_equal(pd[r0:r1].cpu().numpy().view(np.uint64), want_pd), r0
        x_h = np.ascontiguousarray(x[r0:r1].cpu().numpy().view(np.uint64))
        want = np.zeros(3, dtype=np.int64)
        L.bmfo_confusion(OC._ptr(x_h), OC._ptr(want_pd), rows, words, OC._ptr(want), None, None)
        got = device.zeros((3,), torch.int64)
        _native.call("bmf_confusion_factors", x[r0:r1], rows, words, uw[r0:r1], 1, vt, k, -1, got, None, None)
        assert [int(v) for v in got.cpu().numpy()] == [int(v) for v in want], r0
    cb = device.zeros((3,), torch.int64)
    _native.call("bmf_confusion_bits", x, pd, m, words, -1, cb, None, None)
    assert torch.equal(cb, total)                                   # fused == materialised at full size


def test_confusion_triplets(nat):
    _native, device = nat
    rng = np.random.RandomState(8)
    m, n, k, nnz = 120, 90, 6, 5000
    U = _rand01(rng, m, k, 0.3)
    V = _rand01(rng, n, k, 0.3)
    r = rng.randint(0, m, nnz).astype(np.int32)
    c = rng.randint(0, n, nnz).astype(np.int32)
    g = (rng.rand(nnz) < 0.5).astype(np.uint8)
    pd = O.bool_product(U, V)[r, c]
    counts = device.zeros((4,), torch.int64)
    _native.call("bmf_confusion_triplets", _dev(r), _dev(c), _dev(g), nnz, _dev(device.dense_to_words(U, words=1)), 1,
                 _dev(device.dense_to_words(V, words=1)), counts)
    want = [int(((g == 1) & (pd == 1)).sum()), int(((g == 0) & (pd == 1)).sum()),
            int(((g == 1) & (pd == 0)).sum()), int(((g == 0) & (pd == 0)).sum())]
    assert list(counts.cpu().numpy()) == want


@pytest.mark.parametrize("w", [(0.5, 0.5), (0.2, 0.8)])
def test_refine_column(nat, w):
    _native, device = nat
    rng = np.random.RandomState(21)
    m, n, k = 310, 190, 5
    X = _rand01(rng, m, n, 0.3)
    U = _rand01(rng, m, k, 0.3)
    V = _rand01(rng, n, k, 0.3)
    w_fp, w_fn = w
    iw = O.integer_weights(w_fp, w_fn)
    wa, wb = (iw[0], iw[1]) if iw else (0, 0)
    words = device.words_for(n)
    uw = _dev(device.dense_to_words(U, words=1))
    vt = _dev(device.dense_to_words(V.T))
    xb = _dev(device.dense_to_words(X))
    for col in (0, 3):
        out = device.zeros((5,), torch.int64)
        _native.call("bmf_refine_column", xb, m, n, words, uw, 1, vt, k, col, wa, wb, w_fp, w_fn, out)
        idx = [i for i in range(k) if i != col]
        C_old = O.bool_product(U[:, idx], V[:, idx])
        score, use, P, N, tpo, fpo = O.score_candidates(X, C_old, V[:, col].reshape(1, -1), w_fp, w_fn)
        U[:, col] = use[:, 0]
        assert np.array_equal(device.words_to_dense(uw.cpu().numpy(), k), U)
        tp, fp, _ = O.confusion(X, O.bool_product(U, V))
        u = use[:, 0]
        assert list(out.cpu().numpy()) == [int(tp), int(fp), int(u.sum()), int(P[u, 0].sum()), int(N[u, 0].sum())]
