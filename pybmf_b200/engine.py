"""Device-resident state of one Asso fit: bit-packed X and cover, candidate basis, int8 operand
planes, and the per-step launch sequence  score-all -> (all-reduce) -> argmax -> apply.

Rows of X are sharded across ranks when torch.distributed is initialised (one process per
GPU); the candidate basis, V and the greedy decisions are replicated.  Integer partial
gains are summed with ONE all-reduce per greedy step, so every rank sees identical
totals and takes the identical lowest-index strict argmax.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

from . import _native, device

ROW_ALIGN = 256        # shard boundaries are multiples of the MMA N tile (and of 64-bit u words)
F4_ROW_PAD = 496       # BMF_F4_SUPER_ROWS: data rows of the FP4 planes are padded to whole super tiles (256 + 240)


def integer_weights(w_fp, w_fn, max_int=127, max_shift=30):
    """(a, b, s) with w_fp = a/2^s and w_fn = b/2^s exactly and 0 <= a, b <= 127, else None.

    When it exists, every product and sum of the reference's fp64 score expression
    (PyBMF/utils/metrics.py:201) is exact, so `s_new > s_old` <=> b*P - a*N > 0 and the
    score is 2^-s times an integer: one signed int8 contraction decides everything."""
    for s in range(max_shift + 1):
        a, b = w_fp * (1 << s), w_fn * (1 << s)
        if a == int(a) and b == int(b):
            a, b = int(a), int(b)
            if 0 <= a <= max_int and 0 <= b <= max_int and (a | b):
                return a, b, s
            return None
    return None


class ShardPlan:
    """Contiguous row ranges per rank, aligned to ROW_ALIGN (pure host logic, CPU-testable)."""

    def __init__(self, m: int, world: int):
        per = device.round_up(-(-m // max(world, 1)), ROW_ALIGN)
        self.m, self.world = m, world
        self.bounds = [(min(r * per, m), min((r + 1) * per, m)) for r in range(world)]

    def rows(self, rank: int):
        return self.bounds[rank]


def dist_ctx():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_reduce_sum(t):
    """In-place integer SUM across ranks (NCCL on GPU tensors, gloo on CPU tensors in tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class _Trace:
    """BMF_FIT_TRACE=1: wall-clock phases of a fit (with a device sync at every mark), printed by rank 0."""

    def __init__(self):
        self.on = os.environ.get("BMF_FIT_TRACE", "0") == "1"
        self.t = time.perf_counter()
        self.rows = []

    def mark(self, name):
        if not self.on:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        self.rows.append((name, now - self.t))
        self.t = now

    def dump(self, rank):
        if self.on and rank == 0:
            agg = {}
            for name, dt in self.rows:
                agg[name] = agg.get(name, 0.0) + dt
            print("[fit trace] " + "  ".join("%s=%.1fms" % (k, 1e3 * v) for k, v in agg.items()), file=sys.stderr)


class CoverEngine:
    """All device buffers and kernel launches behind Asso.init_model / Asso._fit."""

    def __init__(self, X: sp.csr_matrix, w_fp: float, w_fn: float, scorer: str = "auto", assoc: str = "auto"):
        _native.require_gpu()
        self.trace = _Trace()
        self.rank, self.world = dist_ctx()
        self.m, self.n = X.shape
        self.plan = ShardPlan(self.m, self.world)
        r0, r1 = self.plan.rows(self.rank)
        self.r0, self.r1 = r0, r1
        self.m_loc = r1 - r0
        # Each rank takes only ITS rows (a view, no copy).  The packer ORs bits, so unsorted or duplicate column
        # indices need no host-side canonicalisation; explicitly stored zeros do matter, but scanning 1e8 values
        # takes the host ~0.1 s, so the scan runs in build_basis() WHILE the GPU packs and computes X^T X, and only
        # a matrix that really stores zeros (rare) is cleaned and rebuilt.  |X| is counted on the device.
        self._ctor_args = (w_fp, w_fn, scorer, assoc)
        Xl = device.csr_rows_view(X, r0, r1)
        self._host_rows = Xl
        self.trace.mark("host_csr")
        self.w_fp, self.w_fn = float(w_fp), float(w_fn)
        iw = integer_weights(self.w_fp, self.w_fn)
        self.wa, self.wb, self.shift = iw if iw else (0, 0, 0)
        self.integer_mode = iw is not None
        # scorer: 'auto' / 'tcgen05' pick the fastest exact tensor-core path (FP4 when the operand values are E2M1
        # numbers, else int8); 'tcgen05_i8' / 'tcgen05_f4' force one; 'popc' is the bit-packed AND+POPC variant
        operand = os.environ.get("BMF_OPERAND", "auto")
        if scorer in ("tcgen05_i8", "tcgen05_f4"):
            operand, scorer = scorer[-2:], "tcgen05"
        if scorer == "auto":
            scorer = "tcgen05"
        assert scorer in ("tcgen05", "popc") and assoc in ("auto", "tcgen05", "tcgen05_i8", "tcgen05_f4", "popc")
        assert operand in ("auto", "i8", "f4")
        self.scorer = scorer
        # operand encoding of the rows plane (see include/pybmf_b200.h):
        #   integer mode, all give D = wb*P - wa*N exactly:
        #   "zero"  : uncovered one -> wa+wb, uncovered zero -> 0, covered -> wa, bias wa*|b_j| in the epilogue
        #   "signed": uncovered one -> +wb,   uncovered zero -> -wa, covered -> 0
        #   general (non-dyadic) weights:
        #   "pq"    : interleaved 0/1 planes P = x & ~c and Q = c per 128 rows; the epilogue gets P and
        #             N = |b_j| - Q - P per element and evaluates the reference's fp64 row test literally
        enc = os.environ.get("BMF_PLANE_ENCODING", "zero")
        if enc == "zero" and self.wa + self.wb > 127:
            enc = "signed"
        if not self.integer_mode:
            enc = "pq"
        self.encoding = enc
        self.plane_sign = -1 if enc == "signed-" else 1
        self.cand_pop = None
        self._cand_pop_host = None
        # FP4 (tcgen05 kind::mxf4, twice the int8 rate): exact when the plane values 0 / wa / wa+wb are E2M1 numbers
        lib = _native.load()
        f4_ok = (enc == "pq") or (self.integer_mode and enc == "zero" and lib.bmf_e2m1_code(self.wa) >= 0
                                  and lib.bmf_e2m1_code(self.wa + self.wb) >= 0)       # the P/Q planes are 0/1: always fine
        if max(self.wa + self.wb, 1) * self.n >= (1 << 24):    # every partial sum must stay an exact FP32 integer
            f4_ok = False
        if operand == "f4" and scorer == "tcgen05" and not f4_ok:
            raise ValueError("the FP4 scorer needs integer weights whose values wa=%d and wa+wb=%d are E2M1 numbers "
                             "(0, 1, 2, 3, 4, 6) and (wa+wb)*n < 2^24" % (self.wa, self.wa + self.wb))
        self.operand = "f4" if (scorer == "tcgen05" and f4_ok and operand in ("auto", "f4")) else "i8"
        if assoc in ("tcgen05_i8", "tcgen05_f4"):
            self.assoc_operand, assoc = assoc[-2:], "tcgen05"
        else:
            self.assoc_operand = "i8" if operand == "i8" else "f4"
        if max(self.r1 - self.r0, 1) >= (1 << 24):             # co-occurrence counts must stay below 2^24 for FP32
            self.assoc_operand = "i8"
        self.assoc_kind = "tcgen05" if assoc == "auto" else assoc

        self.words = device.words_for(self.n)
        self.words_m = device.words_for(max(self.m_loc, 1))
        self.ld = device.round_up(self.n, 128)                  # K extent of the int8 cover planes (bytes)
        self.ld4 = device.round_up(self.n, 256) // 2            # ... of the packed FP4 planes (bytes)
        m_alloc = max(self.m_loc, 1)
        self.cnt = None
        self.launches = 0
        self._ip = self._ix = None
        if self.assoc_kind == "tcgen05" and self.assoc_operand == "f4" and self.m_loc > 0:
            self._upload_pack_associate(Xl)                     # chunked: H2D of chunk c+1 overlaps X^T X of chunk c
        else:
            self._ip, self._ix = device.upload_csr(Xl)
            self.x_bits = device.pack_csr(self._ip, self._ix, self.m_loc, self.n)
        ones = device.zeros((3,), torch.int64)
        if self.m_loc > 0:                                      # |X| of this rank's rows: TP of X against itself
            _native.call("bmf_confusion_bits", self.x_bits, self.x_bits, self.m_loc, self.words, -1, ones, None, None)
        all_reduce_sum(ones)
        self._ones = ones                                       # read (one sync) at the end of build_basis()
        self.sum_x = None
        self.trace.mark("h2d_pack")
        self.c_bits = device.zeros((m_alloc, self.words), torch.int64)
        self.tp_old = device.zeros((m_alloc,), torch.int32)
        self.fp_old = device.zeros((m_alloc,), torch.int32)
        self.alive = device.zeros((self.n,), torch.uint8)
        self.basis_bits = device.zeros((self.n, self.words), torch.int64)
        self.cand_pad = device.round_up(self.n, 256)           # multiple of 256 -> the 2-SM (cta_group::2) kernel
        self.cand_plane = None
        self.rows_plane = None
        self.gain_p = device.zeros((self.cand_pad,), torch.int64)
        self.gain_n = device.zeros((self.cand_pad,), torch.int64)
        self.record = device.zeros((8,), torch.int64)           # [winner, score bits, used, sumP, sumN]
        self.u_cols = []                                        # device bit vectors, one per chosen factor
        self.tp_tot = 0
        self.fp_tot = 0
        self.prescored = False

    # ---- association + basis (Asso.py:191-235) -------------------------------------------------
    def build_basis(self, tau: float, prescore: bool = False):
        """Association + basis (+ operand planes).  prescore=True also enqueues the first greedy step's scoring pass
        before the host-side stored-zero scan, so that the scan (0.09 s for 1e8 values) hides behind ~80 ms of GPU
        work; the caller must then skip its first score_all() (`self.prescored`)."""
        self._build_basis(tau)                                 # enqueued, not waited for
        if prescore:
            self.score_all()
        self.prescored = bool(prescore)
        Xl, self._host_rows = self._host_rows, None
        dirty = torch.tensor([1 if (Xl is not None and device.has_stored_zeros(Xl)) else 0], dtype=torch.int64,
                             device=device.dev())
        self.trace.mark("host_zero_scan")
        all_reduce_sum(dirty)                                  # every rank must take the same branch
        if int(dirty.item()):
            launches = self.launches
            clean = device.drop_stored_zeros(Xl)
            full = sp.csr_matrix((self.m, self.n), dtype=clean.dtype)      # this rank's rows in place, others empty
            ip = np.zeros(self.m + 1, dtype=np.int64)
            ip[self.r0 + 1:self.r1 + 1] = clean.indptr[1:]
            ip[self.r1 + 1:] = clean.indptr[-1]
            full.indptr, full.indices, full.data = ip, clean.indices, clean.data
            self.__init__(full, *self._ctor_args)
            self._host_rows = None
            self.launches += launches
            self._build_basis(tau)
            if prescore:
                self.score_all()
            self.prescored = bool(prescore)
        nb = int(self.alive.sum().item())
        self.sum_x = int(self._ones[0].item())
        self.trace.mark("basis_wait")
        return nb

    def _cnt_shape(self):
        n_pad = device.round_up(self.n, 256)
        return n_pad, max(n_pad, device.round_up(self.n, F4_ROW_PAD))   # the FP4 kernel walks data rows in super tiles of 496

    def _upload_pack_associate(self, Xl: sp.csr_matrix):
        """FP4 association path: the csr rows go up in a few chunks on a copy stream; as soon as a chunk has landed the
        main stream packs its bit rows, packs / expands its slice of X^T (K = the chunk's rows) and ACCUMULATES that
        slice's X^T X into cnt, while the (host-blocking, pageable) copy of the next chunk is in flight.  At c4 the copy
        (38 ms) and the association (36 ms) used to run back to back."""
        n, m_loc = self.n, self.m_loc
        n_pad, ldc = self._cnt_shape()
        d = device.dev()
        self.x_bits = device.zeros((m_loc, self.words), torch.int64)
        cnt = device.zeros((n_pad, ldc), torch.int32)
        nchunks = 3 if Xl.nnz >= (1 << 24) else 1
        step = device.round_up(-(-m_loc // nchunks), 256)      # chunk boundaries: multiples of 256 rows (K tiles, bit words)
        main = torch.cuda.current_stream()
        copy = torch.cuda.Stream() if nchunks > 1 else main
        indptr, indices = Xl.indptr, Xl.indices
        first = True
        for a in range(0, m_loc, step):
            b = min(a + step, m_loc)
            ia, ib = int(indptr[a]), int(indptr[b])
            ip_h = torch.from_numpy(np.ascontiguousarray((indptr[a:b + 1] - indptr[a]).astype(np.int64, copy=False)))
            ix_h = torch.from_numpy(np.ascontiguousarray(indices[ia:ib].astype(np.int32, copy=False)))
            with torch.cuda.stream(copy):
                ip_d, ix_d = ip_h.to(d, non_blocking=True), ix_h.to(d, non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(copy)
            main.wait_event(landed)
            ip_d.record_stream(main)
            ix_d.record_stream(main)
            rows = b - a
            if ib > ia:
                _native.call("bmf_pack_csr", ip_d, ix_d, rows, n, 0, self.x_bits[a:b], self.words)
                xt_bits = device.pack_csr(ip_d, ix_d, rows, n, transposed=True)            # [n, words(rows)]
                ldk = device.round_up(rows, 256) // 2
                xt_plane = device.empty((max(n_pad, ldc), ldk), torch.uint8)
                _native.call("bmf_expand_bits_f4", xt_bits, None, n, rows, xt_bits.shape[1], 2, 0, 0, xt_plane,
                             xt_plane.shape[0], ldk)
                _native.call("bmf_gemm_f4_nt", xt_plane, n_pad, xt_plane, device.round_up(n, F4_ROW_PAD), ldk, cnt, ldc,
                             0 if first else 1)
                first = False
                self.launches += 4
                del xt_plane, xt_bits
        self.cnt = cnt

    def _build_basis(self, tau: float):
        n, m_loc = self.n, self.m_loc
        n_pad, ldc = self._cnt_shape()
        cnt = self.cnt if self.cnt is not None else device.zeros((n_pad, ldc), torch.int32)
        if m_loc > 0 and self.cnt is None:
            xt_bits = device.pack_csr(self._ip, self._ix, m_loc, n, transposed=True)
            if self.assoc_kind == "tcgen05" and self.assoc_operand == "f4":
                # X^T as packed E2M1 0/1; A operand = rows padded to 256, B operand = the same plane padded to 496
                ldk = device.round_up(m_loc, 256) // 2
                xt_plane = device.empty((max(n_pad, ldc), ldk), torch.uint8)
                _native.call("bmf_expand_bits_f4", xt_bits, None, n, m_loc, xt_bits.shape[1], 2, 0, 0, xt_plane,
                             xt_plane.shape[0], ldk)
                _native.call("bmf_gemm_f4_nt", xt_plane, n_pad, xt_plane, device.round_up(n, F4_ROW_PAD), ldk, cnt, ldc, 0)
                del xt_plane
            elif self.assoc_kind == "tcgen05":
                xt_plane = device.expand_bits_i8(xt_bits, n, m_loc, 1, 0, 256)
                _native.call("bmf_assoc_counts_i8", xt_plane, n, n_pad, xt_plane.shape[1], cnt, ldc)
                del xt_plane
            else:
                _native.call("bmf_assoc_counts_popc", xt_bits, n, xt_bits.shape[1], cnt, ldc)
            self.launches += 3
            del xt_bits
        self.trace.mark("assoc_counts")
        all_reduce_sum(cnt)
        self.trace.mark("assoc_allreduce")
        self.cnt = cnt
        if self.scorer == "tcgen05" and self.operand == "i8":
            self.cand_plane = device.zeros((self.cand_pad, self.ld), torch.int8)
        self.cand_pop = device.zeros((self.cand_pad,), torch.int32)
        _native.call("bmf_basis_threshold", cnt, ldc, n, float(tau), self.basis_bits, self.words,
                     self.cand_plane, self.ld, self.alive, self.cand_pop)
        self.launches += 1
        if self.scorer == "tcgen05" and self.operand == "f4":   # candidate rows as packed E2M1 0/1
            self.cand_plane = device.empty((self.cand_pad, self.ld4), torch.uint8)
            _native.call("bmf_expand_bits_f4", self.basis_bits, None, n, n, self.words, 2, 0, 0, self.cand_plane,
                         self.cand_pad, self.ld4)
            self.launches += 1
        if self.scorer == "tcgen05":
            self._rebuild_rows_plane()
        self.trace.mark("basis_planes")

    def _rebuild_rows_plane(self):
        """rows_plane[i][k] = 0 if covered, +wb if x, -wa otherwise (the signed operand of D = wb*P - wa*N)."""
        if self.encoding == "pq" and self.operand == "f4":
            if self.rows_plane is None:
                self.rows_plane = device.empty((2 * device.round_up(max(self.m_loc, 1), 120), self.ld4), torch.uint8)
            if self.m_loc > 0:
                _native.call("bmf_expand_bits_pq_f4", self.x_bits, self.c_bits, self.m_loc, self.n, self.words,
                             self.rows_plane, self.ld4)
            self.launches += 1
            return
        if self.encoding == "pq":
            if self.rows_plane is None:
                self.rows_plane = device.empty((2 * device.round_up(max(self.m_loc, 1), 128), self.ld), torch.int8)
            if self.m_loc > 0:
                _native.call("bmf_expand_bits_pq", self.x_bits, self.c_bits, self.m_loc, self.n, self.words,
                             self.rows_plane, self.ld)
            self.launches += 1
            return
        one, zero, covered = self._plane_values()
        if self.operand == "f4":
            lib = _native.load()
            rows_pad = device.round_up(max(self.m_loc, 1), F4_ROW_PAD)
            if self.rows_plane is None:
                self.rows_plane = device.empty((rows_pad, self.ld4), torch.uint8)
            _native.call("bmf_expand_bits_f4", self.x_bits, self.c_bits, self.m_loc, self.n, self.words,
                         lib.bmf_e2m1_code(one), 0, lib.bmf_e2m1_code(covered), self.rows_plane, rows_pad, self.ld4)
            self.launches += 1
            return
        self.rows_plane = device.expand_bits_i8(self.x_bits, self.m_loc, self.n, one, zero, 256,
                                                mask=self.c_bits, out=self.rows_plane, masked=covered)
        self.launches += 1

    def _plane_values(self):
        """(uncovered one, uncovered zero, covered) byte values of the rows plane."""
        if self.encoding == "zero":
            return self.wa + self.wb, 0, self.wa
        sg = self.plane_sign
        return sg * self.wb, -sg * self.wa, 0

    def assoc_host(self):
        """The reference's `assoc` attribute (n x n float64) from the device counts."""
        n = self.n
        cnt = self.cnt[:n, :n].cpu().numpy().astype(np.float64)
        s = np.diag(cnt).copy()
        out = np.zeros_like(cnt)
        nz = s > 0
        out[nz] = cnt[nz] / s[nz][:, None]
        return out

    def basis_host(self):
        """Remaining candidate rows (alive only, original order) as uint8 [nb, n]."""
        B = device.bits_to_host(self.basis_bits, self.n)
        return B[self.alive.cpu().numpy().astype(bool)]

    # ---- one greedy step (Asso.py:62-110) ------------------------------------------------------
    def score_all(self):
        if self.m_loc == 0:                                   # a rank without rows only joins the exchange
            self.gain_p.zero_()
            self.gain_n.zero_()
        elif self.scorer == "tcgen05" and self.encoding == "pq" and self.operand == "f4":
            _native.call("bmf_cover_score_f4_general", self.cand_plane, self.cand_pad, self.rows_plane, self.m_loc,
                         self.ld4, self.cand_pop, self.tp_old, self.fp_old, self.w_fp, self.w_fn, self.gain_p,
                         self.gain_n)
        elif self.scorer == "tcgen05" and self.encoding == "pq":
            _native.call("bmf_cover_score_i8_general", self.cand_plane, self.cand_pad, self.rows_plane, self.m_loc,
                         self.ld, self.cand_pop, self.tp_old, self.fp_old, self.w_fp, self.w_fn, self.gain_p,
                         self.gain_n)
        elif self.scorer == "tcgen05" and self.operand == "f4":
            _native.call("bmf_cover_score_f4", self.cand_plane, self.cand_pad, self.rows_plane,
                         self.rows_plane.shape[0], self.ld4, self.cand_pop, self.wa, self.gain_p)
        elif self.scorer == "tcgen05":
            _native.call("bmf_cover_score_i8", self.cand_plane, self.cand_pad, self.rows_plane,
                         self.rows_plane.shape[0], self.ld, self.plane_sign,
                         self.cand_pop if self.encoding == "zero" else None, self.wa, self.gain_p)
        else:
            _native.call("bmf_cover_score_popc", self.x_bits, self.c_bits, self.m_loc, self.n, self.words,
                         self.basis_bits, self.alive, self.tp_old, self.fp_old, self.wa, self.wb, self.w_fp,
                         self.w_fn, self.gain_p, self.gain_n)
        self.launches += 1
        all_reduce_sum(self.gain_p)
        if not self.integer_mode:
            all_reduce_sum(self.gain_n)

    def select_and_apply(self, best_score: float):
        """argmax + apply without a host round trip in between; one small D2H read at the end.
        Returns (winner, score, n_used, sum_p, sum_n) with winner = -1 when nothing beats best_score."""
        base_int = self.wb * self.tp_tot - self.wa * self.fp_tot
        scale = 1.0 / float(1 << self.shift)
        self.record.zero_()
        _native.call("bmf_select_first_max", self.gain_p, self.gain_n if not self.integer_mode else None, self.alive,
                     self.n, self.wa, self.wb, base_int, scale, self.w_fp, self.w_fn, self.tp_tot, self.fp_tot,
                     float(best_score), self.record)
        u_bits = device.zeros((self.words_m,), torch.int64)
        if self.m_loc > 0 and self.scorer == "tcgen05" and self.encoding == "pq" and self.operand == "f4":
            _native.call("bmf_cover_apply_f4_general", self.x_bits, self.c_bits, self.m_loc, self.n, self.words,
                         self.basis_bits, self.alive, self.record, self.tp_old, self.fp_old, self.w_fp, self.w_fn,
                         self.rows_plane, self.ld4, u_bits, self.record[2:5])
        elif self.m_loc > 0 and self.scorer == "tcgen05" and self.encoding == "pq":
            _native.call("bmf_cover_apply_general", self.x_bits, self.c_bits, self.m_loc, self.n, self.words,
                         self.basis_bits, self.alive, self.record, self.tp_old, self.fp_old, self.w_fp, self.w_fn,
                         self.rows_plane, self.ld, u_bits, self.record[2:5])
        elif self.m_loc > 0 and self.scorer == "tcgen05" and self.operand == "f4":
            _native.call("bmf_cover_apply_f4", self.x_bits, self.c_bits, self.m_loc, self.n, self.words,
                         self.basis_bits, self.alive, self.record, self.tp_old, self.fp_old, self.wa, self.wb,
                         self.rows_plane, self.ld4, _native.load().bmf_e2m1_code(self.wa), u_bits, self.record[2:5])
        elif self.m_loc > 0:
            _native.call("bmf_cover_apply", self.x_bits, self.c_bits, self.m_loc, self.n, self.words, self.basis_bits,
                         self.alive, self.record, self.tp_old, self.fp_old, self.wa, self.wb, self.w_fp, self.w_fn,
                         self.rows_plane, self.ld, self._plane_values()[2] if self.scorer == "tcgen05" else 0,
                         u_bits, self.record[2:5])
        self.launches += 2
        if self.world > 1:
            all_reduce_sum(self.record[2:5])
        rec = self.record.cpu().numpy()
        winner = int(rec[0])
        if winner < 0:
            return -1, float(best_score), 0, 0, 0
        if self.m_loc == 0:
            self.alive[winner] = 0                            # bmf_cover_apply does this on ranks that own rows
        score = float(rec[1:2].view(np.float64)[0])
        used, sp_, sn_ = int(rec[2]), int(rec[3]), int(rec[4])
        self.u_cols.append(u_bits)
        self.tp_tot += sp_
        self.fp_tot += sn_
        return winner, score, used, sp_, sn_

    def basis_row_host(self, j: int) -> np.ndarray:
        return device.words_to_dense(self.basis_bits[j:j + 1].cpu().numpy(), self.n)[0]

    def basis_rows_host(self, js) -> np.ndarray:
        idx = torch.as_tensor(list(js), dtype=torch.int64, device=device.dev())
        return device.words_to_dense(self.basis_bits[idx].cpu().numpy(), self.n)

    def cand_pop_host(self, j: int) -> int:
        """|b_j| from a host copy of the per-candidate popcounts (one D2H per fit instead of one per greedy step)."""
        if self._cand_pop_host is None:
            self._cand_pop_host = self.cand_pop[: self.n].cpu().numpy()
        return int(self._cand_pop_host[j])

    def gather_used_words(self, ids):
        """Bit columns `ids` of U over ALL ranks' rows, still packed: a list of (uint64 words [len(ids), w_r], rows_r),
        one entry per rank in row order (one device all-gather + D2H; unpacking is host work, see unpack_used)."""
        if not ids:
            return []
        local = torch.stack([self.u_cols[i] for i in ids])                       # [c, words_m]
        if self.world == 1:
            return [(local.cpu().numpy(), self.m_loc)]
        import torch.distributed as dist
        wmax = device.words_for(self.plan.rows(0)[1] - self.plan.rows(0)[0])
        padded = device.zeros((len(ids), wmax), torch.int64)
        padded[:, : local.shape[1]] = local
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded)
        both = torch.stack(parts).cpu().numpy()                                  # one D2H
        return [(both[r], self.plan.rows(r)[1] - self.plan.rows(r)[0]) for r in range(self.world)]

    @staticmethod
    def unpack_used(parts, ncols_live) -> np.ndarray:
        """Packed parts of gather_used_words -> uint8 [m, ncols_live]."""
        if not parts:
            return np.zeros((0, ncols_live), np.uint8)
        return np.ascontiguousarray(np.concatenate([device.words_to_dense(w, rows).T for w, rows in parts], axis=0))

    def gather_used_columns(self, ids) -> np.ndarray:
        """Columns `ids` of U over ALL ranks' rows as uint8 [m, len(ids)]."""
        if not ids:
            return np.zeros((self.m, 0), np.uint8)
        return self.unpack_used(self.gather_used_words(ids), len(ids))

    def basis_words_host(self, js) -> np.ndarray:
        idx = torch.as_tensor(list(js), dtype=torch.int64, device=device.dev())
        return self.basis_bits[idx].cpu().numpy()

    # ---- cover rebuilt from a factor list (after the reference's truncation quirk D1) -----------
    def reset_cover(self, kept):
        """Recompute c_bits / tp_old / fp_old / rows_plane from the factors in `kept`, a list of
        (u_col index, basis row index): the reference recomputes X_pd from the (possibly
        truncated) U, V at the top of every step (Asso.py:80, BaseModelTools.py:392-393)."""
        self.c_bits.zero_()
        self.tp_old.zero_()
        self.fp_old.zero_()
        self.tp_tot = self.fp_tot = 0
        if kept and self.m_loc > 0:
            k = len(kept)
            kw = (k + 63) // 64
            cols = [device.words_to_dense(self.u_cols[ui].cpu().numpy().reshape(1, -1), self.m_loc)[0]
                    for (ui, _j) in kept]
            uw = torch.from_numpy(device.dense_to_words(np.stack(cols, axis=1), words=kw)).to(device.dev())
            vt = torch.stack([self.basis_bits[j] for (_ui, j) in kept]).contiguous()
            _native.call("bmf_bool_product", uw, self.m_loc, kw, vt, k, self.words, self.c_bits)
            counts = device.zeros((3,), torch.int64)
            _native.call("bmf_confusion_bits", self.x_bits, self.c_bits, self.m_loc, self.words, -1, counts,
                         self.tp_old, self.fp_old)
            self.launches += 2
            all_reduce_sum(counts)
            c = counts.cpu().numpy()
            self.tp_tot, self.fp_tot = int(c[0]), int(c[1])
        if self.scorer == "tcgen05":
            self._rebuild_rows_plane()
