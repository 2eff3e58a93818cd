"""Device-side containers for the Asso path: bit-packed matrices and int8 operand planes in HBM.

Layout (see include/pybmf_b200.h): a bit matrix is a row-major torch.int64 tensor
[rows, words] (uint64 words viewed as int64), bit c of a row in word c>>6 at position
c&63, `words` even so that every row is 16-byte aligned; pad bits are zero.
torch supplies memory and streams only -- every operation on these buffers is a
kernel of libbmf_b200.so.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from . import _native


def words_for(ncols: int) -> int:
    """Even number of uint64 words covering `ncols` bits (>= 2)."""
    w = (int(ncols) + 63) // 64
    w += w & 1
    return max(w, 2)


def round_up(x: int, mult: int) -> int:
    return ((int(x) + mult - 1) // mult) * mult


def dev():
    return torch.device("cuda", torch.cuda.current_device())


def zeros(shape, dtype):
    return torch.zeros(shape, dtype=dtype, device=dev())


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=dev())


# ---- host <-> bit words (numpy side; used for inputs, outputs and tests) --------------------
def dense_to_words(A: np.ndarray, words: int | None = None) -> np.ndarray:
    """0/1 array [rows, ncols] -> int64 words [rows, words] in the library's bit order."""
    A = (np.asarray(A) != 0).astype(np.uint8)
    rows, ncols = A.shape
    words = words_for(ncols) if words is None else words
    padded = np.zeros((rows, words * 64), dtype=np.uint8)
    padded[:, :ncols] = A
    return np.ascontiguousarray(np.packbits(padded, axis=1, bitorder="little")).view(np.int64).reshape(rows, words)


def words_to_dense(W: np.ndarray, ncols: int) -> np.ndarray:
    W = np.ascontiguousarray(W).view(np.uint8).reshape(W.shape[0], -1)
    return np.unpackbits(W, axis=1, bitorder="little")[:, :ncols]


def to_csr_pattern(X) -> sp.csr_matrix:
    """Non-zero pattern of X as canonical csr (sorted, no duplicates, no explicit zeros).
    No copy is made when X is already a canonical csr without stored zeros (the common case)."""
    X = X if sp.isspmatrix_csr(X) else sp.csr_matrix(X)
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()
    if X.nnz and np.count_nonzero(X.data) != X.nnz:
        X = X.copy()
        X.eliminate_zeros()
    return X


def csr_rows_view(X, r0: int, r1: int) -> sp.csr_matrix:
    """Rows [r0, r1) of a csr matrix WITHOUT copying indices / data (scipy's X[r0:r1] copies both: 0.4 s for half
    of a 1e8-nnz matrix); only the (r1 - r0 + 1) row pointers are rebased."""
    X = X if sp.isspmatrix_csr(X) else sp.csr_matrix(X)
    if r0 == 0 and r1 == X.shape[0]:
        return X
    a, b = int(X.indptr[r0]), int(X.indptr[r1])
    indptr = (X.indptr[r0:r1 + 1] - X.indptr[r0]).astype(np.int64, copy=False)
    out = sp.csr_matrix((r1 - r0, X.shape[1]), dtype=X.dtype)
    out.indptr, out.indices, out.data = indptr, X.indices[a:b], X.data[a:b]
    return out


def has_stored_zeros(X: sp.csr_matrix) -> bool:
    """O(nnz) host scan of the value array (the pattern kernels treat every STORED entry as a one).  Memory bound:
    large arrays are scanned by a few threads (numpy releases the GIL inside count_nonzero)."""
    if not X.nnz:
        return False                                            # (a bool matrix can store explicit False entries too)
    data = X.data
    if data.size < (1 << 22):
        return np.count_nonzero(data) != data.size
    import os
    from concurrent.futures import ThreadPoolExecutor
    nthreads = max(1, min(8, (os.cpu_count() or 1)))
    bounds = np.linspace(0, data.size, nthreads + 1, dtype=np.int64)
    with ThreadPoolExecutor(nthreads) as pool:
        counts = list(pool.map(lambda ab: int(np.count_nonzero(data[ab[0]:ab[1]])), zip(bounds[:-1], bounds[1:])))
    return sum(counts) != data.size


def drop_stored_zeros(X: sp.csr_matrix) -> sp.csr_matrix:
    X = X.copy()
    X.eliminate_zeros()
    return X


# ---- device packing --------------------------------------------------------------------------
def upload_csr(X: sp.csr_matrix):
    """H2D copy of the pattern arrays (indptr int64, indices int32); values are never read."""
    d = dev()
    indptr = torch.from_numpy(np.ascontiguousarray(X.indptr.astype(np.int64, copy=False)))
    indices = torch.from_numpy(np.ascontiguousarray(X.indices.astype(np.int32, copy=False)))
    return indptr.to(d, non_blocking=True), indices.to(d, non_blocking=True)


class _Stager:
    """Host -> device copies of LARGE pageable arrays through a small pinned ring, filled by a few threads.

    `tensor.to(device)` from pageable memory is staged by the driver on ONE thread: 11 GB/s on this pool's hosts
    (35 ms for the 400 MB index array of the Netflix-shaped config).  A threaded memcpy into pinned slots followed by
    async copies reaches the PCIe rate (profiles/r02j_h2d_pinned_probe.log: 5.3 ms memcpy on 8 threads + 7.2 ms DMA,
    overlapped slot by slot).  The ring is 4 x 16 MB: pinning it costs ~40 ms ONCE per process (pinning the whole 400 MB
    would cost 240 ms, registering the array in place 60 ms per fit -- both measured, both worse)."""
    SLOT = 16 << 20
    SLOTS = 4

    def __init__(self, threads):
        import os
        from concurrent.futures import ThreadPoolExecutor
        self.SLOT = int(os.environ.get("BMF_STAGE_SLOT_MB", self.SLOT >> 20)) << 20      # experiments: slot size / thread count
        threads = int(os.environ.get("BMF_STAGE_THREADS", threads))
        self.buf = torch.empty(self.SLOT * self.SLOTS, dtype=torch.uint8, pin_memory=True)
        self.view = self.buf.numpy()
        self.events = [None] * self.SLOTS
        self.next = 0
        self.threads = threads
        self.pool = ThreadPoolExecutor(threads) if threads > 1 else None

    def upload(self, src: np.ndarray, dst, stream):
        """src (contiguous numpy) -> dst (device tensor of the same byte size) on `stream`; returns when every piece has been
        ENQUEUED (the last DMA may still be in flight: order later work on `stream`)."""
        src_u8 = src.reshape(-1).view(np.uint8)
        dst_u8 = dst.reshape(-1).view(torch.uint8)
        nbytes = src_u8.size
        assert dst_u8.numel() == nbytes
        for off in range(0, nbytes, self.SLOT):
            s = self.next
            self.next = (s + 1) % self.SLOTS
            if self.events[s] is not None:
                self.events[s].synchronize()                       # the slot's previous DMA has drained
            ln = min(self.SLOT, nbytes - off)
            base = s * self.SLOT
            if self.pool is None or ln < (1 << 20):
                np.copyto(self.view[base:base + ln], src_u8[off:off + ln])
            else:
                cuts = np.linspace(0, ln, self.threads + 1, dtype=np.int64)
                list(self.pool.map(lambda ab: np.copyto(self.view[base + ab[0]:base + ab[1]], src_u8[off + ab[0]:off + ab[1]]),
                                   zip(cuts[:-1], cuts[1:])))
            with torch.cuda.stream(stream):
                dst_u8[off:off + ln].copy_(self.buf[base:base + ln], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            self.events[s] = ev


_STAGER = None
_COPY_STREAM = {}


def copy_stream():
    """The process-wide side stream for host -> device copies (one per device)."""
    idx = torch.cuda.current_device()
    if idx not in _COPY_STREAM:
        _COPY_STREAM[idx] = torch.cuda.Stream(device=idx)
    return _COPY_STREAM[idx]



def stager(world: int = 1):
    """The process-wide pinned staging ring (created on first use); None when BMF_PINNED_UPLOAD=0."""
    global _STAGER
    import os
    if os.environ.get("BMF_PINNED_UPLOAD", "1") == "0":
        return None
    if _STAGER is None:
        cores = os.cpu_count() or 1
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            pass
        _STAGER = _Stager(max(1, min(8, cores // max(world, 1))))      # the ranks of one box share its cores
    return _STAGER


def pack_csr(indptr_d, indices_d, m: int, n: int, transposed: bool = False):
    """CSR pattern on device -> bit matrix [m, words(n)] (or X^T: [n, words(m)])."""
    rows, cols = (n, m) if transposed else (m, n)
    words = words_for(cols)
    bits = zeros((max(rows, 1), words), torch.int64)
    if m > 0 and indices_d.numel() > 0:
        _native.call("bmf_pack_csr", indptr_d, indices_d, m, n, 1 if transposed else 0, bits, words)
    return bits


def expand_bits_i8(bits, rows: int, ncols: int, one: int, zero: int, row_tile: int, mask=None, out=None, masked=0):
    """Bit matrix -> int8 plane [round_up(rows,row_tile), round_up(ncols,128)]; `masked` where `mask` is set."""
    ld = round_up(max(ncols, 1), 128)
    rows_pad = round_up(max(rows, 1), row_tile)
    plane = empty((rows_pad, ld), torch.int8) if out is None else out
    _native.call("bmf_expand_bits_i8", bits, mask, rows, ncols, bits.shape[1], one, zero, masked, plane, rows_pad, ld)
    return plane


def bits_to_host(bits, ncols: int) -> np.ndarray:
    return words_to_dense(bits.cpu().numpy(), ncols)
