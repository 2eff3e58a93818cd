"""`result_digest(model)`: a compact, order-sensitive fingerprint of a finished Asso fit.

bench.py prints it at every GPU count and tests/test_asso_gpu.py compares it with the fixtures that the CPU
restatement wrote (tests/golden/c*_digest.json) -- the definition is shared, the code is not:
  winners    : per greedy step, the ORIGINAL column index of the chosen association row (Asso.py:94-107)
  score_bits : the float64 score of the step as 16 hex digits (little-endian bytes)
  used       : number of data rows that use the new factor
  tp, fp     : cumulative TP / FP of the cover after the step
  u_sha256   : SHA-256 over the columns of U in factor order, each packed LSB-first into ceil(m/8) bytes
  v_sha256   : the same over the columns of V (ceil(n/8) bytes each)
"""
from __future__ import annotations

import hashlib

import numpy as np


def _column_bytes_dense(col) -> bytes:
    return np.packbits((np.asarray(col).ravel() != 0).astype(np.uint8), bitorder="little").tobytes()


def result_digest(model) -> dict:
    steps = list(getattr(model, "fit_steps_", []))
    d = model.__dict__
    hu, hv = hashlib.sha256(), hashlib.sha256()
    if "_host_factors" in d and "U" not in d:                  # still bit-packed on the host: hash without unpacking
        packed, ncols = d["_host_factors"]
        m, n = model.m, model.n
        zero_u, zero_v = bytes((m + 7) // 8), bytes((n + 7) // 8)
        pos = {} if packed is None else {int(p): i for i, p in enumerate(packed[2])}
        for c in range(ncols):
            if c not in pos:
                hu.update(zero_u)
                hv.update(zero_v)
                continue
            i = pos[c]
            for words, rows in packed[0]:                      # one part per rank, row ranges are multiples of 256
                hu.update(np.ascontiguousarray(words[i]).tobytes()[: (rows + 7) // 8])
            hv.update(np.ascontiguousarray(packed[1][i]).tobytes()[: (n + 7) // 8])
    else:
        U, V = model.U.tocsc(), model.V.tocsc()
        for c in range(U.shape[1]):
            hu.update(_column_bytes_dense(U[:, c].toarray()))
            hv.update(_column_bytes_dense(V[:, c].toarray()))
    return {"winners": [int(s["winner"]) for s in steps],
            "score_bits": [np.float64(s["score"]).tobytes().hex() for s in steps],
            "used": [int(s["used"]) for s in steps],
            "tp": [int(s["tp"]) for s in steps], "fp": [int(s["fp"]) for s in steps],
            "u_sha256": hu.hexdigest(), "v_sha256": hv.hexdigest()}


DIGEST_KEYS = ("winners", "score_bits", "used", "tp", "fp", "u_sha256", "v_sha256")


def digest_matches(got: dict, want: dict, steps=None) -> bool:
    """Equality on the digest keys; steps=K compares only the first K greedy steps (hashes are then skipped)."""
    if steps is None:
        return all(got[k] == want[k] for k in DIGEST_KEYS)
    return all(list(got[k])[:steps] == list(want[k])[:steps] for k in DIGEST_KEYS[:5])
