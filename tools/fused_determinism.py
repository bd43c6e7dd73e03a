#!/usr/bin/env python
"""Race hunt: the fused half-step kernels must be bit-reproducible (fixed summation order).  Runs every shape / variant
several times and reports where repeated runs differ (row blocks, columns) and how far each is from the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_b200 import _native  # noqa: E402
from oracle import cases, klnmf_oracle as O  # noqa: E402

SHAPES = [(513, 333, 100), (4096, 2048, 128), (700, 1000, 50), (4096, 2048, 256), (513, 333, 256), (1300, 2100, 130)]
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 6


def run(X, H, iters, mode, ts, fit):
    os.environ["KLNMF_FUSED"] = "1"
    os.environ["KLNMF_FUSED_TS"] = "1" if ts else "0"
    n, f = X.shape
    with _native.Engine(n, f, H.shape[0], mode=mode) as e:
        e.set_dense(X)
        e.set_dictionary(H)
        e.init_coefficients()
        e.run(iters, 0.0, fit)
        return e.get_coefficients()


for (n, f, k) in SHAPES:
    rs = np.random.RandomState(n + f + k)
    X = rs.random_sample((n, f))
    X[rs.random_sample((n, f)) < 0.2] = 0.0
    np.random.seed(5)
    H = O.init_dictionary(k, f)
    Wr = np.asarray(X.dot(H.T))
    for _ in range(6):
        Wr, _ = O.update(X, Wr, H, fit=False)
    for mode in ("tf32r", "tf32"):
        for ts in ((True, False) if k <= 128 else (True,)):
            for fit in (False, True):
                outs = [run(X, H, 6, mode, ts, fit) for _ in range(REPS)]
                bad = []
                for i, W in enumerate(outs[1:], 1):
                    d = np.argwhere(W != outs[0])
                    if len(d):
                        rows = np.unique(d[:, 0])
                        cols = np.unique(d[:, 1])
                        bad.append("rep%d: %d entries, rows %d..%d (%d rows, blocks %s), cols %d..%d, max rel %.2e" % (
                            i, len(d), rows.min(), rows.max(), len(rows), sorted(set((rows // 128).tolist()))[:8],
                            cols.min(), cols.max(), np.max(np.abs(W - outs[0]) / (np.abs(outs[0]) + 1e-30))))
                err = [cases.rel_fro(W, Wr) for W in outs] if not fit else []
                print("n=%d f=%d k=%d %-5s %s %s: %s  err_vs_oracle %s" % (
                    n, f, k, mode, "tmem" if ts else "smem", "fit" if fit else "transform",
                    "REPRODUCIBLE" if not bad else "DIFFERS " + " | ".join(bad[:3]),
                    " ".join("%.2e" % e for e in err)), flush=True)
