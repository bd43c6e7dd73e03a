#!/usr/bin/env python
"""A wide dense modality next to a sparse one: the hybrid stack (dense block through the contraction engine, CSR block
through the sparse passes; api.cu: HybridSide) against the all-CSR stack the reference would build (array_utils.py:5-9).

    python tools/hybrid_vs_csr.py [n f_dense f_sparse nnz_per_row k iters]
"""
import os
import sys
import time

os.environ.setdefault("KLNMF_PROFILE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from multimodal_b200 import _native
from oracle import klnmf_oracle as O

args = [int(a) for a in sys.argv[1:]]
n, fd, fs, m, k, iters = (args + [65536, 4096, 50000, 250, 256, 10][len(args):])[:6]
rs = np.random.RandomState(0)
dense = rs.random_sample((n, fd)).astype(np.float32)
dense[dense < 0.1] = 0.0                                        # 10 % structural zeros in the dense modality
# stratified pattern: the t-th stored entry of a row falls into the t-th of m equal column strata (sorted, distinct)
lo = (np.arange(m) * fs) // m
width = ((np.arange(m) + 1) * fs) // m - lo
idx = (lo[None, :] + (rs.random_sample((n, m)) * width[None, :]).astype(np.int64)).astype(np.int32)
vals = (1.0 - rs.random_sample((n, m))).astype(np.float32)
sparse = sp.csr_matrix((vals.ravel(), idx.ravel(), np.arange(n + 1, dtype=np.int64) * m), shape=(n, fs))
f = fd + fs
np.random.seed(1)
H0 = O.init_dictionary(k, f)
res = {}
for name, hyb in (("all-CSR stack", "0"), ("hybrid stack", "1")):
    os.environ["KLNMF_HYBRID"] = hyb
    with _native.Engine(n, f, k, mode="tf32r") as e:
        t0 = time.perf_counter()
        e.set_stacked_blocks([dense, sparse], [1.0, 1.0])
        t1 = time.perf_counter()
        assert e.is_hybrid() == (hyb == "1")
        e.set_dictionary(H0)
        e.init_coefficients()
        e.run(2, 0.0, True)
        errs, _ = e.run(iters, 0.0, True)
        ms, _ = e.last_run_profile()
        res[name] = (e.get_dictionary(), np.asarray(errs))
        print("%-14s n=%d dense %d + CSR %d (%d per row) k=%d: %8.3f ms/iteration (rows+dense half %.3f, numerators %.3f, dictionary %.3f); "
              "data set in %.2f s" % (name, n, fd, fs, m, k, ms["total"] / iters, ms["ratio"] / iters, ms["numerator"] / iters,
                                      ms["dictionary"] / iters, t1 - t0))
a, b = res["all-CSR stack"], res["hybrid stack"]
print("hybrid vs all-CSR: dictionary %.2e (rel. Frobenius), objective %.2e (max rel.)"
      % (np.linalg.norm(a[0] - b[0]) / np.linalg.norm(a[0]), np.max(np.abs(a[1] - b[1]) / np.abs(a[1]))))
