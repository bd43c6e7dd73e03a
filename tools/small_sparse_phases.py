#!/usr/bin/env python
"""Per-phase device times of small CSR fits (cfg2's shape: 1000 x 110 450, 1550 stored entries per row, k = 50) --
the regime the reference's own experiments live in (icdl2013: ~1000 samples).

    python tools/small_sparse_phases.py [n f nnz_per_row k iters]
"""
import os
import sys
import time

os.environ.setdefault("KLNMF_PROFILE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multimodal_b200 import _native
from oracle import klnmf_oracle as O

args = [int(a) for a in sys.argv[1:]]
n, f, m, k, iters = (args + [1000, 110450, 1550, 50, 50][len(args):])[:5]
np.random.seed(0)
H0 = O.init_dictionary(k, f)
for mode in ("tf32r", "fp64"):
    with _native.Engine(n, f, k, mode=mode) as e:
        e.fill_csr_synthetic(m, 3)
        e.set_dictionary(H0)
        e.init_coefficients()
        e.run(2, 0.0, True)                      # builds the blocked-CSC copy
        for fit in (True, False):
            t0 = time.perf_counter()
            e.run(iters, 0.0, fit)
            wall = time.perf_counter() - t0
            ms, cnt = e.last_run_profile()
            print("n=%d f=%d nnz/row=%d k=%d %s %-9s %7.3f ms/it (wall %7.3f)  rows %.3f numerator %.3f dictionary %.3f"
                  % (n, f, m, k, mode, "fit" if fit else "transform", ms["total"] / iters, wall * 1e3 / iters,
                     ms["ratio"] / iters, ms["numerator"] / iters, ms["dictionary"] / iters))
