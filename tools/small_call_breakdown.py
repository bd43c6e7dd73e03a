#!/usr/bin/env python
"""Where the time of a SMALL call goes (cfg2's stack: 1000 x 110 450 CSR, k = 50, 50 iterations): the reference's own
experiments make hundreds of such calls, and their device time (0.46 ms per iteration) is a fraction of the call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from multimodal_b200 import _native
from multimodal_b200.lib.nmf import KLdivNMF
from oracle import klnmf_oracle as O

n, f, k, iters = 1000, 110450, 50, 50
rs = np.random.RandomState(0)
X = sp.random(n, f, density=0.014, random_state=rs, format="csr", dtype=np.float64)
X.data = np.ceil(5 * X.data)
np.random.seed(0)
H0 = O.init_dictionary(k, f)
for rep in range(3):
    t = [time.perf_counter()]
    eng = _native.Engine(n, f, k, mode=_native.DEFAULT_MODE); t.append(time.perf_counter())
    eng.set_csr(X); t.append(time.perf_counter())
    eng.check_input(); t.append(time.perf_counter())
    eng.set_dictionary(H0); t.append(time.perf_counter())
    eng.init_coefficients(); t.append(time.perf_counter())
    eng.run(iters, 0.0, True); t.append(time.perf_counter())
    W = eng.get_coefficients(); t.append(time.perf_counter())
    H = eng.get_dictionary(); t.append(time.perf_counter())
    eng.close(); t.append(time.perf_counter())
    names = ["create", "set_csr", "check", "set_dictionary", "init_coefficients", "run %d it" % iters, "get_coefficients",
             "get_dictionary", "close"]
    print("rep", rep, " ".join("%s=%.1fms" % (a, (b - c) * 1e3) for a, b, c in zip(names, t[1:], t[:-1])),
          "total=%.1fms" % ((t[-1] - t[0]) * 1e3))
for rep in range(3):
    t0 = time.perf_counter()
    np.random.seed(1)
    est = KLdivNMF(n_components=k, max_iter=iters, tol=0)
    W = est.fit_transform(X)
    t1 = time.perf_counter()
    W2 = est.transform(X[:100])
    t2 = time.perf_counter()
    print("rep", rep, "KLdivNMF.fit_transform %.1f ms   transform(100 rows, %d it) %.1f ms" % ((t1 - t0) * 1e3, iters, (t2 - t1) * 1e3))
