#!/usr/bin/env python
"""One-time cost of the blocked-CSC copy of the pattern (sparse.cu: bcsc_build) at cfg4's shape: the first fit iteration
of a data set against the later ones.   python tools/bcsc_build_time.py [n]"""
import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np
from multimodal_b200 import _native
n, f, k = (int(sys.argv[1]) if len(sys.argv) > 1 else 2000000), 50000, 256
np.random.seed(0)
H0 = np.abs(np.random.random((k, f))) + .01
H0 /= H0.sum(axis=1, keepdims=True)
with _native.Engine(65536, f, k, mode="tf32r") as e:      # warm the process: kernel images, attributes, allocator
    e.fill_csr_synthetic(250, 3)
    e.set_dictionary(H0)
    e.init_coefficients()
    e.run(2, 0.0, True)
with _native.Engine(n, f, k, mode="tf32r") as e:
    e.fill_csr_synthetic(250, 3)
    e.set_dictionary(H0)
    e.init_coefficients()
    for i in range(3):
        t0 = time.perf_counter()
        e.run(1, 0.0, True)
        print("run(1) call %d: %.1f ms" % (i, (time.perf_counter() - t0) * 1e3))
