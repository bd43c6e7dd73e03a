#!/usr/bin/env python
"""Accuracy of the TF32 modes against the FP64 mode of the same engine, as a function of the shape (device-generated
data, same seed in every mode): norm-relative error of W and H after `iters` fit iterations, relative error of the
objective.  The tensor core accumulates FP32 with truncation, so long contractions of non-negative terms (f = 8192
features, thousands of samples) drift further from float64 than the small parity cases do."""
import sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from multimodal_b200 import _native
from oracle import cases, klnmf_oracle as O

def run(n, f, k, mode, iters, fit=True, sparse=0):
    np.random.seed(11)
    H0 = O.init_dictionary(k, f)
    with _native.Engine(n, f, k, mode=mode) as e:
        if sparse:
            e.fill_csr_synthetic(sparse, 9)
        else:
            e.fill_dense_synthetic(5)
        e.set_dictionary(H0)
        e.init_coefficients()
        errs, _ = e.run(iters, 0.0, fit)
        return e.get_coefficients(), e.get_dictionary(), np.asarray(errs)

if __name__ == "__main__":
    iters = 10
    for (n, f, k) in [(2048, 512, 64), (8192, 1024, 256), (8192, 4096, 256), (8192, 8192, 512), (65536, 8192, 512), (262144, 2048, 128)]:
        ref = run(n, f, k, "fp64", iters)
        for mode in ("tf32x3", "tf32r", "tf32"):
            W, H, e = run(n, f, k, mode, iters)
            print("n=%6d f=%5d k=%3d %-6s  W %.2e  H %.2e  KL %.2e" % (n, f, k, mode, cases.rel_fro(W, ref[0]), cases.rel_fro(H, ref[1]),
                  np.max(np.abs(e - ref[2]) / np.abs(ref[2]))), flush=True)
