"""Where the end-to-end time of KLdivNMF.fit_transform on a pinned host array goes (cfg5 shape, row subsample)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_b200 import _native

n, f, k, steps = int(os.environ.get("N", 1000000)), 8192, 512, 5
MODE = os.environ.get("MODE", _native.DEFAULT_MODE)
WANT_W = os.environ.get("WANT_W", "0") == "1"
X = torch.empty((n, f), dtype=torch.float32, pin_memory=True).numpy()
with _native.Engine(n, f, k, mode=MODE) as e:
    e.fill_dense_synthetic(1)
    e.get_dense(X)
np.random.seed(0)
H0 = np.abs(np.random.random((k, f))) + .01
H0 /= H0.sum(axis=1, keepdims=True)
for rep in range(2):
    t = [time.perf_counter()]
    eng = _native.Engine(n, f, k, mode=MODE); t.append(time.perf_counter())
    eng.set_dense(X); eng.check_input(); t.append(time.perf_counter())
    eng.set_dictionary(H0); eng.init_coefficients(); t.append(time.perf_counter())
    eng.run(steps, 0.0, True); t.append(time.perf_counter())
    W = eng.get_coefficients() if WANT_W else np.zeros((1, 1)); t.append(time.perf_counter())
    H = eng.get_dictionary(); t.append(time.perf_counter())
    eng.close(); t.append(time.perf_counter())
    names = ["create", "set_dense+check (H2D %.1f GB)" % (X.nbytes / 1e9), "dictionary+W0", "run %d it" % steps,
             "get_coefficients (D2H %.1f GB f64)" % (W.nbytes / 1e9), "get_dictionary", "close"]
    print("rep", rep, " ".join("%s=%.3fs" % (a, b - c) for a, b, c in zip(names, t[1:], t[:-1])), "total=%.3fs" % (t[-1] - t[0]))
