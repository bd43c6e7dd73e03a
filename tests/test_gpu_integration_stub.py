"""-m gpu: the binding INTEGRATION.md shows a maintainer of the reference (Option B: a raw ctypes stub that replaces the
loop of nmf.py:201-228) is executed as it is printed there -- only the library path is filled in -- against the oracle."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from multimodal_b200 import build
from oracle import cases
from oracle import klnmf_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_stub():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "_klnmf_loop" in b]
    assert len(stub) == 1, "INTEGRATION.md: expected exactly one Option-B stub"
    code = stub[0].replace('ctypes.CDLL("libklnmf.so")', 'ctypes.CDLL(%r)' % build.LIB)
    ns = {}
    exec(compile(code, "INTEGRATION.md:option-B", "exec"), ns)
    return ns["_klnmf_loop"]


@pytest.mark.parametrize("kind", ["dense", "csr"])
@pytest.mark.parametrize("mode", [2, 3], ids=["fp64", "tf32r"])
def test_option_b_stub_runs_and_matches_the_oracle(within, kind, mode):
    loop = load_stub()
    rs = np.random.RandomState(3)
    n, f, k = 120, 90, 7
    X = rs.gamma(0.5, 1.0, size=(n, f))
    if kind == "csr":
        X[rs.random_sample((n, f)) < 0.8] = 0.0
        X = sp.csr_matrix(X)
    np.random.seed(9)
    H0 = O.init_dictionary(k, f)
    W_ref, H_ref, errs_ref, n_iter_ref = O.fit_transform(X.copy(), k=k, max_iter=10, tol=0, H0=H0)
    W, H, errs, n_iter = loop(X.copy(), H0, 10, 0.0, True, mode)
    assert n_iter == 10 and len(errs) == 10
    tol = 1e-12 if mode == 2 else (8e-4 if kind == "dense" else 2e-6)
    within("W", cases.rel_fro(W, W_ref), tol)
    within("H", cases.rel_fro(H, H_ref), tol)
    within("objective", float(np.max(np.abs(np.asarray(errs) - errs_ref) / np.abs(errs_ref))), tol)
