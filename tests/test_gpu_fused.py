"""-m gpu: the fused coefficient half-step (dense_fused.cu: S = W.H -> Q in the TMEM epilogue -> G += Q.H^T
without writing Q to HBM) against the float64 oracle's transform (nmf.py:275-291, 325-343) and against the
unfused three-kernel form of the same mode, on ragged shapes and both k paddings (64, 128)."""
import os

import numpy as np
import pytest

from multimodal_b200 import _native
from oracle import cases, klnmf_oracle as O

pytestmark = pytest.mark.gpu

SHAPES_256 = [
    (128, 32, 130),      # one row block, one step, k just above the single-CTA limit
    (700, 1000, 200),    # ragged rows / features
    (513, 333, 256),     # full k, features not a multiple of 64 (nor of the 32 per CTA)
    (4096, 2048, 256),   # many row blocks per cluster: barrier phases wrap
    (300, 4100, 160),
]

SHAPES = [
    (128, 32, 8),        # one row block, one step
    (700, 1000, 50),     # ragged rows / features, k padded to 64
    (513, 333, 100),     # k padded to 128, features not a multiple of 32
    (4096, 2048, 128),   # many row blocks per CTA: barrier phases wrap, TMEM accumulator is reused
    (300, 4100, 33),     # long feature sweep
]


def run_transform(X, H, iters, fused, ts=False):
    os.environ["KLNMF_FUSED"] = "1" if fused else "0"
    os.environ["KLNMF_FUSED_TS"] = "1" if ts else "0"
    try:
        n, f = X.shape
        with _native.Engine(n, f, H.shape[0], mode="tf32") as e:
            e.set_dense(X)
            e.set_dictionary(H)
            e.init_coefficients()
            c0 = e.counters()["launches"]
            errs, _ = e.run(iters, 0.0, False)
            launches = e.counters()["launches"] - c0
            return e.get_coefficients(), np.asarray(errs), launches
    finally:
        os.environ.pop("KLNMF_FUSED", None)
        os.environ.pop("KLNMF_FUSED_TS", None)


@pytest.mark.parametrize("ts", [False, True], ids=["w_in_smem", "w_in_tmem"])
@pytest.mark.parametrize("n,f,k", SHAPES)
def test_fused_transform_matches_oracle_and_unfused(n, f, k, ts):
    rs = np.random.RandomState(n + f + k)
    X = rs.random_sample((n, f))
    X[rs.random_sample((n, f)) < 0.2] = 0.0          # exact zeros: q = eps/(s+eps) there (nmf.py:336)
    np.random.seed(5)
    H = O.init_dictionary(k, f)
    iters = 6
    W_ref = np.asarray(X.dot(H.T))
    errs_ref = []
    for _ in range(iters):
        errs_ref.append(O.error(X, W_ref, H))
        W_ref, _ = O.update(X, W_ref, H, fit=False)
    Wf, ef, lf = run_transform(X, H, iters, True, ts)
    Wu, eu, lu = run_transform(X, H, iters, False)
    assert lf < lu, "the fused path must be the one that ran (one kernel per iteration)"
    assert np.isfinite(Wf).all()
    # stated tolerance of the single-pass TF32 mode (tests/test_gpu_parity.py): 3e-3 on W, 1e-2 on the objective
    assert cases.rel_fro(Wf, W_ref) < 3e-3, cases.rel_fro(Wf, W_ref)
    np.testing.assert_allclose(ef, errs_ref, rtol=1e-2)
    # fused and unfused run the same arithmetic up to the order of the FP32 accumulation
    assert cases.rel_fro(Wf, Wu) < 1e-3, cases.rel_fro(Wf, Wu)
    np.testing.assert_allclose(ef, eu, rtol=1e-4)


@pytest.mark.parametrize("n,f,k", SHAPES_256)
def test_fused_pair_transform_k256(n, f, k):
    # 128 < k <= 256: the cluster-of-two kernel (dense_fused256.cu), Q exchanged through distributed shared memory
    rs = np.random.RandomState(n + f + k)
    X = rs.random_sample((n, f))
    X[rs.random_sample((n, f)) < 0.2] = 0.0
    np.random.seed(5)
    H = O.init_dictionary(k, f)
    iters = 5
    W_ref = np.asarray(X.dot(H.T))
    errs_ref = []
    for _ in range(iters):
        errs_ref.append(O.error(X, W_ref, H))
        W_ref, _ = O.update(X, W_ref, H, fit=False)
    Wf, ef, lf = run_transform(X, H, iters, True)
    Wu, eu, lu = run_transform(X, H, iters, False)
    assert lf < lu, "the fused path must be the one that ran"
    assert np.isfinite(Wf).all()
    assert cases.rel_fro(Wf, W_ref) < 3e-3, cases.rel_fro(Wf, W_ref)
    np.testing.assert_allclose(ef, errs_ref, rtol=1e-2)
    assert cases.rel_fro(Wf, Wu) < 1e-3, cases.rel_fro(Wf, Wu)
    np.testing.assert_allclose(ef, eu, rtol=1e-4)


def test_fused_transform_through_the_estimator():
    # the public API takes the fused path for transform in tf32 mode (k <= 128)
    from multimodal_b200.lib.nmf import KLdivNMF
    rs = np.random.RandomState(0)
    X = rs.random_sample((257, 190))
    np.random.seed(1)
    H = O.init_dictionary(20, 190)
    est = KLdivNMF(n_components=20, max_iter=25, tol=0, mode="tf32")
    est.components_ = H
    W = est.transform(X)
    W_ref = np.asarray(X.dot(H.T))
    for _ in range(25):
        W_ref, _ = O.update(X, W_ref, H, fit=False)
    assert cases.rel_fro(W, W_ref) < 3e-3


def run_fit(X, H, iters, fused):
    os.environ["KLNMF_FUSED"] = "1" if fused else "0"
    try:
        n, f = X.shape
        with _native.Engine(n, f, H.shape[0], mode="tf32") as e:
            e.set_dense(X)
            e.set_dictionary(H)
            e.init_coefficients()
            c0 = e.counters()["launches"]
            errs, _ = e.run(iters, 0.0, True)
            launches = e.counters()["launches"] - c0
            return e.get_coefficients(), e.get_dictionary(), np.asarray(errs), launches
    finally:
        os.environ.pop("KLNMF_FUSED", None)


@pytest.mark.parametrize("n,f,k", [(700, 1000, 50), (513, 333, 100), (2000, 96, 8)])
def test_fused_fit_matches_oracle_and_unfused(n, f, k):
    # fit: the fused kernel also writes the ratio panel (TMA store) for the numerator N += W'^T.Q (nmf.py:349)
    rs = np.random.RandomState(n * 3 + f + k)
    X = rs.random_sample((n, f))
    X[rs.random_sample((n, f)) < 0.3] = 0.0
    np.random.seed(9)
    H0 = O.init_dictionary(k, f)
    W_ref, H_ref = np.asarray(X.dot(H0.T)), H0
    errs_ref = []
    for _ in range(10):
        errs_ref.append(O.error(X, W_ref, H_ref))
        W_ref, H_ref = O.update(X, W_ref, H_ref, fit=True)
    Wf, Hf, ef, lf = run_fit(X, H0, 10, True)
    Wu, Hu, eu, lu = run_fit(X, H0, 10, False)
    assert lf <= lu       # one fused kernel (+ the H^T refresh) replaces the ratio and coefficient contractions
    assert cases.rel_fro(Wf, W_ref) < 3e-3 and cases.rel_fro(Hf, H_ref) < 3e-3
    np.testing.assert_allclose(ef, errs_ref, rtol=1e-2)
    assert cases.rel_fro(Wf, Wu) < 1e-3 and cases.rel_fro(Hf, Hu) < 1e-3


def test_large_download_roundtrip():
    # coefficient downloads >= 32 MB take the pipelined path (pinned staging + host copy threads, api.cu)
    n, f, k = 70001, 40, 72
    rs = np.random.RandomState(2)
    W = rs.random_sample((n, k))
    for mode in ("tf32", "tf32x3", "fp64"):
        with _native.Engine(n, f, k, mode=mode) as e:
            e.set_dense(np.ones((n, f), dtype=np.float32))
            e.set_coefficients(W)
            back = e.get_coefficients()
        assert back.shape == W.shape
        if mode == "fp64":
            assert np.array_equal(back, W)
        else:
            np.testing.assert_allclose(back, W, rtol=1.2e-7)     # one float32 rounding (hi + lo keeps more in tf32x3)
