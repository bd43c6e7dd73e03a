"""-m gpu: the fused coefficient half-step (dense_fused.cu, k <= 128; dense_fused256.cu on CTA pairs, k <= 256:
S = W.H -> Q in the TMEM epilogue -> G += Q.H^T without writing Q to HBM; fit: Q written once for the numerator)
against the float64 oracle (nmf.py:275-291, 325-351) and against the unfused three-kernel form of the same mode, in
both one-pass modes (tf32r, the default, and tf32), on ragged shapes and every k padding (64, 128, 256).
Tolerances: about 3 x the worst measured value (profiles/r2_parity_measured.json)."""
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


MODES = ["tf32r", "tf32"]
TOL_W = {"tf32r": 6e-4, "tf32": 1.3e-3}        # against the oracle; measured worst 2.1e-4 / 4.3e-4 (k = 8, f = 32)
TOL_KL = {"tf32r": 5e-5, "tf32": 6e-3}         # measured worst 1.7e-5 / 2.0e-3
TOL_SAME = {"tf32r": 1.7e-4, "tf32": 1.1e-4}   # fused against unfused (same arithmetic up to the FP32 summation order): 5.6e-5 / 3.6e-5


def maxrel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.abs(b)))


def run_transform(X, H, iters, fused, ts=True, mode="tf32"):
    os.environ["KLNMF_FUSED"] = "1" if fused else "0"
    try:
        n, f = X.shape
        with _native.Engine(n, f, H.shape[0], mode=mode) as e:
            e.set_dense(X)
            e.set_dictionary(H)
            e.init_coefficients()
            c0 = e.counters()["launches"]
            errs, _ = e.run(iters, 0.0, False)
            launches = e.counters()["launches"] - c0
            return e.get_coefficients(), np.asarray(errs), launches
    finally:
        os.environ.pop("KLNMF_FUSED", None)


def oracle_transform(X, H, iters):
    W_ref = np.asarray(X.dot(H.T))
    errs_ref = []
    for _ in range(iters):
        errs_ref.append(O.error(X, W_ref, H))
        W_ref, _ = O.update(X, W_ref, H, fit=False)
    return W_ref, errs_ref


def zero_heavy(n, f, k, frac=0.2):
    rs = np.random.RandomState(n + f + k)
    X = rs.random_sample((n, f))
    X[rs.random_sample((n, f)) < frac] = 0.0          # exact zeros: q = eps/(s+eps) there (nmf.py:336)
    return X


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,f,k", SHAPES)
def test_fused_transform_matches_oracle_and_unfused(within, n, f, k, mode, ts=True):
    X = zero_heavy(n, f, k)
    np.random.seed(5)
    H = O.init_dictionary(k, f)
    iters = 6
    W_ref, errs_ref = oracle_transform(X, H, iters)
    Wf, ef, lf = run_transform(X, H, iters, True, ts, mode)
    Wu, eu, lu = run_transform(X, H, iters, False, mode=mode)
    assert lf < lu, "the fused path must be the one that ran (one kernel per iteration)"
    assert np.isfinite(Wf).all()
    within("W", cases.rel_fro(Wf, W_ref), TOL_W[mode])
    within("objective", maxrel(ef, errs_ref), TOL_KL[mode])
    within("W_fused_vs_unfused", cases.rel_fro(Wf, Wu), TOL_SAME[mode])
    within("objective_fused_vs_unfused", maxrel(ef, eu), 5e-6)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,f,k", SHAPES_256)
def test_fused_pair_transform_k256(within, n, f, k, mode):
    # 128 < k <= 256: the cluster-of-two kernel (dense_fused256.cu), Q exchanged through distributed shared memory
    X = zero_heavy(n, f, k)
    np.random.seed(5)
    H = O.init_dictionary(k, f)
    iters = 5
    W_ref, errs_ref = oracle_transform(X, H, iters)
    Wf, ef, lf = run_transform(X, H, iters, True, mode=mode)
    Wu, eu, lu = run_transform(X, H, iters, False, mode=mode)
    assert lf < lu, "the fused path must be the one that ran"
    assert np.isfinite(Wf).all()
    within("W", cases.rel_fro(Wf, W_ref), TOL_W[mode])
    within("objective", maxrel(ef, errs_ref), TOL_KL[mode])
    within("W_fused_vs_unfused", cases.rel_fro(Wf, Wu), TOL_SAME[mode])
    within("objective_fused_vs_unfused", maxrel(ef, eu), 5e-6)


@pytest.mark.parametrize("mode", MODES)
def test_fused_transform_through_the_estimator(within, mode):
    # the public API takes the fused path for transform in the one-pass modes (k <= 256)
    from multimodal_b200.lib.nmf import KLdivNMF
    rs = np.random.RandomState(0)
    X = rs.random_sample((257, 190))
    np.random.seed(1)
    H = O.init_dictionary(20, 190)
    est = KLdivNMF(n_components=20, max_iter=25, tol=0, mode=mode)
    est.components_ = H
    W = est.transform(X)
    W_ref = np.asarray(X.dot(H.T))
    for _ in range(25):
        W_ref, _ = O.update(X, W_ref, H, fit=False)
    within("W", cases.rel_fro(W, W_ref), TOL_W[mode])


def run_fit(X, H, iters, fused, mode="tf32"):
    os.environ["KLNMF_FUSED"] = "1" if fused else "0"
    try:
        n, f = X.shape
        with _native.Engine(n, f, H.shape[0], mode=mode) as e:
            e.set_dense(X)
            e.set_dictionary(H)
            e.init_coefficients()
            c0 = e.counters()["launches"]
            errs, _ = e.run(iters, 0.0, True)
            launches = e.counters()["launches"] - c0
            return e.get_coefficients(), e.get_dictionary(), np.asarray(errs), launches
    finally:
        os.environ.pop("KLNMF_FUSED", None)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,f,k", [(700, 1000, 50), (513, 333, 100), (2000, 96, 8),
                                   (700, 1000, 200), (513, 333, 256), (1300, 2100, 130), (4096, 512, 192)])
def test_fused_fit_matches_oracle_and_unfused(within, n, f, k, mode):
    # fit: the fused kernel also writes the ratio panel (TMA store) for the numerator N += W'^T.Q (nmf.py:349);
    # k <= 128 on one CTA per row block, 128 < k <= 256 on CTA pairs (every warp stores its 32 x 32 box)
    X = zero_heavy(n * 3, f, k, 0.3)[:n]
    np.random.seed(9)
    H0 = O.init_dictionary(k, f)
    W_ref, H_ref = np.asarray(X.dot(H0.T)), H0
    errs_ref = []
    for _ in range(10):
        errs_ref.append(O.error(X, W_ref, H_ref))
        W_ref, H_ref = O.update(X, W_ref, H_ref, fit=True)
    Wf, Hf, ef, lf = run_fit(X, H0, 10, True, mode)
    Wu, Hu, eu, lu = run_fit(X, H0, 10, False, mode)
    assert lf <= lu      # one fused kernel (+ the H^T refresh) replaces the ratio and coefficient contractions
    within("W", cases.rel_fro(Wf, W_ref), TOL_W[mode])
    within("H", cases.rel_fro(Hf, H_ref), TOL_W[mode])
    within("objective", maxrel(ef, errs_ref), TOL_KL[mode])
    within("W_fused_vs_unfused", cases.rel_fro(Wf, Wu), TOL_SAME[mode])
    within("H_fused_vs_unfused", cases.rel_fro(Hf, Hu), TOL_SAME[mode])


def test_fused_fit_multi_panel_k256():
    """Several row panels of the ratio scratch with the cluster kernel's Q store: the panel boundaries must not show."""
    rs = np.random.RandomState(4)
    n, f, k = 1500, 700, 256
    X = rs.gamma(0.5, 1.0, size=(n, f))
    np.random.seed(3)
    H0 = O.init_dictionary(k, f)
    outs = []
    for limit in (None, 512 * 704 * 4):           # second run: 512-row panels
        with _native.Engine(n, f, k, mode="tf32r", scratch_limit=limit) as e:
            e.set_dense(X)
            e.set_dictionary(H0)
            e.init_coefficients()
            errs, _ = e.run(4, 0.0, True)
            outs.append((e.get_coefficients(), e.get_dictionary(), np.asarray(errs)))
    assert cases.rel_fro(outs[1][0], outs[0][0]) < 1e-5 and cases.rel_fro(outs[1][1], outs[0][1]) < 1e-5
    np.testing.assert_allclose(outs[1][2], outs[0][2], rtol=1e-6)


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
