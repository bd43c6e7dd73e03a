"""Generate tests/golden/*.npz by running the REAL reference (test infrastructure).

Run in the build container only (the reference tree cannot travel to the GPU box):

    python oracle/make_golden.py            # needs /root/reference

The vectors pin `oracle/klnmf_oracle.py` (tests/test_oracle.py) and are what the
`-m gpu` parity tests compare the CUDA path with.  Inputs are regenerated from the
seeds recorded here by `oracle/cases.py`, so only outputs are stored.
"""
import os
import sys

import numpy as np

np.Inf = np.inf          # numpy>=2 shim: nmf.py:206 uses np.Inf
np.alltrue = np.all

REF = os.environ.get("KLNMF_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from multimodal.lib.nmf import KLdivNMF            # noqa: E402  (the reference)
from multimodal.learner import MultimodalLearner   # noqa: E402
from oracle import cases                            # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden")


def run_fit(X, k, iters, seed, tol=0, fit=True, H0=None):
    nmf = KLdivNMF(n_components=k, max_iter=iters, tol=tol)
    if H0 is not None:
        nmf.components_ = H0.copy()
        nmf._init_dictionary = nmf.components_
    np.random.seed(seed)
    W, errs = nmf.fit_transform(X.copy(), _fit=fit, return_errors=True)
    return np.asarray(W), np.asarray(nmf.components_), np.asarray(errs, dtype=np.float64)


def fit_case(name, X, k, seed, long_iters=200):
    """W, H after 10 iterations; objective trace over `long_iters`."""
    W10, H10, e10 = run_fit(X, k, 10, seed)
    _, _, elong = run_fit(X, k, long_iters, seed)
    nmf = KLdivNMF(n_components=k, max_iter=long_iters, tol=0)
    np.random.seed(seed)
    Wl = nmf.fit_transform(X.copy())
    final = nmf.error(X.copy(), Wl)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), W10=W10, H10=H10, errors10=e10,
                        errors_long=elong, final_error=np.float64(final),
                        W_long_rows=np.asarray(Wl)[:8], H_long=np.asarray(nmf.components_))
    print(name, "e0=%.6g e9=%.6g e_last=%.6g final=%.6g" % (e10[0], e10[-1], elong[-1], final))


def main():
    os.makedirs(OUT, exist_ok=True)

    # --- SURVEY 8c known answers (dense 6x2, CSR 5x4) -------------------------
    X = cases.kat_dense()
    W, H, e = run_fit(X, 2, 10, 0)
    nmf = KLdivNMF(n_components=2, max_iter=10, tol=0)
    nmf.components_ = H
    np.savez_compressed(os.path.join(OUT, "kat_dense.npz"), W=W, H=H, errors=e,
                        after=np.float64(nmf.error(X, W)))
    Xs = cases.kat_csr()
    W, H, e = run_fit(Xs, 2, 10, 0)
    Wd, Hd, ed = run_fit(Xs.toarray(), 2, 10, 0)
    np.savez_compressed(os.path.join(OUT, "kat_csr.npz"), W=W, H=H, errors=e,
                        W_densepath=Wd, H_densepath=Hd, errors_densepath=ed)

    # --- cfg1 (BASELINE.json configs[0]) --------------------------------------
    fit_case("cfg1_dense", cases.cfg1_X(), 10, 1, long_iters=200)

    # --- mid-size dense with ragged shapes (not multiples of any tile) --------
    fit_case("ragged_dense", cases.ragged_dense_X(), 13, 5, long_iters=200)

    # --- dense with exact zeros (dense path keeps eps/(WH+eps) there) ---------
    fit_case("zeros_dense", cases.zeros_dense_X(), 8, 7, long_iters=60)

    # --- sparse CSR ------------------------------------------------------------
    fit_case("sparse_mid", cases.sparse_mid_X(), 16, 11, long_iters=200)

    # --- transform with a fixed (sub-)dictionary -------------------------------
    X = cases.cfg1_X()[:64]
    H0 = cases.sub_dictionary(10, 200)
    W, H, e = run_fit(X, 10, 30, 0, fit=False, H0=H0)
    assert np.array_equal(H, H0)
    np.savez_compressed(os.path.join(OUT, "transform_dense.npz"), W=W, errors=e)
    Xs = cases.sparse_mid_X()[:50]
    H0 = cases.sub_dictionary(16, Xs.shape[1])
    W, H, e = run_fit(Xs, 16, 30, 0, fit=False, H0=H0)
    np.savez_compressed(os.path.join(OUT, "transform_sparse.npz"), W=W, errors=e)

    # --- early stop with tol > 0 ------------------------------------------------
    X = cases.cfg1_X()[:120, :60]
    W, H, e = run_fit(X, 6, 500, 2, tol=1e-5)
    np.savez_compressed(os.path.join(OUT, "early_stop.npz"), W=W, H=H, errors=e)
    print("early_stop n_errors", len(e))

    # --- tol = 0 and an objective that rises: the reference breaks at the second evaluation (nmf.py:215) ---
    X, H0 = cases.rise_case()
    W, H, e = run_fit(X, H0.shape[0], 50, 0, tol=0, H0=H0)
    assert len(e) == 1, len(e)
    np.savez_compressed(os.path.join(OUT, "rise_tol0.npz"), W=W, H=H, errors=e)
    print("rise_tol0 n_errors", len(e), "shape", X.shape, "k", H0.shape[0])

    # --- learner: two modalities (dense motion + CSR sound), cfg2 in miniature ---
    mot, snd, coefs = cases.learner_small()
    lr = MultimodalLearner(['motion', 'sound'], [mot.shape[1], snd.shape[1]], coefs, 8)
    np.random.seed(3)
    lr.train([mot, snd.copy()], 20)
    internal = lr.reconstruct_internal('sound', snd[:25].copy(), 15)
    internal_m = lr.reconstruct_internal('motion', mot[:25], 15)
    m2s = lr.modality_to_modality('motion', 'sound', mot[:25], 15)
    both = lr.reconstruct_internal_multi(['motion', 'sound'], [mot[:25], snd[:25].copy()], 15)
    np.savez_compressed(os.path.join(OUT, "learner_small.npz"), dico=lr.dico,
                        internal_sound=internal, internal_motion=internal_m,
                        motion_to_sound=m2s, internal_both=both)
    print("learner dico", lr.dico.shape, "m2s", m2s.shape)

    # --- reference unit-test known answers --------------------------------------
    from multimodal.lib.metrics import generalized_KL
    from multimodal.lib.array_utils import normalize_sum
    x = np.array([[1., 2.], [3., 4.]])
    np.savez_compressed(
        os.path.join(OUT, "primitives.npz"),
        kl_known=np.float64(generalized_KL(np.array([1., 2.]), np.array([2., 1.]))),
        norm_axis0=normalize_sum(x, axis=0), norm_axis1=normalize_sum(x, axis=1))


if __name__ == "__main__":
    main()
