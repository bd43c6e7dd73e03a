"""-m gpu: the CUDA path (through the drop-in Python API -> C ABI) against the golden vectors
produced by the real reference and against the float64 oracle on the same seeds.

Stated tolerances: about 3 x the worst value measured on the B200 (profiles/r2_parity_measured.json, written by the
`within` fixture of conftest.py); the table is TOL_* below, DESIGN.md section 2 explains each mode.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from multimodal_b200.lib.nmf import KLdivNMF
from multimodal_b200.learner import MultimodalLearner, fit_coefficients
from oracle import cases
from oracle import klnmf_oracle as O

pytestmark = pytest.mark.gpu

MODES = ["fp64", "tf32x3", "tf32r", "tf32"]
# stated tolerances = about 3 x the worst value measured on the B200 (profiles/r2_parity_measured.json, written by the
# `within` fixture of conftest.py); k, f >= 6 cases.  The 6 x 2, k = 2 known-answer case has its own row: nothing
# averages the operand rounding of the one-pass modes there.
TOL_WH = {"fp64": 1e-12, "tf32x3": 5e-6, "tf32r": 8e-4, "tf32": 2.5e-3}      # measured: 1e-15, 1.6e-6, 2.7e-4, 8.4e-4
TOL_KL = {"fp64": 1e-12, "tf32x3": 3e-6, "tf32r": 2e-5, "tf32": 2e-3}        # measured: 4e-16, 9.1e-7, 5.1e-6, 7.3e-4
TOL_KL_LONG = {"fp64": 1e-12, "tf32x3": 4e-6, "tf32r": 1e-3, "tf32": 1e-2}   # converged fit (objective small against sum(X)): 2e-15, 1.1e-6, 3.6e-4, 3.2e-3
TOL_KAT_WH = {"fp64": 1e-12, "tf32x3": 5e-7, "tf32r": 8e-4, "tf32": 1.2e-3}  # measured: 2.7e-16, 1.4e-7, 2.6e-4, 3.7e-4
TOL_KAT_KL = {"fp64": 1e-12, "tf32x3": 6e-6, "tf32r": 6e-3, "tf32": 5e-3}    # measured: 5.4e-15, 2.1e-6, 2.2e-3, 1.7e-3
# shapes with k >= 256 (test_mid_size_dense_vs_oracle, test_k512_ragged_dense_vs_oracle): the FP32 accumulation of the
# objective over k terms shows in tf32x3 (4.2e-5 at k = 512)
TOL_KL_WIDE = {"fp64": 1e-12, "tf32x3": 2.5e-5, "tf32r": 2.5e-5, "tf32": 5e-3}    # measured: 6e-16, 7.0e-6, 6.8e-6, 1.6e-3
TOL_WH_WIDE = {"fp64": 1e-12, "tf32x3": 1.6e-5, "tf32r": 4.5e-4, "tf32": 5.5e-4}  # measured: 3e-15, 5.3e-6, 1.4e-4, 1.8e-4
# the CSR path: FP32 FMA in every TF32 mode, FP64 accumulation of the objective across rows
TOL_CSR_WH = 1.5e-6        # measured 4.4e-7 (300 x 400, k = 16), 6e-8 (5 x 4, k = 2); 1.0e-5 at the cfg4 shape (own test)
TOL_CSR_KL = 1.2e-6        # measured 3.9e-7 (k = 2), 7e-9 (k = 16)


def maxrel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.abs(b))) if a.size else 0.0


def fit(X, k, iters, seed, mode, tol=0, **kw):
    est = KLdivNMF(n_components=k, max_iter=iters, tol=tol, mode=mode)
    np.random.seed(seed)
    W, errs = est.fit_transform(X, return_errors=True, **kw)
    return est, W, np.asarray(errs)


@pytest.mark.parametrize("mode", MODES)
def test_kat_dense(golden, within, mode):
    g = golden("kat_dense")
    est, W, errs = fit(cases.kat_dense(), 2, 10, 0, mode)
    within("W", cases.rel_fro(W, g["W"]), TOL_KAT_WH[mode])
    within("H", cases.rel_fro(est.components_, g["H"]), TOL_KAT_WH[mode])
    # a 6 x 2 matrix with k = 2: nothing averages the 2^-12 operand rounding of the one-pass modes, and the objective
    # after 10 iterations (0.185, from 5.75) magnifies the resulting 2e-4 on W
    within("objective", maxrel(errs, g["errors"]), TOL_KAT_KL[mode])
    within("objective_after", maxrel(est.error(cases.kat_dense(), W), g["after"]), TOL_KAT_KL[mode])


@pytest.mark.parametrize("mode", MODES)
def test_kat_csr(golden, within, mode):
    g = golden("kat_csr")
    est, W, errs = fit(cases.kat_csr(), 2, 10, 0, mode)
    tol = 1e-12 if mode == "fp64" else TOL_CSR_WH        # the sparse path is FP32 FMA in every TF32 mode
    within("W", cases.rel_fro(W, g["W"]), tol)
    within("H", cases.rel_fro(est.components_, g["H"]), tol)
    within("objective", maxrel(errs, g["errors"]), 1e-12 if mode == "fp64" else TOL_CSR_KL)
    # and the dense path on the same matrix is the reference's OTHER algorithm (eps at X == 0)
    est, W, errs = fit(cases.kat_csr().toarray(), 2, 10, 0, mode)
    within("objective_densepath", maxrel(errs, g["errors_densepath"]), TOL_KAT_KL[mode])   # k = 2


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,maker,k,seed,long_iters", [
    ("cfg1_dense", cases.cfg1_X, 10, 1, 200),
    ("ragged_dense", cases.ragged_dense_X, 13, 5, 200),
    ("zeros_dense", cases.zeros_dense_X, 8, 7, 60),
])
def test_dense_fit_golden(golden, within, mode, name, maker, k, seed, long_iters):
    g = golden(name)
    est, W, errs = fit(maker(), k, 10, seed, mode)
    assert W.dtype == np.float64 and W.flags.c_contiguous and W.shape == g["W10"].shape
    assert len(errs) == 10
    within("W10", cases.rel_fro(W, g["W10"]), TOL_WH[mode])
    within("H10", cases.rel_fro(est.components_, g["H10"]), TOL_WH[mode])
    within("objective10", maxrel(errs, g["errors10"]), TOL_KL[mode])
    est, W, errs = fit(maker(), k, long_iters, seed, mode)
    assert len(errs) == long_iters
    # on a converged fit the objective is small against sum(X) and follows the 1e-5..1e-4 of W and H
    within("objective_long", maxrel(errs[-1], g["errors_long"][-1]), TOL_KL_LONG[mode])
    within("final_error", maxrel(est.error(maker(), W), g["final_error"]), TOL_KL_LONG[mode])


@pytest.mark.parametrize("mode", MODES)
def test_sparse_fit_golden(golden, within, mode):
    g = golden("sparse_mid")
    tol = 1e-12 if mode == "fp64" else TOL_CSR_WH
    tol_kl = 1e-12 if mode == "fp64" else TOL_CSR_KL
    est, W, errs = fit(cases.sparse_mid_X(), 16, 10, 11, mode)
    within("W10", cases.rel_fro(W, g["W10"]), tol)
    within("H10", cases.rel_fro(est.components_, g["H10"]), tol)
    within("objective10", maxrel(errs, g["errors10"]), tol_kl)
    est, W, errs = fit(cases.sparse_mid_X(), 16, 200, 11, mode)
    within("objective_long", maxrel(errs[-1], g["errors_long"][-1]), tol_kl)
    within("final_error", maxrel(est.error(cases.sparse_mid_X(), W), g["final_error"]), tol_kl)


@pytest.mark.parametrize("mode", MODES)
def test_transform_golden(golden, within, mode):
    H0 = cases.sub_dictionary(10, 200)
    est = KLdivNMF(n_components=10, max_iter=30, tol=0, mode=mode)
    est.components_ = H0
    np.random.seed(0)
    W = est.transform(cases.cfg1_X()[:64])
    assert est.components_ is H0 and est._init_dictionary is H0      # untouched + sticky (nmf.py:289)
    within("W_dense", cases.rel_fro(W, golden("transform_dense")["W"]), TOL_WH[mode])
    Xs = cases.sparse_mid_X()[:50]
    W = fit_coefficients(Xs, cases.sub_dictionary(16, Xs.shape[1]), iter_nmf=30, mode=mode)
    within("W_csr", cases.rel_fro(W, golden("transform_sparse")["W"]), 1e-12 if mode == "fp64" else TOL_CSR_WH)


@pytest.mark.parametrize("mode", MODES)
def test_early_stop_golden(golden, within, mode, capsys):
    g = golden("early_stop")
    est, W, errs = fit(cases.cfg1_X()[:120, :60], 6, 500, 2, mode, tol=1e-5)
    if mode == "fp64":
        assert len(errs) == len(g["errors"])          # stops at the same iteration as the reference
        within("W", cases.rel_fro(W, g["W"]), 1e-8)
        within("H", cases.rel_fro(est.components_, g["H"]), 1e-8)
    else:
        # the stop test compares an improvement of ~tol_abs with the mode's objective noise: the iteration differs
        # (measured: the same iteration in every mode; W 4.6e-7 / 1.7e-3 / 4.1e-4)
        within("iterations_off", abs(len(errs) - len(g["errors"])) + 0.5, {"tf32x3": 3, "tf32r": 8, "tf32": 12}[mode])
        within("W", cases.rel_fro(W, g["W"]), {"tf32x3": 2e-6, "tf32r": 5e-3, "tf32": 1.5e-3}[mode])
    assert "Iteration limit" not in capsys.readouterr().err
    # and the warning text when the limit IS reached with tol > 0 (nmf.py:224-225)
    fit(cases.cfg1_X()[:120, :60], 6, 5, 2, mode, tol=1e-5)
    assert "Warning: Iteration limit reached during fit\n" in capsys.readouterr().err


@pytest.mark.parametrize("mode", MODES)
def test_tol0_breaks_on_a_rise_like_the_reference(golden, within, mode):
    """tol = 0 (the learner's idiom, learner.py:12,39): the reference breaks iff the objective RISES (nmf.py:215).
    With this adversarial initial dictionary it does, at the second objective evaluation, by 73 % -- far above any
    mode's noise floor, so every mode has to stop exactly there and hand back the state after ONE update."""
    g = golden("rise_tol0")
    X, H0 = cases.rise_case()
    est = KLdivNMF(n_components=H0.shape[0], max_iter=50, tol=0, mode=mode)
    est._init_dictionary = H0
    W, errs = est.fit_transform(X, return_errors=True)
    assert len(g["errors"]) == 1 and len(errs) == 1
    # the ratios of this case reach 1e3 and the dictionary spans 8 decades: W.H itself is what limits the one-pass modes
    tol = {"fp64": 1e-12, "tf32x3": 3e-7, "tf32r": 7e-4, "tf32": 1.1e-3}[mode]     # measured: 2e-16, 8.3e-8, 2.2e-4, 3.6e-4
    within("objective", maxrel(errs, g["errors"]), tol)
    within("W", cases.rel_fro(W, g["W"]), tol)
    within("H", cases.rel_fro(est.components_, g["H"]), tol)
    # ... and the objective of the returned state is indeed the larger one the reference saw and refused
    assert est.error(X, W) > errs[0] * 1.5


@pytest.mark.parametrize("mode", MODES)
def test_learner_golden(golden, within, mode):
    g = golden("learner_small")
    tol = 1e-12 if mode == "fp64" else 1e-5        # the CSR path is FP32 FMA in every TF32 mode: measured 3.4e-6
    mot, snd, coefs = cases.learner_small()
    lr = MultimodalLearner(['motion', 'sound'], [mot.shape[1], snd.shape[1]], coefs, 8, mode=mode)
    np.random.seed(3)
    lr.train([mot, snd.copy()], 20)
    assert lr.dico.shape == g["dico"].shape
    within("dico", cases.rel_fro(lr.dico, g["dico"]), tol)
    within("internal_sound", cases.rel_fro(lr.reconstruct_internal('sound', snd[:25].copy(), 15), g["internal_sound"]), tol)
    # a dense-only sub-problem goes through the dense engine of the mode
    within("internal_motion", cases.rel_fro(lr.reconstruct_internal('motion', mot[:25], 15), g["internal_motion"]),
           max(tol, TOL_WH[mode]))
    m2s = lr.modality_to_modality('motion', 'sound', mot[:25], 15)
    assert m2s.shape == g["motion_to_sound"].shape
    within("motion_to_sound", cases.rel_fro(m2s, g["motion_to_sound"]), max(tol, TOL_WH[mode]))
    both = lr.reconstruct_internal_multi(['motion', 'sound'], [mot[:25], snd[:25].copy()], 15)
    within("internal_both", cases.rel_fro(both, g["internal_both"]), tol)
    with pytest.raises(AssertionError):
        lr.reconstruct_internal('sound', mot[:25], 5)          # wrong width (learner.py:73)


@pytest.mark.parametrize("mode", MODES)
def test_mid_size_dense_vs_oracle(within, mode):
    """cfg3/cfg5-like tile shapes (f, k multiples of the MMA tile) on a row subsample."""
    rs = np.random.RandomState(5)
    n, f, k = 1000, 1024, 256
    X = rs.gamma(0.5, 1.0, size=(n, f))
    np.random.seed(9)
    Wr, Hr, er, _ = O.fit_transform(X, k=k, max_iter=10, tol=0)
    est, W, errs = fit(X, k, 10, 9, mode)
    within("W", cases.rel_fro(W, Wr), TOL_WH_WIDE[mode])
    within("H", cases.rel_fro(est.components_, Hr), TOL_WH_WIDE[mode])
    within("objective", maxrel(errs, er), TOL_KL_WIDE[mode])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,f", [(391, 1000), (1030, 333)])
def test_k512_ragged_dense_vs_oracle(within, mode, n, f):
    """k = 512 (cfg5's component count) on ragged rows / features: the ratio contraction takes its in-place form
    (the ratio overwrites the X chunk in shared memory, dense_tc.cu) for k >= 512, the coefficient and numerator
    contractions their wide tiles (256 x 512 / 512 x 256), and tf32x3 / tf32r the centered ratio."""
    rs = np.random.RandomState(n + f)
    X = rs.gamma(0.5, 1.0, size=(n, f))
    X[rs.random_sample((n, f)) < 0.25] = 0.0
    np.random.seed(21)
    Wr, Hr, er, _ = O.fit_transform(X, k=512, max_iter=10, tol=0)
    est, W, errs = fit(X, 512, 10, 21, mode)
    within("W", cases.rel_fro(W, Wr), TOL_WH_WIDE[mode])
    within("H", cases.rel_fro(est.components_, Hr), TOL_WH_WIDE[mode])
    within("objective", maxrel(errs, er), TOL_KL_WIDE[mode])


@pytest.mark.parametrize("n,f,k", [(600, 700, 300), (520, 300, 700)])
def test_ragged_component_counts_beyond_one_tile(within, n, f, k):
    """256 < k not a multiple of anything: the second 256-column MMA of the wide coefficient tile and the second
    256-row block of the wide numerator tile are partly (k = 300) or more than once (k = 700: two tiles) used."""
    rs = np.random.RandomState(n + k)
    X = rs.gamma(0.5, 1.0, size=(n, f))
    np.random.seed(22)
    Wr, Hr, er, _ = O.fit_transform(X, k=k, max_iter=6, tol=0)
    for mode in ("tf32r", "tf32"):
        est, W, errs = fit(X, k, 6, 22, mode)
        within(mode + "_W", cases.rel_fro(W, Wr), TOL_WH_WIDE[mode])
        within(mode + "_H", cases.rel_fro(est.components_, Hr), TOL_WH_WIDE[mode])
        within(mode + "_objective", maxrel(errs, er), TOL_KL_WIDE[mode])


def test_multi_panel_equals_single_panel():
    """Row panels (the ratio scratch) must not change the result."""
    from multimodal_b200 import _native
    X = cases.ragged_dense_X()
    np.random.seed(5)
    H0 = O.init_dictionary(13, X.shape[1])
    outs = []
    for limit in (None, 128 * 160 * 8):          # second run: 128-row panels
        with _native.Engine(X.shape[0], X.shape[1], 13, mode="fp64", scratch_limit=limit) as e:
            e.set_dense(X)
            e.set_dictionary(H0)
            e.init_coefficients()
            errs, _ = e.run(5, 0.0, True)
            outs.append((e.get_coefficients(), e.get_dictionary(), errs))
    assert cases.rel_fro(outs[1][0], outs[0][0]) < 1e-12 and cases.rel_fro(outs[1][1], outs[0][1]) < 1e-12
    np.testing.assert_allclose(outs[1][2], outs[0][2], rtol=1e-12)


def test_cfg2_full_shapes_learner_vs_oracle(within):
    """BASELINE.json configs[1] at its real shapes: 1000 samples, motion 450 dense + sound 110 000
    CSR at 1 % -> a 1000 x 110 450 CSR stack (learner.py:53-56), k = 50; a few training iterations and a
    reconstruction of the missing modality against the float64 oracle (3 iterations keep the oracle's
    nnz x k temporaries, 3 x 0.6 GB, and its runtime in seconds)."""
    motion, sound, coefs = cases.cfg2_inputs()
    mods, dims = ['motion', 'sound'], [motion.shape[1], sound.shape[1]]
    lr = MultimodalLearner(mods, dims, coefs, 50)
    np.random.seed(3)
    lr.train([motion, sound.copy()], 3)
    ref = O.Learner(mods, dims, coefs, 50)
    np.random.seed(3)
    ref.train([motion, sound.copy()], 3)
    assert lr.dico.shape == (50, 110450)
    within("dico", cases.rel_fro(lr.dico, ref.dico), 5e-6)                     # measured 1.0e-6
    internal = lr.reconstruct_internal('motion', motion[:40], 5)
    internal_ref = ref.reconstruct_internal_multi(['motion'], [motion[:40]], 5)
    within("internal", cases.rel_fro(internal, internal_ref), 7e-4)            # dense 40 x 450 transform in tf32r: 2.3e-4
    snd = lr.reconstruct_modality('sound', internal)
    assert snd.shape == (40, 110000)
    within("reconstruction", cases.rel_fro(snd, internal_ref.dot(ref.get_dico('sound'))), 1.5e-4)   # measured 4.7e-5


@pytest.mark.parametrize("mode", MODES)
def test_learner_dense_stack_formed_on_the_device(within, mode):
    """Three dense modalities (float64, float32, integer counts): learner.stack_data hands the estimator the blocks
    and klnmf_set_dense_blocks_host scales and concatenates them on the device; same dictionary as the reference
    formula safe_hstack([c * m]) (learner.py:53-56) fed to the estimator, and as the float64 oracle."""
    rs = np.random.RandomState(17)
    n = 300
    mats = [rs.gamma(0.5, 1.0, size=(n, 70)), rs.random_sample((n, 45)).astype(np.float32),
            rs.poisson(1.5, size=(n, 33))]
    mods, dims = ['sound', 'image', 'motion'], [70, 45, 33]
    coefs = [1. / np.mean(np.sum(m, axis=1)) for m in mats]          # experiment.py:70-72
    lr = MultimodalLearner(mods, dims, coefs, 9, mode=mode)
    np.random.seed(4)
    lr.train(mats, 10)
    V = np.hstack([c * m for m, c in zip(mats, coefs)])
    est = KLdivNMF(n_components=9, max_iter=10, tol=0, mode=mode)
    np.random.seed(4)
    est.fit(V)
    within("blocks_vs_hstack", cases.rel_fro(lr.dico, est.components_) + 1e-300,
           1e-12 if mode == "fp64" else (1e-5 if mode == "tf32r" else 1e-6))
    ref = O.Learner(mods, dims, coefs, 9)
    np.random.seed(4)
    ref.train(mats, 10)
    within("dico", cases.rel_fro(lr.dico, ref.dico), TOL_WH[mode])
    internal = lr.reconstruct_internal_multi(['sound', 'motion'], [mats[0][:50], mats[2][:50]], 10)
    internal_ref = ref.reconstruct_internal_multi(['sound', 'motion'], [mats[0][:50], mats[2][:50]], 10)
    within("internal", cases.rel_fro(internal, internal_ref), TOL_WH[mode])


@pytest.mark.parametrize("mode", ["fp64", "tf32r"])
def test_batched_multi_subset_transform_vs_oracle(within, mode):
    """the coefficients of the same test samples through every subset of three dense modalities
    (experiment.py:350-369, `_get_all_internals`: six subsets for the twelve tested combinations) -- one upload of the
    scaled stack, one device-side column view per subset (learner.reconstruct_internal_batch) -- against six separate
    oracle transforms (learner.py:71-78 -> fit_coefficients, learner.py:11-15)."""
    rs = np.random.RandomState(23)
    n, mods, dims, k = 150, ['sound', 'image', 'motion'], [70, 45, 33], 9
    mats = [rs.gamma(0.5, 1.0, size=(n, 70)), rs.random_sample((n, 45)).astype(np.float32), rs.poisson(1.5, size=(n, 33))]
    coefs = [1. / np.mean(np.sum(m, axis=1)) for m in mats]
    ref = O.Learner(mods, dims, coefs, k)
    np.random.seed(4)
    ref.train(mats, 10)
    lr = MultimodalLearner(mods, dims, coefs, k, mode=mode)
    lr.dico = ref.dico                                             # the same dictionary on both sides
    subsets = [['sound'], ['image'], ['motion'], ['image', 'motion'], ['sound', 'motion'], ['sound', 'image']]
    test = [m[:60] for m in mats]
    got = lr.reconstruct_internal_batch(test, subsets, 12)
    assert sorted(got) == sorted(tuple(s) for s in subsets)
    for s in subsets:
        want = ref.reconstruct_internal_multi(s, [test[mods.index(name)] for name in s], 12)
        assert got[tuple(s)].shape == want.shape == (60, k)
        # a float32-only stack stays float32 in the reference, whose (X + eps) is then a FLOAT32 sum (numpy keeps the
        # array's type against a Python float, nmf.py:336) -- eps is rounded away for x > 0.3; the engine adds eps in the
        # arithmetic of its mode: 2e-8 apart in fp64
        tol = 1e-7 if (mode == "fp64" and s == ['image']) else TOL_WH[mode]
        within("+".join(s), cases.rel_fro(got[tuple(s)], want), tol)
        # and the batched path is the one-by-one path of the same mode, up to nothing but the upload route
        one = lr.reconstruct_internal_multi(s, [test[mods.index(name)] for name in s], 12)
        within("+".join(s) + "_vs_one_by_one", cases.rel_fro(got[tuple(s)], one) + 1e-300, 1e-12 if mode == "fp64" else 1e-6)
    with pytest.raises(ValueError, match="Negative values"):
        bad = [m.copy() for m in test]
        bad[2] = bad[2].astype(np.float64)
        bad[2][3, 4] = -1.0
        lr.reconstruct_internal_batch(bad, subsets, 2)


@pytest.mark.parametrize("mode", ["fp64", "tf32r"])
def test_mixed_dense_and_csr_stack_built_on_the_device(within, mode):
    """a dense modality next to a sparse one.  The reference sparsifies, scales and stacks on the host
    (learner.py:53-56 -> scipy.sparse.hstack, array_utils.py:5-9); here the blocks are uploaded as they are and the
    scaled CSR stack is built on the device (klnmf_set_stacked_blocks_host).  Same dictionary and coefficients as the
    reference formula fed to the estimator, and as the float64 oracle; zeros of the dense block and explicit zeros of
    the CSR block are dropped like eliminate_zeros() drops them (nmf.py:66)."""
    rs = np.random.RandomState(31)
    n = 260
    motion = rs.dirichlet(0.1 * np.ones(40), n)
    motion[motion < 1e-3] = 0.0                                   # real zeros in the dense modality
    motion[7, :] = 0.0                                            # and an all-zero row of it
    sound = sp.random(n, 300, density=0.05, random_state=rs, format='csr')
    sound.data = np.ceil(5 * sound.data)
    sound.data[::17] = 0.0                                        # explicit zeros in the CSR modality
    image = rs.random_sample((n, 33)).astype(np.float32)
    mods, dims = ['motion', 'sound', 'image'], [40, 300, 33]
    coefs = [1. / np.mean(np.sum(motion, axis=1)), 1. / np.mean(np.asarray(sound.sum(axis=1))), np.float32(0.5)]
    lr = MultimodalLearner(mods, dims, coefs, 7, mode=mode)
    stack = lr.stack_data(mods, [motion, sound, image])
    from multimodal_b200.lib.array_utils import MixedBlocks
    assert isinstance(stack, MixedBlocks)
    n_explicit = sound.nnz
    np.random.seed(5)
    lr.train([motion, sound, image], 10)
    assert sound.nnz == n_explicit                               # the caller's modality matrix is left alone
    V = sp.hstack([c * m for m, c in zip([motion, sound, image], coefs)]).tocsr()     # the reference's stack
    est = KLdivNMF(n_components=7, max_iter=10, tol=0, mode=mode)
    np.random.seed(5)
    est.fit(V)
    within("device_stack_vs_host_stack", cases.rel_fro(lr.dico, est.components_) + 1e-300, 1e-12 if mode == "fp64" else 2e-6)
    ref = O.Learner(mods, dims, coefs, 7)
    np.random.seed(5)
    ref.train([motion, sound.copy(), image], 10)
    within("dico", cases.rel_fro(lr.dico, ref.dico), 1e-12 if mode == "fp64" else 2e-6)            # measured 5.9e-7
    internal = lr.reconstruct_internal_multi(['motion', 'sound'], [motion[:50], sound[:50]], 10)
    internal_ref = ref.reconstruct_internal_multi(['motion', 'sound'], [motion[:50], sound[:50]], 10)
    within("internal", cases.rel_fro(internal, internal_ref), 1e-12 if mode == "fp64" else 2e-6)   # measured 6.0e-7
    with pytest.raises(ValueError, match="Negative values"):
        bad = motion.copy()
        bad[3, 3] = -0.5
        lr.train([bad, sound, image], 2)
