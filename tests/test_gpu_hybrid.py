"""-m gpu: hybrid stacks -- the dense modalities of a mixed stack kept dense and run through the contraction
engine (tcgen05 / DMMA), the CSR modality through the sparse passes, one coefficient matrix and one dictionary
normaliser over both (api.cu: HybridSide).  The answer is the reference's for its all-sparse stack (learner.py:53-56,
array_utils.py:5-9, nmf.py:52-70): zeros of a dense block carry no ratio.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from multimodal_b200 import _native
from multimodal_b200.learner import MultimodalLearner
from multimodal_b200.lib.array_utils import MixedBlocks
from multimodal_b200.lib.nmf import KLdivNMF
from oracle import cases
from oracle import klnmf_oracle as O

pytestmark = pytest.mark.gpu

# The dense block multiplies in the mode's one-pass form: tcgen05 kind::tf32 on round-to-nearest TF32 copies of its
# operands (api.cu: HybridSide::Hr, Wr).  Stated = about 3 x the worst value measured (profiles/r2_parity_measured.json):
# W0 3.9e-5, W 1.2e-4, H 7.2e-5, coefficients of test samples 1.6e-4 (2.0e-4 motion -> sound), objective 6.2e-7.
TOL = {"fp64": 1e-12, "tf32r": 6e-4, "tf32": 6e-4}
TOL_KL = {"fp64": 1e-12, "tf32r": 2e-5, "tf32": 2e-5}
# the golden learner: 20 fit iterations on a 60-column dense block with k = 8 -- nothing averages the 2^-12 operand
# rounding over a contraction this short, and the fit amplifies it (measured 2.1e-3 on the dictionary; the default
# threshold keeps such narrow modalities in the CSR stack, whose FP32 FMA gives 1e-6)
TOL_GOLDEN = {"fp64": 1e-12, "tf32r": 6e-3}


@pytest.fixture
def hybrid_on():
    old = os.environ.get("KLNMF_HYBRID")
    os.environ["KLNMF_HYBRID"] = "1"
    yield
    if old is None:
        del os.environ["KLNMF_HYBRID"]
    else:
        os.environ["KLNMF_HYBRID"] = old


def modalities(n, seed, f_motion=96, f_sound=400, f_image=70):
    rs = np.random.RandomState(seed)
    motion = rs.dirichlet(0.1 * np.ones(f_motion), n)
    motion[motion < 1e-3] = 0.0                                   # real zeros in the dense modality
    motion[7 % n, :] = 0.0                                        # and an all-zero row of it
    sound = sp.random(n, f_sound, density=0.05, random_state=rs, format='csr')
    sound.data = np.ceil(5 * sound.data)
    sound.data[::17] = 0.0                                        # explicit zeros in the CSR modality
    image = rs.random_sample((n, f_image)).astype(np.float32)
    coefs = [1. / np.mean(np.sum(motion, axis=1)), 1. / np.mean(np.asarray(sound.sum(axis=1))), np.float32(0.5)]
    return [motion, sound, image], coefs


@pytest.mark.parametrize("mode", ["fp64", "tf32r", "tf32"])
def test_hybrid_engine_against_the_oracle(hybrid_on, within, mode):
    """The engine level: W0, ten fit iterations, the objective history and the objective after, against the oracle on
    the reference's all-sparse stack; dense - CSR - dense block order, so the dictionary is scattered over both parts."""
    n, k = 300, 24
    mats, coefs = modalities(n, 31)
    V = sp.hstack([c * m for m, c in zip(mats, coefs)]).tocsr()
    V.eliminate_zeros()
    f = V.shape[1]
    np.random.seed(4)
    H0 = O.init_dictionary(k, f)
    W_ref, H_ref, errs_ref, _ = O.fit_transform(V.copy(), k=k, max_iter=10, tol=0, H0=H0)
    with _native.Engine(n, f, k, mode=mode) as e:
        e.set_stacked_blocks(MixedBlocks(mats, coefs).canonical().blocks, coefs)
        assert e.is_hybrid()
        assert e.check_input() == (0, 0)
        e.set_dictionary(H0)
        within("H0_roundtrip", cases.rel_fro(e.get_dictionary(), H0) + 1e-300, 1e-12 if mode == "fp64" else 1e-7)
        e.init_coefficients()
        within("W0", cases.rel_fro(e.get_coefficients(), V.dot(H0.T)), TOL[mode])
        errs, n_iter = e.run(10, 0.0, True)
        W, H = e.get_coefficients(), e.get_dictionary()
        after = e.error()
    assert n_iter == 10 and len(errs) == 10
    np.testing.assert_allclose(H.sum(axis=1), 1.0, rtol=1e-5)
    within("W", cases.rel_fro(W, W_ref), TOL[mode])
    within("H", cases.rel_fro(H, H_ref), TOL[mode])
    within("objective", float(np.max(np.abs(np.asarray(errs) - errs_ref) / np.abs(errs_ref))), TOL_KL[mode])
    within("objective_after", abs(after - O.error(V, W_ref, H_ref)) / O.error(V, W_ref, H_ref), TOL_KL[mode])


@pytest.mark.parametrize("mode", ["fp64", "tf32r"])
def test_hybrid_learner_train_and_reconstruct(hybrid_on, within, mode):
    """Through the drop-in API: MultimodalLearner.train on three modalities, coefficients of test samples from two of
    them (a dense + CSR sub-stack: hybrid again) and from the CSR one alone, modality to modality."""
    mats, coefs = modalities(260, 32)
    mods, dims = ['motion', 'sound', 'image'], [m.shape[1] for m in mats]
    lr = MultimodalLearner(mods, dims, coefs, 9, mode=mode)
    assert isinstance(lr.stack_data(mods, mats), MixedBlocks)
    np.random.seed(5)
    lr.train(mats, 10)
    ref = O.Learner(mods, dims, coefs, 9)
    np.random.seed(5)
    ref.train([mats[0], mats[1].copy(), mats[2]], 10)
    within("dico", cases.rel_fro(lr.dico, ref.dico), TOL[mode])
    test = [mats[0][:50], mats[1][:50]]
    # the same dictionary on both sides: what is compared is the transform
    ref.dico = np.array(lr.dico)
    internal = lr.reconstruct_internal_multi(['motion', 'sound'], test, 10)
    within("internal", cases.rel_fro(internal, ref.reconstruct_internal_multi(['motion', 'sound'], test, 10)), TOL[mode])
    out = lr.modalities_to_modalities(['motion', 'sound'], ['image'], test, 10)
    out_ref = ref.modalities_to_modalities(['motion', 'sound'], ['image'], test, 10)
    within("to_image", cases.rel_fro(out[0], out_ref[0]), TOL[mode])
    with pytest.raises(ValueError, match="Negative values"):
        bad = mats[0].copy()
        bad[3, 3] = -0.5
        lr.train([bad, mats[1], mats[2]], 2)


def test_hybrid_is_off_below_the_width_threshold_and_in_tf32x3(within):
    """Default threshold: 1024 dense columns; tf32x3 keeps the all-CSR stack (FP32 FMA: its FP32-grade promise)."""
    mats, coefs = modalities(64, 33)
    blocks = MixedBlocks(mats, coefs).canonical().blocks
    f = sum(m.shape[1] for m in mats)
    with _native.Engine(64, f, 5, mode="tf32r") as e:
        e.set_stacked_blocks(blocks, coefs)
        assert not e.is_hybrid()
    with _native.Engine(64, f, 5, mode="tf32r") as e:
        e.set_hybrid_min_cols(100)
        e.set_stacked_blocks(blocks, coefs)
        assert e.is_hybrid()
    with _native.Engine(64, f, 5, mode="tf32x3") as e:
        e.set_hybrid_min_cols(100)
        e.set_stacked_blocks(blocks, coefs)
        assert not e.is_hybrid()


@pytest.mark.parametrize("mode", ["fp64", "tf32r"])
def test_hybrid_walks_the_samples_in_panels(hybrid_on, within, mode):
    """The ratio panel of the dense block is bounded by the scratch limit: with 128-row panels a 300-sample fit takes
    three (the last one ragged); same results as with one panel and as the oracle."""
    n, k = 300, 16
    mats, coefs = modalities(n, 35)
    blocks = MixedBlocks(mats, coefs).canonical().blocks
    V = sp.hstack([c * m for m, c in zip(mats, coefs)]).tocsr()
    V.eliminate_zeros()
    f = V.shape[1]
    np.random.seed(8)
    H0 = O.init_dictionary(k, f)
    W_ref, H_ref, errs_ref, _ = O.fit_transform(V.copy(), k=k, max_iter=6, tol=0, H0=H0)
    out = {}
    for panels, limit in (("one", None), ("three", 128 * 192 * 8)):
        with _native.Engine(n, f, k, mode=mode, scratch_limit=limit) as e:
            e.set_stacked_blocks(blocks, coefs)
            assert e.is_hybrid()
            e.set_dictionary(H0)
            e.init_coefficients()
            errs, _ = e.run(6, 0.0, True)
            out[panels] = (e.get_coefficients(), e.get_dictionary(), np.asarray(errs), e.error())
    for a, b in zip(out["three"][:3], (W_ref, H_ref, errs_ref)):
        within("vs_oracle", cases.rel_fro(np.atleast_2d(a), np.atleast_2d(b)), TOL[mode])
    same = 1e-12 if mode == "fp64" else 6e-5                       # the panels change the order of the numerator sums only (measured 1.9e-5)
    within("W_vs_one_panel", cases.rel_fro(out["three"][0], out["one"][0]) + 1e-300, same)
    within("H_vs_one_panel", cases.rel_fro(out["three"][1], out["one"][1]) + 1e-300, same)
    within("objective_after", abs(out["three"][3] - out["one"][3]) / abs(out["one"][3]) + 1e-300, same)


@pytest.mark.parametrize("mode", ["fp64", "tf32r"])
def test_hybrid_against_the_reference_golden_learner(hybrid_on, golden, within, mode):
    """tests/golden/learner_small.npz was written by the REAL reference's MultimodalLearner (oracle/make_golden.py): a
    dense motion modality next to a CSR sound one, i.e. its all-sparse stack (learner.py:53-56).  The hybrid stack must
    give that dictionary and those coefficients."""
    g = golden("learner_small")
    mot, snd, coefs = cases.learner_small()
    lr = MultimodalLearner(['motion', 'sound'], [mot.shape[1], snd.shape[1]], coefs, 8, mode=mode)
    np.random.seed(3)
    lr.train([mot, snd.copy()], 20)
    tol = TOL[mode]
    within("dico", cases.rel_fro(lr.dico, g["dico"]), TOL_GOLDEN[mode])
    lr.dico = np.array(g["dico"])                                  # the same dictionary on both sides from here on
    both = lr.reconstruct_internal_multi(['motion', 'sound'], [mot[:25], snd[:25].copy()], 15)
    within("internal_both", cases.rel_fro(both, g["internal_both"]), tol)
    m2s = lr.modality_to_modality('motion', 'sound', mot[:25], 15)
    within("motion_to_sound", cases.rel_fro(m2s, g["motion_to_sound"]), tol)


def test_hybrid_cfg2_full_shapes_learner_vs_oracle(hybrid_on, within):
    """BASELINE.json configs[1] at its real shapes in the hybrid form: 1000 samples, the 450-column
    dense motion modality through the contraction engine, the 110 000-column CSR sound modality through the sparse
    passes, k = 50; training iterations and the coefficients of test samples from BOTH modalities against the float64
    oracle on the reference's all-sparse stack (3 iterations keep the oracle's runtime in seconds)."""
    motion, sound, coefs = cases.cfg2_inputs()
    mods, dims = ['motion', 'sound'], [motion.shape[1], sound.shape[1]]
    lr = MultimodalLearner(mods, dims, coefs, 50)
    np.random.seed(3)
    lr.train([motion, sound.copy()], 3)
    ref = O.Learner(mods, dims, coefs, 50)
    np.random.seed(3)
    ref.train([motion, sound.copy()], 3)
    assert lr.dico.shape == (50, 110450)
    within("dico", cases.rel_fro(lr.dico, ref.dico), TOL["tf32r"])
    ref.dico = np.array(lr.dico)
    test = [motion[:40], sound[:40]]
    internal = lr.reconstruct_internal_multi(mods, test, 5)
    within("internal", cases.rel_fro(internal, ref.reconstruct_internal_multi(mods, [motion[:40], sound[:40].copy()], 5)), TOL["tf32r"])
