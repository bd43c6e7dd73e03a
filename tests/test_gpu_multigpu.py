"""-m gpu, two devices (skipped on a one-GPU box): the samples sharded over several GPUs from ONE process through the
drop-in API -- `KLdivNMF(device=[0, 1])`, `MultimodalLearner(..., device=[0, 1])` -- against the float64 oracle and
against the single-device run of the same mode (rows shard, the k x f numerator and the objective partials
are all-reduced, everything else is local; learner.py:31-41 keeps its signature)."""
import numpy as np
import pytest
import scipy.sparse as sp

from multimodal_b200.lib.nmf import KLdivNMF
from multimodal_b200.learner import MultimodalLearner
from oracle import cases, klnmf_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

DEVICES = [0, 1]
TOL = {"fp64": 1e-12, "tf32x3": 5e-6, "tf32r": 8e-4, "tf32": 2.5e-3}          # tests/test_gpu_parity.py TOL_WH


def maxrel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.mark.parametrize("mode", ["fp64", "tf32x3", "tf32r"])
@pytest.mark.parametrize("kind", ["dense", "csr"])
def test_sharded_fit_matches_oracle_and_single_device(within, mode, kind):
    rs = np.random.RandomState(12)
    n, f, k = 1001, 333, 24                  # an odd sample count: shards of 501 and 500 rows
    X = rs.gamma(0.5, 1.0, size=(n, f))
    if kind == "csr":
        X[rs.random_sample((n, f)) < 0.85] = 0.0
        X = sp.csr_matrix(X)
    np.random.seed(6)
    Wr, Hr, er, _ = O.fit_transform(X.copy(), k=k, max_iter=10, tol=0)
    out = {}
    for dev in (0, DEVICES):
        est = KLdivNMF(n_components=k, max_iter=10, tol=0, mode=mode, device=dev)
        np.random.seed(6)
        W, errs = est.fit_transform(X.copy(), return_errors=True)
        out[str(dev)] = (W, est.components_, np.asarray(errs))
    W, H, errs = out[str(DEVICES)]
    assert W.shape == (n, k) and W.dtype == np.float64 and len(errs) == 10
    tol = TOL[mode] if kind == "dense" else (1e-12 if mode == "fp64" else 2e-5)
    within("W", cases.rel_fro(W, Wr), tol)
    within("H", cases.rel_fro(H, Hr), tol)
    within("objective", maxrel(errs, er), tol)
    # sharding changes the order of the numerator's sum only
    same = {"fp64": 1e-12, "tf32x3": 1e-5, "tf32r": 2e-4}[mode]      # tf32r: the numerator's rounding follows its sum order (4e-5)
    within("W_vs_single", cases.rel_fro(W, out["0"][0]) + 1e-300, same)
    within("H_vs_single", cases.rel_fro(H, out["0"][1]) + 1e-300, same)


@pytest.mark.parametrize("chunks", ["1", "4", "8"])
def test_chunked_numerator_all_reduce(within, monkeypatch, chunks):
    """Several ranks keep the dense numerator in contiguous f-chunks and all-reduce every finished chunk on a side
    stream while the next one is contracted (api.cu: setup_num_chunks, dense_iteration).  Ragged f (chunks of unequal
    width), several row panels per shard, a shard count that leaves one shard a row short -- against the oracle and the
    one-all-reduce form."""
    from multimodal_b200 import _native
    monkeypatch.setenv("KLNMF_AR_CHUNKS", chunks)
    monkeypatch.setenv("KLNMF_AR_CHUNK_MIN_F", "64")
    rs = np.random.RandomState(14)
    n, f, k = 901, 333, 40
    X = rs.gamma(0.5, 1.0, size=(n, f))
    np.random.seed(7)
    Wr, Hr, er, _ = O.fit_transform(X, k=k, max_iter=8, tol=0)
    for mode in ("fp64", "tf32r"):
        est = KLdivNMF(n_components=k, max_iter=8, tol=0, mode=mode, device=DEVICES)
        np.random.seed(7)
        W, errs = est.fit_transform(X, return_errors=True)
        within(mode + "_W", cases.rel_fro(W, Wr), TOL[mode])
        within(mode + "_H", cases.rel_fro(est.components_, Hr), TOL[mode])
        within(mode + "_objective", maxrel(errs, er), 1e-12 if mode == "fp64" else 2e-5)
    # fewer samples than shards would leave one engine without rows: it still joins every chunk's all-reduce
    est = KLdivNMF(n_components=3, max_iter=4, tol=0, mode="fp64", device=DEVICES)
    np.random.seed(7)
    W1 = est.fit_transform(X[:1, :])
    ref = KLdivNMF(n_components=3, max_iter=4, tol=0, mode="fp64", device=0)
    np.random.seed(7)
    W0 = ref.fit_transform(X[:1, :])
    within("one_sample_W", cases.rel_fro(W1, W0) + 1e-300, 1e-12)
    within("one_sample_H", cases.rel_fro(est.components_, ref.components_) + 1e-300, 1e-12)


def test_sharded_transform_needs_no_exchange_and_matches(within):
    rs = np.random.RandomState(13)
    X = rs.random_sample((700, 400))
    np.random.seed(2)
    H = O.init_dictionary(40, 400)
    Wr = np.asarray(X.dot(H.T))
    for _ in range(8):
        Wr, _ = O.update(X, Wr, H, fit=False)
    est = KLdivNMF(n_components=40, max_iter=8, tol=0, device=DEVICES)
    est.components_ = H
    W = est.transform(X)
    assert est.components_ is H and est._init_dictionary is H
    within("W", cases.rel_fro(W, Wr), 8e-4)


def test_sharded_input_guards():
    X = np.ones((64, 8))
    X[50, 3] = -1.0                           # lives in the second shard
    with pytest.raises(ValueError, match="Negative values in data passed to NMF.fit"):
        KLdivNMF(n_components=2, max_iter=2, device=DEVICES).fit(X)
    X[50, 3] = np.nan
    with pytest.raises(ValueError, match="array contains NaN or infinity"):
        KLdivNMF(n_components=2, max_iter=2, device=DEVICES).fit(X)


def test_learner_trains_over_two_devices(golden, within):
    g = golden("learner_small")
    mot, snd, coefs = cases.learner_small()
    lr = MultimodalLearner(['motion', 'sound'], [mot.shape[1], snd.shape[1]], coefs, 8, device=DEVICES)
    np.random.seed(3)
    lr.train([mot, snd.copy()], 20)
    within("dico", cases.rel_fro(lr.dico, g["dico"]), 1e-5)
    within("internal_sound", cases.rel_fro(lr.reconstruct_internal('sound', snd[:25].copy(), 15), g["internal_sound"]), 1e-5)
    m2s = lr.modality_to_modality('motion', 'sound', mot[:25], 15)
    within("motion_to_sound", cases.rel_fro(m2s, g["motion_to_sound"]), 8e-4)
    # three dense modalities: the blocks are sliced per shard and stacked on each device
    rs = np.random.RandomState(17)
    mats = [rs.gamma(0.5, 1.0, size=(300, 70)), rs.random_sample((300, 45)).astype(np.float32), rs.poisson(1.5, size=(300, 33))]
    c3 = [1. / np.mean(np.sum(m, axis=1)) for m in mats]
    lr3 = MultimodalLearner(['sound', 'image', 'motion'], [70, 45, 33], c3, 9, device=DEVICES)
    np.random.seed(4)
    lr3.train(mats, 10)
    ref = O.Learner(['sound', 'image', 'motion'], [70, 45, 33], c3, 9)
    np.random.seed(4)
    ref.train(mats, 10)
    within("dico_dense_blocks", cases.rel_fro(lr3.dico, ref.dico), 8e-4)


@pytest.mark.parametrize("mode", ["fp64", "tf32r"])
def test_sharded_hybrid_stack_through_the_learner(within, monkeypatch, mode):
    """A mixed (dense + CSR) stack on two GPUs in its hybrid form: every shard keeps the dense modalities
    dense (distributed._hybrid_for_all_shards decides once for all), the numerators of both parts are all-reduced in
    one NCCL group; against the oracle's learner on the reference's all-sparse stack (learner.py:31-41, 53-56)."""
    monkeypatch.setenv("KLNMF_HYBRID", "1")
    rs = np.random.RandomState(41)
    n = 301
    motion = rs.dirichlet(0.1 * np.ones(96), n)
    motion[motion < 1e-3] = 0.0
    sound = sp.random(n, 400, density=0.05, random_state=rs, format='csr')
    sound.data = np.ceil(5 * sound.data)
    image = rs.random_sample((n, 70)).astype(np.float32)
    mats = [motion, sound, image]
    mods, dims = ['motion', 'sound', 'image'], [96, 400, 70]
    coefs = [1. / np.mean(np.sum(motion, axis=1)), 1. / np.mean(np.asarray(sound.sum(axis=1))), np.float32(0.5)]
    ref = O.Learner(mods, dims, coefs, 9)
    np.random.seed(5)
    ref.train([motion, sound.copy(), image], 10)
    out = {}
    for dev in (0, DEVICES):
        lr = MultimodalLearner(mods, dims, coefs, 9, mode=mode, device=dev)
        np.random.seed(5)
        lr.train(mats, 10)
        out[str(dev)] = np.array(lr.dico)
    tol = 1e-12 if mode == "fp64" else 6e-4                       # tests/test_gpu_hybrid.py
    within("dico", cases.rel_fro(out[str(DEVICES)], ref.dico), tol)
    within("dico_vs_single", cases.rel_fro(out[str(DEVICES)], out["0"]) + 1e-300, 1e-12 if mode == "fp64" else 2e-4)
