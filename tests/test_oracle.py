"""Pin the CPU oracle against (a) the reference's own known answers and (b) the
golden vectors produced by running the real reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import cases
from oracle import klnmf_oracle as O

TIGHT = 1e-12


def _fit(X, k, iters, seed, **kw):
    np.random.seed(seed)
    return O.fit_transform(X, k=k, max_iter=iters, tol=kw.pop("tol", 0), **kw)


# ---- the reference's unit-test known answers --------------------------------

def test_kl_known_answer():
    # reference tests/test_metrics.py:48-54
    x = np.zeros((4, 2)); x[1, 1] = 1
    y = .5 * np.ones((4, 2))
    np.testing.assert_array_almost_equal(O.generalized_KL(x, y), np.log(2.) + 3.)


def test_kl_properties():
    # reference tests/test_metrics.py:30-46
    rs = np.random.RandomState(42)
    x, y = rs.random_sample((10, 15)), rs.random_sample((10, 15))
    assert O.generalized_KL(x, y) >= 0
    np.testing.assert_array_almost_equal(O.generalized_KL(x, x), 0)
    np.testing.assert_array_almost_equal(.3 * O.generalized_KL(x, y),
                                         O.generalized_KL(.3 * x, .3 * y), decimal=5)
    assert O.generalized_KL(x, y, axis=1).shape == (10,)


def test_normalize_sum_known_answers():
    # reference tests/test_array_utils.py:31-41 (exact equality there too)
    a = np.array([[0., 1., 3.], [2., 3., 3.]])
    assert np.all(O.normalize_sum(a, axis=0) == np.array([[0., .25, .5], [1., .75, .5]]))
    assert np.all(O.normalize_sum(a, axis=1) == np.array([[0., .25, .75], [.25, .375, .375]]))
    z = np.random.random((2, 4)); z[1] *= 0
    assert not np.any(np.isnan(O.normalize_sum(z, axis=1)))
    with pytest.raises(ValueError):
        O.normalize_sum(np.zeros((2, 3, 4)), axis=3)


def test_scale_known_answers():
    # reference tests/test_nmf_kl.py:56-68
    m = np.array([[1, 2, 3], [4, 5, 6]])
    np.testing.assert_array_almost_equal(O.scale(m, np.array([2, 3]), axis=1),
                                         [[2, 4, 6], [12, 15, 18]])
    np.testing.assert_array_almost_equal(O.scale(m, np.array([3, 2, 1]), axis=0),
                                         [[3, 4, 3], [12, 10, 6]])
    with pytest.raises(ValueError):
        O.scale(np.zeros((3, 4)), np.zeros(4), axis=3)
    with pytest.raises(ValueError):
        O.scale(np.zeros((3, 4, 6)), np.zeros(3), axis=1)


def test_primitives_golden(golden):
    g = golden("primitives")
    assert O.generalized_KL(np.array([1., 2.]), np.array([2., 1.])) == g["kl_known"]
    x = np.array([[1., 2.], [3., 4.]])
    assert np.array_equal(O.normalize_sum(x, axis=0), g["norm_axis0"])
    assert np.array_equal(O.normalize_sum(x, axis=1), g["norm_axis1"])


# ---- SURVEY 8c known answers captured from the running reference ------------

def test_kat_dense_survey_values(golden):
    W, H, errs, _ = _fit(cases.kat_dense(), 2, 10, 0)
    survey = [5.753249649715822, 0.7721442137111796, 0.7546991235810249, 0.728220792261248,
              0.6884958855743429, 0.6314718306102629, 0.5552480169289469, 0.46268159818644683,
              0.3624412729006138, 0.2664785316848004]
    np.testing.assert_allclose(errs, survey, rtol=1e-13)
    np.testing.assert_allclose(H, [[0.5983523500812611, 0.40164764991873897],
                                   [0.9152582047011343, 0.0847417952988657]], rtol=1e-13)
    np.testing.assert_allclose(W[0], [1.6948227329739203, 0.3051772644195952], rtol=1e-13)
    np.testing.assert_allclose(W[5], [2.1216681318026986, 4.878331870080324], rtol=1e-13)
    np.testing.assert_allclose(O.error(cases.kat_dense(), W, H), 0.18506954199762637, rtol=1e-13)
    g = golden("kat_dense")
    assert cases.rel_fro(W, g["W"]) < TIGHT and cases.rel_fro(H, g["H"]) < TIGHT
    np.testing.assert_allclose(O.error(cases.kat_dense(), W, H), g["after"], rtol=1e-13)


def test_kat_csr_survey_values(golden):
    W, H, errs, _ = _fit(cases.kat_csr(), 2, 10, 0)
    np.testing.assert_allclose(errs[0], 17.71856814634861, rtol=1e-13)
    np.testing.assert_allclose(errs[9], 5.712029130170944, rtol=1e-13)
    g = golden("kat_csr")
    np.testing.assert_allclose(errs, g["errors"], rtol=1e-13)
    assert cases.rel_fro(W, g["W"]) < TIGHT and cases.rel_fro(H, g["H"]) < TIGHT
    # the dense path on the same matrix is a DIFFERENT algorithm at X == 0
    Wd, Hd, ed, _ = _fit(cases.kat_csr().toarray(), 2, 10, 0)
    np.testing.assert_allclose(ed[9], 5.712029206342203, rtol=1e-13)
    np.testing.assert_allclose(ed, g["errors_densepath"], rtol=1e-13)
    assert cases.rel_fro(Wd, g["W_densepath"]) < TIGHT


# ---- golden vectors from the real reference ----------------------------------

@pytest.mark.parametrize("name,maker,k,seed,long_iters", [
    ("cfg1_dense", cases.cfg1_X, 10, 1, 200),
    ("ragged_dense", cases.ragged_dense_X, 13, 5, 200),
    ("zeros_dense", cases.zeros_dense_X, 8, 7, 60),
    ("sparse_mid", cases.sparse_mid_X, 16, 11, 200),
])
def test_fit_golden(golden, name, maker, k, seed, long_iters):
    g = golden(name)
    W, H, errs, _ = _fit(maker(), k, 10, seed)
    assert cases.rel_fro(W, g["W10"]) < TIGHT
    assert cases.rel_fro(H, g["H10"]) < TIGHT
    np.testing.assert_allclose(errs, g["errors10"], rtol=1e-12)
    W, H, errs, _ = _fit(maker(), k, long_iters, seed)
    np.testing.assert_allclose(errs, g["errors_long"], rtol=1e-10)
    np.testing.assert_allclose(O.error(maker(), W, H), g["final_error"], rtol=1e-10)
    assert cases.rel_fro(H, g["H_long"]) < 1e-9


def test_transform_golden(golden):
    H0 = cases.sub_dictionary(10, 200)
    W, H, errs, _ = _fit(cases.cfg1_X()[:64], 10, 30, 0, H0=H0, fit=False)
    g = golden("transform_dense")
    assert H is H0                                  # _fit=False never touches the dictionary
    assert cases.rel_fro(W, g["W"]) < TIGHT
    np.testing.assert_allclose(errs, g["errors"], rtol=1e-12)
    Xs = cases.sparse_mid_X()[:50]
    W, _, errs, _ = _fit(Xs, 16, 30, 0, H0=cases.sub_dictionary(16, Xs.shape[1]), fit=False)
    g = golden("transform_sparse")
    assert cases.rel_fro(W, g["W"]) < TIGHT
    np.testing.assert_allclose(errs, g["errors"], rtol=1e-12)


def test_early_stop_golden(golden):
    g = golden("early_stop")
    W, H, errs, n_iter = _fit(cases.cfg1_X()[:120, :60], 6, 500, 2, tol=1e-5)
    assert len(errs) == len(g["errors"]) and n_iter == len(errs) + 1
    assert cases.rel_fro(W, g["W"]) < 1e-10 and cases.rel_fro(H, g["H"]) < 1e-10


def test_rise_with_tol0_golden(golden):
    # nmf.py:215 with tol = 0: the loop breaks when the objective rises; the reference does so at the second evaluation
    g = golden("rise_tol0")
    X, H0 = cases.rise_case()
    W, H, errs, n_iter = O.fit_transform(X, k=H0.shape[0], max_iter=50, tol=0, H0=H0)
    assert n_iter == 2 and len(errs) == len(g["errors"]) == 1
    np.testing.assert_allclose(errs, g["errors"], rtol=TIGHT)
    assert cases.rel_fro(W, g["W"]) < TIGHT and cases.rel_fro(H, g["H"]) < TIGHT
    assert O.error(X, W, H) > 1.5 * errs[0]          # the state handed back is the one whose objective rose


def test_learner_golden(golden):
    g = golden("learner_small")
    mot, snd, coefs = cases.learner_small()
    lr = O.Learner(['motion', 'sound'], [mot.shape[1], snd.shape[1]], coefs, 8)
    np.random.seed(3)
    lr.train([mot, snd.copy()], 20)
    assert cases.rel_fro(lr.dico, g["dico"]) < TIGHT
    assert cases.rel_fro(lr.reconstruct_internal_multi(['sound'], [snd[:25].copy()], 15),
                         g["internal_sound"]) < TIGHT
    assert cases.rel_fro(lr.reconstruct_internal_multi(['motion'], [mot[:25]], 15),
                         g["internal_motion"]) < TIGHT
    assert cases.rel_fro(lr.modalities_to_modalities(['motion'], ['sound'], [mot[:25]], 15),
                         g["motion_to_sound"]) < TIGHT
    assert cases.rel_fro(
        lr.reconstruct_internal_multi(['motion', 'sound'], [mot[:25], snd[:25].copy()], 15),
        g["internal_both"]) < TIGHT


# ---- properties the reference tests (tests/test_nmf_kl.py) -------------------

def test_sparse_error_equals_dense_error():
    # reference tests/test_nmf_kl.py:92-95
    rs = np.random.RandomState(1)
    X = sp.random(20, 30, density=.5, random_state=rs, format='csr')
    W, H = rs.random_sample((20, 5)), rs.random_sample((5, 30))
    np.testing.assert_array_almost_equal(O.error(X, W, H), O.error(X.toarray(), W, H), decimal=6)


def test_sddmm_structure_and_values():
    # reference tests/test_nmf_kl.py:175-192
    rs = np.random.RandomState(2)
    ref = sp.random(12, 17, density=.3, random_state=rs, format='csr')
    a, b = rs.random_sample((12, 4)), rs.random_sample((4, 17))
    out = O.sddmm(a, b, ref)
    assert np.array_equal(out.indptr, ref.indptr) and np.array_equal(out.indices, ref.indices)
    np.testing.assert_array_almost_equal(out.toarray(), a.dot(b) * (ref.toarray() != 0))


@pytest.mark.parametrize("sparse", [False, True])
def test_update_properties(sparse):
    # reference tests/test_nmf_kl.py:118-134
    rs = np.random.RandomState(3)
    X = np.abs(rs.random_sample((10, 15)))
    if sparse:
        X = sp.csr_matrix(X * (rs.random_sample((10, 15)) < .5))
    W, H = np.abs(rs.random_sample((10, 4))), np.abs(rs.random_sample((4, 15)))
    e0 = O.error(X, W, H)
    W1, H1 = O.update(X, W, H, fit=True)
    assert (W1 >= 0).all() and (H1 >= 0).all()
    assert O.error(X, W1, H1) < e0
    W2, H2 = O.update(X, W, H, fit=False)
    assert H2 is H


def test_input_guards():
    with pytest.raises(ValueError):
        O.as_input(np.array([[1., -1.]]))
    with pytest.raises(ValueError):
        O.as_input(np.array([[1., np.nan]]))


def test_sharded_update_equals_unsharded():
    X = cases.cfg1_X()[:97]
    np.random.seed(0)
    H = O.init_dictionary(6, X.shape[1])
    W = X.dot(H.T)
    Wref, Href = O.update(X, W, H)
    b = O.row_partition(X.shape[0], 3)
    assert b == [0, 33, 65, 97]
    Ws, Hs = O.sharded_update([X[b[i]:b[i + 1]] for i in range(3)],
                              [W[b[i]:b[i + 1]] for i in range(3)], H)
    assert cases.rel_fro(np.vstack(Ws), Wref) < 1e-14 and cases.rel_fro(Hs, Href) < 1e-13


# ---- the size-independent properties tests/test_gpu_properties.py relies on, established on the oracle itself ----
@pytest.mark.parametrize("sparse", [False, True], ids=["dense", "csr"])
def test_oracle_scale_equivariance(sparse):
    # X -> cX gives W -> cW and the same dictionary (W0 = X.H0^T scales, the ratio does not, up to eps)
    rs = np.random.RandomState(8)
    X = rs.gamma(0.7, 1.0, size=(60, 40)) + 0.05
    if sparse:
        X[rs.random_sample(X.shape) < 0.8] = 0.0
        X = sp.csr_matrix(X)
    outs = []
    for c in (1.0, 8.0):
        np.random.seed(2)
        W, H, _, _ = O.fit_transform(X * c, k=5, max_iter=8, tol=0)
        outs.append((np.asarray(W), np.asarray(H)))
    assert cases.rel_fro(outs[1][0], 8.0 * outs[0][0]) < 1e-6
    assert cases.rel_fro(outs[1][1], outs[0][1]) < 1e-6


def test_oracle_rows_are_independent_given_the_dictionary():
    # what sample sharding and the row-subset GPU test rest on
    rs = np.random.RandomState(3)
    X = rs.gamma(0.5, 1.0, size=(90, 30))
    np.random.seed(4)
    H = O.init_dictionary(6, 30)
    W = np.asarray(X.dot(H.T))
    Wsub = W[10:40].copy()
    for _ in range(6):
        W, _ = O.update(X, W, H, fit=False)
        Wsub, _ = O.update(X[10:40], Wsub, H, fit=False)
    np.testing.assert_allclose(Wsub, W[10:40], rtol=1e-12)


def test_oracle_centered_ratio_identities():
    # the algebra behind the centered ratio of the split-TF32 modes:
    #   Q.H^T = (Q-1).H^T + rowsum(H)   and   W'^T.Q = W'^T.(Q-1) + colsum(W')
    rs = np.random.RandomState(5)
    X = rs.gamma(0.5, 1.0, size=(50, 35))
    np.random.seed(6)
    H = O.init_dictionary(4, 35)
    W = np.asarray(X.dot(H.T))
    Q = np.asarray(O.ratio(X, W, H))
    G = Q.dot(H.T)
    np.testing.assert_allclose((Q - 1.0).dot(H.T) + H.sum(axis=1), G, rtol=1e-12)
    Wn = W * G
    np.testing.assert_allclose(Wn.T.dot(Q - 1.0) + Wn.sum(axis=0)[:, None], Wn.T.dot(Q), rtol=1e-12)
    assert (G >= 0).all()          # a sum of non-negative terms: what the clamp in the centered epilogue restores


def test_oracle_cancellation_free_objective():
    # x log q - x + s  ==  x P(u) + (x - s)(x - s - eps)/d,  u = (x - s)/d,  d = s + eps,  P(u) = log1p(u) - u
    rs = np.random.RandomState(7)
    x = rs.gamma(0.5, 1.0, size=5000)
    x[rs.random_sample(x.size) < 0.3] = 0.0
    s = x * np.exp(rs.normal(0, 0.5, x.size)) + rs.random_sample(x.size) * 1e-3
    eps = O.EPS
    d = s + eps
    u = (x - s) / d
    lhs = x * np.log((x + eps) / d) - x + s
    rhs = x * (np.log1p(u) - u) + (x - s) * (x - s - eps) / d
    np.testing.assert_allclose(rhs, lhs, rtol=1e-9, atol=1e-12)
    assert abs(rhs.sum() - O.generalized_KL(x, s)) < 1e-9 * abs(lhs.sum())
