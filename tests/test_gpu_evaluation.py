"""-m gpu: nearest-example evaluation on the device (csrc/evaluation.cu) against the reference's formulas
(evaluation.py:103-116 with metrics.py:58-86), restated in numpy on the broadcast arrays."""
import numpy as np
import pytest
import scipy.sparse as sp

from multimodal_b200 import evaluation as ev
from multimodal_b200.lib import metrics as M

pytestmark = pytest.mark.gpu

MEASURES = [M.kl_div, M.rev_kl_div, M.sym_kl_div, M.frobenius, M.cosine_diff]


def broadcast(reco, ex, measure):
    return measure(reco[:, np.newaxis, :], ex[np.newaxis, :, :], axis=-1)


@pytest.mark.parametrize("measure", MEASURES, ids=lambda m: m.__name__)
@pytest.mark.parametrize("nt,ne,d", [(1, 1, 1), (37, 53, 50), (200, 17, 333), (16, 16, 32), (5, 40, 7)])
def test_all_distances_match_the_broadcast_form(measure, nt, ne, d):
    rs = np.random.RandomState(nt + ne + d)
    A, B = rs.random_sample((nt, d)), rs.random_sample((ne, d))
    A[rs.random_sample(A.shape) < 0.3] = 0.0
    B[rs.random_sample(B.shape) < 0.3] = 0.0
    if nt > 3:
        A[2] = 0.0                      # an all-zero vector: cosine_diff must return -0 / (0 + 1) = 0
    ref = broadcast(A, B, measure)
    got = ev.all_distances(A, B, measure)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13)
    labels = list(range(100, 100 + ne))
    assert ev.classify_NN(A, B, labels, measure) == [labels[m] for m in np.argmin(ref, axis=1)]


def test_ties_take_the_first_example_like_argmin():
    A = np.array([[1., 2., 3.]])
    B = np.array([[9., 9., 9.], [1., 2., 3.], [1., 2., 3.]] * 7)
    assert ev.classify_NN(A, B, list(range(len(B))), M.frobenius) == [1]


def test_sparse_inputs_and_scores():
    rs = np.random.RandomState(0)
    ex = sp.csr_matrix(rs.random_sample((12, 30)) * (rs.random_sample((12, 30)) < 0.4))
    reco = ex[[3, 7, 7, 1]].toarray() + 1e-3
    labels = [i % 4 for i in range(12)]
    found = ev.classify_NN(reco, ex, labels, M.cosine_diff)
    assert found == [labels[3], labels[7], labels[7], labels[1]]
    assert ev.evaluate_NN_label(reco, ex, [labels[3], labels[7], 99, labels[1]], labels, M.cosine_diff) == 0.75
    D = ev.all_distances(reco, ex, M.kl_div)
    assert ev.scores_from_dists(D, [labels[3], labels[7], labels[7], labels[1]], labels) == 1.0


def test_unknown_measure_is_refused():
    with pytest.raises(TypeError):
        ev.all_distances(np.ones((2, 2)), np.ones((2, 2)), lambda a, b, axis: 0)
