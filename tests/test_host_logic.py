"""Host-side logic: the reference-API helpers and the world_size-2 sharding plumbing (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from multimodal_b200 import distributed as D
from multimodal_b200.lib import nmf as M
from multimodal_b200.lib.array_utils import normalize_sum, safe_hstack
from multimodal_b200.lib.metrics import generalized_KL
from multimodal_b200.lib.sklearn_utils import atleast2d_or_csr
from oracle import cases
from oracle import klnmf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    assert D.shard_bounds(10, 3) == [0, 4, 7, 10] == O.row_partition(10, 3)
    assert D.shard_bounds(4_000_000, 8)[-1] == 4_000_000
    assert D.shard_bounds(2, 4) == [0, 1, 2, 2, 2]


def test_reference_helpers_known_answers():
    # reference tests/test_array_utils.py:31-41, tests/test_metrics.py:48-54, tests/test_nmf_kl.py:25-68
    a = np.array([[0., 1., 3.], [2., 3., 3.]])
    assert np.all(normalize_sum(a, axis=1) == np.array([[0., .25, .75], [.25, .375, .375]]))
    x = np.zeros((4, 2)); x[1, 1] = 1
    np.testing.assert_array_almost_equal(generalized_KL(x, .5 * np.ones((4, 2))), np.log(2.) + 3.)
    m = np.array([[1, 2, 3], [4, 5, 6]])
    np.testing.assert_array_almost_equal(M._scale(m, np.array([2, 3]), axis=1), [[2, 4, 6], [12, 15, 18]])
    for bad in [lambda: M._scale(np.zeros((3, 4)), np.zeros(4), axis=3),
                lambda: M._scale(np.zeros((3, 4, 6)), np.zeros(3), axis=1),
                lambda: M._scale(np.zeros((3,)), np.zeros(3), axis=1),
                lambda: M._scale(np.zeros((3, 4)), np.zeros(2), axis=1)]:
        with pytest.raises(ValueError):
            bad()
    with pytest.raises(ValueError):
        M.check_non_negative(np.array([1., -2.]), "NMF.fit")
    assert sp.issparse(safe_hstack([np.ones((2, 2)), sp.csr_matrix(np.ones((2, 3)))]))
    assert atleast2d_or_csr(np.matrix([[1., 2.]])).__class__ is np.ndarray
    with pytest.raises(ValueError):
        atleast2d_or_csr(np.array([[np.inf, 1.]]))


def test_estimator_constructor_matches_reference_defaults():
    e = M.KLdivNMF()
    assert (e.n_components, e.tol, e.max_iter, e.eps, e.subit, e.random_state) == (None, 1e-6, 200, 1e-8, 10, None)
    assert e._init_dictionary is None
    s_W, s_H = M.KLdivNMF(eps=0.).scale(np.ones((2, 3)), np.ones((3, 4)), np.array([1., 2., 4.]))
    assert np.allclose(s_W[0], [1, 2, 4]) and np.allclose(s_H[:, 0], [1, .5, .25])


# ---- world_size 2 over gloo --------------------------------------------------------------------

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch
    from multimodal_b200 import _native, distributed as DD
    from oracle import cases as C, klnmf_oracle as OO
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # (1) the 128-byte NCCL id travels from rank 0 (NCCL itself is not needed for the plumbing)
    _native.nccl_unique_id = lambda *a: bytes((7 * i) % 251 for i in range(128))
    uid = DD.broadcast_unique_id()
    # (2) every rank gets rank 0's host-drawn H0
    np.random.seed(100 + rank)
    H0 = DD.draw_shared_dictionary(6, 200)
    # (3) one sharded fit iteration == the unsharded oracle iteration
    X = C.cfg1_X()[:97]
    b = DD.shard_bounds(X.shape[0], world)
    Xs = X[b[rank]:b[rank + 1]]
    Ws = Xs.dot(H0.T)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    Wn, Hn = OO.sharded_update([Xs], [Ws], H0, allreduce=allreduce)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), uid=np.frombuffer(uid, dtype=np.uint8), H0=H0, W=Wn[0], H=Hn)
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(r0["uid"], r1["uid"]) and len(r0["uid"]) == 128
    assert np.array_equal(r0["H0"], r1["H0"])
    np.random.seed(100)
    assert np.array_equal(r0["H0"], O.init_dictionary(6, 200))          # bit-identical with nmf.py:150-151
    X = cases.cfg1_X()[:97]
    Wref, Href = O.update(X, X.dot(r0["H0"].T), r0["H0"])
    assert cases.rel_fro(np.vstack([r0["W"], r1["W"]]), Wref) < 1e-14
    assert cases.rel_fro(r0["H"], Href) < 1e-13 and np.array_equal(r0["H"], r1["H"])


def test_measures_known_answers():
    # reference tests/test_metrics.py:48-54 (KL known answer) and the zero guard of cosine_similarity (metrics.py:73-78)
    from multimodal_b200.lib import metrics as M
    x, y = np.array([1., 2., 0.]), np.array([2., 2., 3.])
    assert abs(M.generalized_KL(x, y) - (np.log(.5) + 1. + 3.)) < 1e-7
    assert abs(M.kl_div(x, y) - M.generalized_KL(x, y)) == 0
    assert M.rev_kl_div(x, y) == M.kl_div(y, x)
    assert M.sym_kl_div(x, y) == .5 * (M.kl_div(x, y) + M.kl_div(y, x))
    assert M.frobenius(np.array([3., 0.]), np.array([0., 4.])) == 5.0
    assert M.cosine_similarity(np.zeros(3), y) == 0.0 and M.cosine_diff(y, y) == pytest.approx(-1.0)
    a = np.array([[1., 3.]])
    M.kl_div(a, a.copy(), normalize=True)
    assert np.allclose(a, [[.25, .75]])          # normalisation happens in place, as in the reference


def test_evaluation_label_bookkeeping():
    # host-side helpers of the evaluation mirror (reference evaluation.py:8-91): pure list / small-array logic
    from multimodal_b200 import evaluation as E
    labels = [0, 1, 2, 0, 1, 2, 2]
    assert sorted(E.chose_examples(labels, number=2)) == [0, 1, 2, 3, 4, 5]
    assert sorted(E.chose_examples(labels, label_set={2}, number=3)) == [2, 5, 6]
    with pytest.raises(ValueError):
        E.chose_examples(labels, label_set={0}, number=3)
    reco = np.array([[.9, .1, .8], [.2, .7, .1]])
    true = np.array([[1, 0, 1], [1, 0, 0]])
    assert list(E.compare_labels_given_nb(reco, true)) == [True, False]
    assert E.score_labels_given_nb(reco, true) == .5
    assert list(E.compare_labels_given_nb(reco[0], true[0])) == [True]
    assert E.score_labels_threshold(reco, true, .5) == .5
    assert E.evaluate_label_reco(reco, [0, 1]) == 1.0
    d = np.array([[3., 1., 1.], [0., 5., 5.]])
    assert E.dists_to_found_labels(d, ['a', 'b', 'c']) == ['b', 'a']          # first minimum, like np.argmin
    assert E.found_labels_to_score(['b', 'x'], ['b', 'a']) == .5
    conf = E.found_labels_to_confusion([0, 0, 1], [0, 0, 1], 2)
    assert conf[0, 0] == 1 and conf[1, 1] == 1                                 # repeated pairs count once (reference quirk)
    with pytest.raises(TypeError):
        E._measure_key(lambda a, b, axis=-1: 0)


def test_stack_data_of_dense_modalities_is_lazy_and_equals_the_reference_formula():
    # learner.py:53-56: safe_hstack([c * m]); dense-only stacks are formed on the device (StackedBlocks)
    from multimodal_b200.learner import MultimodalLearner
    from multimodal_b200.lib.array_utils import StackedBlocks
    rs = np.random.RandomState(0)
    a, b, c = rs.random_sample((7, 3)), rs.random_sample((7, 5)).astype(np.float32), rs.randint(0, 4, (7, 2))
    lr = MultimodalLearner(['a', 'b', 'c'], [3, 5, 2], [2., .5, 3.], 4)
    st = lr.stack_data(['a', 'b', 'c'], [a, b, c])
    assert isinstance(st, StackedBlocks) and st.shape == (7, 10) and st.dtype == np.float64
    np.testing.assert_array_equal(np.asarray(st), np.hstack([2. * a, .5 * b, 3. * c]))
    sub = lr.stack_data(['c', 'a'], [c, a])                      # modality order of the call, coefficients by name
    np.testing.assert_array_equal(sub.toarray(), np.hstack([3. * c, 2. * a]))
    assert isinstance(lr.stack_data(['b'], [b]), np.ndarray)     # one modality: a plain scaled copy, as before
    # any sparse block makes the reference's stack sparse (array_utils.py:5-9); here the blocks stay apart until the
    # device builds that CSR matrix, and `tocsr()` is the reference's formula
    from multimodal_b200.lib.array_utils import MixedBlocks
    mixed = lr.stack_data(['a', 'b'], [a, sp.csr_matrix(b)])
    assert isinstance(mixed, MixedBlocks) and mixed.shape == (7, 8)
    ref = sp.hstack([2. * a, .5 * sp.csr_matrix(b)]).tocsr()
    assert sp.issparse(mixed.tocsr()) and np.array_equal(mixed.toarray(), ref.toarray())
    assert sp.issparse(lr.stack_data(['b'], [sp.csr_matrix(b)]))   # one sparse modality: a plain scaled matrix


def test_device_group_host_helpers():
    # the host side of KLdivNMF(device=[...]) (distributed.DeviceGroup): row blocks of every input kind, the thread
    # helper's error propagation (the cause wins over the BrokenBarrierError it gives the other threads)
    import threading
    from multimodal_b200 import distributed as D
    from multimodal_b200.lib.array_utils import StackedBlocks
    rs = np.random.RandomState(1)
    a, b = rs.random_sample((9, 3)), rs.random_sample((9, 2)).astype(np.float32)
    st = D._rows(StackedBlocks([a, b], [2., .5]), 4, 9)
    assert isinstance(st, StackedBlocks) and st.shape == (5, 5)
    np.testing.assert_array_equal(st.toarray(), np.hstack([2. * a[4:9], .5 * b[4:9]]))
    assert D._rows(a, 2, 5).base is a                                   # dense: a view, no copy
    Xs = sp.csr_matrix(a)
    np.testing.assert_array_equal(D._rows(Xs, 3, 7).toarray(), a[3:7])
    assert D.shard_bounds(1001, 2) == [0, 501, 1001] and D.shard_bounds(3, 8)[-1] == 3
    # speed-proportional shards: a GPU 5 % slower gets 5 % fewer samples; the blocks still tile [0, n)
    bw = D.shard_bounds(4000000, 4, weights=[1.0, 0.95, 1.02, 0.97])
    assert bw[0] == 0 and bw[-1] == 4000000 and all(b1 > b0 for b0, b1 in zip(bw, bw[1:]))
    rows = np.diff(bw)
    np.testing.assert_allclose(rows / rows[0], [1.0, 0.95, 1.02, 0.97], rtol=1e-5)
    assert D.shard_bounds(10, 3, weights=[1, 1, 1]) == [0, 3, 7, 10]
    assert D._in_threads([lambda: 1, lambda: 2]) == [1, 2]
    gate = threading.Barrier(2)

    def boom():
        gate.abort()
        raise MemoryError("cause")

    def waits():
        gate.wait()
    with pytest.raises(MemoryError, match="cause"):
        D._in_threads([waits, boom])
