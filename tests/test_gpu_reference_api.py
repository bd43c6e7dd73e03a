"""-m gpu: the reference's own unit tests for the path (tests/test_nmf_kl.py), re-stated
against the drop-in module -- same class names, same assertions."""
import numpy as np
import pytest
import scipy.sparse as sp
from numpy.testing import assert_array_almost_equal

from multimodal_b200.lib import nmf
from multimodal_b200.lib.metrics import generalized_KL

pytestmark = pytest.mark.gpu


def random_NN_matrix(shape):
    return np.abs(np.random.random(shape))


def random_NN_sparse(h, w, density):
    r = sp.rand(h, w, density)
    r.data = np.abs(r.data)
    return r


def is_NN(a):
    return np.all(a >= 0)


@pytest.fixture(params=["fp64", "tf32x3", "tf32r"])
def mode(request):
    return request.param


class TestError:
    # reference tests/test_nmf_kl.py:71-101
    n_samples, n_components, n_features = 20, 3, 30

    def setup_method(self):
        np.random.seed(0)
        self.X = random_NN_sparse(self.n_samples, self.n_features, .1)
        self.W = random_NN_matrix((self.n_samples, self.n_components))
        self.H = random_NN_matrix((self.n_components, self.n_features))

    def _nmf(self, mode):
        e = nmf.KLdivNMF(n_components=3, tol=1e-4, max_iter=200, eps=1.e-8, subit=10, mode=mode)
        e.components_ = random_NN_matrix((self.n_components, self.n_features))
        return e

    def test_error_is_gen_kl(self, mode):
        Xdense = self.X.todense()
        err = self._nmf(mode).error(Xdense, self.W, H=self.H)
        kl = generalized_KL(np.asarray(Xdense), self.W.dot(self.H))
        assert_array_almost_equal(err, kl, decimal=6 if mode == "fp64" else 3)

    def test_error_sparse(self, mode):
        e = self._nmf(mode)
        err_dense = e.error(self.X.todense(), self.W, H=self.H)
        err_sparse = e.error(self.X, self.W, H=self.H)
        assert_array_almost_equal(err_dense, err_sparse, decimal=6 if mode == "fp64" else 3)

    def test_error_is_gen_kl_with_compenents(self, mode):
        e = self._nmf(mode)
        Xdense = self.X.todense()
        err = e.error(Xdense, self.W)
        kl = generalized_KL(np.asarray(Xdense), self.W.dot(e.components_))
        assert_array_almost_equal(err, kl, decimal=6 if mode == "fp64" else 3)


class TestUpdates:
    # reference tests/test_nmf_kl.py:104-134
    n_samples, n_components, n_features = 20, 3, 30
    sparse = False

    def setup_method(self):
        np.random.seed(1)
        if self.sparse:
            self.X = random_NN_sparse(self.n_samples, self.n_features, .5).tocsr()
        else:
            self.X = random_NN_matrix((self.n_samples, self.n_features))
        self.W = random_NN_matrix((self.n_samples, self.n_components))
        self.H = random_NN_matrix((self.n_components, self.n_features))

    def _nmf(self, mode):
        e = nmf.KLdivNMF(n_components=3, tol=1e-4, max_iter=200, eps=1.e-8, subit=10, mode=mode)
        e.components_ = self.H
        return e

    def test_W_remains_NN(self, mode):
        assert is_NN(self._nmf(mode)._updated_W(self.X, self.W, self.H, mode=mode))

    def test_H_remains_NN(self, mode):
        assert is_NN(self._nmf(mode)._updated_H(self.X, self.W, self.H, mode=mode))

    def test_decreases_KL(self, mode):
        e = self._nmf(mode)
        dkl_prev = e.error(self.X, self.W)
        W = e._update(self.X, self.W, _fit=True)
        dkl_next = e.error(self.X, W)
        assert dkl_prev > dkl_next

    def test_no_compenents_update(self, mode):
        e = self._nmf(mode)
        e._update(self.X, self.W, _fit=False)
        assert (e.components_ == self.H).all()

    def test_building_blocks_agree_with_the_fused_update(self, mode):
        # Q -> _updated_W(Q=Q) -> _updated_H(W_new, Q=Q) is what _update fuses (nmf.py:251-256)
        e = self._nmf(mode)
        Q = e._Q(self.X, self.W, self.H, mode=mode)
        Wn = e._updated_W(self.X, self.W, self.H, Q=Q, mode=mode)
        Hn = e._updated_H(self.X, Wn, self.H, Q=Q, mode=mode)
        W2 = e._update(self.X, self.W, _fit=True)
        # the _Q hook keeps the plain, unrounded ratio; the loop of tf32r contracts the rounded, centered one: 1.7e-4
        tol = 1e-10 if mode == "fp64" else (5e-4 if mode == "tf32r" else 1e-4)
        assert np.linalg.norm(Wn - W2) <= tol * np.linalg.norm(W2)
        assert np.linalg.norm(Hn - e.components_) <= tol * np.linalg.norm(Hn)


class TestSparseUpdates(TestUpdates):
    # reference tests/test_nmf_kl.py:137-147
    sparse = True


class TestFitTransform:
    # reference tests/test_nmf_kl.py:150-172
    def _nmf(self, mode):
        return nmf.KLdivNMF(n_components=3, tol=1e-6, max_iter=200, eps=1.e-8, subit=10, mode=mode)

    def test_cv(self, mode):
        np.random.seed(2)
        X = random_NN_matrix((10, 5))
        W, errors = self._nmf(mode).fit_transform(X, return_errors=True)
        assert abs(errors[-1] - errors[-2]) < errors[0] * 1.e-2

    def test_zero_error_on_fact_data(self, mode):
        np.random.seed(3)
        X = np.dot(random_NN_matrix((5, 2)), random_NN_matrix((2, 3)))
        W, errors = self._nmf(mode).fit_transform(X, return_errors=True)
        assert errors[-1] < errors[0] * 1.e-3

    def test_no_compenents_update(self, mode):
        np.random.seed(4)
        components = random_NN_matrix((3, 5))
        e = self._nmf(mode)
        e.components_ = components
        e.fit_transform(random_NN_matrix((10, 5)), components, _fit=False)
        assert (e.components_ == components).all()


class TestSparseDot:
    # reference tests/test_nmf_kl.py:175-192
    def setup_method(self):
        np.random.seed(5)
        self.ref = sp.rand(5, 6, .3).tocsr()
        self.a = np.random.random((5, 7))
        self.b = np.random.random((7, 6))

    def test_indices(self, mode):
        ab = nmf._special_sparse_dot(self.a, self.b, self.ref, mode=mode)
        assert (ab.indptr == self.ref.indptr).all() and (ab.indices == self.ref.indices).all()

    def test_correct(self, mode):
        ok = np.multiply(np.dot(self.a, self.b), (self.ref.toarray() != 0))
        ans = nmf._special_sparse_dot(self.a, self.b, self.ref, mode=mode).toarray()
        assert_array_almost_equal(ans, ok, decimal=6 if mode == "fp64" else 5)


class TestInputGuards:
    # nmf.py:193-194, sklearn_utils.py:59-69
    def test_negative(self, mode):
        with pytest.raises(ValueError, match="Negative values in data passed to NMF.fit"):
            nmf.KLdivNMF(n_components=2, max_iter=2, mode=mode).fit(np.array([[1., -1.], [1., 1.]]))

    def test_non_finite(self, mode):
        with pytest.raises(ValueError, match="array contains NaN or infinity"):
            nmf.KLdivNMF(n_components=2, max_iter=2, mode=mode).fit(np.array([[1., np.nan], [1., 1.]]))

    def test_dictionary_shape_assert(self, mode):
        e = nmf.KLdivNMF(n_components=2, max_iter=2, mode=mode)
        e.components_ = np.ones((3, 4))
        with pytest.raises(AssertionError):
            e.transform(np.ones((5, 4)))

    def test_float32_and_int_inputs_give_float64(self, mode):
        np.random.seed(6)
        X = (np.random.random((12, 7)) * 5).astype(np.float32)
        W = nmf.KLdivNMF(n_components=2, max_iter=3, tol=0, mode=mode).fit_transform(X)
        assert W.dtype == np.float64
        W = nmf.KLdivNMF(n_components=2, max_iter=3, tol=0, mode=mode).fit_transform((X * 3).astype(np.int64))
        assert W.dtype == np.float64

    def test_explicit_zeros_are_removed_from_the_callers_matrix(self, mode):
        X = sp.csr_matrix(np.array([[1., 0., 2.], [0., 3., 0.]]))
        X.data[0] = 0.0
        nmf.KLdivNMF(n_components=2, max_iter=2, tol=0, mode=mode).fit(X)
        assert X.nnz == 2                      # nmf.py:66 side effect


class TestStoreOnDevice:
    """SURVEY 8f-4 on the GPU: device-resident features (one upload, several estimator calls) and dictionary
    checkpoints written while a fit runs, in the reference logger's layout (logger.py:79-136)."""

    def test_device_resident_features_equal_host_arrays(self, tmp_path):
        from multimodal_b200 import store
        rs = np.random.RandomState(5)
        motion = rs.dirichlet(0.1 * np.ones(45), 130)
        np.savez(str(tmp_path / "motion.npz"), Xmotion=motion)
        host = store.load_npz_features(str(tmp_path / "motion.npz"), pinned=False)
        with store.load_npz_features_device(str(tmp_path / "motion.npz")) as dev:
            assert dev.shape == (130, 45)
            outs = []
            for X in (host, dev, dev):                     # the resident copy serves any number of calls
                est = nmf.KLdivNMF(n_components=6, max_iter=15, tol=0)
                np.random.seed(8)
                W = est.fit_transform(X)
                outs.append((W, est.components_))
            for W, H in outs[1:]:
                assert np.array_equal(W, outs[0][0]) and np.array_equal(H, outs[0][1])
            est = nmf.KLdivNMF(n_components=6, max_iter=10, tol=0)
            est.components_ = outs[0][1]
            assert np.array_equal(est.transform(dev), est.transform(host))
            with pytest.raises(ValueError):
                nmf.KLdivNMF(n_components=6, max_iter=2, mode="fp64").fit(dev)      # uploaded for another mode
        bad = motion.copy()
        bad[3, 3] = -1.0
        with store.to_device(bad) as dev:
            with pytest.raises(ValueError, match="Negative values"):
                nmf.KLdivNMF(n_components=2, max_iter=2).fit(dev)

    @pytest.mark.parametrize("tol", [0, 1e-4])
    def test_checkpointed_fit_is_the_same_fit(self, tmp_path, tol, mode):
        from multimodal_b200 import store
        rs = np.random.RandomState(6)
        X = rs.gamma(0.5, 1.0, size=(150, 80))
        plain = nmf.KLdivNMF(n_components=5, max_iter=60, tol=tol, mode=mode)
        np.random.seed(9)
        W0, e0 = plain.fit_transform(X, return_errors=True)
        ck = store.DictionaryCheckpoint(str(tmp_path / "ck"), every=7)
        est = nmf.KLdivNMF(n_components=5, max_iter=60, tol=tol, mode=mode, checkpoint=ck)
        np.random.seed(9)
        W1, e1 = est.fit_transform(X, return_errors=True)
        # the pieces are the same iterations and the same stop decision; two runs of the same fit differ by the order
        # of their atomic sums only (FP64 objective partials, FP32 numerator partials in the TF32 modes)
        same = {"fp64": 1e-11, "tf32x3": 1e-5, "tf32r": 2e-4}[mode]
        assert abs(len(e1) - len(e0)) <= (0 if mode == "fp64" or tol == 0 else 2)
        m = min(len(e0), len(e1))
        np.testing.assert_allclose(np.asarray(e1)[:m], np.asarray(e0)[:m], rtol=same)
        if len(e1) == len(e0):
            assert np.linalg.norm(W1 - W0) <= same * np.linalg.norm(W0)
            assert np.linalg.norm(est.components_ - plain.components_) <= same * np.linalg.norm(plain.components_)
        back = ck.resume()
        assert back['iterations_done'] == len(e1) and back['objective'] == e1[-1]
        assert np.array_equal(back['dictionary'], est.components_)
        assert np.array_equal(store.load_run_dictionary(str(tmp_path / "ck")), est.components_)   # the reference's reader
        # a restarted fit picks the dictionary up: its coefficients start again from W0 = X.H^T (nmf.py:156), whose scale
        # is off whatever the dictionary -- but one update later it is ahead of the cold start, and a few iterations
        # bring it back to where the fit stood
        again = nmf.KLdivNMF(n_components=5, max_iter=10, tol=0, mode=mode)
        again._init_dictionary = back['dictionary']
        np.random.seed(1)
        _, e2 = again.fit_transform(X, return_errors=True)
        assert e2[1] < e1[1] and e2[-1] <= e1[-1] * 1.02
