"""CPU: the disk formats either side of the path (multimodal_b200/store.py).

tests/golden/logger_store.{json,npz} was written by the reference's own `Logger.save` (oracle/make_store_golden.py):
the dictionary store must read it, and must write files the reference would read back the same way."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from multimodal_b200 import store
from oracle import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "logger_store")


def test_reads_the_reference_logger_files():
    dicos = cases.store_dictionaries()
    assert np.array_equal(store.load_run_dictionary(GOLDEN), dicos[-1])          # get_last_value('dictionary')
    assert np.array_equal(store.load_run_dictionary(GOLDEN, run=0), dicos[0])
    assert np.array_equal(store.load_run_dictionary(GOLDEN, run=1, key='train'), [0, 1, 2, 4])   # a json-stored value
    with pytest.raises(KeyError):
        store.load_run_dictionary(GOLDEN, key='no-such-key')
    with pytest.raises(IndexError):
        store.load_run_dictionary(GOLDEN, run=2)


def test_writes_what_the_reference_logger_writes(tmp_path):
    out = str(tmp_path / "store")
    extra = [{'train': [0, 1, 2, 3 + r], 'test': [4, 5]} for r in range(2)]
    store.save_run_dictionaries(out, cases.store_dictionaries(),
                                glob={'sample-pairing': np.arange(12).reshape(6, 2), 'k': 4}, extra=extra)
    ours, ref = json.load(open(out + ".json")), json.load(open(GOLDEN + ".json"))
    assert ours['glob'] == ref['glob'] and ours['exps'] == ref['exps'] and ours['has_np'] == ref['has_np']
    assert sorted(ours['exp_keys']) == sorted(ref['exp_keys'])        # a set in the reference: order is not defined
    assert sorted(ours['result_keys']) == sorted(ref['result_keys'])
    with np.load(out + ".npz") as a, np.load(GOLDEN + ".npz") as b:
        assert sorted(a.files) == sorted(b.files)
        for k in b.files:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k])


def test_attach_dictionary_checks_the_shape():
    class L(object):
        k, dim, dico = 4, [4, 5], None
    lr = store.attach_dictionary(L(), GOLDEN)
    assert lr.dico.shape == (4, 9) and lr.dico.flags.c_contiguous and lr.dico.dtype == np.float64
    L.dim = [4, 4]
    with pytest.raises(AssertionError):
        store.attach_dictionary(L(), GOLDEN)


def test_feature_loaders(tmp_path):
    from scipy.io import savemat
    rs = np.random.RandomState(0)
    hac = sp.random(30, 200, density=0.05, random_state=rs, format='csc')
    hac.data = np.ceil(5 * hac.data)
    savemat(str(tmp_path / "feat.mat"), {'hac': hac})
    X = store.load_mat_features(str(tmp_path / "feat.mat"), pinned=False)
    assert sp.isspmatrix_csr(X) and X.shape == (30, 200)
    assert X.data.dtype == np.float32 and X.indices.dtype == np.int32
    assert X.has_sorted_indices and np.array_equal(X.toarray(), hac.toarray())
    motion = rs.dirichlet(0.1 * np.ones(45), 30)
    np.savez(str(tmp_path / "motion.npz"), Xmotion=motion)
    M = store.load_npz_features(str(tmp_path / "motion.npz"), pinned=False)
    assert M.dtype == np.float32 and M.flags.c_contiguous and np.allclose(M, motion, rtol=1e-7)
    # pinned=True degrades to plain host memory without a CUDA device and still returns the same values
    assert np.array_equal(store.load_npz_features(str(tmp_path / "motion.npz")), M)
