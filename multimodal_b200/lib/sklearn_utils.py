"""Input validation with the reference's names and error behaviour
(reference: multimodal/lib/sklearn_utils.py:59-110)."""
import numpy as np
from scipy import sparse


def assert_all_finite(X):
    if X.dtype.char in np.typecodes['AllFloat'] and not np.isfinite(X.sum()) \
            and not np.isfinite(X.data if sparse.issparse(X) else X).all():
        raise ValueError("array contains NaN or infinity")


def array2d(X, dtype=None, order=None, copy=False):
    if sparse.issparse(X):
        raise TypeError('A sparse matrix was passed, but dense data '
                        'is required. Use X.todense() to convert to dense.')
    X_2d = np.asarray(np.atleast_2d(X), dtype=dtype, order=order)
    if X is X_2d and copy:
        X_2d = np.copy(X_2d, order='K')
    return X_2d


def atleast2d_or_csr(X, dtype=None, order=None, copy=False, check_finite=True):
    """Like the reference (sklearn_utils.py:83-97).  `check_finite=False` lets the
    estimator run the same finite test on the device copy instead of a host pass."""
    if sparse.issparse(X):
        if dtype is None or X.dtype == dtype:
            X = X.tocsr()
        else:
            X = sparse.csr_matrix(X, dtype=dtype)
    else:
        X = array2d(X, dtype=dtype, order=order, copy=copy)
    if check_finite:
        assert_all_finite(X)
    return X
