"""CPU oracle for the KL-NMF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 numpy/scipy restatement of the reference algorithm
(omangin/multimodal, `multimodal/lib/nmf.py` + `multimodal/learner.py`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product package
`multimodal_b200` never does (it fails loudly when the CUDA library is absent).

Parity status: PINNED.  `oracle/make_golden.py` runs the *real* reference
(imported from /root/reference in the build container) and the vectors it wrote
to `tests/golden/*.npz` are compared with this restatement in
`tests/test_oracle.py` (bit-for-bit identical iteration order, so agreement is
~1e-15), together with the reference's own known answers
(`tests/test_metrics.py:48-54`, `tests/test_array_utils.py:31-41`,
`tests/test_nmf_kl.py:56-68`).

Letters follow the reference: X n x f data, W n x k coefficients, H k x f
dictionary (`components_`).
"""

import numpy as np
import scipy.sparse as sp

EPS = 1.e-8          # literal default of _update/_Q/error (nmf.py:232,297,325)
NORM_EPS = 1.e-16    # normalize_sum default (array_utils.py:19)


# --------------------------------------------------------------------------- #
# primitives
# --------------------------------------------------------------------------- #

def generalized_KL(x, y, eps=EPS, axis=None):
    """metrics.py:18-20."""
    return (np.multiply(x, np.log(np.divide(x + eps, y + eps))) - x + y
            ).sum(axis=axis)


def normalize_sum(a, axis=0, eps=NORM_EPS):
    """array_utils.py:19-22."""
    if axis >= len(a.shape):
        raise ValueError
    return a / (eps + np.expand_dims(np.sum(a, axis=axis), axis))


def scale(matrix, factors, axis=0):
    """nmf.py:29-49 (`_scale`)."""
    if not (len(matrix.shape) == 2):
        raise ValueError("Wrong array shape: %s" % str(matrix.shape))
    if axis not in (0, 1):
        raise ValueError("Wrong axis")
    factors = np.squeeze(np.asarray(factors))
    if axis == 1:
        factors = factors[:, np.newaxis]
    return np.multiply(matrix, factors)


def sddmm(a, b, refmat):
    """nmf.py:52-70 (`_special_sparse_dot`): (a @ b) sampled on refmat's
    non-zeros, CSR with refmat's structure.  Mutates refmat
    (eliminate_zeros) exactly like the reference.  Row-blocked so the
    temporaries stay small; the arithmetic per entry is the same sum over k.
    """
    refmat.eliminate_zeros()
    refmat = refmat.tocsr()
    indptr, indices = refmat.indptr, refmat.indices
    out = np.empty(indices.shape[0], dtype=np.float64)
    bt = np.ascontiguousarray(b.T)
    n = refmat.shape[0]
    step = max(1, int(4e6 // max(1, a.shape[1])))
    row_of = np.repeat(np.arange(n), np.diff(indptr))
    for s in range(0, indices.shape[0], step):
        e = min(indices.shape[0], s + step)
        out[s:e] = np.multiply(a[row_of[s:e], :], bt[indices[s:e], :]).sum(axis=1)
    return sp.csr_matrix((out, indices.copy(), indptr.copy()), shape=refmat.shape)


def as_input(X):
    """atleast2d_or_csr + check_non_negative (sklearn_utils.py:83-97,
    nmf.py:23-26)."""
    if sp.issparse(X):
        X = X.tocsr()
        data = X.data
    else:
        X = np.asarray(np.atleast_2d(X))
        data = X
    if data.dtype.kind == 'f' and not np.isfinite(data.sum()) \
            and not np.isfinite(data).all():
        raise ValueError("array contains NaN or infinity")
    if (data < 0).any():
        raise ValueError("Negative values in data passed to NMF.fit")
    return X


# --------------------------------------------------------------------------- #
# one iteration (nmf.py:232-257, 297-351)
# --------------------------------------------------------------------------- #

def error(X, W, H, eps=EPS):
    """nmf.py:297-310."""
    if sp.issparse(X):
        WH = sddmm(W, H, X)
        WH_sum = np.sum(np.multiply(np.sum(W, axis=0), np.sum(H, axis=1)))
        return (np.multiply(X.data, np.log(np.divide(X.data + eps, WH.data + eps)))
                ).sum() - X.data.sum() + WH_sum
    return generalized_KL(X, np.dot(W, H))


def ratio(X, W, H, eps=EPS):
    """nmf.py:325-336 (`_Q`): dense (X+eps)/(WH+eps) everywhere; sparse only on
    the stored non-zeros (structural zeros stay zero)."""
    if sp.issparse(X):
        WH = sddmm(W, H, X)
        WH.data = (X.data + eps) / (WH.data + eps)
        return WH
    return np.divide(X + eps, np.dot(W, H) + eps)


def updated_W(W, H, Q):
    """nmf.py:338-343: no denominator."""
    if sp.issparse(Q):
        return np.multiply(W, Q @ H.T)
    return np.multiply(W, np.dot(Q, H.T))


def updated_H(W_new, H, Q):
    """nmf.py:345-351: uses the NEW W with the STALE Q, then row-normalises."""
    if sp.issparse(Q):
        num = np.asarray((Q.T @ W_new).T)
    else:
        num = np.dot(W_new.T, Q)
    return normalize_sum(np.multiply(H, num), axis=1)


def update(X, W, H, fit=True, eps=EPS):
    """nmf.py:232-257 (the dead scale_W branch omitted)."""
    Q = ratio(X, W, H, eps)
    W = updated_W(W, H, Q)
    if fit:
        H = updated_H(W, H, Q)
    return W, H


def init_dictionary(k, f):
    """nmf.py:150-151: draws from the GLOBAL legacy numpy RNG."""
    return normalize_sum(np.abs(np.random.random((k, f))) + .01, axis=1)


def fit_transform(X, k=None, max_iter=200, tol=1e-6, H0=None, fit=True):
    """nmf.py:159-230.  Returns (W, H, errors, n_iter).  errors[i] is the
    objective *before* update i+1; on a break the pre-update W, H are returned.
    """
    X = as_input(X)
    n, f = X.shape
    if not k:
        k = f
    if H0 is None:
        H0 = init_dictionary(k, f)
    assert H0.shape == (k, f)
    H = H0
    W = X.dot(H0.T)
    W = np.asarray(W)
    prev = np.inf
    tol_abs = tol * n * f
    errors = []
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        e = error(X, W, H)
        if prev - e < tol_abs:
            break
        prev = e
        errors.append(e)
        W, H = update(X, W, H, fit=fit)
    return W, H, errors, n_iter


# --------------------------------------------------------------------------- #
# sample-sharded iteration: what the multi-GPU path must equal
# --------------------------------------------------------------------------- #

def row_partition(n, world):
    """Contiguous row blocks; rank r owns [bounds[r], bounds[r+1])."""
    base, rem = divmod(n, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def sharded_update(X_shards, W_shards, H, allreduce=None):
    """One fit iteration on row shards: local Q, W update, local numerator
    W_new^T Q; ONE all-reduce of the k x f numerator; replicated H update.
    `allreduce(arr)` sums over ranks (defaults to summing the list given)."""
    nums, W_new = [], []
    for Xs, Ws in zip(X_shards, W_shards):
        Q = ratio(Xs, Ws, H)
        Wn = updated_W(Ws, H, Q)
        W_new.append(Wn)
        if sp.issparse(Q):
            nums.append(np.asarray((Q.T @ Wn).T))
        else:
            nums.append(np.dot(Wn.T, Q))
    num = allreduce(nums[0]) if allreduce is not None else sum(nums)
    return W_new, normalize_sum(np.multiply(H, num), axis=1)


# --------------------------------------------------------------------------- #
# learner (learner.py)
# --------------------------------------------------------------------------- #

def safe_hstack(blocks):
    """array_utils.py:5-9."""
    if any(sp.issparse(b) for b in blocks):
        return sp.hstack(blocks)
    return np.hstack(blocks)


def fit_coefficients(data_obs, dictionary, iter_nmf=100):
    """learner.py:11-15: transform with a fixed dictionary, tol=0."""
    W, _, _, _ = fit_transform(data_obs, k=dictionary.shape[0], max_iter=iter_nmf,
                               tol=0, H0=dictionary, fit=False)
    return W


class Learner(object):
    """learner.py:18-94, state reduced to what the path needs."""

    def __init__(self, modalities, dimensions, coefficients, k):
        self.mod, self.dim, self.coef, self.k = modalities, dimensions, coefficients, k
        self.dico = None

    def get_index(self, m):
        return self.mod.index(m)

    def get_axis_range(self, m):
        i = self.get_index(m)
        start = sum(self.dim[:i])
        return start, start + self.dim[i]

    def get_dico(self, m=None):
        if m is None:
            return self.dico
        a, b = self.get_axis_range(m)
        return self.dico[:, a:b]

    def get_stacked_dicos(self, mods):
        return safe_hstack([self.get_dico(m) for m in mods])

    def stack_data(self, mods, mats):
        return safe_hstack([self.coef[self.get_index(m)] * x for m, x in zip(mods, mats)])

    def train(self, mats, iterations):
        V = self.stack_data(self.mod, mats)
        _, H, _, _ = fit_transform(V, k=self.k, max_iter=iterations, tol=0)
        self.dico = H

    def reconstruct_internal_multi(self, mods, mats, iterations):
        return fit_coefficients(self.stack_data(mods, mats),
                                self.get_stacked_dicos(mods), iter_nmf=iterations)

    def modalities_to_modalities(self, orig, dest, mats, iterations):
        internal = self.reconstruct_internal_multi(orig, mats, iterations)
        return internal.dot(self.get_stacked_dicos(dest))
