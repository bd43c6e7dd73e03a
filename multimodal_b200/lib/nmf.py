"""KL-divergence NMF estimator, B200-native.

Drop-in for `multimodal/lib/nmf.py` of omangin/multimodal: same class, method names,
signatures, attributes and quirks (cited inline as nmf.py:LINE); the arithmetic runs in
libklnmf's sm_100a kernels (tcgen05 split-TF32 / TF32, or DMMA FP64) -- there is no
numpy fallback, and importing this module on a machine without the CUDA library or a
B200 raises as soon as an estimator method needs the device.

Extra, keyword-only constructor arguments (not in the reference): `mode` in
{"tf32r" (default), "tf32x3", "tf32", "fp64"} or the KLNMF_MODE environment variable
(see _native.MODES / DEFAULT_MODE and DESIGN.md section 2 for what each computes and its
stated tolerance), and `device` (CUDA ordinal, or a list of ordinals to shard the samples
over several GPUs of the box).
"""
import sys

import numpy as np
import scipy.sparse as sp

from .. import _native
from .array_utils import normalize_sum, StackedBlocks, MixedBlocks
from .sklearn_utils import atleast2d_or_csr

_EPS = 1.e-8      # the loop's literal eps (nmf.py:232, 297, 325); the kernels hard-wire it


def check_non_negative(X, whom):
    """nmf.py:23-26."""
    X = X.data if sp.issparse(X) else X
    if (X < 0).any():
        raise ValueError("Negative values in data passed to %s" % whom)


def _scale(matrix, factors, axis=0):
    """Scales lines or columns of a matrix (nmf.py:29-49)."""
    if not (len(matrix.shape) == 2):
        raise ValueError(
            "Wrong array shape: %s, should have only 2 dimensions."
            % str(matrix.shape))
    if axis not in (0, 1):
        raise ValueError('Wrong axis, should be 0 (scaling lines)\
                or 1 (scaling columns).')
    factors = np.squeeze(np.asarray(factors))
    if axis == 1:
        factors = factors[:, np.newaxis]
    return np.multiply(matrix, factors)


def _canonical_csr(X):
    """The reference's SDDMM calls eliminate_zeros() on the caller's matrix
    (nmf.py:66) -- same mutation here.  Duplicates are undefined behaviour in the
    reference; they are summed on a private copy."""
    X.eliminate_zeros()
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()
    return X


def _first_device(device):
    """`device` is a CUDA ordinal or a list of ordinals (sample sharding, distributed.DeviceGroup)."""
    if isinstance(device, (list, tuple)):
        return int(device[0])
    return int(device)


def _is_device_features(X):
    from ..store import DeviceFeatures
    return isinstance(X, DeviceFeatures)


def _load(eng, X):
    """Hand X (dense array, CSR matrix or a stack of dense modality blocks) to an engine."""
    if isinstance(X, StackedBlocks):
        eng.set_dense_blocks(X.blocks, X.coefs)
    elif isinstance(X, MixedBlocks):
        eng.set_stacked_blocks(X.blocks, X.coefs)
    elif sp.issparse(X):
        eng.set_csr(X)
    else:
        eng.set_dense(X)


def _engine_for(X, k, mode, device):
    """Context holding X (validated on the device with the reference's messages)."""
    n, f = X.shape
    if _is_device_features(X):      # resident on the GPU already (store.DeviceFeatures): a device-side copy, no upload
        if _native.resolve_mode(mode) != X.mode:
            raise ValueError("the device-resident features were uploaded for another arithmetic mode")
        return X.view(k)
    eng = _native.Engine(n, f, k, mode=mode, device=_first_device(device))
    try:
        _load(eng, X)
    except Exception:
        eng.close()
        raise
    return eng


def _host_prepare(X):
    """atleast2d_or_csr (nmf.py:193) and the dtype the device copy is made from; no O(n f) scan on the host."""
    if isinstance(X, MixedBlocks):
        return X.canonical()
    if not isinstance(X, StackedBlocks) and not _is_device_features(X):
        X = atleast2d_or_csr(X, check_finite=False)
        if sp.issparse(X):
            X = _canonical_csr(X)
        elif X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
    return X


def _raise_invalid(neg, bad, whom):
    """The reference's messages (sklearn_utils.py:59-69, nmf.py:23-26); non-finite wins, as atleast2d_or_csr runs first."""
    if bad:
        raise ValueError("array contains NaN or infinity")
    if neg:
        raise ValueError("Negative values in data passed to %s" % whom)


def _validated(X, whom, mode, device, k):
    """atleast2d_or_csr + check_non_negative (nmf.py:193-194), with the O(n f) scans run
    on the device copy.  Returns (X, engine)."""
    X = _host_prepare(X)
    eng = _engine_for(X, k, mode, device)
    neg, bad = eng.check_input()
    if bad or neg:
        eng.close()
        _raise_invalid(neg, bad, whom)
    return X, eng


def _special_sparse_dot(a, b, refmat, mode=None, device=0):
    """(a.b) on the non-zeros of refmat, CSR with refmat's structure (nmf.py:52-70)."""
    refmat.eliminate_zeros()
    ref = refmat.tocsr()
    if not ref.has_canonical_format:
        ref = ref.copy()
        ref.sum_duplicates()
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    with _engine_for(ref, a.shape[1], mode, device) as eng:
        eng.set_dictionary(b)
        eng.set_coefficients(a)
        vals = eng.sddmm(ref.nnz)
    return sp.csr_matrix((vals, ref.indices.copy(), ref.indptr.copy()), shape=ref.shape)


class KLdivNMF(object):
    """Non negative factorization with Kullback Leibler divergence cost
    (Lee & Seung multiplicative updates) -- see the reference class nmf.py:73-134.

    Parameters are the reference's (nmf.py:136-145): `subit` and `random_state` are
    accepted and, as there, never read; `eps` is only read by `scale()`.
    """

    def __init__(self, n_components=None, tol=1e-6, max_iter=200, eps=1.e-8,
                 subit=10, random_state=None, *, mode=None, device=0, checkpoint=None):
        self.n_components = n_components
        self._init_dictionary = None
        self.random_state = random_state
        self.tol = tol
        self.max_iter = max_iter
        self.eps = eps
        self.subit = subit
        self.mode = mode
        self.device = device
        self.checkpoint = checkpoint     # store.DictionaryCheckpoint: write the dictionary every so many fit iterations

    # -- initialisation (nmf.py:147-157) ----------------------------------------------
    def _draw_dictionary(self, n_features):
        if self._init_dictionary is None:
            # GLOBAL legacy numpy RNG, float64 -- bit-identical H0 with the reference
            return normalize_sum(np.abs(np.random.random(
                (self.n_components, n_features))) + .01, axis=1)
        assert(self._init_dictionary.shape ==
               (self.n_components, n_features))
        return self._init_dictionary

    def _init(self, X):
        X = atleast2d_or_csr(X, check_finite=False)
        if sp.issparse(X):
            X = _canonical_csr(X)
        H_init = self._draw_dictionary(X.shape[1])
        with _engine_for(X, self.n_components, self.mode, self.device) as eng:
            eng.set_dictionary(H_init)
            eng.init_coefficients()
            W_init = eng.get_coefficients()
        return W_init, H_init

    # -- fit / transform (nmf.py:159-291) ------------------------------------------------
    def fit_transform(self, X, y=None, weights=1., _fit=True,
                      return_errors=False, scale_W=False):
        """Learn a NMF model for the data X and return the transformed data.

        `y`, `weights` and `scale_W` are accepted and ignored exactly as in the
        reference (nmf.py:222 never forwards scale_W).
        """
        Xv = X if (isinstance(X, (StackedBlocks, MixedBlocks)) or _is_device_features(X)) \
            else atleast2d_or_csr(X, check_finite=False)
        n_samples, n_features = Xv.shape
        if not self.n_components:
            self.n_components = n_features
        if isinstance(self.device, (list, tuple)) and len(self.device) > 1:
            return self._fit_transform_sharded(Xv, _fit, return_errors)
        Xv, eng = _validated(Xv, "NMF.fit", self.mode, self.device, self.n_components)
        try:
            H_init = self._draw_dictionary(n_features)
            eng.set_dictionary(H_init)
            eng.init_coefficients()                       # W0 = X.H0^T   (nmf.py:156)
            if _fit:
                self.components_ = H_init
            elif self.components_ is not H_init:
                # reference quirk: W0 comes from H_init, the loop uses self.components_
                H_loop = np.asarray(self.components_)
                assert H_loop.shape == (self.n_components, n_features)
                eng.set_dictionary(H_loop)
            if self.max_iter < 1:
                # the reference dies here with an unbound `n_iter` (nmf.py:224)
                raise NameError("name 'n_iter' is not defined (max_iter < 1)")
            tol = self.tol * n_samples * n_features
            errors, n_iter = self._run(eng, tol, _fit)
            # fit() throws the coefficients away (nmf.py:259-273): do not bring them back from the device for that
            W = None if getattr(self, "_discard_coefficients", False) else eng.get_coefficients()
            if _fit:
                self.components_ = eng.get_dictionary()
        finally:
            eng.close()
        if n_iter == self.max_iter and tol > 0:
            sys.stderr.write("Warning: Iteration limit reached during fit\n")
        if return_errors:
            return W, [e for e in errors]
        return W

    def _run(self, eng, tol, _fit):
        """The loop (nmf.py:212-222), in one piece -- or, with a dictionary checkpoint, in pieces of `every` iterations
        with the stop test carried across them (klnmf_run_resume)."""
        ck = self.checkpoint if _fit else None
        if ck is None:
            return eng.run(self.max_iter, tol, _fit)
        errors, done, prev = [], 0, float("inf")
        while done < self.max_iter:
            chunk = min(ck.every, self.max_iter - done)
            errs, _ = eng.run(chunk, tol, _fit, prev)
            errors.extend(errs)
            done += len(errs)
            if len(errs):
                prev = errs[-1]
                ck.write(eng.get_dictionary(), done, prev)
            if len(errs) < chunk:                       # the stop test fired inside this piece
                return np.asarray(errors), done + 1
        return np.asarray(errors), self.max_iter

    def _fit_transform_sharded(self, Xv, _fit, return_errors):
        """The same call with the samples sharded over `self.device` (a list of CUDA ordinals): one engine and one
        host thread per GPU, the k x f numerator and the objective partials all-reduced over NCCL every fit
        iteration; identical state transitions and return values as the single-device path above."""
        from ..distributed import DeviceGroup
        if self.checkpoint is not None or _is_device_features(Xv):
            raise NotImplementedError("dictionary checkpoints and device-resident features are single-device features")
        n_samples, n_features = Xv.shape
        Xv = _host_prepare(Xv)
        H_init = self._draw_dictionary(n_features)
        H_loop = None
        if _fit:
            self.components_ = H_init
        elif self.components_ is not H_init:
            H_loop = np.asarray(self.components_)
            assert H_loop.shape == (self.n_components, n_features)
        if self.max_iter < 1:
            raise NameError("name 'n_iter' is not defined (max_iter < 1)")
        tol = self.tol * n_samples * n_features
        group = DeviceGroup(self.device, mode=self.mode)
        W, H, errors, n_iter, (neg, bad) = group.run(
            Xv, self.n_components, H_init, H_loop, self.max_iter, tol, _fit,
            want_coefficients=not getattr(self, "_discard_coefficients", False), set_data=_load)
        _raise_invalid(neg, bad, "NMF.fit")
        if _fit:
            self.components_ = H
        if n_iter == self.max_iter and tol > 0:
            sys.stderr.write("Warning: Iteration limit reached during fit\n")
        if return_errors:
            return W, [e for e in errors]
        return W

    def _update(self, X, W, _fit=True, scale_W=False, eps=1.e-8):
        """One update iteration (nmf.py:232-257); updates `components_` if _fit."""
        self._need_default_eps(eps)
        X = atleast2d_or_csr(X, check_finite=False)
        if sp.issparse(X):
            X = _canonical_csr(X)
        W = np.asarray(W, dtype=np.float64)
        if scale_W:
            W = _scale(normalize_sum(W, axis=1), np.asarray(X.sum(axis=1)).ravel(), axis=1)
        H = np.asarray(self.components_, dtype=np.float64)
        with _engine_for(X, H.shape[0], self.mode, self.device) as eng:
            eng.set_dictionary(H)
            eng.set_coefficients(W)
            eng.run(1, -np.inf, _fit)             # tol=-inf: never stops before the update
            W = eng.get_coefficients()
            if _fit:
                self.components_ = eng.get_dictionary()
        return W

    def fit(self, X, y=None, **params):
        self._discard_coefficients = True
        try:
            self.fit_transform(X, **params)
        finally:
            self._discard_coefficients = False
        return self

    def transform(self, X, **params):
        """nmf.py:275-291 -- note the sticky `_init_dictionary` side effect."""
        self._init_dictionary = self.components_
        params['_fit'] = False
        return self.fit_transform(X, **params)

    # -- objective (nmf.py:297-310) ---------------------------------------------------------
    def error(self, X, W, H=None, weights=1., eps=1.e-8):
        self._need_default_eps(eps)
        X = atleast2d_or_csr(X)
        if H is None:
            H = self.components_
        if sp.issparse(X):
            X = _canonical_csr(X)
        H = np.asarray(H, dtype=np.float64)
        with _engine_for(X, H.shape[0], self.mode, self.device) as eng:
            eng.set_dictionary(H)
            eng.set_coefficients(np.asarray(W, dtype=np.float64))
            return eng.error()

    # -- projections (nmf.py:314-321) -----------------------------------------------------------
    def scale(self, W, H, factors):
        safe_factors = factors + self.eps
        s_W = _scale(W, safe_factors, axis=0)
        s_H = _scale(H, 1. / safe_factors, axis=1)
        return s_W, s_H

    # -- update rules, kept as classmethods for API parity (nmf.py:323-351) -----------------------
    @staticmethod
    def _need_default_eps(eps):
        if eps != _EPS:
            raise ValueError("the CUDA kernels hard-wire eps=1e-8, the value the reference's "
                             "loop always uses (nmf.py:232); got eps=%r" % (eps,))

    @classmethod
    def _Q(cls, X, W, H, eps=1.e-8, mode=None, device=0):
        """(X+eps)/(WH+eps): dense everywhere, CSR only on the stored entries."""
        cls._need_default_eps(eps)
        X = atleast2d_or_csr(X, check_finite=False)
        H = np.asarray(H, dtype=np.float64)
        if sp.issparse(X):
            X = _canonical_csr(X)
        with _engine_for(X, H.shape[0], mode, device) as eng:
            eng.set_dictionary(H)
            eng.set_coefficients(np.asarray(W, dtype=np.float64))
            if sp.issparse(X):
                vals = eng.ratio(X.nnz)
                return sp.csr_matrix((vals, X.indices.copy(), X.indptr.copy()), shape=X.shape)
            return eng.ratio()

    @classmethod
    def _updated_W(cls, X, W, H, weights=1., Q=None, eps=1.e-8, mode=None, device=0):
        cls._need_default_eps(eps)
        W = np.asarray(W, dtype=np.float64)
        H = np.asarray(H, dtype=np.float64)
        if Q is None:
            X = atleast2d_or_csr(X, check_finite=False)
            if sp.issparse(X):
                X = _canonical_csr(X)
            with _engine_for(X, H.shape[0], mode, device) as eng:
                eng.set_dictionary(H)
                eng.set_coefficients(W)
                eng.run(1, -np.inf, False)
                return eng.get_coefficients()
        # explicit Q: W (.) (Q.H^T); the product is the engine's W0 = X.H0^T with X := Q
        Qm = atleast2d_or_csr(Q, check_finite=False)
        with _engine_for(Qm, H.shape[0], mode, device) as eng:
            eng.set_dictionary(H)
            eng.init_coefficients()
            return np.multiply(W, eng.get_coefficients())

    @classmethod
    def _updated_H(cls, X, W, H, weights=1., Q=None, eps=1.e-8, mode=None, device=0):
        cls._need_default_eps(eps)
        W = np.asarray(W, dtype=np.float64)
        H = np.asarray(H, dtype=np.float64)
        if Q is None:
            X = atleast2d_or_csr(X, check_finite=False)
            if sp.issparse(X):
                X = _canonical_csr(X)
            with _engine_for(X, H.shape[0], mode, device) as eng:
                eng.set_dictionary(H)
                eng.set_coefficients(W)
                eng.dictionary_step()
                return eng.get_dictionary()
        # explicit Q: W^T.Q = (Q^T.W)^T -- again the engine's X.H0^T with X := Q^T, H0 := W^T
        Qt = atleast2d_or_csr(Q.T if not sp.issparse(Q) else Q.T.tocsr(), check_finite=False)
        if not sp.issparse(Qt):
            Qt = np.ascontiguousarray(Qt)
        with _engine_for(Qt, W.shape[1], mode, device) as eng:
            eng.set_dictionary(np.ascontiguousarray(W.T))
            eng.init_coefficients()
            num = eng.get_coefficients().T
        return normalize_sum(np.multiply(H, num), axis=1)
