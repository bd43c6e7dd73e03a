"""Host-side measures with the reference's names and meanings (reference: multimodal/lib/metrics.py:15-20 and
:58-86).  They are small-array utilities for callers and tests -- the estimator's own objective
(`KLdivNMF.error`) is computed on the GPU, and `multimodal_b200.evaluation` recognises these callables by
identity and evaluates them on the GPU (csrc/evaluation.cu) without the n_test x n_ex x d broadcast."""
import numpy as np

EPSILON = 1.e-8


def generalized_KL(x, y, eps=EPSILON, axis=None):
    """sum of x log((x+eps)/(y+eps)) - x + y."""
    ratio = (x + eps) / (y + eps)
    terms = x * np.log(ratio) - x + y
    return terms.sum(axis=axis)


def _unit_sum_inplace(v, axis):
    v /= np.expand_dims(v.sum(axis=axis), axis)


def kl_div(a, b, axis=-1, eps=EPSILON, normalize=False):
    """KL(a || b) along `axis`; `normalize` rescales both arguments IN PLACE first.  As in the reference the `eps`
    argument is accepted and not used: the module constant is."""
    if normalize:
        _unit_sum_inplace(a, axis)
        _unit_sum_inplace(b, axis)
    return generalized_KL(a, b, eps=EPSILON, axis=axis)


def rev_kl_div(a, b, **kwargs):
    """KL(b || a)."""
    return kl_div(b, a, **kwargs)


def sym_kl_div(*args, **kwargs):
    """mean of the two directions."""
    forward = kl_div(*args, **kwargs)
    backward = rev_kl_div(*args, **kwargs)
    return .5 * (forward + backward)


def frobenius(a, b, axis=-1):
    """Euclidean distance along `axis`."""
    diff = a - b
    return np.sqrt((diff * diff).sum(axis=axis))


def cosine_similarity(a, b, axis=-1):
    """<a, b> / (|a| |b|), and 0 when either vector is all zero (the dot product is then exactly 0 and the
    boolean guard adds 1 to the denominator)."""
    dot = (a * b).sum(axis=axis)
    norms = np.sqrt((a * a).sum(axis=axis) * (b * b).sum(axis=axis))
    return dot / (norms + (dot == 0))


def cosine_diff(a, b, axis=-1):
    return -cosine_similarity(a, b, axis=axis)


DEVICE_MEASURES = {kl_div: "kl_div", rev_kl_div: "rev_kl_div", sym_kl_div: "sym_kl_div", frobenius: "frobenius",
                   cosine_diff: "cosine_diff"}
