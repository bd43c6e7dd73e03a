"""Array helpers on the path (reference: multimodal/lib/array_utils.py:5-22)."""
import numpy as np
import scipy.sparse as sp


def safe_hstack(blocks):
    """Modality concatenation; any sparse block makes the whole stack sparse
    (reference array_utils.py:5-9)."""
    if any([sp.issparse(b) for b in blocks]):
        return sp.hstack(blocks)
    else:
        return np.hstack(blocks)


def normalize_sum(a, axis=0, eps=1.e-16):
    """a / (eps + sum(a, axis)) (reference array_utils.py:19-22)."""
    if axis >= len(a.shape):
        raise ValueError
    return a / (eps + np.expand_dims(np.sum(a, axis=axis), axis))
